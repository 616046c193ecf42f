set -x
python tools/prof_one.py --rank 255 --fmt 1 > gpurun_out/r2_p255_f8.log 2>&1
python tools/prof_one.py --rank 255 --fmt 0 > gpurun_out/r2_p255_f16.log 2>&1
cat gpurun_out/r2_p255_f8.log gpurun_out/r2_p255_f16.log
ncu --set full --clock-control none --import-source on -k regex:k_score_f8 -s 1 -c 1 -o gpurun_out/r2_f8_r255 -f python tools/prof_one.py --rank 255 --fmt 1 > gpurun_out/r2_ncu_f8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_lr -s 1 -c 1 -o gpurun_out/r2_lr_r255 -f python tools/prof_one.py --rank 255 --fmt 0 > gpurun_out/r2_ncu_lr.log 2>&1
tail -3 gpurun_out/r2_ncu_f8.log gpurun_out/r2_ncu_lr.log
