python tools/prof_one.py --rank 255 --fmt 2 > gpurun_out/r2_p255_pair.log 2>&1
cat gpurun_out/r2_p255_pair.log
ncu --set full --clock-control none --import-source on -k regex:k_score_f8 -s 1 -c 1 -o gpurun_out/r2_pair_r255 -f python tools/prof_one.py --rank 255 --fmt 2 > gpurun_out/r2_ncu_pair.log 2>&1
tail -n 3 gpurun_out/r2_ncu_pair.log
