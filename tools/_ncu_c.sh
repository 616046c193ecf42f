set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-c2 --no-side > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_lr -s 1 -c 1 -o gpurun_out/r2_lr_r60 -f python tools/prof_one.py --rank 60 > gpurun_out/r2_ncu_lr60.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_f8 -s 1 -c 1 -o gpurun_out/r2_f8_r255b -f python tools/prof_one.py --rank 255 > gpurun_out/r2_ncu_f8_255b.log 2>&1
tail -n 2 gpurun_out/r2_ncu_lr60.log gpurun_out/r2_ncu_f8_255b.log
