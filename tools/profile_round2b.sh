# Round 2, second capture (after the pruning votes): ncu --set full of the search kernel on configs 2
# and 3, the launch list of a bench run (taken after the same command had exited 0 without ncu).
set -x
cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vmvo_window_search -s 2 -c 1 -f -o gpurun_out/r02b_search_cfg2 python tools/profile_search.py config2_single_drive_10k_32x32_w30 4 > gpurun_out/ncu_cfg2.log 2>&1; tail -2 gpurun_out/ncu_cfg2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vmvo_window_search -s 1 -c 1 -f -o gpurun_out/r02b_search_cfg3 python tools/profile_search.py config3_dense_256x256_w60 3 > gpurun_out/ncu_cfg3.log 2>&1; tail -2 gpurun_out/ncu_cfg3.log
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02b_bench_for_ncu.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep
