set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 2 -c 1 -o gpurun_out/prof_path_r1 python tools/profile_frame.py --frames 4 > gpurun_out/ncu_full_path.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_progressive" -s 2 -c 1 -o gpurun_out/prof_prog_r1 python tools/profile_frame.py --frames 4 > gpurun_out/ncu_full_prog.log 2>&1
ls -la gpurun_out
