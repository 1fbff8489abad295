set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 120 python tools/profile_frame.py --frames 6 2>&1 | tee gpurun_out/frames_demo.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -5 gpurun_out/bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(trace|shade|progressive)" -c 200 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_frame.py --frames 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 8 -c 4 -o gpurun_out/prof_trace_r1 python tools/profile_frame.py --frames 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
