set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 1 -c 1 -o gpurun_out/prof_c4_k_path python tools/profile_frame.py --scene instanced --width 1920 --height 1080 --frames 3 > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/ncu_c4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_temporal" -s 2 -c 1 -o gpurun_out/prof_k_temporal python tools/profile_frame.py --frames 4 --mode 1 > gpurun_out/ncu_k_temporal.log 2>&1
tail -3 gpurun_out/ncu_k_temporal.log
python tools/profile_frame.py --frames 6 --mode 1
