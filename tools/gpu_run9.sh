set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 2 -c 1 -o gpurun_out/prof_k_path_fast python tools/profile_frame.py --frames 4 > gpurun_out/ncu_k_path_fast.log 2>&1
tail -3 gpurun_out/ncu_k_path_fast.log
