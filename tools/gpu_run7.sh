set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 2 -c 1 -o gpurun_out/prof_path_r1b python tools/profile_frame.py --frames 4 > gpurun_out/ncu_full_path_b.log 2>&1
GDPT_SCHEDULE=4 GDPT_MUX_K=4 python tools/profile_frame.py --frames 4 2>&1 | tail -1
GDPT_SCHEDULE=4 GDPT_MUX_K=2 python tools/profile_frame.py --frames 4 2>&1 | tail -1
