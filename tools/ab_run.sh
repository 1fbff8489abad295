ncu --set full --clock-control none -k regex:k_progressive -s 3 -c 1 -o gpurun_out/prof_k2_1080 -f python tools/profile_frame.py --frames 6 > gpurun_out/prof_k2.log 2>&1
ncu --set full --clock-control none -k regex:k_progressive -s 3 -c 1 -o gpurun_out/prof_k2_4k -f python tools/profile_frame.py --frames 6 --width 3840 --height 2160 > gpurun_out/prof_k2b.log 2>&1
tail -1 gpurun_out/prof_k2b.log
