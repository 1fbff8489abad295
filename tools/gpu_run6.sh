set -x
for k in 2 4; do
GDPT_MUX_K=$k timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path_mux" -s 2 -c 1 -o gpurun_out/prof_mux_k$k python tools/profile_frame.py --frames 4 > gpurun_out/ncu_mux_k$k.log 2>&1
done
