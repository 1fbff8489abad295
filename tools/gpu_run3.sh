set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/ab.log
for cfg in "0 0 20 4" "0 1 24 16" "1 0 20 4" "1 1 24 16" "1 1 28 16" "1 1 16 32" "1 0 28 8" "1 0 12 8"; do
  set -- $cfg
  echo "== cull $1 schedule $2 refill_below $3 burst $4" | tee -a gpurun_out/ab.log
  GDPT_CULL=$1 GDPT_SCHEDULE=$2 GDPT_REFILL_BELOW=$3 GDPT_BURST=$4 timeout 120 python tools/profile_frame.py --frames 5 2>&1 | tail -2 | tee -a gpurun_out/ab.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 8 -c 3 -o gpurun_out/prof_trace_r3 python tools/profile_frame.py --frames 2 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | head -30
