set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
GDPT_SCHEDULE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trace_parity or full_size" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_sched0.log
for cfg in "0 20 4" "1 24 16" "1 28 16" "1 16 16" "1 24 4" "1 24 64"; do
  set -- $cfg
  echo "== schedule $1 refill_below $2 burst $3" | tee -a gpurun_out/ab.log
  GDPT_SCHEDULE=$1 GDPT_REFILL_BELOW=$2 GDPT_BURST=$3 timeout 120 python tools/profile_frame.py --frames 4 2>&1 | tail -2 | tee -a gpurun_out/ab.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 8 -c 3 -o gpurun_out/prof_trace_r2 python tools/profile_frame.py --frames 2 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | head -30
