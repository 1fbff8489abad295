set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
GDPT_SCHEDULE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trace_parity or full_size or progressive" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_sched1.log
rm -f gpurun_out/ab.log
for cfg in "1 0 12 8 8" "1 2 24 16 8" "1 2 28 16 8" "1 2 16 16 8" "1 2 24 16 4" "1 2 24 16 16" "1 2 24 16 1" "0 2 24 16 8"; do
  set -- $cfg
  echo "== cull $1 schedule $2 refill_below $3 burst $4 shade_at $5" | tee -a gpurun_out/ab.log
  GDPT_CULL=$1 GDPT_SCHEDULE=$2 GDPT_REFILL_BELOW=$3 GDPT_BURST=$4 GDPT_SHADE_AT=$5 timeout 120 python tools/profile_frame.py --frames 5 2>&1 | tail -2 | tee -a gpurun_out/ab.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 1 -c 1 -o gpurun_out/prof_path_r4 python tools/profile_frame.py --frames 2 > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out | head -30
