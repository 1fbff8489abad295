# Produces the measured artefacts of a round: GPU tests, bench line, reference-arm line, ncu launch list and full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 2 -c 1 -o gpurun_out/prof_k_path python tools/profile_frame.py --frames 4 > gpurun_out/ncu_k_path.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_primary_cull" -s 2 -c 1 -o gpurun_out/prof_k_primary_cull python tools/profile_frame.py --frames 4 > gpurun_out/ncu_k_primary_cull.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_progressive" -s 2 -c 1 -o gpurun_out/prof_k_progressive python tools/profile_frame.py --frames 4 > gpurun_out/ncu_k_progressive.log 2>&1
python tools/warp_profile.py --frames 4 > gpurun_out/warp_profile.json 2>&1
ls -la gpurun_out
# the other BASELINE configs on one GPU (C1, C3, C4 at 4K) and the adversarial tier's log
bash tools/gpu_configs.sh > gpurun_out/configs.log 2>&1; tail -4 gpurun_out/configs.log
python -m pytest tests/test_gpu_adversarial.py -q -s 2>&1 | grep "re-traced\|passed\|failed" > gpurun_out/pytest_gpu_adversarial.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
