#!/bin/bash
# Turn the raw files of tools/gpu_round_artifacts.sh (gpurun_out/) into the tracked summaries under profiles/.
# usage: tools/collect_profiles.sh r01
set -e
tag=${1:-r01}
out=profiles
mkdir -p $out
cp gpurun_out/bench.json $out/${tag}_bench.json
cp gpurun_out/bench_reference.json $out/${tag}_bench_reference.json
cp gpurun_out/launches.csv $out/${tag}_launches.csv
cp gpurun_out/warp_profile.json $out/${tag}_warp_profile.json
cp gpurun_out/pytest_gpu.log $out/${tag}_pytest_gpu.log
for k in k_path k_primary_cull k_progressive; do
  ncu -i gpurun_out/prof_$k.ncu-rep --page details > $out/${tag}_ncu_${k}_details.txt 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
want=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__block_size','sm__cycles_active.avg','sm__cycles_elapsed.max')
for h,u,v in zip(hdr,units,vals):
    if h in want: print(f'{h},{u},{v}')
" > $out/${tag}_ncu_${k}_raw.csv
done
python3 tools/ncu_regions.py gpurun_out/prof_k_path.ncu-rep 40 > $out/${tag}_ncu_k_path_regions.txt
ls -la $out
for c in c1 c3 c4; do [ -f gpurun_out/bench_$c.json ] && cp gpurun_out/bench_$c.json $out/${tag}_bench_$c.json; done
[ -f gpurun_out/pytest_gpu_adversarial.log ] && cp gpurun_out/pytest_gpu_adversarial.log $out/${tag}_pytest_gpu_adversarial.log
[ -f gpurun_out/sanitizer_memcheck.log ] && cp gpurun_out/sanitizer_memcheck.log $out/${tag}_sanitizer_memcheck.log
[ -f gpurun_out/warp_profile_c4.json ] && cp gpurun_out/warp_profile_c4.json $out/${tag}_warp_profile_c4.json
true
