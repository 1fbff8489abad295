#!/usr/bin/env python3
"""Summarise an ncu report of a path kernel: headline counters + instructions / lanes / stall samples per code region.
usage: ncu_regions.py report.ncu-rep"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Elapsed Cycles", "SM Active Cycles", "Executed Instructions", "Issue Slots Busy", "Avg. Active Threads", "Registers Per",
        "Achieved Active Warps", "L1/TEX Hit", "L2 Hit Rate", "No Eligible", "Executed Ipc Active", "DRAM Throughput", "Theoretical Active Warps")
for ln in det.splitlines():
    if any(k in ln for k in keys) and "OPT" not in ln:
        print(" ".join(ln.split()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; cur = None; curline = None; seen = set()
per = collections.defaultdict(lambda: [0, 0, 0])
noinst = collections.Counter()
stall = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or len(r) < 8: continue
    if r[0] != '': curline = (cur, int(r[0]), r[1].strip()[:100]); continue
    if r[2] in seen: continue
    seen.add(r[2])
    try:
        per[curline][0] += int(r[iI]); per[curline][1] += int(r[iT]); per[curline][2] += int(r[iS])
    except ValueError:
        continue
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                stall[h] += int(r[i])
                if h == 'stall_no_inst': noinst[curline] += int(r[i])
            except ValueError: pass
tot = sum(v[0] for v in per.values()); tt = sum(v[1] for v in per.values()); ts = sum(v[2] for v in per.values())
print(f"warp instructions {tot}  lanes/instr {tt / max(tot, 1):.1f}  samples {ts}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for k, v in per.items():
    for i in range(3): byfile[k[0]][i] += v[i]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} inst {v[0] / tot * 100:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  samples {v[2] / max(ts, 1) * 100:5.1f}%")
print("top lines:")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"  {k[0]}:{k[1]:4d} inst {v[0] / tot * 100:5.2f}% lanes {v[1] / max(v[0], 1):5.1f} samp {v[2] / max(ts, 1) * 100:5.2f}% | {k[2]}")
s = sum(stall.values())
print("stalls:", ", ".join(f"{k[6:]} {v / s * 100:.1f}%" for k, v in stall.most_common(8)))

tn = sum(noinst.values())
byf = collections.Counter()
for k, v in noinst.items(): byf[k[0]] += v
print("no_inst samples by file:", ", ".join(f"{f} {v / max(tn, 1) * 100:.1f}%" for f, v in byf.most_common(6)))
print("no_inst top lines:")
for k, v in noinst.most_common(25):
    print(f"  {k[0]}:{k[1]:4d} {v / max(tn, 1) * 100:5.2f}% | {k[2]}")
