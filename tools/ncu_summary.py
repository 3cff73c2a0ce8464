"""Summarise an .ncu-rep: key raw metrics + instruction mix + loop buckets (reads `ncu -i` CSV pages).
usage: python tools/ncu_summary.py file.ncu-rep [out.txt] [launch index in the report, default 0]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
sel = ["-s", sys.argv[3], "-c", "1"] if len(sys.argv) > 3 else []     # pick one launch of a multi-launch report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or "issue_stalled" in h and "per_issue_active" in h:
        print(f"{h:90s} {v} {u}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
byop, cnt, recs = collections.Counter(), collections.Counter(), []
for r in rows[2:]:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    s = r[isrc]
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    byop[op] += n
    cnt[n] += 1
    recs.append((n, op))
tot = sum(byop.values())
print(f"\ninstruction mix (warp instructions executed, total {tot}):", file=out)
for op, n in byop.most_common(16):
    print(f"  {op:10s} {n:12d} {100 * n / tot:5.1f}%", file=out)
print("\nloop buckets (exec count x #SASS instructions, share):", file=out)
for v, k in sorted(cnt.items(), key=lambda x: -x[0] * x[1])[:8]:
    mix = collections.Counter(op for n, op in recs if n == v)
    print(f"  {v:10d} x {k:4d} {100 * v * k / tot:5.1f}%  {dict(mix.most_common(10))}", file=out)
