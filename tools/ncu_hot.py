"""Reads an .ncu-rep here (no GPU): headline metrics + hottest SASS lines by stall samples.
Usage: python tools/ncu_hot.py <report.ncu-rep> [n_lines]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum"]
for r in rows[2:3]:
    for i, h in enumerate(hdr):
        if h in want: print(f"{h:75s} {units[i]:12s} {r[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h and r[i] not in ("0", ""): print(f"   {h[33:]:30s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if "Address" in r and "Source" in r: h = r; 
    elif h and len(r) == len(h):
        try: data.append((int(r[h.index("# Samples")]), int(r[h.index("Instructions Executed")]), r[h.index("Source")]))
        except ValueError: pass
    if h and data and r and r[0] == "Kernel Name": break
# first kernel only
seen = len(data)
tot = sum(d[0] for d in data) or 1
print(f"\n{len(data)} SASS lines, {tot} samples; hottest:")
idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:n]
for i in sorted(idx):
    s, e, t = data[i]
    print(f"  [{i:5d}] {s:6d} {s/tot*100:5.1f}%  exec {e:9d}  {t[:100]}")
