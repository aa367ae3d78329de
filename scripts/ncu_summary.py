"""Summarise an .ncu-rep: per kernel launch duration, tensor/dram/L2 utilisation, top stall PCs.
usage: ncu_summary.py report.ncu-rep [launch-index-for-source-page]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__icc_request_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]
idx = {}
for w in want:
    for i, h in enumerate(hdr):
        if h == w or h.endswith(w):
            idx[w] = i
            break
for r in rows[2:]:
    print(" | ".join("%s=%s" % (w.split(".")[0][-28:], r[idx[w]][:40]) for w in want if w in idx))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    h = {n: i for i, n in enumerate(rd[0])}
    stalls = [n for n in rd[0] if n.startswith("stall_") and "Not Issued" not in n]
    data = []
    for r in rd[1:]:
        try:
            data.append((int(r[h["# Samples"]]), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data)
    agg = {s: sum(int(r[h[s]] or 0) for _, r in data) for s in stalls}
    print("instructions", len(data), "samples", tot)
    print(sorted(agg.items(), key=lambda x: -x[1])[:8])
    for s, r in sorted(data, key=lambda x: -x[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        top = sorted([(int(r[h[st]] or 0), st) for st in stalls], reverse=True)[:2]
        print("%6d %5.1f%% %s %-64s x%s %s" % (s, 100.0 * s / tot, r[h["Address"]][-5:], r[h["Source"]][:64], r[h["Instructions Executed"]], top))
