"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one optimizer step."""
import csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
seq = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"]) / 1000) for r in rows
       if "spin_kernel" not in r["Kernel Name"]]   # torch.cuda._sleep of the profiling phase
# one complete step: the shortest run between two consecutive 'proj' (first-of-step) kernels (other runs also contain the
# bench's own bookkeeping launches: state roll-back, re-pack)
starts = [i for i, s in enumerate(seq) if "prep" in s[0] or "proj_kernel" in s[0]]
pairs = list(zip(starts, starts[1:]))
a, b = min(pairs, key=lambda ab: ab[1] - ab[0]) if pairs else (0, len(seq))
step = seq[a:b]
tot = sum(s[2] for s in step)
for name, grid, us in step:
    short = name.replace("msf::<unnamed>::", "").replace("msf::", "").split("(")[0][:44]
    print("%-46s %-16s %8.1f us %5.1f%%" % (short, grid, us, 100 * us / tot))
print("TOTAL %.1f us over %d launches" % (tot, len(step)))
