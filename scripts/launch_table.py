"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one optimizer step."""
import csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
seq = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"]) / 1000) for r in rows]
# last complete step: from the last 'prep' kernel backwards
starts = [i for i, s in enumerate(seq) if "prep" in s[0]]
a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(seq))
step = seq[a:b]
tot = sum(s[2] for s in step)
for name, grid, us in step:
    short = name.replace("msf::<unnamed>::", "").replace("msf::", "").split("(")[0][:44]
    print("%-46s %-16s %8.1f us %5.1f%%" % (short, grid, us, 100 * us / tot))
print("TOTAL %.1f us over %d launches" % (tot, len(step)))
