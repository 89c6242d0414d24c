"""Per-kernel share of one steady step from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(scripts/profile_step.py runs warm-up + 1 step; the second half of the launches is the steady step).
usage: python scripts/ncu_launch_shares.py launches.csv [command line shown in the header]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]])
    u = r[ix["Metric Unit"]]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    seq.append((r[ix["Kernel Name"]], v))
# the steady step starts at the last launch of the first kernel of a step (k_edge_lengths opens every constraint set)
firsts = [i for i, (n, _) in enumerate(seq) if "k_edge_lengths" in n]
begin = firsts[-2] if len(firsts) >= 2 else 0  # constraint set + CCD both call it: take the last constraint-set one
half = seq[begin:]
agg = collections.OrderedDict()
for n, v in half:
    key = n.split("(")[0][:84]
    agg[key] = agg.get(key, 0.0) + v
tot = sum(agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none, %s" % (sys.argv[2] if len(sys.argv) > 2 else "scripts/profile_step.py"))
print("# last (steady) step only; per-launch times are cold-cache and serialised: compare SHARES")
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:32]:
    print("%9.3f ms %5.1f%%  %s" % (v, 100 * v / tot, k))
print("total %.3f ms over %d launches" % (tot, len(half)))
