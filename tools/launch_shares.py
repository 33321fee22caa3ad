"""Per-kernel share of a `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised: compare SHARES).
    python tools/launch_shares.py gpurun_out/r02b_bench_launches.csv > profiles/r02b_bench_launch_shares.txt"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[r[ki]] += v
    cnt[r[ki]] += 1
s = sum(tot.values())
print(f"# total {sum(cnt.values())} launches, {s} ns")
for k, v in tot.most_common():
    print(f"{cnt[k]:4d} launches {v:14.1f} ns {100 * v / s:5.1f}%  {k[:150]}")
