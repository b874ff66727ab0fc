"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: time and share per kernel name.
usage: python tools/launch_shares.py launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    name = r[ik].replace("void ", "").split("(")[0][:50]
    tot[name] += float(r[iv].replace(",", "")) / 1e3
    cnt[name] += 1
total = sum(tot.values())
for name, us in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print(f"{us:9.1f} us {100*us/total:5.1f}%  {cnt[name]:3d}x {name}")
print(f"total {total:.1f} us over {sum(cnt.values())} launches")
