"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("uf::", "").replace("(int)", "").replace("(bool)", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v  # -> us
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':50s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k[:50]:50s} {cnt[k]:8d} {v:10.1f} {v / cnt[k]:8.1f} {100 * v / total:6.1f}%")
print(f"{'TOTAL':50s} {sum(cnt.values()):8d} {total:10.1f}")
