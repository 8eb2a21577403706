"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("dgp::<unnamed>::", "")
    tot[name] += v * scale; cnt[name] += 1
total = sum(tot.values())
print(f"# per-kernel device time over {sum(cnt.values())} launches, total {total:.1f} ms (ncu serialised, cold cache: compare shares)")
for name, t in tot.most_common():
    print(f"{name:70s} launches {cnt[name]:6d}  {t:10.2f} ms  {100*t/total:6.2f} %")
