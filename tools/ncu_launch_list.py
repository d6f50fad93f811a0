"""Launch list of an `ncu --metrics gpu__time_duration.sum --csv` log: kernel, launches, total ms, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1.0)
    a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("kernel, launches, total ms, share of GPU time (cold-cache, serialised: shares only)")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {n:4d} {ms:10.3f} {100*ms/tot:6.2f}%")
