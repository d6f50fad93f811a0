"""Per (file, source line) stall samples by reason from `ncu --page source --print-source cuda,sass --csv`, without the
barrier samples (warps parked at a barrier say where the OTHERS are slow): python tools/ncu_stalls_by_line.py X.csv [top] [file-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ""
fn, hdr, agg = "", None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fn = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 49 or not r[0].isdigit() or r[2] != "-":
        continue
    reasons = {hdr[i][6:]: int(r[i] or 0) for i in range(32, 49)}
    a = agg.setdefault((fn, int(r[0])), [r[1].strip(), 0, {}])
    a[1] += int(r[7] or 0)
    for k, v in reasons.items():
        a[2][k] = a[2].get(k, 0) + v
tot = sum(sum(v for k, v in a[2].items() if k != "barrier") for a in agg.values())
print("non-barrier samples", tot, " barrier samples", sum(a[2].get("barrier", 0) for a in agg.values()))
allr = {}
for a in agg.values():
    for k, v in a[2].items():
        allr[k] = allr.get(k, 0) + v
print("by reason:", {k: v for k, v in sorted(allr.items(), key=lambda kv: -kv[1]) if v})
items = [(k, a) for k, a in agg.items() if only in k[0]]
for (f, ln), (src, inst, rs) in sorted(items, key=lambda kv: -sum(v for k, v in kv[1][2].items() if k != "barrier"))[:top]:
    nb = sum(v for k, v in rs.items() if k != "barrier")
    tops = ", ".join(f"{k} {v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:4] if v and k != "barrier")
    print(f"{f}:{ln:4d} {100*nb/tot:5.1f}%  inst {inst:9d}  [{tops}]  {src[:80]}")
