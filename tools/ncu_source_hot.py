"""Aggregates `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line:
prints the lines with the most executed warp-instructions and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 9:
        continue
    if r[0].isdigit() and r[2] == "-":          # a source line summary row
        line, src = int(r[0]), r[1]
        try:
            samples, inst = int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(line, [src, 0, 0])
        a[1] += samples; a[2] += inst
tot_s = sum(a[1] for a in agg.values()); tot_i = sum(a[2] for a in agg.values())
print("total samples", tot_s, "total warp-inst", tot_i)
print("---- by instructions executed")
for line, (src, s, i) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{line:5d} inst {100*i/tot_i:5.1f}% samp {100*s/tot_s:5.1f}%  {src.strip()[:110]}")
print("---- by stall samples")
for line, (src, s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{line:5d} inst {100*i/tot_i:5.1f}% samp {100*s/tot_s:5.1f}%  {src.strip()[:110]}")
