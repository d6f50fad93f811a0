"""Region/line breakdown of an ncu source-page csv (cuda,sass view) for deb_core.cuh."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cur = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try: s = int(r[4]); i = int(r[7])
        except ValueError: continue
        a = agg.setdefault((cur, int(r[0])), [r[1], 0, 0]); a[1] += s; a[2] += i
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print(f"samples {ts} warp-inst {ti} inst/step {ti/nsteps:.0f}")
import os
CS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'disco-eb_b200', 'csrc')
MARKS = {}
for fn, first in (('deb_core.cuh', 700), ('deb_team.cuh', 60)):
    marks = []
    for i, l in enumerate(open(os.path.join(CS, fn)).read().split('\n'), 1):
        if re.match(r'\s*// (----|====)', l) and i > first: marks.append((i, l.strip()[:70]))
        if re.match(r'DEB_DEV .*\b(compute_bg|chain_coeffs_lane|compute_metric|head_row|tail_row|spl_pos|spl_locate|fill_slots|spl_at|team_helper_compute)\(', l): marks.append((i, l.strip()[:70]))
    MARKS[fn] = sorted(marks)
R = {}
for (f, ln), (s_, sm, ins) in agg.items():
    if f not in MARKS:
        a = R.setdefault(f, [0, 0]); a[0] += sm; a[1] += ins; continue
    name = f'{f} (helpers: dual ops, exp/log)'
    for b, nm in MARKS[f]:
        if ln >= b: name = f"{f}:{b}:{nm}"
    a = R.setdefault(name, [0, 0]); a[0] += sm; a[1] += ins
for k, (sm, ins) in sorted(R.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{100*sm/ts:5.1f}% samp {100*ins/ti:5.1f}% inst  {k}")
print("---- hottest lines by samples")
for (f, ln), (s_, sm, ins) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{100*sm/ts:5.1f}% samp {100*ins/ti:5.1f}% inst {f}:{ln}  {s_.strip()[:100]}")
