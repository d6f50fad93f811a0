"""Static SASS instruction count per CUDA source region from `nvdisasm -g` output of one kernel (stdin or file).
Usage: nvdisasm -g X.cubin | sed -n '/.text.<kernel>:/,/^\/\/----/p' | python tools/sass_by_line.py deb_lane.cuh"""
import re, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = open(sys.argv[2]) if len(sys.argv) > 2 else sys.stdin
focus = sys.argv[1] if len(sys.argv) > 1 else "deb_lane.cuh"
cnt = collections.Counter(); cur = None
for line in txt:
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        cnt[cur] += 1
tot = sum(cnt.values()); print("total SASS", tot, "=", tot * 16 // 1024, "KB")
byf = collections.Counter()
for (f, l), c in cnt.items(): byf[f] += c
print(dict(byf))
src = open(os.path.join(ROOT, "disco-eb_b200", "csrc", focus)).read().split("\n")
marks = [(i + 1, l.strip()[:80]) for i, l in enumerate(src) if re.match(r"\s*// (----|====)", l)]
R = collections.OrderedDict()
for (f, l), c in sorted(cnt.items()):
    if f != focus: continue
    name = "(before)"
    for b, nm in marks:
        if l >= b: name = f"{b}:{nm}"
    R[name] = R.get(name, 0) + c
for k, v in R.items(): print(f"{v:6d} {k}")
print("top lines")
for (f, l), c in cnt.most_common(30): print(c, f, l)
