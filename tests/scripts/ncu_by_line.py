"""Join an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) with nvdisasm -gi line info of the profiled cubin:
executed instructions, local loads / stores and stall samples per source line of jb_quad.cuh (the line of the first
frame of jb_quad.cuh in the inline chain).  python tests/scripts/ncu_by_line.py source.csv sass_lines.txt [top]"""
import collections, csv, re, sys

src_csv, sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
off2 = {}
chain, in_text = [], False
for ln in open(sass).read().splitlines():
    if ln.startswith('//--------------------- .text.'):
        in_text = 'jb_sim_kernel' in ln
        continue
    if not in_text:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        if chain:
            # innermost frame first; prefer the first frame that is in jb_quad.cuh (helpers in jb_math.cuh are everywhere)
            loc = next((c for c in chain if c[0] == 'jb_quad.cuh'), chain[0])
            last = loc
        off2[int(m.group(1), 16)] = (last, m.group(2))
        chain = []
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
base = min(int(r[ix['Address']], 16) for r in data)


def f(r, k):
    try:
        return float(r[ix[k]])
    except ValueError:
        return 0.0


ex, loc, smp, lsb, bar = (collections.Counter() for _ in range(5))
for r in data:
    (line, op) = off2.get(int(r[ix['Address']], 16) - base, (("?", 0), ""))
    n = f(r, 'Instructions Executed')
    ex[line] += n; smp[line] += f(r, '# Samples'); lsb[line] += f(r, 'stall_long_sb'); bar[line] += f(r, 'stall_barrier')
    if re.search(r'\b(LDL|STL)', op):
        loc[line] += n
T, TL, TS, TB = sum(ex.values()), sum(loc.values()), sum(smp.values()), sum(lsb.values())
print(f"executed {T/1e9:.2f} G warp instructions, local ld/st {100*TL/T:.1f} %, samples {TS:.0f}")
print("line                     exec%   local%  samples%  long_sb%")
for k, v in ex.most_common(top):
    print(f"{k[0]:14s}{k[1]:6d} {100*v/T:7.2f} {100*loc[k]/max(TL,1):7.2f} {100*smp[k]/TS:8.2f} {100*lsb[k]/max(TB,1):8.2f}")
print("top local-memory lines")
for k, v in loc.most_common(12):
    print(f"{k[0]:14s}{k[1]:6d} {100*v/TL:7.2f} %")
