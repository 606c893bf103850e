#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if len(r) > 5 and '# Samples' in r)
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
out = []
for r in rows:
    if len(r) > ii and r[0].isdigit():
        try:
            st = {k: int(r[i]) for k, i in stall.items() if r[i].isdigit()}
            out.append((int(r[ii]), int(r[si]), int(r[0]), r[1].strip()[:80], st))
        except ValueError:
            pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print('total warp-instructions', tot, 'samples', tots)
agg = {}
for o in out:
    for k, v in o[4].items(): agg[k] = agg.get(k, 0) + v
print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for o in sorted(out, key=lambda o: -o[1])[:top]:
    st = sorted(o[4].items(), key=lambda kv: -kv[1])[:2]
    print(f"{o[1]/tots*100:5.1f}% smp {o[0]/tot*100:5.1f}% inst  L{o[2]:4d} {o[3]:80s} {st}")
