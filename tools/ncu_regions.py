#!/usr/bin/env python
"""Attribute the warp-instruction counts / stall samples of one profiled kernel to source REGIONS (device functions),
looking through inlined helpers.

    ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > sass.csv
    cuobjdump -xelf all gdss_fused.o ; nvdisasm -gi gdss_fused.sm_100a.cubin > all_gi.txt
    python tools/ncu_regions.py sass.csv all_gi.txt <mangled kernel name> [lines]

Both listings are in address order, so SASS row k of the ncu page is instruction k of the nvdisasm function.
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, gi_txt, kname = sys.argv[1:4]
by_line = len(sys.argv) > 4

HELPER_NAMES = {"tanh_fast", "fma2", "dup2", "lo2", "hi2", "FastDiv", "div", "ldw4", "split_tf32", "mma_tf32", "mma3", "sm_addr",
                "tma_fetch", "mbar_wait_parity", "mm_make"}
SRC = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "gdss_b200", "csrc",
                                 "gdss_fused.cu")


def scan_functions(path):
    """(first line, last line, name) of every top-level function body of the source file."""
    out, name, start, depth = [], None, 0, 0
    sig = re.compile(r"^(?:DEVINL|__device__|__global__|static|template)\b.*?([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    lines = open(path).read().split("\n")
    for k, line in enumerate(lines, 1):
        if name is None:
            m = sig.match(line)
            if m and not line.startswith("template"):
                name, start = m.group(1), k
            elif line.startswith("  DEVINL") or line.startswith("  DEVINL explicit"):
                pass
            if name and "{" in line and line.rstrip().endswith("}") and line.count("{") == line.count("}"):
                out.append((start, k, name)); name = None          # one-line body
            continue
        if line.startswith("}"):
            out.append((start, k, name))
            name = None
    return out


FUNCS = scan_functions(SRC)
KSTART = next((a for a, b, nm in FUNCS if nm == 'fused_layers_kernel'), 10 ** 9)
HELPERS = [(a, b) for a, b, nm in FUNCS if nm in HELPER_NAMES] + [(146, 156)]
REGIONS = [(a, b, nm) for a, b, nm in FUNCS if nm not in HELPER_NAMES]


def region_of(chain):
    for f, ln in chain:                                 # innermost first
        if not f.endswith("gdss_fused.cu"):
            continue
        if FUNCS and ln >= KSTART:                      # kernel body: split by the call-site comment blocks
            return ("kernel:glue", ln)
        if any(a <= ln <= b for a, b in HELPERS):
            continue
        for a, b, name in REGIONS:
            if a <= ln <= b:
                return name, ln
        return "other", ln
    return "other", 0


# ---- nvdisasm: instruction k -> inline chain
chains, cur, last, infn = [], [], [], False
pat = re.compile(r'//## File "([^"]+)", line (\d+)')
for line in open(gi_txt, errors="replace"):
    if line.startswith("//---------------------"):
        infn = (".text." + kname + " ") in line
        continue
    if not infn:
        continue
    m = pat.search(line)
    if m:
        cur.append((m.group(1), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        if cur:                                          # an instruction without its own note inherits the previous chain
            last, cur = cur, []
        chains.append(last)

# ---- ncu sass page
rows = list(csv.reader(open(sass_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
data = [r for r in rows if r and r[0].startswith("0x")]
if len(data) % max(len(chains), 1) == 0 and len(data) > len(chains):
    data = data[:len(chains)]                             # the page lists every selected launch: keep the first
if len(data) != len(chains):
    print(f"warning: {len(data)} profiled SASS rows vs {len(chains)} disassembled instructions", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int)])
tot_i = tot_s = 0
for k, r in enumerate(data):
    name, ln = region_of(chains[k]) if k < len(chains) else ("other", 0)
    key = (name, ln) if by_line else name
    ins, smp = int(r[ii] or 0), int(r[si] or 0)
    a = agg[key]
    a[0] += ins
    a[1] += smp
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    a[2][op.split(".")[0]] += ins
    for h, i in stall_cols.items():
        if r[i].isdigit():
            a[3][h[6:]] += int(r[i])
    tot_i += ins
    tot_s += smp
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    ops = ", ".join(f"{o} {v / a[0] * 100:.0f}%" for o, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:6]) if a[0] else ""
    st = ", ".join(f"{o} {v / max(a[1], 1) * 100:.0f}%" for o, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{str(key):28s} inst {a[0] / tot_i * 100:5.1f}%  samples {a[1] / tot_s * 100:5.1f}%   [{ops}]   <{st}>")
