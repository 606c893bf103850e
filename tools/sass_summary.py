"""Static evidence for the built library (no GPU needed): per kernel the registers / spills / shared memory ptxas reports
(gdss_b200/csrc/*.ptxas.log) and how often the SASS of libgdss_b200.so carries the Blackwell tensor-memory / TMA / MMA mnemonics.

    python tools/sass_summary.py > profiles/rNN_sass_static.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gdss_b200", "csrc")
WATCH = ["UTCHMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA2", "FFMA", "MUFU", "LDS", "STS",
         "LDG", "STG", "BAR", "REDUX", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    print("# ptxas -v per kernel (registers, spill bytes, static shared memory)")
    rows = []
    for log in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        if log.endswith("_tf32.ptxas.log"):
            continue
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = {"name": m.group(1), "spill": 0}
                rows.append(cur)
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                cur["spill"] = int(m.group(1)) + int(m.group(2))
            m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                cur["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(s.group(1)) if s else 0
    names = demangle([r["name"] for r in rows])
    for r in sorted(rows, key=lambda r: short(names[r["name"]])):
        if "regs" in r:
            print(f"{r['regs']:4d} regs {r['spill']:5d} B spill {r['smem']:6d} B smem  {short(names[r['name']])}")
    print("\n# SASS mnemonic counts per kernel of libgdss_b200.so (cuobjdump -sass), kernels with tensor / TMA instructions first")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, "libgdss_b200.so")], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur[w] += 1
                    break
    names = demangle(list(counts))
    total = collections.Counter()
    tensor = lambda c: c["UTCHMMA"] + c["UTCMMA"] + c["LDTM"] + c["STTM"] + c["UBLKCP"] + c["HMMA"]
    for k, c in sorted(counts.items(), key=lambda kv: -tensor(kv[1])):
        total.update(c)
        if tensor(c):
            print(short(names[k]))
            print("    " + "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]) + f"  (of {c['total']} instructions)")
    print("\n# whole library: " + "  ".join(f"{w} {total[w]}" for w in WATCH if total[w]) + f"  (of {total['total']} instructions, "
          f"{len(counts)} kernels)")


if __name__ == "__main__":
    main()
