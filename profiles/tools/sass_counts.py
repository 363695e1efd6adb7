"""Static SASS instruction counts per kernel of libgple_b200.so (markdown table on stdout).
usage: python profiles/tools/sass_counts.py [path to the .so]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gaussian_process_liouville_equation_b200", "libgple_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = {}
try:
    filt = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True, check=True).stdout.splitlines()
    names = dict(zip(re.findall(r"Function : (\S+)", sass), filt))
except Exception:
    pass
COLS = [("DMMA", "DMMA"), ("LDGSTS", "LDGSTS"), ("TMA", ("UTMALDG", "UBLKCP", "UTMASTG")), ("DFMA", "DFMA"), ("DMUL", "DMUL"), ("DADD", "DADD"), ("MUFU", "MUFU"), ("LDS", "LDS"), ("STS", "STS"), ("LDG", "LDG"), ("STG", "STG"), ("BAR", "BAR")]
rows, cur = [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = {"name": names.get(m.group(1), m.group(1)), "n": 0, **{c: 0 for c, _ in COLS}}
        rows.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["n"] += 1
        for c, pre in COLS:
            pre = pre if isinstance(pre, tuple) else (pre,)
            if any(op == p or op.startswith(p + ".") for p in pre):
                cur[c] += 1
other = sum(len(re.findall(r"\b" + p, sass)) for p in ("HMMA", "IMMA", "UTCHMMA", "UTCIMMA"))
print("| kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |\n|---|---:|" + "---:|" * len(COLS))
for r in sorted(rows, key=lambda r: -r["n"]):
    name = re.sub(r"\(.*$", "", r["name"])
    name = re.sub(r"^void ", "", name)
    print(f"| `{name}` | {r['n']} | " + " | ".join(str(r[c]) for c, _ in COLS) + " |")
print(f"\nWhole library: {sum(r['n'] for r in rows)} instructions, DMMA {sum(r['DMMA'] for r in rows)}, LDGSTS {sum(r['LDGSTS'] for r in rows)}, TMA {sum(r['TMA'] for r in rows)}, HMMA/IMMA/UTCHMMA {other} (no reduced-precision tensor instruction anywhere).")
