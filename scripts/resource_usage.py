"""Registers / stack (spill) / shared memory per kernel of a built library, names demangled (development aid).
usage: python scripts/resource_usage.py <lib.so> [substring]"""
import re, subprocess, sys
lib = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
lines = out.splitlines()
names, rows = [], []
for i, l in enumerate(lines):
    m = re.match(r"\s*Function (\S+):", l)
    if m and i + 1 < len(lines):
        names.append(m.group(1))
        rows.append(lines[i + 1])
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
for n, d, r in zip(names, dem, rows):
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = re.sub(r">\(.*", ">", d)
    if pat in d:
        g = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", r))
        print(f"{d:60s} REG {g.get('REG'):>4s} STACK {g.get('STACK'):>4s} SHARED {g.get('SHARED'):>6s}")
