"""Static SASS statistics per kernel: instruction count and opcode histogram (development aid).
usage: python scripts/sass_stats.py <lib.so> <substring of mangled kernel name> [top]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
stats = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        stats[cur][m.group(2).split(".")[0]] += 1
for name, c in stats.items():
    if pat in name:
        n = sum(c.values())
        print(name, "total", n, "bytes", n * 16)
        print("  ", ", ".join(f"{k}:{v}" for k, v in c.most_common(top)))
