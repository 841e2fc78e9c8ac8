"""Group the SASS instructions of a kernel in an .ncu-rep by how often a warp executes them (the per-warp
multiplicity separates prologue, per-sweep code and per-star code of mh_kernel) and print, per group, the warp
instructions executed, the opcode mix and the stall samples (development aid).
usage: python scripts/ncu_regions.py <report.ncu-rep> [kernel-substring]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be concatenated: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["rows"][0]
    iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = [(d[iS].strip(), int(d[iE]), int(d[iSm])) for d in b["rows"][1:] if len(d) > iE and d[iE].isdigit()]
    warps = data[0][1]
    print(b["name"][:100], "warps", warps, "instructions", len(data))
    groups = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
    for s, e, sm in data:
        mult = e / warps
        key = "~0" if mult < 0.5 else ("1-3" if mult < 4 else ("4-40" if mult < 40 else ("40-150" if mult < 150 else ("150-400" if mult < 400 else ">400"))))
        op = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9]+)?)", s)
        op = op.group(1) if op else "?"
        if not op.startswith("MUFU"):
            op = op.split(".")[0]
        g = groups[key]
        g[0] += e; g[1] += sm; g[2][op] += e; g[3] += 1
    tot_e = sum(g[0] for g in groups.values()); tot_s = sum(g[1] for g in groups.values())
    for key in ["~0", "1-3", "4-40", "40-150", "150-400", ">400"]:
        if key not in groups:
            continue
        e, sm, ops, n = groups[key]
        print(f"  per-warp multiplicity {key:8s}: {n:5d} static instr, {e / warps:9.1f} executed per warp ({100 * e / tot_e:5.1f}%), samples {100 * sm / max(1, tot_s):5.1f}%")
        print("      " + ", ".join(f"{o} {v / warps:.0f}" for o, v in ops.most_common(14)))
