"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python scripts/launch_list_summary.py <launches.csv> <out.txt> "<command that was profiled>" """
import collections, csv, re, sys
src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"^(at::native::|smcdet::|\(anonymous namespace\)::)+", "", name)
    name = re.sub(r"\(.*$", "", name)[:80]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
with open(out, "w") as f:
    f.write(f"# every launch of `{cmd}`\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# full list: {src.replace('gpurun_out', 'profiles')}\n# {sum(cnt.values())} launches, {total:.2f} ms of kernel time\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"{v:10.3f} ms {100 * v / total:6.2f}%  n={cnt[k]:4d}  {k}\n")
