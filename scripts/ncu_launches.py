#!/usr/bin/env python3
"""Turn an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>` launch list into the
per-kernel table kept under profiles/ (launch count, total time, share of the step).

    python scripts/ncu_launches.py gpurun_out/launches.csv "command line that was profiled" > profiles/rNN_launches_bench.md
"""
import collections, csv, sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr):
        continue
    if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[hdr["Metric Unit"]]
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    rows.append((r[hdr["Kernel Name"]], ms))
tot = collections.OrderedDict()
for k, ms in rows:
    name = k.split("(")[0][:96]
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1; c[1] += ms
total = sum(v[1] for v in tot.values()) or 1.0
print("# ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none)" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("# per-launch times are cold-cache and serialised: compare SHARES")
print("kernel | launches | total ms | share | ms per launch")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s | %d | %.3f | %.1f%% | %.3f" % (name, n, ms, 100 * ms / total, ms / n))
