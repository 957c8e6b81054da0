#!/usr/bin/env python3
"""Split an .ncu-rep's per-instruction stall samples into kernel phases (by SASS landmarks)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
ins = []
for r in data:
    s = r[ix["Source"]].split()
    if not s: continue
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    ins.append((op, int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), r))
ops = [o for o, _, _, _ in ins]
def first(op): return next((i for i, o in enumerate(ops) if o == op), None)
def last(op): return max((i for i, o in enumerate(ops) if o == op), default=None)
# only the executed copy (the compiler also emits a never-taken WARPSYNC.COLLECTIVE clone)
exe = [i for i, (_, _, e, _) in enumerate(ins) if e > 0]
lo, hi = min(exe), max(exe)
marks = {}
r0, r1 = first("CREDUX"), None
red = [i for i in exe if ops[i] == "CREDUX"]
mu = [i for i in exe if ops[i] == "MUFU"]
ff = [i for i in exe if ops[i] == "FFMA2"]
phases = []
if red:
    phases += [("stage-in", lo, red[0] - 8), ("pre-pass", red[0] - 8, red[-1] + 30), ("reg-load", red[-1] + 30, mu[0] - 5)]
else:
    phases += [("stage-in+reg-load", lo, mu[0] - 5)]
phases += [("eliminate", mu[0] - 5, ff[-1] + 1), ("store+stage-out", ff[-1] + 1, hi + 1)]
tot = sum(s for _, s, _, _ in ins)
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot)
for name, a, b in phases:
    seg = ins[a:b]
    s = sum(x[1] for x in seg); e = sum(x[2] for x in seg)
    st = collections.Counter()
    for x in seg:
        for c in stall_cols:
            v = x[3][ix[c]]
            if v: st[c.replace("stall_", "")] += int(v)
    top = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in st.most_common(5))
    print("%-18s samples %6d (%4.1f%%)  warp-inst executed %11d  | %s" % (name, s, 100.0 * s / tot, e, top))
