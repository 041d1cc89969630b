#!/usr/bin/env python3
"""Hot spots of a kernel from an ncu report's source page: tools/ncu_hot.py <report> [kernel-substring] [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]: continue
    h = b["hdr"]; si = h.index("Source"); ie = h.index("Instructions Executed"); ss = h.index("# Samples")
    tot_i = sum(float(r[ie]) for r in b["rows"]); tot_s = sum(float(r[ss]) for r in b["rows"])
    print("==", b["name"][:100], "instructions %.3e samples %d" % (tot_i, tot_s))
    # aggregate by opcode
    agg = {}
    for r in b["rows"]:
        op = r[si].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        op = op.split(".")[0]
        a = agg.setdefault(op, [0.0, 0.0]); a[0] += float(r[ie]); a[1] += float(r[ss])
    print("  by opcode (inst%, stall-sample%):", ", ".join("%s %.1f/%.1f" % (k, 100 * v[0] / tot_i, 100 * v[1] / tot_s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]))
    top = sorted(b["rows"], key=lambda r: -float(r[ss]))[:N]
    for r in top:
        print("  %5.2f%% samples  %5.2f%% inst  %s" % (100 * float(r[ss]) / tot_s, 100 * float(r[ie]) / tot_i, r[si].strip()[:100]))
