"""Warp-stall samples of one kernel aggregated by SASS opcode (ncu -i rep --page source --csv --print-source sass):
    python tools/ncu_stalls_by_opcode.py file.ncu-rep [title]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
byop, execs, bystall = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    if not src:
        continue
    n, ex = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    byop[op] += n; execs[op] += ex; tot += n
    for c in stallcols:
        bystall[c] += int(r[ix[c]] or 0)
if len(sys.argv) > 2:
    print(sys.argv[2])
print(rows[0][1] if len(rows[0]) > 1 else "")
print(f"warp-stall samples: {tot}")
for op, n in byop.most_common(20):
    print(f"{op:12s} samples {n:7d} {100 * n / max(tot, 1):5.1f} %   warp-instructions executed {execs[op]:11d}")
print("by stall reason:", ", ".join(f"{k[6:]} {v}" for k, v in bystall.most_common(10)))
