"""Per-source-line executed-instruction and stall-sample shares from an .ncu-rep captured with --import-source on (first kernel in the report)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg, nfunc, fname = None, None, {}, 0, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        if nfunc and r[1] != fname:
            break
        fname = r[1]
        nfunc += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and "Instructions Executed" in hdr:
        d = dict(zip(hdr, r))
        try:
            e, s = int(d["Instructions Executed"] or 0), int(d["Warp Stall Sampling (All Samples)"] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur_file.split("/")[-1], int(r[0]), r[1].strip()[:100]), [0, 0])
        a[0] += e
        a[1] += s
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print(f"{tot} warp instructions, {tots} stall samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% exec {v[1] / max(1, tots) * 100:5.1f}% stall  {k[0]}:{k[1]}  {k[2]}")
