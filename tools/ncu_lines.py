"""Per-source-line instruction / stall-sample totals of one kernel launch of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep launch_index [top]"""
import csv
import io
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, acc = None, None, {}
kernel = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":  # keep the per-line summary rows (Address == "-")
        continue
    try:
        inst = int(r[hdr["Instructions Executed"]])
        samp = int(r[hdr["# Samples"]])
    except ValueError:
        continue
    key = (cur, int(r[0]), r[1].strip()[:100])
    a = acc.setdefault(key, [0, 0])
    a[0] += inst
    a[1] += samp
tot_i = sum(a[0] for a in acc.values()) or 1
tot_s = sum(a[1] for a in acc.values()) or 1
print(kernel[:160])
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("--- by instructions")
for (f, ln, src), (i, s) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% samp  {f}:{ln}  {src}")
print("--- by stall samples")
for (f, ln, src), (i, s) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * s / tot_s:5.1f}% samp {100 * i / tot_i:5.1f}% inst  {f}:{ln}  {src}")
