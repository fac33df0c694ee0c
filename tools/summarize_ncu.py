"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md  [steps]
  python tools/summarize_ncu.py kernels  gpurun_out/prof_r1b.ncu-rep profiles/r1_kernels.md
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("phb::", "")[:90]


def launches(src, dst, steps_hint=None):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, order, total = {}, [], 0.0
    for r in rows[1:]:
        k, v = short(r[ik]), float(r[iv].replace(",", ""))
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += v
        total += v
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(src)})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 1 "
                "--no-e2e --no-cpu` (3 warm-up + 2 timed steps + initialisation, config 5, one B200). Per-launch times "
                "are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"total launches: {len(rows) - 1}, total kernel time: {total / 1e6:.2f} ms\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k in sorted(order, key=lambda k: -agg[k][1]):
            n, t = agg[k]
            f.write(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / n / 1e6:.4f} | {100 * t / total:.1f} % |\n")
    print(open(dst).read())


def kernels(src, dst, particles=64 ** 3 * 64, what="python tools/microbench.py c5s` (3-D, order 1, 64^3 cells x 64 ppc = 16.8 M particles, > L2)"):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, traffic = set(), {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(src)})\n\n"
                f"`ncu --set full --clock-control none --import-source on` over `{what}, one launch per kernel shown.\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            if name in seen:
                continue
            seen.add(name)
            f.write(f"\n## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in KEEP:
                if m in idx:
                    f.write(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |\n")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                rd *= mult.get(units[idx["dram__bytes_read.sum"]], 1)
                wr *= mult.get(units[idx["dram__bytes_write.sum"]], 1)
                traffic[name] = dict(dram_bytes_per_launch=rd + wr, particles=particles,
                                     dram_bytes_per_particle=(rd + wr) / particles)
            except Exception:
                pass
    tfile = os.path.join(os.path.dirname(dst), "traffic_c5s.json")
    merged = json.load(open(tfile)) if os.path.exists(tfile) else {}
    merged.update(traffic)  # keep the kernels captured in earlier reports
    json.dump(merged, open(tfile, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif len(sys.argv) > 4:  # kernels <rep> <md> <particles> <what>
        kernels(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
    else:
        kernels(sys.argv[2], sys.argv[3])
