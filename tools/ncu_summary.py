#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under
profiles/ (the .ncu-rep files themselves are scratch).

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep profiles/ncu_x.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def to_ms(value, unit):
    v = float(value.replace(",", ""))
    return {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v, "s": v * 1e3, "second": v * 1e3}.get(unit, v / 1e6)


def launches(src, dst, note=""):
    with open(src) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += to_ms(r["Metric Value"], r["Metric Unit"])
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list ({src})\n\n{note}\n\n")
        out.write("`gpu__time_duration.sum` per launch, `--clock-control none`; times are cold-cache and serialised, "
                  "so compare SHARES with the live CUDA-event numbers of bench.py, not absolutes.\n\n")
        out.write("| kernel | launches | total ms | share | avg ms | grid (last) | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for name, (n, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{name}` | {n} | {ms:.3f} | {ms / total:.3f} | {ms / n:.4f} | {grid} | {block} |\n")
        out.write(f"\nTotal: {len(rows)} launches, {total:.3f} ms\n")


def full(src, dst, note=""):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary ({src})\n\n{note}\n\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            out.write(f"## `{d.get('Kernel Name', '?')}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    out.write(f"| {k} | {d[k]} | {u[k]} |\n")
            out.write("\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "full": full}[mode](src, dst, note)
