#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name.txt [--top 12]

Per kernel launch in the report: duration, DRAM bytes read/written and GB/s, L2 bytes, tensor-pipe
and issue activity, registers / shared memory / grid, the warp-stall breakdown and the hottest
SASS lines by stall samples (source page; needs -lineinfo, which the build always passes).
"""
import csv
import io
import subprocess
import sys

RAW = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor hmma inst %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none; cold-cache, serialised replays)", ""]
    for r in data:
        lines.append(f"## {r[ix['Kernel Name']]}  (launch id {r[ix['ID']]})")
        for key, label in RAW:
            if key in ix:
                lines.append(f"  {label:24s} {r[ix[key]]} {units[ix[key]]}")
        stalls = []
        for h, i in ix.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("  stalls (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
        lines.append("")
    # source page: hottest SASS per kernel
    src = run(["-i", rep, "--page", "source", "--csv"])
    block, name = [], None
    def flush():
        if not block:
            return
        rr = list(csv.reader(io.StringIO("\n".join(block))))
        h = {c: i for i, c in enumerate(rr[0])}
        if "# Samples" not in h:
            return
        body = rr[1:]
        tot = sum(int(x[h["# Samples"]] or 0) for x in body) or 1
        lines.append(f"## hottest SASS of {name} ({tot} stall samples)")
        order = sorted(range(len(body)), key=lambda i: -int(body[i][h["# Samples"]] or 0))[:top]
        for i in sorted(order):
            x = body[i]
            lines.append(f"  [{i:5d}] {x[h['Source']][:70]:70s} samples={x[h['# Samples']]:>7s} ({100*int(x[h['# Samples']])/tot:4.1f}%) exec={x[h['Instructions Executed']]}")
        lines.append("")
    for ln in src.splitlines():
        if ln.startswith('"Kernel Name"'):
            flush()
            block, name = [], ln.split(",", 1)[1].strip('",')
        elif name is not None:
            block.append(ln)
    flush()
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"wrote {out} ({len(lines)} lines)")


if __name__ == "__main__":
    main()
