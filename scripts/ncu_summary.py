"""Summarise an .ncu-rep (read here on the CPU box) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_step_kernel.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    lines = [f"ncu summary of {rep}", ""]
    raw = ncu(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    kcol = hdr.index("Kernel Name")
    lines.append("kernels: " + "; ".join(sorted({r[kcol] for r in rows})))
    lines.append(f"{'metric':78s} {'unit':14s} per captured launch")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k:78s} {units[i]:14s} " + "  ".join(r[i] for r in rows))
    lines.append("")
    lines.append("warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio, first launch):")
    st = [(hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(rows[0][i]))
          for i in range(len(hdr))
          if hdr[i].startswith("smsp__average_warps_issue_stalled_") and hdr[i].endswith("_per_issue_active.ratio")
          and "not_issued" not in hdr[i]]
    for name, v in sorted(st, key=lambda kv: -kv[1]):
        if v > 0.01:
            lines.append(f"  {name:28s} {v:6.2f}")
    src = ncu(rep, "source")
    data = [r for r in src[2:] if len(r) >= 6 and r[0].startswith("0x")]
    # first launch only: addresses restart when the next launch's listing begins
    first = []
    seen = set()
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    tot = sum(int(r[2]) for r in first) or 1
    execs = Counter(int(r[5]) for r in first)
    lines.append("")
    lines.append(f"SASS (first launch): {len(first)} instructions, {sum(int(r[5]) for r in first)} warp-instructions executed, "
                 f"{tot} stall samples")
    lines.append("executed-count histogram (count: #instructions): " +
                 ", ".join(f"{k}: {v}" for k, v in sorted(execs.items(), key=lambda kv: -kv[1])[:8]))
    lines.append("top stall sites:")
    for r in sorted(first, key=lambda r: -int(r[2]))[:25]:
        lines.append(f"  {100 * int(r[2]) / tot:5.1f}%  exec {int(r[5]):8d}  {r[1].strip()[:100]}")
    mn = Counter()
    for r in first:
        toks = r[1].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        mn[op.split(".")[0]] += int(r[5])
    lines.append("executed warp-instructions by opcode: " + ", ".join(f"{k} {v}" for k, v in mn.most_common(18)))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
