"""Condense an .ncu-rep into a small committed summary: python tools/ncu_summary.py rep out.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__inst_executed.avg",
        "smsp__inst_executed.max", "smsp__cycles_active.avg", "smsp__cycles_active.max", "sm__cycles_elapsed.max",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (per launch).", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines += ["## %s" % d.get("Kernel Name", "?"), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        st = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and "per_issue_active" in k and d[k] not in ("", "n/a")), reverse=True)
        lines += ["", "Warp stall reasons (cycles per issued instruction): " +
                  ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in st[:9]), ""]
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
