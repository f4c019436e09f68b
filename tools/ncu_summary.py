"""Summarise an .ncu-rep (ncu --set full) into a small CSV of the metrics DESIGN.md / bench.py cite:
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_ncu_x.csv"""
import csv, io, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size" , "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = csv.writer(sys.stdout)
    out.writerow(["launch", "metric", "value", "unit"])
    for n, r in enumerate(rows[2:]):
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.writerow([n, k, r[i], units[i]])
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(r[i]) >= 0.05:
                        out.writerow([n, h[len(STALL):-len("_per_issue_active.ratio")] + " (stalled warps per issue)", r[i], ""])
                except ValueError:
                    pass


if __name__ == "__main__":
    main(sys.argv[1])
