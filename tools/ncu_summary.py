"""Prints the handful of ncu metrics this project reads from a .ncu-rep (run where ncu is installed, no GPU needed):
python tools/ncu_summary.py <file.ncu-rep> [--csv out.csv]"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.per_cycle_active', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum']


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    table = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = []
        for k in KEYS:
            if k in d:
                rec.append((k, d[k], units[hdr.index(k)]))
        st = {k: v for k, v in d.items() if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio')}
        for k, v in sorted(st.items(), key=lambda kv: -float(kv[1] or 0))[:9]:
            rec.append(("stall_" + k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v, "warps/issue"))
        table.append(rec)
        for k, v, u in rec:
            print("  %-70s %s %s" % (k, v, u))
        print()
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w") as f:
            w = csv.writer(f)
            w.writerow(["launch", "metric", "value", "unit"])
            for i, rec in enumerate(table):
                for k, v, u in rec:
                    w.writerow([i, k, v, u])


if __name__ == "__main__":
    main()
