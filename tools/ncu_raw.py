"""Key raw metrics of every kernel instance in an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(r[h.index("Kernel Name")][:70])
        for w in WANT:
            if w in h:
                print("   %-66s %s %s" % (w, r[h.index(w)], units[h.index(w)]))


if __name__ == "__main__":
    main()
