"""One line per kernel of an .ncu-rep: time, DRAM bytes, SM/DRAM %, occupancy, issue-active, registers."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, vals = rows[0], rows[2:]
def g(d, k, default=""):
    return d.get(k, default)
print("%-34s %-14s %5s %8s %8s %7s %6s %6s %6s %6s %5s" % ("kernel", "grid", "blk", "us", "dramMB", "dram%", "sm%", "occ%", "issue%", "waves", "regs"))
tot = 0
for v in vals:
    d = dict(zip(hdr, v))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:34]
    def f(k):
        try: return float(d[k].replace(",", ""))
        except Exception: return float("nan")
    us = f("gpu__time_duration.sum")
    u = dict(zip(hdr, rows[1]))
    if u.get("gpu__time_duration.sum") == "ns": us /= 1e3
    elif u.get("gpu__time_duration.sum") == "ms": us *= 1e3
    def mb(k):
        x = f(k); un = u.get(k, "")
        return x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(un, 1)
    tot += us
    print("%-34s %-14s %5s %8.1f %8.2f %7.1f %6.1f %6.1f %6.1f %6.2f %5s" % (
        name, d["Grid Size"].replace(" ", ""), d["Block Size"].split(",")[0].strip("("), us,
        mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
        f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        f("sm__warps_active.avg.pct_of_peak_sustained_active"), f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f("launch__waves_per_multiprocessor"), d.get("launch__registers_per_thread", "")))
print("total %.1f us" % tot)
