"""Summarise an .ncu-rep (from `ncu --set full`) into a small CSV for profiles/: one row per captured launch."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["%s [%s]" % (n, units[i]) if units[i] else n for n, i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split("(")[0].replace("<unnamed>::", "") if n == "Kernel Name" else r[i] for n, i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
