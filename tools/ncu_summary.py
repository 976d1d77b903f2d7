"""Summarise an ncu report (`ncu --set full ... -o x`) into the per-kernel CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full_....csv
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__waves_per_multiprocessor", "waves"),
]


def main():
    out = csv.writer(sys.stdout)
    header_done = False
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = [(hdr.index(m), label) for m, label in COLS if m in hdr]
        if not header_done:
            out.writerow(["report", "kernel"] + [f"{label} [{units[i]}]" if units[i] else label for i, label in idx])
            header_done = True
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            out.writerow([rep.split("/")[-1], r[ki][:70]] + [r[i] for i, _ in idx])


if __name__ == "__main__":
    main()
