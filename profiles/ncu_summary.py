"""Per-launch summary of an `ncu --set full` report (read here on the CPU box: `ncu -i rep --page raw --csv`).

    python profiles/ncu_summary.py gpurun_out/tc_conv.ncu-rep > profiles/rN/ncu_tc_conv_summary.txt
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(n) if n in hdr else None, short) for n, short in COLS]
    print("# " + path)
    print("# " + " | ".join(f"{s}[{units[i]}]" if i is not None and units[i] else s for i, s in idx))
    for r in rows[2:]:
        out = []
        for i, s in idx:
            v = r[i] if i is not None else "-"
            if s == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            else:
                try:
                    v = f"{float(v):.2f}"
                except ValueError:
                    pass
            out.append(v)
        print(" | ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
