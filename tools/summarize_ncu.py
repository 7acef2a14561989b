"""Compact per-launch table from an `ncu --page raw --csv` dump (run in the dev container, no GPU needed).
usage: python tools/summarize_ncu.py gpurun_out/prof_block_TAG_raw.csv > profiles/TAG_block_ncu_summary.csv"""
import csv
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_KB"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in WANT if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow(["launch"] + [f"{n}[{u}]" if u else n for _, n, u in cols])
    for i, r in enumerate(rows[2:]):
        out = [i]
        for c, n, _ in cols:
            v = r[c]
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "")
            else:
                try:
                    v = f"{float(v):.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1])
