"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per captured launch with the metrics the roofline
discussion needs (duration, DRAM bytes, DRAM / tensor / SM utilisation, occupancy, registers, launch shape).
Usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "inst")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    print(" | ".join("%s[%s]" % (n, units[i]) if units[i] and n not in ("dram_rd", "dram_wr") else n
                     for i, n in cols))
    for r in body:
        out = []
        for i, n in cols:
            v = r[i]
            if n == "kernel":
                v = v.replace("void ", "").replace("mc::<unnamed>::", "").replace("unnamed>::", "").split("(")[0][:34]
            elif n in ("dram_rd", "dram_wr"):
                v = "%.0f B" % (float(v.replace(",", "")) * SCALE.get(units[i], 1.0))
            else:
                try:
                    v = "%.2f" % float(v.replace(",", "")) if "." in v else v
                except ValueError:
                    pass
            out.append(v)
        print(" | ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
