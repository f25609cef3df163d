"""Turns `ncu --page raw --csv` of a run into the per-launch table bench.py's roofline.traffic reads
(profiles/r01_ncu_all_launches_one_forward_*.csv): one row per launch of ONE forward (from a mask_synth launch up to the next).
usage: python tools/ncu_all_launches.py raw.csv out.csv"""
import csv
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
           "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum"]
SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    kcol, gcol = hdr.index("Kernel Name"), hdr.index("Grid Size")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    starts = [i for i, r in enumerate(data) if "mask_synth" in r[kcol]]
    if len(starts) >= 2:
        data = data[starts[-2]:starts[-1]]
    elif starts:
        data = data[starts[-1]:]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["idx", "kernel", "grid"] + [m for m, _ in cols])
        for i, r in enumerate(data):
            name = r[kcol].replace("void ", "").replace("nib::", "")[:60]
            vals = [float(r[c].replace(",", "")) * SCALE.get(units[c], 1.0) for _, c in cols]
            w.writerow([i, name, r[gcol]] + vals)
    print(f"{dst}: {len(data)} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
