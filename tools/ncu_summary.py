"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the handful of columns the roofline discussion uses.
usage: python tools/ncu_summary.py raw.csv out.csv"""
import csv
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(dst, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in cols])
        for r in data:
            wr.writerow([r[i] for _, i in cols])
    print(f"{dst}: {len(data)} launches x {len(cols)} metrics")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
