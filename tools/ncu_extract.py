"""Pick the judged columns out of `ncu -i X.ncu-rep --page raw --csv` files.  Usage: python tools/ncu_extract.py title=file.csv [...]"""
import csv
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
for arg in sys.argv[1:]:
    title, path = arg.split("=", 1)
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    if len(rows) < 3:
        print(f"== {title}: no rows in {path}"); continue
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    print(f"== {title} ( {path.split('/')[-1]} )")
    print(" | ".join(f"{hdr[i]} [{units[i]}]" for i in idx))
    for r in rows[2:]:
        print(" | ".join(r[i][:44] for i in idx))
    print()
