"""Condense `ncu --page raw --csv` output into the per-kernel table kept under profiles/.
usage: python tools/ncu_summary.py RAW.csv > profiles/NAME.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__inst_executed.sum", "warp_insts"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct")]
cols = [(k, n) for k, n in cols if k in ix]
w = csv.writer(sys.stdout)
w.writerow([n + (f" [{units[ix[k]]}]" if units[ix[k]] else "") for k, n in cols])
for r in data:
    out = []
    for k, n in cols:
        v = r[ix[k]]
        if n == "kernel":
            v = v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        out.append(v)
    w.writerow(out)
