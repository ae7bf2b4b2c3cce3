"""Summarises an `ncu --set full` capture of one step's kernels (raw page) into a small table for profiles/.
    ncu -i x.ncu-rep --page raw --csv > x.csv;  python tools/ncu_step_summary.py x.csv [cells bytes_per_cell]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
U = dict(zip(h, units))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0,
         "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
def val(d, k):
    try:
        return float(d[k].replace(",", "")) * SCALE.get(U.get(k, ""), 1.0)
    except Exception:
        return float("nan")
tot_t = tot_r = tot_w = tot_s = 0.0
for v in rows[2:]:
    d = dict(zip(h, v))
    t, r, w, s = val(d, KEYS[0]), val(d, KEYS[1]), val(d, KEYS[2]), val(d, KEYS[3])
    tot_t += t; tot_r += r; tot_w += w; tot_s += s
    print(f"{d['Kernel Name'][:46]:46s} grid {d.get('launch__grid_size','?'):>6s} x {d.get('launch__block_size','?'):>4s}  "
          f"{t*1e6:7.2f} us  dram read {r/1e6:7.2f} MB  write {w/1e6:7.2f} MB  L2 sectors {s:10.0f} ({s*32/1e6:6.1f} MB)  "
          f"warps active {val(d, KEYS[4]):5.1f} %  IPC/SM {val(d, KEYS[6]):4.2f}  regs {d.get(KEYS[7])}  dram {val(d, KEYS[10]):4.1f} % of peak")
print(f"{'sum over the step':46s} {'':21s}{tot_t*1e6:7.2f} us  dram read {tot_r/1e6:7.2f} MB  write {tot_w/1e6:7.2f} MB  L2 sectors {tot_s:10.0f} ({tot_s*32/1e6:6.1f} MB)")
if len(sys.argv) >= 4:
    alg = float(sys.argv[2]) * float(sys.argv[3])
    print(f"algorithmic bytes {alg/1e6:.2f} MB; dram traffic / algorithmic {(tot_r+tot_w)/alg:.2f}; L2 traffic / algorithmic {tot_s*32/alg:.2f}")
