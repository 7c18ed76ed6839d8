"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the per-launch summary tables committed under profiles/
(the columns bench.py's `traffic` reads).    python scripts/ncu_summarize.py raw.csv out_prefix
Writes <out_prefix>_{gemm,recfwd,recbwd,other}_full_summary.csv."""
import csv
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(c) if c in hdr else None for c in COLS]


def to_mbyte(v, unit):
    v = float(v.replace(",", "")) if v else 0.0
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


groups = {"gemm": [], "recfwd": [], "recbwd": [], "other": []}
for r in data:
    name = r[idx[0]]
    short = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    out = []
    for c, i in zip(COLS, idx):
        if i is None:
            out.append("")
        elif c == "Kernel Name":
            out.append(short)
        elif c.startswith("dram__bytes"):
            out.append("%.6f" % to_mbyte(r[i], units[i]))
        else:
            out.append(r[i].replace(",", ";"))
    key = "gemm" if "gemm" in short else "recfwd" if "lstmp_fwd" in short else "recbwd" if "lstmp_bwd" in short else "other"
    groups[key].append(out)
unit_row = ["" if i is None else ("Mbyte" if c.startswith("dram__bytes") else units[i]) for c, i in zip(COLS, idx)]
for k, v in groups.items():
    if not v:
        continue
    with open("%s_%s_full_summary.csv" % (sys.argv[2], k), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLS)
        w.writerow(unit_row)
        w.writerows(v)
    print(k, len(v), "launches")
