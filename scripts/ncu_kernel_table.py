"""Per-kernel means (launches, grid, duration, DRAM bytes) of an `ncu --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ii, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
d, grid = collections.defaultdict(dict), {}
for r in data:
    if len(r) > vi:
        d[(r[ii], r[ki])][r[mi]] = float(r[vi].replace(",", ""))
        grid[(r[ii], r[ki])] = r[gi]
agg = collections.defaultdict(list)
for (i, k), m in d.items():
    agg[(k.split("(")[0].replace("void ", "").replace("<unnamed>::", ""), grid[(i, k)])].append(m)
print("kernel,grid,launches,mean_us,mean_dram_read_MB,mean_dram_write_MB,read_GBps")
for (k, g), ms in agg.items():
    n = len(ms)
    t = sum(m.get("gpu__time_duration.sum", 0) for m in ms) / n
    rd = sum(m.get("dram__bytes_read.sum", 0) for m in ms) / n
    wr = sum(m.get("dram__bytes_write.sum", 0) for m in ms) / n
    print('%s,"%s",%d,%.2f,%.2f,%.2f,%.0f' % (k, g, n, t / 1e3, rd / 1e6, wr / 1e6, rd / t if t else 0))
