"""Print per-kernel times of the LAST iteration from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list.  Usage: python tools/kernel_times.py <csv> <iterations in the capture>"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
d = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
    elif r[mi].startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r[ui], 1.0)
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = v
ids = sorted(d)
n = len(ids) // int(sys.argv[2]) if len(sys.argv) > 2 else len(ids)
for k in ids[-n:]:
    m = d[k]
    name = k[1].replace("sba::<unnamed>::", "").replace("void ", "")[:70]
    print(f"{m['gpu__time_duration.sum']:10.1f} us  rd {m.get('dram__bytes_read.sum', 0):8.1f} MB  wr {m.get('dram__bytes_write.sum', 0):8.1f} MB  {name}")
