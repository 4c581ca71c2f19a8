"""Summarise an ncu --csv launch list: python scripts/launches.py gpurun_out/launches_TAG.csv [first_row]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
ki, ni, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
rec = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) <= vi:
        continue
    d = rec.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})
    d[r[ni]] = float(r[vi].replace(",", ""))
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for k, d in rec.items():
    if k < tail:
        continue
    print("%4d %-28s %9.2f us  rd %8.2f MB  wr %8.2f MB" % (k, d["name"][:28], d.get("gpu__time_duration.sum", 0) / 1e3,
          d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
