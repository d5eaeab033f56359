"""Summarise an `ncu --csv --metrics ...` launch list of tools/prof_page.py: one line per layer of the
LAST page in the log.   python tools/ncu_launches.py gpurun_out/launches.csv gpurun_out/prof_page.log"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ID, KN, MN, MV = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
launches = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= MV or not r[ID].isdigit():
        continue
    d = launches.setdefault(int(r[ID]), {"kernel": r[KN]})
    try:
        d[r[MN]] = float(r[MV].replace(",", ""))
    except ValueError:
        pass
names = None
for line in open(sys.argv[2], errors="replace"):
    if line.startswith("LAYERS "):
        names = line.strip()[7:].split(",")
ls = [v for v in launches.values() if "memset" not in v["kernel"].lower()]
ls = ls[-len(names):]
short = {"gpu__time_duration.sum": "us", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor%",
         "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "tc_smem%",
         "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed": "xbar_rd%",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts%",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
         "dram__bytes_read.sum": "rdMB", "dram__bytes_write.sum": "wrMB",
         "sm__cycles_elapsed.avg.per_second": "GHz"}
print(f"{'layer':18s}" + "".join(f"{v:>10s}" for v in short.values()))
tot = 0.0
for n, d in zip(names, ls):
    line = f"{n:18s}"
    for k, s in short.items():
        v = d.get(k, float("nan"))
        if s == "us":
            v /= 1e3; tot += v
        if s in ("rdMB", "wrMB"):
            v /= 1e6
        if s == "GHz":
            v /= 1e9
        line += f"{v:10.2f}"
    print(line)
print(f"sum of kernel durations: {tot / 1e3:.3f} ms over {len(ls)} launches")
