"""Top stall sites of an .ncu-rep (SASS view of the source page): which instructions the sampled
warps sit on.   python tools/ncu_hot.py gpurun_out/x.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
A, S, N, IE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        n = int(r[N])
    except (ValueError, IndexError):
        continue
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    data.append((n, idx, r[S], r[IE], top))
tot = sum(d[0] for d in data)
print(f"total samples {tot}")
for n, idx, src, ie, top in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100 * n / tot:5.1f}%  #{idx:5d} exec={ie:>9s}  {src[:70]:70s} {top}")
