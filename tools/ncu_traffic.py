"""DRAM traffic per kernel group from an ncu launch list of tools/prof_page.py (the list must carry
dram__bytes_read.sum and dram__bytes_write.sum): writes the profiles/*_traffic.json that bench.py reads
for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/launches.csv gpurun_out/prof_page.log profiles/r01i_traffic.json"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_digest, kernel_group  # noqa: E402

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
ls = [v for v in launches.values() if "memset" not in v["kernel"].lower()][-len(names):]  # the LAST page in the log
groups = OrderedDict()
for n, d in zip(names, ls):
    g = groups.setdefault(kernel_group(n), {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    g["launches"] += 1
    g["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    g["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
def git_head():
    try:
        import subprocess
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except OSError:
        return None


# the kernel sources this list was measured with: bench.py refuses the file once csrc/ differs (the GPU box has no
# .git, so the digest is the binding stamp; SBB_GIT_HEAD lets the caller pass the commit the snapshot was taken at)
# what bench.kernel_group depends on, with the defaults in effect NOW: recorded explicitly so that the grouping of
# this list can be reproduced after the defaults move on
PLAN_KNOBS = {"SBB_PAIR": "2", "SBB_PAIR_HEAD": "1", "SBB_PAIR64": "1", "SBB_DEC4_MERGED": "1", "SBB_DEC5_MERGED": "1"}
out = {"csrc_digest": csrc_digest(), "git_head": os.environ.get("SBB_GIT_HEAD") or git_head(),
       "plan_env": {k: os.environ.get(k, d) for k, d in PLAN_KNOBS.items()},
       "source": f"{sys.argv[1]} (ncu --metrics ...,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 f"one 2800x2000 page; per-launch times are cold-cache and serialised)", "groups": groups}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
