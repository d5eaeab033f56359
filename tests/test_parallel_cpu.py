"""world_size-2 gloo tests of the N>1 path: page sharding, the one init-time weight broadcast,
max-over-ranks timing reduction.  No data-path collective exists (pages are independent)."""
import os
import subprocess
import sys

from conftest import ROOT
from sbb_textline_detection_b200.parallel import shard_pages

WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, sys.argv[1])
import numpy as np
from sbb_textline_detection_b200 import parallel, weights
rank, world, local = parallel.init_distributed("gloo")
blob = None
if rank == 0:
    w = weights.random_init(7, 2)
    small = {k: v for k, v in w.items()}
    blob = weights.pack_blob(small, 2)
got = parallel.broadcast_blob(blob, src=0)
h = hashlib.sha256(got).hexdigest()
pages = parallel.shard_pages(7, rank, world)
mx = parallel.all_reduce_max(10.0 + rank)
sm = parallel.all_reduce_sum(float(len(pages)))
print(f"RESULT rank={rank} world={world} n={len(got)} sha={h} pages={"-".join(map(str,pages))} max={mx} sum={sm}", flush=True)
'''


def test_shard_pages_partition():
    for world in (1, 2, 4, 8):
        parts = [shard_pages(64, r, world) for r in range(world)]
        assert sorted(p for part in parts for p in part) == list(range(64))
        assert all(len(part) == 64 // world for part in parts)
    assert shard_pages(7, 1, 2) == [1, 3, 5]


def test_gloo_world2_broadcast_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = sorted(l for o, _ in outs for l in o.splitlines() if l.startswith("RESULT"))
    assert len(lines) == 2
    f0, f1 = (dict(kv.split("=", 1) for kv in l.split()[1:]) for l in lines)
    assert f0["sha"] == f1["sha"] and f0["n"] == f1["n"] and int(f0["n"]) > 100_000_000
    assert f0["pages"] == "0-2-4-6" and f1["pages"] == "1-3-5"
    assert f0["max"] == f1["max"] == "11.0" and f0["sum"] == f1["sum"] == "7.0"
