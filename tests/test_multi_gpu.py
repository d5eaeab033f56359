"""Latency mode (one page across GPUs): needs >= 2 GPUs, skipped otherwise.  The N>1 host logic is covered
on CPU below and in tests/test_parallel_cpu.py (gloo)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from sbb_textline_detection_b200 import parallel
from sbb_textline_detection_b200.model import compute_tile_grid


def test_tile_ranges_partition_the_loop_order():
    for n, w in [(48, 8), (63, 8), (154, 8), (5, 8), (48, 1), (7, 2)]:
        r = parallel.tile_ranges(n, w)
        assert len(r) == w and r[0][0] == 0 and sum(c for _, c in r) == n
        assert all(r[k][0] + r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(c for _, c in r) - min(c for _, c in r) <= 1


def test_disjoint_owners_make_max_a_union(built_lib):
    """Why a MAX all-reduce (and unordered peer stores) stitch correctly: every pixel has exactly one owner tile."""
    nx, ny, org, ox, oy = compute_tile_grid(4600, 3400, 672, 672, -1)
    owner = ox[None, :].astype(np.int32) * ny + oy[:, None].astype(np.int32)
    assert (ox >= 0).all() and (oy >= 0).all()
    parts = []
    for first, count in parallel.tile_ranges(nx * ny, 8):
        parts.append(np.where((owner >= first) & (owner < first + count), owner + 1, 0))
    assert (np.maximum.reduce(parts) == owner + 1).all()
    assert (np.sum([p > 0 for p in parts], axis=0) == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,T", [(1300, 1000, 448)])
def test_page_sharded_across_gpus_equals_single_gpu(built_lib, H, W, T):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(H), str(W), str(T)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert line, out.stdout[-2000:] + out.stderr[-4000:]
    res = json.loads(line[0][len("MGPU_RESULT "):])
    assert res["p2p_equal"] and res["allreduce_equal"] and res["abi_broadcast_equal"], res
