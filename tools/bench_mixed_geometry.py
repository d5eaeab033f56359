"""Page-geometry cache under production-like input: every page of a run has its own border crop (main.py:2061 ->
2072), i.e. its own (H, W).  Times device-resident page calls for (a) one fixed 2800x2000 geometry, (b) 8 different
geometries in rotation (all cache hits after the first round), (c) 24 different geometries in rotation (every call
a miss that refills the least recently used of the 8 slots).   python tools/bench_mixed_geometry.py [--steps 48]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=48)
a = ap.parse_args()
w, nc = synthetic_weights("textline")
m = SbbModel(w, 448, 448, nc, max_batch=48)
base = torch.from_numpy(synth.document_page(2800, 2000, seed=0)).cuda()
st = torch.cuda.Stream()


def crops(n):  # same 6 x 8 = 48 tile grid, different sizes: what different border crops of one scan size give
    return [base[3 * k:2800 - 2 * k, 2 * k:2000 - 5 * k] for k in range(n)]


def run(pages, label):
    outs = [torch.empty(p.shape[:2], dtype=torch.uint8, device="cuda") for p in pages]
    for k in range(max(len(pages), 4)):
        m.predict_page(pages[k % len(pages)], out=outs[k % len(pages)], stream=st.cuda_stream)
    torch.cuda.synchronize()
    h0, m0 = m.geom_cache_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for k in range(a.steps):
        m.predict_page(pages[k % len(pages)], out=outs[k % len(pages)], stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    h1, m1 = m.geom_cache_stats()
    ms = e0.elapsed_time(e1) / a.steps
    return {"case": label, "ms_per_page": ms, "pages_per_s": 1e3 / ms, "cache_hits": h1 - h0, "cache_misses": m1 - m0}


res = [run([base], "fixed 2800x2000"), run(crops(8), "8 geometries in rotation (hits)"),
       run(crops(24), "24 geometries in rotation (every call refills a slot)"), run([base], "fixed 2800x2000 (again)")]
fixed = 0.5 * (res[0]["pages_per_s"] + res[3]["pages_per_s"])
for r in res:
    r["vs_fixed"] = r["pages_per_s"] / fixed
    print(json.dumps(r), flush=True)
m.close()
