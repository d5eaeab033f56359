"""Single-GPU timings of the BASELINE.json configurations that are not the bench.py line:

  cfg3  border + region (Otsu) + textline models on one 2800x2000 page through the drop-in class' stage
        drivers (host image in, host label maps out; models cached like a serving process would)
  cfg5  4600x3400 page, 672x672 tiles: the reference's margin rule (67 -> 7x9 = 63 tiles) and the
        "50 % overlap" reading of BASELINE.json (margin 168 -> stride 336 -> 10x13 = 130 tiles)

    python tools/bench_configs.py [--steps 10]          (prints one JSON line per configuration)
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sbb_textline_detection_b200 import arch, detector as D, synth  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel, compute_tile_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()

# ------------------------------------------------------------------ cfg3
os.environ["SBB_SYNTHETIC_MODELS"] = "1"
tmp = tempfile.mkdtemp()
page = synth.document_page(2800, 2000, seed=3)
png = os.path.join(tmp, "p.png")
cv2.imwrite(png, page)
det = D.textline_detector(png, tmp, "p", tmp)   # tile 448, models cached per process


stage_s = [0.0, 0.0, 0.0]


def three_stages():
    det.image = page
    t = [time.perf_counter()]
    image_page, coord = det.extract_page()
    t.append(time.perf_counter())
    # the synthetic border model crops arbitrarily; run the two tiled models on the FULL page so that the
    # workload is the configured 1 + 48 + 48 tiles
    reg = det.extract_text_regions(page)
    t.append(time.perf_counter())
    tl = det.textline_contours(page)
    t.append(time.perf_counter())
    for k in range(3):
        stage_s[k] += t[k + 1] - t[k]
    return coord, reg, tl


for _ in range(2):
    three_stages()
torch.cuda.synchronize()
stage_s = [0.0, 0.0, 0.0]
t0 = time.perf_counter()
for _ in range(a.steps):
    three_stages()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.steps
fl = sum(arch.conv_flops_per_tile(448, 448, nc)[0] * n for nc, n in ((2, 1), (4, 48), (2, 48)))
print(json.dumps({"config": "cfg3: border + region(Otsu) + textline on one 2800x2000 page, 97 tiles, host in / host out, "
                            "stage drivers of the drop-in class (page uploaded once; resize / Otsu / dilate on the GPU; border contour on the host)",
                  "ms_per_page": dt * 1e3, "pages_per_s": 1 / dt, "alg_tflops": fl / dt / 1e12,
                  "stage_ms": {"extract_page": stage_s[0] / a.steps * 1e3, "extract_text_regions": stage_s[1] / a.steps * 1e3,
                               "textline_contours": stage_s[2] / a.steps * 1e3},
                  "note": "extract_page includes the reference's host contour pass (findContours + a Python contourArea loop, "
                          "main.py:398-402) over the border label map; with random-init weights that map is noise with "
                          "thousands of contours, a trained border model yields a handful"}), flush=True)
D._MODEL_CACHE.clear()

# ------------------------------------------------------------------ cfg5
w, nc = D.synthetic_weights("textline")
big = synth.document_page(4600, 3400, seed=5)
for margin, what in ((-1, "reference margin rule int(0.1*672)=67"), (168, "margin 168 = 50% overlap (stride 336)")):
    nx, ny, org, _, _ = compute_tile_grid(4600, 3400, 672, 672, margin)
    m = SbbModel(w, 672, 672, nc, max_batch=min(nx * ny, 48))
    d_page = torch.from_numpy(big).cuda()
    d_out = torch.empty((4600, 3400), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    for _ in range(2):
        m.predict_page(d_page, margin=margin, out=d_out, stream=st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(a.steps):
        m.predict_page(d_page, margin=margin, out=d_out, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    host = m.predict_page(big, margin=margin)
    assert (host == d_out.cpu().numpy()).all()
    t0 = time.perf_counter()
    for _ in range(max(a.steps // 2, 1)):
        m.predict_page(big, margin=margin)
    e2e = (time.perf_counter() - t0) / max(a.steps // 2, 1)
    fl = arch.conv_flops_per_tile(672, 672, nc)[0] * nx * ny
    print(json.dumps({"config": f"cfg5: 4600x3400 page, 672x672 tiles, {what}: {nx}x{ny}={nx * ny} tiles, textline model",
                      "ms_per_page_device": ms, "pages_per_s_device": 1e3 / ms, "ms_per_page_host_in_out": e2e * 1e3,
                      "alg_tflops": fl / (ms * 1e-3) / 1e12, "label_fraction_class1": float((host == 1).mean())}), flush=True)
    m.close()
