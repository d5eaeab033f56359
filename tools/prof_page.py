"""Run N page forwards (2800x2000, device-resident) -- the command ncu wraps for launch lists and
per-kernel captures.   python tools/prof_page.py [--pages 2] [--precision fp16x3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=2)
ap.add_argument("--precision", default="fp16x3")
a = ap.parse_args()
w, nc = synthetic_weights("textline")
m = SbbModel(w, 448, 448, nc, precision=a.precision, max_batch=48)
page = torch.from_numpy(synth.document_page(2800, 2000, seed=0)).cuda()
out = torch.empty((2800, 2000), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(a.pages):
    m.predict_page(page, out=out, stream=st)
torch.cuda.synchronize()
names = [n for n, _, _ in m.layer_times()]
print("LAYERS " + ",".join(names))
m.close()
