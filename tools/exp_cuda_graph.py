"""Does replaying the page forward as a CUDA graph beat 58 stream launches?  (launch gaps / CPU launch cost)
    python tools/exp_cuda_graph.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402

w, nc = synthetic_weights("textline")
m = SbbModel(w, 448, 448, nc, max_batch=48)
page = torch.from_numpy(synth.document_page(2800, 2000, seed=0)).cuda()
out = torch.empty((2800, 2000), dtype=torch.uint8, device="cuda")
ref = m.predict_page(page).clone()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3):
        m.predict_page(page, out=out, stream=st.cuda_stream)
torch.cuda.synchronize()


def timed(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    m.predict_page(page, out=out, stream=st.cuda_stream)
torch.cuda.synchronize()
out.zero_()
with torch.cuda.stream(st):
    g.replay()
torch.cuda.synchronize()
print("graph replay reproduces the page call:", bool((out == ref).all()))
for rnd in range(2):
    a = timed(lambda: m.predict_page(page, out=out, stream=st.cuda_stream))
    with torch.cuda.stream(st):
        b = timed(lambda: g.replay())
    print(f"stream launches {a:.3f} ms/page   graph replay {b:.3f} ms/page")
m.close()
