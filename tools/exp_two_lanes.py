"""Experiment: two pages in flight per GPU (two model workspaces on two streams) so that the partial last wave
and the pipeline fill/drain of one page's launches are covered by the other page's kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sbb_textline_detection_b200 import synth
from sbb_textline_detection_b200.detector import synthetic_weights
from sbb_textline_detection_b200.model import SbbModel

w, nc = synthetic_weights("textline")
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = 40
models = [SbbModel(w, 448, 448, nc, max_batch=48) for _ in range(lanes)]
streams = [torch.cuda.Stream() for _ in range(lanes)]
pages = [torch.from_numpy(synth.document_page(2800, 2000, seed=i)).cuda() for i in range(4)]
outs = [torch.empty((2800, 2000), dtype=torch.uint8, device="cuda") for _ in range(lanes)]
ref = models[0].predict_page(pages[1]).clone()
for i in range(2 * lanes):
    models[i % lanes].predict_page(pages[i % 4], out=outs[i % lanes], stream=streams[i % lanes].cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
done = [torch.cuda.Event() for _ in range(lanes)]
for s in streams[1:]:
    pass
e0.record(streams[0])
for s in streams[1:]:
    s.wait_event(e0)
for i in range(steps):
    l = i % lanes
    models[l].predict_page(pages[i % 4], out=outs[l], stream=streams[l].cuda_stream)
for l in range(1, lanes):
    done[l].record(streams[l])
    streams[0].wait_event(done[l])
e1.record(streams[0])
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
last = (steps - 1)
ok = bool((outs[last % lanes] == models[0].predict_page(pages[last % 4])).all())
print(f"lanes={lanes}: {ms:.3f} ms/page -> {1000/ms:.2f} pages/s, last output equal to a fresh single-lane run: {ok}")
