"""How much of a page's time is per-launch overhead / wave tails?  One forward over 96 tiles against two forwards over
48 (device-resident float tiles, full decoder -- no margin crop, so absolute times exceed the page call's).
    python tools/exp_batch96.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sbb_textline_detection_b200 import _lib  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402

w, nc = synthetic_weights("textline")
x = torch.rand((96, 448, 448, 3), device="cuda")
lab = torch.empty((96, 448, 448), dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
for nb in (48, 96, 48, 96):
    m = SbbModel(w, 448, 448, nc, max_batch=nb)

    def run(n_tiles):
        _lib.check(_lib.lib().sbb_predict_tiles(m._handle(), C.c_void_p(x.data_ptr()), n_tiles, C.c_void_p(lab.data_ptr()), None, None,
                                                _lib.SBB_MEM_DEVICE, C.c_void_p(st.cuda_stream)))
    for _ in range(3):
        run(96)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        run(96)          # max_batch 48: two forwards of 48 tiles; max_batch 96: one forward of 96
    e1.record(st)
    torch.cuda.synchronize()
    print(f"max_batch {nb}: {e0.elapsed_time(e1) / 10 / 2:.3f} ms per 48 tiles")
    m.close()
