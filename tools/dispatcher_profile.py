"""Where a page's time goes on the HOST in the three-model pipeline (one thread, document-like synthetic models):
    python tools/dispatcher_profile.py"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SBB_SYNTHETIC_MODELS"] = "semantic"
import cProfile  # noqa: E402
import pstats  # noqa: E402

import torch  # noqa: E402

from sbb_textline_detection_b200 import synth  # noqa: E402
from sbb_textline_detection_b200.pipeline import PageDispatcher  # noqa: E402

tmp = tempfile.mkdtemp()
pages = [synth.framed_page(2800, 2000, seed=i, frame=100 + 12 * i) for i in range(4)]
for workers in (1, 2, 3, 4, 6):
    with PageDispatcher(tmp, tmp, workers=workers) as d:
        list(d.map(pages[:workers + 1]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 24
        list(d.map([pages[i % 4] for i in range(n)]))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
    print(f"workers {workers}: {dt * 1e3:.2f} ms/page  {1 / dt:.1f} pages/s", flush=True)
d = PageDispatcher(tmp, tmp, workers=1)
pr = cProfile.Profile()
pr.enable()
list(d.map([pages[i % 4] for i in range(8)]))
pr.disable()
d.close()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
