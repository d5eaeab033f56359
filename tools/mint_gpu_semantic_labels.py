"""Run on a B200 (gpurun): the GPU stage drivers on the page of tests/golden/ref_semantic_run.npz with the
document-like synthetic models -> gpurun_out/gpu_semantic_labels.npz.  Copied to tests/golden/ it is what
tests/test_semantic_run.py feeds the reference's host glue in the build container (where there is no GPU)."""
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SBB_SYNTHETIC_MODELS"] = "semantic"
from sbb_textline_detection_b200 import detector as D, synth  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ref_semantic_run.npz"))
h, w, seed, frame = (int(v) for v in g["page"])
tmp = tempfile.mkdtemp()
png = os.path.join(tmp, "page.png")
cv2.imwrite(png, synth.framed_page(h, w, seed=seed, frame=frame))
det = D.textline_detector(png, tmp, "page", tmp, cache_models=False)
coord, regions, textline = det.run_segmentation()
H, W, _ = (int(v) for v in g["image_page_shape"])
want_r = np.unpackbits(g["regions_packed"])[:H * W].reshape(H, W)
want_t = np.unpackbits(g["textline_packed"])[:H * W].reshape(H, W)
print("page_coord", list(coord), "golden", g["page_coord"].tolist())
print("region mismatch", np.mean((regions[:, :, 0] == 1) != (want_r == 1)), "textline mismatch", np.mean((textline != 0) != (want_t != 0)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "gpu_semantic_labels.npz"), page_coord=np.array(coord),
                    regions_packed=np.packbits(regions[:, :, 0] == 1), textline_packed=np.packbits(textline != 0))
