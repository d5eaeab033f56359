import os, sys, time, tempfile
sys.path.insert(0, '.')
import cv2, numpy as np, torch
os.environ["SBB_SYNTHETIC_MODELS"] = "1"
from sbb_textline_detection_b200 import detector as D, synth, prepost
tmp = tempfile.mkdtemp()
page = synth.document_page(2800, 2000, seed=3)
det = D.textline_detector(os.path.join(tmp, "p.png"), tmp, "p", tmp)
det.image = page
def T(label, f, n=3):
    for _ in range(2): r = f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); print(f"{label:40s} {(time.perf_counter()-t)/n*1e3:8.2f} ms", flush=True)
    return r
T("extract_page", lambda: (setattr(det, "image", page), det.extract_page())[1])
det.image = page
T("extract_text_regions", lambda: det.extract_text_regions(page))
T("textline_contours", lambda: det.textline_contours(page))
d = det._device_page()
T("_device_view", lambda: det._device_view(page))
T("otsu_copy dev", lambda: prepost.otsu_copy(d))
m, _ = det.start_new_session_and_model(det.model_region_dir)
b = prepost.otsu_copy(d)
T("region predict_page dev", lambda: m.predict_page(b))
T("region predict_page dev + cpu", lambda: m.predict_page(b).cpu().numpy())
seg = m.predict_page(b).cpu().numpy()
T("np.repeat", lambda: np.repeat(seg[:, :, None], 3, axis=2))
mp, _ = det.start_new_session_and_model(det.model_page_dir)
T("resize down", lambda: prepost.resize_nearest(d, 448, 448))
sm = prepost.resize_nearest(d, 448, 448)
T("predict_full dev", lambda: mp.predict_full(sm))
sg = mp.predict_full(sm)
T("resize up", lambda: prepost.resize_nearest(sg, 2800, 2000))
fu = prepost.resize_nearest(sg, 2800, 2000)
T("dilate6", lambda: prepost.dilate(fu, iterations=6))
g = prepost.dilate(fu, iterations=6).cpu().numpy()
T("threshold+findContours host", lambda: cv2.findContours(cv2.threshold(g, 0, 255, 0)[1], cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE))
