import time, numpy as np, torch, sys
sys.path.insert(0, '.')
from sbb_textline_detection_b200 import deskew
from oracle import deskew as odk
rng = np.random.default_rng(0)
for (h, w) in [(600, 900), (1500, 1100)]:
    m = np.zeros((h, w), np.uint8)
    for y in range(20, h - 20, 30): m[y:y+10, 40:w-40] = 1
    a = np.linspace(-25, 25, 80)
    t = time.time(); ref = odk.rotation_profiles_cv2(m, a); t_cpu = time.time() - t
    d = torch.from_numpy(m).cuda()
    deskew.rotation_profiles(d, a)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5): got = deskew.rotation_profiles(d, a)
    torch.cuda.synchronize(); t_gpu = (time.time() - t) / 5
    print(f"deskew profiles {h}x{w} x80 angles: cv2 {t_cpu*1e3:.0f} ms, GPU {t_gpu*1e3:.2f} ms (incl. host matrices + D2H), equal={bool((got==ref).all())}")
