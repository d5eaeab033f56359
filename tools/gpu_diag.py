"""GPU bring-up diagnostics (run on the B200 box through gpurun): per-layer error of each backend
against the CPU oracle, logits/labels parity, per-layer timing.  Prints, never asserts.

    python tools/gpu_diag.py --stage layers --backend simt
    python tools/gpu_diag.py --stage layers --backend tcgen05
    python tools/gpu_diag.py --stage page
    python tools/gpu_diag.py --stage time
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sbb_textline_detection_b200 import synth, weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402


def load_weights(name="textline", seed=1234, nc=2):
    st = np.load(os.path.join(ROOT, "sbb_textline_detection_b200", "data", f"bn_stats_{name}.npz"))
    return weights.apply_bn_stats(weights.random_init(seed, nc), st)


def stage_layers(args):
    import torch
    from oracle.resnet50_unet import OracleNet
    w = load_weights()
    page = synth.document_page(2800, 2000, seed=0)
    T = args.tile
    tiles = np.stack([page[360:360 + T, 360:360 + T], synth.uniform_page(T, T, 0)]).astype(np.float32) / 255.0
    net = OracleNet(w, 2, torch.float32)
    net.taps = {}
    t0 = time.time()
    with torch.no_grad():
        z_ref = net.logits(tiles).numpy()
    print(f"oracle forward of 2 tiles: {time.time() - t0:.1f}s  threads={torch.get_num_threads()}")
    taps = {k: v.permute(0, 2, 3, 1).numpy() for k, v in net.taps.items()}
    m = SbbModel(w, T, T, 2, backend=args.backend, precision=args.precision, max_batch=2)
    t0 = time.time()
    labels, probs, logits = m.predict_tiles(tiles, True, True, True)
    print(f"{args.backend}/{args.precision} predict_tiles: {time.time() - t0:.3f}s launches={m.last_launch_count()}")
    print(f"{'layer':10s} {'shape':>16s} {'max|ref|':>10s} {'max|err|':>10s} {'mean|err|':>10s} {'rel':>9s}")
    for i, (name, h, wd, c) in enumerate(m.activations()):
        ref = taps[name]
        errs, refs, means = [], [], []
        for t in range(2):
            a = m.read_activation(i, t)
            d = np.abs(a - ref[t])
            errs.append(d.max()); means.append(d.mean()); refs.append(np.abs(ref[t]).max())
        print(f"{name:10s} {str((h, wd, c)):>16s} {max(refs):10.4f} {max(errs):10.3e} {np.mean(means):10.3e} {max(errs) / max(refs):9.2e}")
    d = np.abs(logits - z_ref)
    lab_ref = z_ref.argmax(-1)
    p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
    print(f"logits: max|err|={d.max():.3e} mean={d.mean():.3e}   probs max|err|={np.abs(probs - p_ref).max():.3e}")
    print(f"labels: mismatch={np.mean(labels != lab_ref):.3e}  class1 frac ref={lab_ref.mean():.3f} gpu={labels.mean():.3f}")
    inter = np.logical_and(labels == 1, lab_ref == 1).sum(); union = np.logical_or(labels == 1, lab_ref == 1).sum()
    print(f"IoU(class1)={inter / max(union, 1):.6f}")
    if args.backend == "tcgen05":
        m.set_profiling(True)
        m.predict_tiles(tiles, True, False, False)
        tot = 0.0
        for name, ms, fl in m.layer_times():
            tot += ms
            print(f"  {name:22s} {ms:8.3f} ms  {2 * fl / max(ms, 1e-6) / 1e9:9.1f} TFLOP/s(alg, 2 tiles)")
        print(f"  total {tot:.3f} ms for 2 tiles")
    m.close()


def stage_page(args):
    import torch
    from oracle.do_prediction import do_prediction
    from oracle.resnet50_unet import OracleNet
    w = load_weights()
    T = args.tile
    H, W = args.page_h, args.page_w
    page = synth.document_page(H, W, seed=3)
    m = SbbModel(w, T, T, 2, backend=args.backend, precision=args.precision, max_batch=args.batch)
    t0 = time.time()
    lab = m.predict_page(page)
    print(f"predict_page {H}x{W}: {time.time() - t0:.3f}s launches={m.last_launch_count()}  class1 frac={lab.mean():.3f}")
    net = OracleNet(w, 2, torch.float32).as_keras_like(T, T)
    t0 = time.time()
    ref = do_prediction(True, page, net, predict_batch=4)[:, :, 0]
    print(f"oracle do_prediction: {time.time() - t0:.1f}s")
    mism = np.mean(lab != ref)
    inter = np.logical_and(lab == 1, ref == 1).sum(); union = np.logical_or(lab == 1, ref == 1).sum()
    print(f"page labels: mismatch={mism:.3e} IoU(class1)={inter / max(union, 1):.6f}")
    m.close()


def stage_time(args):
    import torch
    w = load_weights()
    T = args.tile
    page = synth.document_page(2800, 2000, seed=0)
    m = SbbModel(w, T, T, 2, backend="tcgen05", precision=args.precision, max_batch=args.batch)
    dpage = torch.from_numpy(page).cuda()
    out = torch.empty((2800, 2000), dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    st = ts.cuda_stream
    for _ in range(2):
        m.predict_page(dpage, out=out, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(args.iters):
        m.predict_page(dpage, out=out, stream=st)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"page 2800x2000 {args.precision} batch={args.batch}: {ms:.2f} ms/page  -> {1000 / ms:.2f} pages/s, "
          f"{48 * 87.85e9 / (ms * 1e-3) / 1e12:.1f} TFLOP/s(alg)")
    m.set_profiling(True)
    m.predict_page(dpage, out=out, stream=st)
    torch.cuda.synchronize()
    rows = m.layer_times()
    tot = sum(r[1] for r in rows)
    for name, ms_, fl in rows:
        print(f"  {name:22s} {ms_:8.3f} ms {100 * ms_ / tot:5.1f}%  {48 * fl / max(ms_, 1e-6) / 1e9:9.1f} TFLOP/s(alg)")
    print(f"  sum of layers {tot:.2f} ms")
    t0 = time.time()
    host = m.predict_page(page)
    print(f"host-buffer call: {1000 * (time.time() - t0):.1f} ms; equal to device result: {bool((host == out.cpu().numpy()).all())}")
    m.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True)
    ap.add_argument("--backend", default="tcgen05")
    ap.add_argument("--precision", default="fp16x3")
    ap.add_argument("--tile", type=int, default=448)
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--page-h", type=int, default=1000)
    ap.add_argument("--page-w", type=int, default=900)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    {"layers": stage_layers, "page": stage_page, "time": stage_time}[a.stage](a)
