#!/usr/bin/env python
"""bench.py -- pages/sec of the tiled textline segmentation hot path (BASELINE.json configs[1]:
one synthetic 2800x2000x3 uint8 page, textline model, 448x448 tiles, 48 tiles/page).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one page per GPU through do_prediction(patches=True): /255, tiling, 61-conv forward per
tile, argmax, margin-crop stitch.  N>1 is launched by torchrun, one rank per GPU; pages are
independent, so ranks share nothing after the one init-time NCCL weight broadcast (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGE_H, PAGE_W, TILE, N_CLASSES = 2800, 2000, 448, 2
TILES_PER_PAGE = 48
METRIC = "pages/sec (2800x2000 textline seg)"
WORKLOAD = ("configs[1]: single 2800x2000x3 uint8 synthetic page, textline model (ResNet50-U-Net, 2 classes, "
            "random-init calibrated), 448x448 tiles, margin 44 -> 6x8=48 tiles/page")


def ncu_traffic(group: str):
    """DRAM bytes (read + write) per launch of a kernel group from the committed ncu launch list of the
    SAME workload (profiles/*_traffic.json, written by tools after an `ncu --metrics dram__bytes_*` pass)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    g = d["groups"].get(group)
    if not g:
        return None, None
    return g["dram_bytes"] / g["launches"], {"file": os.path.relpath(files[-1], ROOT), "launches_per_page": g["launches"],
                                             "dram_bytes_per_page": g["dram_bytes"]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["bf16_tflops"], "measured (MEASURED_PEAKS.json, bf16 sustained)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def oracle_tiles_per_sec(n_tiles: int, threads: int):
    """The reference's schedule on the CPU oracle (kind 'port'): batch-1 predict per tile, sequential,
    fp32, model resident.  Returns (seconds for n_tiles, n_tiles)."""
    import torch
    from oracle.resnet50_unet import OracleNet
    from sbb_textline_detection_b200 import synth
    from sbb_textline_detection_b200.detector import synthetic_weights
    torch.set_num_threads(threads)
    w, nc = synthetic_weights("textline")
    net = OracleNet(w, nc).as_keras_like(TILE, TILE)
    page = synth.document_page(PAGE_H, PAGE_W, seed=0).astype(np.float64) / 255.0
    from oracle.do_prediction import tile_grid
    _, _, _, tiles = tile_grid(PAGE_H, PAGE_W, TILE, TILE)
    net.predict(page[None, :TILE, :TILE])  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for (_, _, x0, y0) in tiles[:n_tiles]:
        p = net.predict(page[None, y0:y0 + TILE, x0:x0 + TILE])
        np.argmax(p, axis=3)
    return time.perf_counter() - t0, n_tiles


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    t1, _ = oracle_tiles_per_sec(1, threads)
    # bounded sample per step: about 5 s of CPU work, at most one full page
    per_step = int(max(1, min(TILES_PER_PAGE, 5.0 / max(t1, 1e-3))))
    for _ in range(args.warmup):
        oracle_tiles_per_sec(min(per_step, 2), threads)
    times = []
    for _ in range(args.steps):
        t, n = oracle_tiles_per_sec(per_step, threads)
        times.append(t)
    tot = sum(times)
    pages = args.steps * per_step / TILES_PER_PAGE
    value = pages / tot
    sample = (f"{per_step} of {TILES_PER_PAGE} tiles of the 2800x2000 page per step, batch-1 sequential model.predict "
              f"per tile (the reference's schedule, main.py:259-288), fp32, PyTorch-CPU oracle port with the model "
              f"kept resident; pages/s = tiles/48/s")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD},
           "cpu_baseline": {"value": value, "unit": "pages/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def kernel_group(name: str) -> str:
    if name.startswith("dec5"):  # one merged-parity N = 128 GEMM unless SBB_DEC5_MERGED=0 (four N = 32 variants)
        return "conv_gemm_tc<BN=32,head>" if os.environ.get("SBB_DEC5_MERGED") == "0" else "conv_gemm_tc<BN=128,head>"
    if name in ("stem_pad", "bn_relu_maxpool"):
        return name
    if name.startswith("conv1") or name.startswith("dec4") or \
            (name.startswith("res2") and ("branch2a" in name or "branch2b" in name)):
        return "conv_gemm_tc<BN=64>"
    return "conv_gemm_tc<BN=128>"


def run_ours(args):
    import torch
    from sbb_textline_detection_b200 import arch, parallel, synth, weights
    from sbb_textline_detection_b200.detector import synthetic_weights
    from sbb_textline_detection_b200.model import SbbModel

    rank, world, local = parallel.init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the hot path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # frozen weights: packed on rank 0, one broadcast at init (NCCL over NVLink), excluded from timing
    t0 = time.perf_counter()
    blob = None
    if rank == 0:
        w, nc = synthetic_weights("textline")
        blob = weights.pack_blob(w, nc)
    blob = parallel.broadcast_blob(blob, src=0, device=dev)
    bcast_s = time.perf_counter() - t0
    model = SbbModel(blob, TILE, TILE, N_CLASSES, device=local, precision=args.precision, max_batch=TILES_PER_PAGE)
    del blob

    # input pool: different pages per step (seeded per rank), resident in HBM for the `value` leg and
    # in pinned host memory for the `e2e` leg.  Per-step activation working set (11 GB) >> 126 MB L2.
    pool = 4
    pages = [synth.document_page(PAGE_H, PAGE_W, seed=100 * rank + i) for i in range(pool)]
    d_pages = [torch.from_numpy(p).to(dev) for p in pages]
    d_out = torch.empty((PAGE_H, PAGE_W), dtype=torch.uint8, device=dev)
    h_pages = [torch.from_numpy(p).pin_memory() for p in pages]
    h_out = torch.empty((PAGE_H, PAGE_W), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.Stream(dev)  # the launching stream: kernels and the timing events share it
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- leg 1: device-resident inputs -> `value`
    for i in range(args.warmup):
        model.predict_page(d_pages[i % pool], out=d_out, stream=sp)
    barrier()
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get('SBB_BENCH_NO_CLOCKS')) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        model.predict_page(d_pages[i % pool], out=d_out, stream=sp)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = parallel.all_reduce_max(e0.elapsed_time(e1))
    launches = parallel.all_reduce_sum(float(model.last_launch_count() * args.steps))
    value = world * args.steps / (ms_total * 1e-3)

    # ---- leg 2: host buffers through the public API (H2D of the page + D2H of the label map inside)
    for i in range(min(args.warmup, 3)):
        model.predict_page(h_pages[i % pool].numpy(), out=h_out.numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        model.predict_page(h_pages[i % pool].numpy(), out=h_out.numpy())
    torch.cuda.synchronize(dev)
    e2e_sync_s = parallel.all_reduce_max(time.perf_counter() - t0)
    e2e_sync = world * args.steps / e2e_sync_s
    # the batch form of the same public call: host pages in, host label maps out, every page's H2D and
    # D2H inside the timed region, copies of neighbouring pages overlapping the forward
    h_outs = [torch.empty((PAGE_H, PAGE_W), dtype=torch.uint8).pin_memory() for _ in range(pool)]
    batch_in = [h_pages[i % pool] for i in range(args.steps)]
    batch_out = [h_outs[i % pool] for i in range(args.steps)]
    model.predict_pages(batch_in[:min(args.warmup, 3)], outs=batch_out[:min(args.warmup, 3)])
    barrier()
    t0 = time.perf_counter()
    model.predict_pages(batch_in, outs=batch_out)
    torch.cuda.synchronize(dev)
    e2e_s = parallel.all_reduce_max(time.perf_counter() - t0)
    e2e = world * args.steps / e2e_s

    # ---- leg 3: per-kernel durations (CUDA event pair around every launch, same stream, same steps)
    model.set_profiling(True)
    groups: dict = {}
    parts: dict = {}  # encoder (stem + ResNet50 stages) / decoder (v4, v5, dec1..dec5 + head), SURVEY 8(d)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        model.predict_page(d_pages[i % pool], out=d_out, stream=sp)
        for name, ms, flops in model.layer_times():
            g = groups.setdefault(kernel_group(name), [0.0, 0.0, 0])
            g[0] += ms; g[1] += flops * TILES_PER_PAGE; g[2] += 1
            q = parts.setdefault("decoder" if name.startswith("dec") else "encoder", [0.0, 0.0])
            q[0] += ms; q[1] += flops * TILES_PER_PAGE
    model.set_profiling(False)
    tot_ms = sum(g[0] for g in groups.values())
    dom = max(groups, key=lambda k: groups[k][0])
    sustained, burst, how = peaks()
    ach = groups[dom][1] / (groups[dom][0] * 1e-3) / 1e12
    flop_page = arch.conv_flops_per_tile(TILE, TILE, N_CLASSES)[0] * TILES_PER_PAGE
    mma_factor = 3 if args.precision == "fp16x3" else 1
    traffic, traffic_src = ncu_traffic(dom)
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
        "frac": ach / sustained, "traffic": traffic, "traffic_source": traffic_src, "peak_source": how,
        "share_of_step": groups[dom][0] / tot_ms, "launches_per_page": groups[dom][2] // prof_steps,
        "note": (f"achieved = algorithmic conv FLOPs (2*MACs, SURVEY 8d) of this kernel's launches / their summed "
                 f"CUDA-event durations; the kernel issues {mma_factor}x that in tcgen05 MMA FLOPs (fp16 hi/lo split) "
                 f"= {ach * mma_factor:.0f} TFLOP/s = {ach * mma_factor / sustained:.2f} of peak"),
        "whole_step": {"alg_tflops": flop_page * value / world / 1e12,
                       "frac": flop_page * value / world / 1e12 / sustained},
        "groups": {k: {"ms_per_page": v[0] / prof_steps, "alg_tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[1] else 0.0}
                   for k, v in groups.items()},
        "parts": {k: {"ms_per_page": v[0] / prof_steps, "alg_tflops": v[1] / (v[0] * 1e-3) / 1e12,
                      "frac": v[1] / (v[0] * 1e-3) / 1e12 / sustained,
                      "issued_frac": mma_factor * v[1] / (v[0] * 1e-3) / 1e12 / sustained}
                  for k, v in parts.items()},
    }
    model.close()

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            t1, _ = oracle_tiles_per_sec(1, threads)
            n = int(max(2, min(TILES_PER_PAGE, 15.0 / max(t1, 1e-3))))
            t, n = oracle_tiles_per_sec(n, threads)
            cpu_baseline = {"value": n / TILES_PER_PAGE / t, "unit": "pages/s", "cores": torch.get_num_threads(),
                            "kind": "port",
                            "sample": f"{n} of 48 tiles of the same 2800x2000 page, batch-1 sequential predict per tile "
                                      f"(reference schedule), fp32 PyTorch-CPU oracle, model resident; {t:.1f} s"}
        out = {
            "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "fp16x3 (fp16 hi+lo operand pairs, 3 tcgen05 MMAs per K step, fp32 accumulate; fp32-grade)"
            if args.precision == "fp16x3" else "fp16 (NOT within the reference tolerance)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_step": world, "parallelism": f"page-per-gpu x{world}",
                       "l2": "per-step activation working set ~11 GB >> 126 MB L2; input pool of 4 different pages",
                       "weights_broadcast_s": bcast_s},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pages/s", "h2d_bytes_per_step": PAGE_H * PAGE_W * 3 * world,
                    "d2h_bytes_per_step": PAGE_H * PAGE_W * world,
                    "sync_call_value": e2e_sync,
                    "note": "SbbModel.predict_pages(host pages) -> host label maps, pinned buffers, wall clock: every page's "
                            "H2D + forward + D2H inside the timed region, copies of neighbouring pages overlap the forward; "
                            "sync_call_value = one blocking SbbModel.predict_page(numpy) per page"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
