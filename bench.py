#!/usr/bin/env python
"""bench.py -- pages/sec of the tiled segmentation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|5|5o]

--config 2 (default; BASELINE.json configs[1], the configuration the headline metric is quoted on):
    one synthetic 2800x2000x3 uint8 page, textline model, 448x448 tiles, 48 tiles/page.
    A "step" = one page per GPU through do_prediction(patches=True): /255, tiling, 61-conv forward per tile,
    argmax, margin-crop stitch.
--config 3: border + region (Otsu) + textline models on one 2800x2000 scan with a scanner border (97 tiles),
    every page with its own crop geometry; e2e goes through the page dispatcher (host image in, host label maps out).
--config 5 / 5o: 4600x3400 page, 672x672 tiles, the reference's margin rule (63 tiles) / 50 % overlap (130 tiles).

N>1 is launched by torchrun, one rank per GPU; pages are independent, so ranks share nothing after the one
init-time NCCL weight broadcast (weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
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

CONFIGS = {
    # name: page h, w, tile, margin (-1: the reference's int(0.1*tile)), tiles per page and model (page, region, textline)
    "2": dict(H=2800, W=2000, tile=448, margin=-1, tiles=(0, 0, 48), metric="pages/sec (2800x2000 textline seg)",
              workload="configs[1]: single 2800x2000x3 uint8 synthetic page, textline model (ResNet50-U-Net, 2 classes, "
                       "random-init calibrated), 448x448 tiles, margin 44 -> 6x8=48 tiles/page"),
    "3": dict(H=2800, W=2000, tile=448, margin=-1, tiles=(1, 48, 48), metric="pages/sec (2800x2000 border+region+textline)",
              workload="configs[2]: border + region (Otsu) + textline models on one 2800x2000x3 synthetic scan with a scanner "
                       "border (document-like synthetic weights, ResNet50-U-Net), 448x448 tiles: 1 + <=48 + <=48 tiles/page, "
                       "every page its own crop geometry"),
    "5": dict(H=4600, W=3400, tile=672, margin=-1, tiles=(0, 0, 63), metric="pages/sec (4600x3400 textline seg, 672 tiles)",
              workload="configs[4]: 4600x3400x3 synthetic page, textline model, 672x672 tiles, reference margin rule "
                       "int(0.1*672)=67 -> 7x9=63 tiles/page"),
    "5o": dict(H=4600, W=3400, tile=672, margin=168, tiles=(0, 0, 130), metric="pages/sec (4600x3400 textline seg, 672 tiles, 50% overlap)",
               workload="configs[4]: 4600x3400x3 synthetic page, textline model, 672x672 tiles, margin 168 = 50 % overlap "
                        "(stride 336) -> 10x13=130 tiles/page"),
}


def csrc_digest() -> str:
    """Content hash of the kernel sources: profiles/*_traffic.json files carry the digest they were measured with."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sbb_textline_detection_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(group: str):
    """DRAM bytes (read + write) per launch of a kernel group from the committed ncu launch list of the SAME
    workload (profiles/*_traffic.json, written by tools/ncu_traffic.py after an `ncu --metrics dram__bytes_*` pass).
    A file measured with other kernel sources than the ones in the tree is refused (traffic: null)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))   # r01.. < r02..: the newest round's
    if not files:
        return None, {"refused": "no profiles/*_traffic.json"}
    d = json.load(open(files[-1]))
    src = {"file": os.path.relpath(files[-1], ROOT), "git_head": d.get("git_head"), "csrc_digest": d.get("csrc_digest")}
    if d.get("csrc_digest") != csrc_digest():
        src["refused"] = f"measured with csrc digest {d.get('csrc_digest')}, the tree has {csrc_digest()}: stale"
        return None, src
    g = d["groups"].get(group)
    if not g:
        src["refused"] = f"no group {group}"
        return None, src
    src.update(launches_per_page=g["launches"], dram_bytes_per_page=g["dram_bytes"])
    return g["dram_bytes"] / g["launches"], src


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


def config_block(cfg, world):
    """The `config` object: the SAME keys and values in both arms (the driver compares them)."""
    return {"workload": cfg["workload"], "parallelism": f"page-per-gpu x{world}",
            "l2": "per-step activation working set ~11 GB >> 126 MB L2; input pool of 4 different pages"}


# ------------------------------------------------------------------------------------------------ CPU side
_WEIGHTS: dict = {}   # (config is pipeline, kind) -> (weights dict, n_classes): generated once per process
_NETS: dict = {}      # same key + tile -> resident oracle model


def oracle_nets(cfg, kinds, fresh: bool = False):
    """Resident oracle models of a config (cached across calls: the "model resident" figure must not pay for the
    synthetic weight generation).  fresh=True rebuilds the model OBJECTS from the cached weight dicts -- the
    re-creation the reference performs for every stage of every page."""
    from oracle.resnet50_unet import OracleNet
    from sbb_textline_detection_b200 import semantic
    from sbb_textline_detection_b200.detector import synthetic_weights
    out = {}
    for k in kinds:
        key = (cfg is CONFIGS["3"], k)
        if key not in _WEIGHTS:
            _WEIGHTS[key] = semantic.semantic_weights(k) if cfg is CONFIGS["3"] else synthetic_weights(k)
        nk = key + (cfg["tile"],)
        if fresh or nk not in _NETS:
            w, nc = _WEIGHTS[key]
            _NETS[nk] = OracleNet(w, nc).as_keras_like(cfg["tile"], cfg["tile"])
        out[k] = _NETS[nk]
    return out


def oracle_sample(cfg, n_tiles: int, threads: int, recreate: bool = False):
    """The reference's schedule on the CPU oracle (kind 'port'): batch-1 model.predict per tile, sequential, fp32
    (main.py:259-288).  `n_tiles` tiles of the config's page, spread over the models the config runs in the
    proportion of its tiles.  recreate=True: the model object is rebuilt from the weight dict once per model and
    sample, as the reference reloads each .h5 for every stage of every page (main.py:386, 442, 492).
    Returns (seconds, tiles done, seconds spent re-creating models)."""
    import torch
    from oracle.do_prediction import tile_grid
    from sbb_textline_detection_b200 import synth
    torch.set_num_threads(threads)
    T = cfg["tile"]
    kinds = [k for k, n in zip(("page", "region", "textline"), cfg["tiles"]) if n]
    page = (synth.framed_page if cfg is CONFIGS["3"] else synth.document_page)(cfg["H"], cfg["W"], seed=0)
    page = page.astype(np.float64) / 255.0
    _, _, _, tiles = tile_grid(cfg["H"], cfg["W"], T, T, None if cfg["margin"] < 0 else cfg["margin"])
    t_load = 0.0
    nets = oracle_nets(cfg, kinds)                       # weights generated / cached outside any timing
    if recreate:
        t0 = time.perf_counter()
        nets = oracle_nets(cfg, kinds, fresh=True)
        t_load = time.perf_counter() - t0
    total = sum(cfg["tiles"])
    done = 0
    t0 = time.perf_counter()
    for k, n in zip(("page", "region", "textline"), cfg["tiles"]):
        if not n:
            continue
        share = max(1, round(n_tiles * n / total))
        for (_, _, x0, y0) in tiles[:share]:
            p = nets[k].predict(page[None, y0:y0 + T, x0:x0 + T])
            np.argmax(p, axis=3)
            done += 1
    return time.perf_counter() - t0, done, t_load


def cpu_baseline(cfg, budget_s: float):
    import torch
    threads = os.cpu_count() or 1
    total = sum(cfg["tiles"])
    oracle_sample(cfg, 1, threads)                       # warm-up (thread pool, allocator)
    t1, n1, _ = oracle_sample(cfg, 1, threads)
    n = int(max(2, min(total, budget_s / max(t1 / n1, 1e-3))))
    t, n, t_load = oracle_sample(cfg, n, threads, recreate=True)
    resident = n / total / t
    n_models = sum(1 for v in cfg["tiles"] if v)
    as_is = 1.0 / (total * t / n + t_load)               # one re-creation of every model per page
    return {"value": resident, "unit": "pages/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} of the {total} tiles of one page, batch-1 sequential model.predict per tile (the reference's "
                      f"schedule, main.py:259-288), fp32 PyTorch-CPU oracle port, model resident; {t:.1f} s",
            "as_is_value": as_is,
            "as_is_note": f"same sample plus re-creating the {n_models} model object(s) once per page as the reference does "
                          f"for every stage (main.py:386, 442, 492): {t_load:.2f} s per page here (building the oracle from "
                          "the weight dict; the reference's keras load_model of a .h5 additionally parses the file and "
                          "builds a TF graph, so this is a lower bound on its reload cost)"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    total = sum(cfg["tiles"])
    t1, n1, _ = oracle_sample(cfg, 1, threads)
    per_step = int(max(1, min(total, 5.0 / max(t1 / n1, 1e-3))))   # bounded sample: about 5 s of CPU work per step
    for _ in range(args.warmup):
        oracle_sample(cfg, min(per_step, 2), threads)
    tot, done = 0.0, 0
    for _ in range(args.steps):
        t, n, _ = oracle_sample(cfg, per_step, threads)
        tot += t
        done += n
    value = done / total / tot
    sample = (f"{per_step} of the {total} tiles of the page per step, batch-1 sequential model.predict per tile (the "
              f"reference's schedule, main.py:259-288), fp32, PyTorch-CPU oracle port with the model kept resident; "
              f"pages/s = tiles/{total}/s")
    out = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": "pages/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_block(cfg, args.gpus),
           "cpu_baseline": {"value": value, "unit": "pages/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ GPU side
def kernel_group(name: str) -> str:
    """Kernel a layer runs on under the default plan (csrc/sbb_net.cu: op.pair): every N = 128 launch with at least
    4 K chunks on the CTA-pair kernel -- all of them except the stage-2 expand convs."""
    mode = os.environ.get("SBB_PAIR", "2")
    pair = mode != "0"
    multi_tap = name in ("dec1", "dec2", "dec3") or (name[:4] in ("res3", "res4", "res5") and name.endswith("branch2b"))
    one_by_one = name in ("dec_v4", "dec_v5") or (name[:4] in ("res3", "res4", "res5") and not name.endswith("branch2b"))
    dec4_pair = pair and os.environ.get("SBB_DEC4_MERGED", "1") != "0"
    if pair and (multi_tap or (dec4_pair and name == "dec4") or (mode == "2" and one_by_one)):
        return "conv_gemm_pair<BN=128>"   # CTA pairs, cta_group::2
    if name.startswith("dec5") and pair and os.environ.get("SBB_PAIR_HEAD", "1") != "0" and os.environ.get("SBB_DEC5_MERGED") != "0":
        return "conv_gemm_pair<BN=128,head>"
    if name.startswith("dec5"):  # one merged-parity N = 128 GEMM unless SBB_DEC5_MERGED=0 (four N = 32 variants)
        return "conv_gemm_tc<BN=32,head>" if os.environ.get("SBB_DEC5_MERGED") == "0" else "conv_gemm_tc<BN=128,head>"
    if name in ("stem_pad", "bn_relu_maxpool"):
        return name
    if pair and os.environ.get("SBB_PAIR64", "1") != "0" and \
            (name == "conv1" or (name.startswith("res2") and (name.endswith("branch2b") or name in ("res2b_branch2a", "res2c_branch2a")))):
        return "conv_gemm_pair<BN=64>"    # N = 64 launches with >= 4 K chunks
    if name.startswith("conv1") or name.startswith("dec4") or \
            (name.startswith("res2") and ("branch2a" in name or "branch2b" in name)):
        return "conv_gemm_tc<BN=64>"
    return "conv_gemm_tc<BN=128>"


def broadcast_weights(kinds, semantic_models, rank, dev):
    """rank 0 packs the frozen weights, ONE broadcast per model at init; returns ({kind: blob}, timing split)."""
    from sbb_textline_detection_b200 import parallel, semantic, weights
    from sbb_textline_detection_b200.detector import synthetic_weights
    split = {"pack_s": 0.0, "broadcast_s": 0.0}
    blobs = {}
    for k in kinds:
        blob = None
        if rank == 0:
            t0 = time.perf_counter()
            w, nc = semantic.semantic_weights(k) if semantic_models else synthetic_weights(k)
            blob = weights.pack_blob(w, nc)
            split["pack_s"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        blobs[k] = parallel.broadcast_blob(blob, src=0, device=dev)
        split["broadcast_s"] += time.perf_counter() - t0
    return blobs, split


def latency_mode_check(model, cfg, dev, world):
    """N>1, outside the timed region: ONE page across all ranks (parallel.PageSharder: every rank's head epilogue
    stores its tiles' pixels into rank 0's label map over NVLink; NCCL MAX all-reduce as the alternative) must equal
    the single-GPU page call bit for bit."""
    import torch
    import torch.distributed as dist
    from sbb_textline_detection_b200 import parallel, synth
    H, W = cfg["H"], cfg["W"]
    page = torch.from_numpy(synth.document_page(H, W, seed=77)).to(dev)
    out = {}
    want = model.predict_page(page, margin=cfg["margin"])
    torch.cuda.synchronize(dev)
    for mode in ("p2p", "allreduce"):
        sh = parallel.PageSharder(model, H, W, owner=0, mode=mode, margin=cfg["margin"])
        sh.run(page)                                     # warm-up (IPC mapping, geometry cache)
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        got = sh.run(page)
        torch.cuda.synchronize(dev)
        ms = (time.perf_counter() - t0) * 1e3
        eq = bool((got == want).all().item()) if dist.get_rank() == 0 else True
        out[mode] = {"equal": eq, "ms": parallel.all_reduce_max(ms)}
        sh.close()
    return {"equal": all(v["equal"] for v in out.values()), "ms": out["p2p"]["ms"], "allreduce_ms": out["allreduce"]["ms"],
            "what": f"one {H}x{W} page across {world} GPUs (tile ranges per rank, stitched through peer memory / NCCL MAX "
                    "all-reduce), incl. the NCCL broadcast of the page; compared with the single-GPU page call"}


def run_ours(args, cfg):
    import torch
    from sbb_textline_detection_b200 import arch, parallel, synth
    from sbb_textline_detection_b200.model import SbbModel

    t_init = time.perf_counter()
    rank, world, local = parallel.init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the hot path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.barrier()                      # first collective: NCCL communicator setup
    init_s = time.perf_counter() - t_init

    H, W, TILE, MARGIN = cfg["H"], cfg["W"], cfg["tile"], cfg["margin"]
    pipeline = cfg is CONFIGS["3"]
    kinds = [k for k, n in zip(("page", "region", "textline"), cfg["tiles"]) if n]
    tiles_per_page = sum(cfg["tiles"])
    n_classes = {"page": 2, "region": 4, "textline": 2}

    # frozen weights: packed on rank 0, one broadcast per model at init (NCCL over NVLink), excluded from timing
    blobs, split = broadcast_weights(kinds, pipeline, rank, dev)
    t0 = time.perf_counter()
    # workspace for TWO pages' tiles where that stays moderate: the many-page entry points (predict_pages,
    # sbb_predict_pages_stacked) then run two same-size pages per forward; the single-page call is unaffected
    per_page = max(cfg["tiles"])
    max_batch = 2 * per_page if (2 * per_page <= 128 and not pipeline) else min(65, per_page)
    models = {k: SbbModel(blobs[k], TILE, TILE, n_classes[k], device=local, precision=args.precision,
                          max_batch=max_batch) for k in kinds}
    split["create_s"] = time.perf_counter() - t0
    split["dist_init_s"] = init_s
    del blobs
    model = models["textline"]

    # input pool: different pages per step (seeded per rank), resident in HBM for the `value` leg and in pinned
    # host memory for the `e2e` leg.  Per-step activation working set (11 GB) >> 126 MB L2.
    pool = 4
    if pipeline:   # every page its own scanner border -> its own crop geometry
        pages = [synth.framed_page(H, W, seed=100 * rank + i, frame=100 + 12 * i) for i in range(pool)]
    else:
        pages = [synth.document_page(H, W, seed=100 * rank + i) for i in range(pool)]
    d_pages = [torch.from_numpy(p).to(dev) for p in pages]
    h_pages = [torch.from_numpy(p).pin_memory() for p in pages]
    stream = torch.cuda.Stream(dev)  # the launching stream: kernels and the timing events share it
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    if pipeline:
        import cv2
        from sbb_textline_detection_b200 import detector as D, prepost
        from sbb_textline_detection_b200.pipeline import PageDispatcher
        # hand the broadcast models to the drop-in class' per-process cache under the reference's file names
        tmp = tempfile.mkdtemp()
        det0 = D.textline_detector("<array>", tmp, "page", tmp, device=local, precision=args.precision)
        for k, attr in (("page", "model_page_dir"), ("region", "model_region_dir"), ("textline", "model_textline_dir")):
            key = (os.path.abspath(getattr(det0, attr)), local, None, args.precision, os.environ.get("SBB_SYNTHETIC_MODELS"))
            D._MODEL_CACHE[key] = models[k]
        # crop boxes once (host contour pass), so that the device-resident leg is pure GPU work
        crops = []
        for p in pages:
            det = D.textline_detector("<array>", tmp, "page", tmp, device=local, precision=args.precision)
            det.image = p
            _, coord = det.extract_page()
            crops.append(coord)

        def device_step(i):
            dp, c = d_pages[i % pool], crops[i % pool]
            with torch.cuda.stream(stream):
                small = prepost.resize_nearest(dp, TILE, TILE)
                seg = models["page"].predict_full(small, stream=sp)
                prepost.dilate(prepost.resize_nearest(seg, H, W), iterations=6)
                crop = dp[c[0]:c[1], c[2]:c[3]]
                models["region"].predict_page(prepost.otsu_copy(crop), stream=sp)
                models["textline"].predict_page(crop, stream=sp)
    else:
        d_out = torch.empty((H, W), dtype=torch.uint8, device=dev)

        def device_step(i):
            model.predict_page(d_pages[i % pool], margin=MARGIN, out=d_out, stream=sp)

    # ---- leg 1: device-resident inputs -> `value`
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get('SBB_BENCH_NO_CLOCKS')) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = 0
    e0.record(stream)
    for i in range(args.steps):
        device_step(i)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = parallel.all_reduce_max(e0.elapsed_time(e1))
    per_page_launches = sum(m.last_launch_count() for m in models.values()) + (9 if pipeline else 0)  # + byte-op kernels
    launches = parallel.all_reduce_sum(float(per_page_launches * args.steps))
    value = world * args.steps / (ms_total * 1e-3)

    # ---- leg 2: host buffers through the public API (H2D of the page + D2H of the label maps inside)
    if pipeline:
        disp = PageDispatcher(tmp, tmp, workers=args.workers, device=local, precision=args.precision)
        list(disp.map([h_pages[i % pool].numpy() for i in range(min(args.warmup, 3) + args.workers)]))
        barrier()
        t0 = time.perf_counter()
        res = list(disp.map([h_pages[i % pool].numpy() for i in range(args.steps)]))
        torch.cuda.synchronize(dev)
        e2e_s = parallel.all_reduce_max(time.perf_counter() - t0)
        d2h = sum(r[1].nbytes + r[2].nbytes for r in res) / len(res) + H * W    # region (x3) + textline maps + border map
        disp1 = PageDispatcher(tmp, tmp, workers=1, device=local, precision=args.precision)
        t0 = time.perf_counter()
        list(disp1.map([h_pages[i % pool].numpy() for i in range(max(args.steps // 3, 2))]))
        e2e_sync = world * max(args.steps // 3, 2) / parallel.all_reduce_max(time.perf_counter() - t0)
        disp.close(); disp1.close()
        batched = None
        e2e_note = (f"PageDispatcher(workers={args.workers}).map(host pages) -> (crop box, region label image, textline mask) "
                    "on the host, wall clock: upload, three models, byte ops, the host contour pass of the border stage and "
                    "all D2H copies inside the timed region, pages overlapped across workers; sync_call_value = one worker")
    else:
        h_out = torch.empty((H, W), dtype=torch.uint8).pin_memory()
        for i in range(min(args.warmup, 3)):
            model.predict_page(h_pages[i % pool].numpy(), margin=MARGIN, out=h_out.numpy())
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            model.predict_page(h_pages[i % pool].numpy(), margin=MARGIN, out=h_out.numpy())
        torch.cuda.synchronize(dev)
        e2e_sync = world * args.steps / parallel.all_reduce_max(time.perf_counter() - t0)
        # the batch form of the same public call: host pages in, host label maps out, every page's H2D and
        # D2H inside the timed region, copies of neighbouring pages overlapping the forward
        h_outs = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in range(pool)]
        ppf_warm = model.pages_per_forward(H, W, MARGIN)
        batch_in = [h_pages[i % pool] for i in range(args.steps)]
        batch_out = [h_outs[i % pool] for i in range(args.steps)]
        # warm-up with the same grouping as the timed call (both double buffers at their final shape)
        model.predict_pages(batch_in[:2 * ppf_warm], outs=batch_out[:2 * ppf_warm], margin=MARGIN)
        barrier()
        t0 = time.perf_counter()
        model.predict_pages(batch_in, outs=batch_out, margin=MARGIN)
        torch.cuda.synchronize(dev)
        e2e_s = parallel.all_reduce_max(time.perf_counter() - t0)
        d2h = H * W
        ppf = model.pages_per_forward(H, W, MARGIN)
        e2e_note = ("SbbModel.predict_pages(host pages) -> host label maps, pinned buffers, wall clock: every page's "
                    "H2D + forward + D2H inside the timed region, copies of neighbouring pages overlap the forward, "
                    f"{ppf} same-size page(s) per forward (sbb_predict_pages_stacked); "
                    "sync_call_value = one blocking SbbModel.predict_page(numpy) per page")
        # the same batching with device-resident pages (what `value` would be with pages_per_step = ppf per GPU)
        batched = None
        if ppf > 1:
            stack = torch.cat([d_pages[i % pool] for i in range(ppf)], dim=0)
            s_out = torch.empty((ppf * H, W), dtype=torch.uint8, device=dev)
            for _ in range(2):
                model.predict_pages_stacked(stack, ppf, margin=MARGIN, out=s_out, stream=sp)
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nb_steps = max(args.steps // ppf, 2)
            b0.record(stream)
            for _ in range(nb_steps):
                model.predict_pages_stacked(stack, ppf, margin=MARGIN, out=s_out, stream=sp)
            b1.record(stream)
            barrier()
            bms = parallel.all_reduce_max(b0.elapsed_time(b1))
            batched = {"pages_per_forward": ppf, "value": world * nb_steps * ppf / (bms * 1e-3), "unit": "pages/s",
                       "ms_per_page": bms / (nb_steps * ppf),
                       "note": "device-resident, sbb_predict_pages_stacked: the per-launch costs of the 58 launches are "
                               "paid once per forward; `value` above stays at one page per step (the named config)"}
    e2e = world * args.steps / e2e_s

    # ---- leg 3: per-kernel durations (CUDA event pair around every launch, same stream, same steps)
    for m in models.values():
        m.set_profiling(True)
    groups: dict = {}
    parts: dict = {}  # encoder (stem + ResNet50 stages) / decoder (v4, v5, dec1..dec5 + head), SURVEY 8(d)
    prof_steps = min(args.steps, 5)
    flop_page = 0.0
    for i in range(prof_steps):
        device_step(i)
        torch.cuda.synchronize(dev)
        for k, m in models.items():
            per_tile = arch.conv_flops_per_tile(TILE, TILE, n_classes[k])[0]
            n_t = cfg["tiles"][("page", "region", "textline").index(k)]
            if i == 0:
                flop_page += per_tile * n_t
            for name, ms, flops in m.layer_times():
                g = groups.setdefault(kernel_group(name), [0.0, 0.0, 0])
                g[0] += ms; g[1] += flops * n_t; g[2] += 1
                q = parts.setdefault("decoder" if name.startswith("dec") else "encoder", [0.0, 0.0])
                q[0] += ms; q[1] += flops * n_t
    for m in models.values():
        m.set_profiling(False)
    # encoder / decoder split of the UNDISTURBED forward: three events per forward instead of a pair per launch
    # (the per-launch pairs above add a gap to each of the 58 launches; they give the per-kernel SHARES)
    part_split = None
    if not pipeline:
        enc_fl = arch.conv_flops_per_tile(TILE, TILE, n_classes["textline"])[1] * cfg["tiles"][2]
        dec_fl = arch.conv_flops_per_tile(TILE, TILE, n_classes["textline"])[2] * cfg["tiles"][2]
        model.set_profiling(2)
        for i in range(3):
            device_step(i)
        model.part_times()                                # warm-up forwards discarded
        n_split = min(args.steps, 48)
        for i in range(n_split):
            device_step(i)
        enc_ms, dec_ms, n_fw = model.part_times()
        model.set_profiling(0)
        part_split = {"encoder": {"ms_per_page": enc_ms / n_fw, "alg_tflops": enc_fl / (enc_ms / n_fw * 1e-3) / 1e12},
                      "decoder": {"ms_per_page": dec_ms / n_fw, "alg_tflops": dec_fl / (dec_ms / n_fw * 1e-3) / 1e12}}
    tot_ms = sum(g[0] for g in groups.values())
    dom = max(groups, key=lambda k: groups[k][0])
    sustained, burst, how = peaks()
    ach = groups[dom][1] / (groups[dom][0] * 1e-3) / 1e12
    mma_factor = 3 if args.precision == "fp16x3" else 1
    traffic, traffic_src = ncu_traffic(dom) if cfg is CONFIGS["2"] else (None, {"refused": "launch list is of config 2"})
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
        "frac": ach / sustained, "frac_of_burst_peak": ach / burst, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": how, "share_of_step": groups[dom][0] / tot_ms, "launches_per_page": groups[dom][2] // prof_steps,
        "note": (f"achieved = algorithmic conv FLOPs (2*MACs, SURVEY 8d) of this kernel's launches / their summed "
                 f"CUDA-event durations; the kernel issues up to {mma_factor}x that in tcgen05 MMA FLOPs (fp16 hi/lo "
                 f"split; less where the decoder skips margin work and merges up-sampled taps)"),
        "whole_step": {"alg_tflops": flop_page * value / world / 1e12,
                       "frac": flop_page * value / world / 1e12 / sustained},
        "groups": {k: {"ms_per_page": v[0] / prof_steps, "alg_tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[1] else 0.0}
                   for k, v in groups.items()},
        "parts": ({k: dict(v, frac=v["alg_tflops"] / sustained) for k, v in part_split.items()} if part_split else
                  {k: {"ms_per_page": v[0] / prof_steps, "alg_tflops": v[1] / (v[0] * 1e-3) / 1e12,
                       "frac": v[1] / (v[0] * 1e-3) / 1e12 / sustained} for k, v in parts.items()}),
        "parts_how": ("three CUDA events per forward (start | first decoder launch | end) over back-to-back forwards, no "
                      "synchronisation or per-launch instrumentation"
                      if part_split else "sums of the per-launch event pairs (each pair adds a few microseconds)"),
    }
    latency = None
    if world > 1 and not pipeline and not args.no_latency_check:
        latency = latency_mode_check(model, cfg, dev, world)
    geom = {k: dict(zip(("hits", "misses"), m.geom_cache_stats())) for k, m in models.items()}
    for m in models.values():
        m.close()

    if rank == 0:
        base = None
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_baseline(cfg, 12.0)
        out = {
            "metric": cfg["metric"], "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "fp16x3 (fp16 hi+lo operand pairs, 3 tcgen05 MMAs per K step, fp32 accumulate; fp32-grade)"
            if args.precision == "fp16x3" else "fp16 (NOT within the reference tolerance)",
            "data": "synthetic",
            "config": config_block(cfg, world),
            "pages_per_step": world,
            "init": dict(split, note="excluded from the timing: rank 0 packs the weights (pack_s), torch.distributed / NCCL "
                                     "setup incl. the first barrier (dist_init_s), one broadcast per model (broadcast_s), "
                                     "sbb_model_create = weight split + upload + plan + TMA descriptors (create_s)"),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pages/s", "h2d_bytes_per_step": H * W * 3 * world,
                    "d2h_bytes_per_step": int(d2h) * world, "sync_call_value": e2e_sync, "note": e2e_note},
            "batched": batched,
            "gpu_launches": int(launches),
            "geometry_cache": geom,
            "roofline": roofline,
            "cpu_baseline": base,
        }
        if latency is not None:
            out["latency_mode"] = latency
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
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--workers", type=int, default=4, help="pages in flight in the config-3 dispatcher (e2e leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency-check", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.config == "3":
        os.environ["SBB_SYNTHETIC_MODELS"] = "semantic"
    if a.impl == "reference":
        run_reference(a, CONFIGS[a.config])
    else:
        run_ours(a, CONFIGS[a.config])
