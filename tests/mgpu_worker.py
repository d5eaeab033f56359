"""Worker of tests/test_multi_gpu.py (one process per GPU, launched by torchrun): the latency mode of
parallel.PageSharder must reproduce the single-GPU page call bit for bit in both exchange modes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sbb_textline_detection_b200 import parallel, synth, weights  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402


def main():
    H, W, T = (int(v) for v in sys.argv[1:4])
    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    blob = None
    if rank == 0:
        w, nc = synthetic_weights("textline")
        blob = weights.pack_blob(w, nc)
    # the weight broadcast through the C ABI (sbb_nccl_comm_create + sbb_model_broadcast) must deliver the same
    # bytes as the torch.distributed one
    via_torch = parallel.broadcast_blob(blob, src=0, device=dev)
    comm = parallel.AbiCommunicator(local)
    via_abi = comm.broadcast_blob(blob, src=0)
    comm.close()
    abi_equal = torch.tensor([1 if via_abi == via_torch else 0], device=dev)
    dist.all_reduce(abi_equal, op=dist.ReduceOp.MIN)
    blob = via_abi
    model = SbbModel(blob, T, T, 2, device=local, max_batch=48)
    page = torch.from_numpy(synth.document_page(H, W, seed=5)).to(dev) if rank == 0 else \
        torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    want = model.predict_page(page).cpu().numpy() if rank == 0 else None
    res = {"world": world, "page": [H, W], "tile": T, "abi_broadcast_equal": bool(abi_equal.item())}
    for mode in ("p2p", "allreduce"):
        sh = parallel.PageSharder(model, H, W, owner=0, mode=mode)
        out = sh.run(page)
        if rank == 0:
            res[mode + "_equal"] = bool(np.array_equal(out.cpu().numpy(), want))
        for _ in range(2):
            sh.run(page)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            sh.run(page)
        torch.cuda.synchronize()
        res[mode + "_ms_per_page"] = (time.perf_counter() - t0) / n * 1e3
        sh.close()
    if rank == 0:
        model.predict_page(page)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            model.predict_page(page)
        torch.cuda.synchronize()
        res["single_gpu_ms_per_page"] = (time.perf_counter() - t0) / 5 * 1e3
        print("MGPU_RESULT " + json.dumps(res), flush=True)
    model.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
