"""BASELINE.json configs[4] at its stated size: 10 M instances x 128 features, bg=512, nsamples=4096, sharded over the GPUs
of one box (torchrun, one process per GPU).  Every rank generates its shard of X ON THE DEVICE (seed + rank), explains it in
row chunks with the engine's device-resident call (shared plan of M=128: two-word coalition rows), keeps phi on the device
and takes part in ONE all-gather of the phi blocks at the end (NCCL; the gathered matrix is 20.5 GB per rank).  Prints one
JSON line on rank 0.  usage: torchrun --nproc-per-node 8 scripts/gpu_cfg4_multi.py [total_instances]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedkernelshap_b200.datasets import dense_tabular  # noqa: E402
from distributedkernelshap_b200.engine import GpuKernelExplainer  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D, NBG, NS, CHUNK = 128, 512, 4096, 65536
wl = dense_tabular(4, D, NBG, seed=0)                       # model + background (replicated); X comes from the device below
eng = GpuKernelExplainer(wl["predictor"].predict_proba, wl["background"], link="logit", seed=0, device=local)
eng.shap_values(wl["X_explain"], nsamples=NS, l1_reg=False)   # shared plan of M = 128 built + uploaded
n_r = total // world + (1 if rank < total % world else 0)
gen = torch.Generator(device="cuda")
gen.manual_seed(1234 + rank)
X = torch.randn((n_r, D), dtype=torch.float64, device="cuda", generator=gen)
phi = torch.empty((2, n_r, D), dtype=torch.float64, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t0 = time.perf_counter()
e0.record(stream)
for lo in range(0, n_r, CHUNK):
    hi = min(n_r, lo + CHUNK)
    part = torch.empty((2, hi - lo, D), dtype=torch.float64, device="cuda")
    eng.explain_device(X[lo:hi].data_ptr(), hi - lo, part.data_ptr(), nsamples=NS)
    phi[:, lo:hi].copy_(part)
e1.record(stream)
gathered = None
if world > 1:
    pad = (total + world - 1) // world
    send = phi if n_r == pad else torch.cat([phi, torch.zeros((2, pad - n_r, D), dtype=torch.float64, device="cuda")], dim=1)
    gathered = torch.empty((world,) + tuple(send.shape), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(gathered.view(-1), send.contiguous().view(-1))
e2.record(stream)
torch.cuda.synchronize()
eng.check_status()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e2)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# additivity on a sample of this rank's rows (link(f(x)) - expected value), float64 on the host
idx = torch.arange(0, min(n_r, 4096), device="cuda")
fx = wl["predictor"].predict_proba(X[idx].cpu().numpy())
add_err = float(np.abs(phi[1, idx].sum(1).cpu().numpy() - (np.log(fx[:, 1] / fx[:, 0]) - eng.expected_value[1])).max())
if rank == 0:
    print(json.dumps({"workload": "BASELINE.json configs[4]: synthetic, 128 features, bg=512, nsamples=4096, l1_reg=False, logit link, "
                                  "X generated on the device per rank", "instances": total, "n_gpus": world,
                      "instances_per_gpu": n_r, "plan": "shared per M (two-word rows)", "compute_ms": float(ms[0]),
                      "compute_plus_allgather_ms": float(ms[1]), "value": total / (float(ms[1]) / 1e3), "unit": "instances/s",
                      "compute_only_value": total / (float(ms[0]) / 1e3), "wall_s": wall, "gathered_bytes_per_rank": int(gathered.numel() * 8) if gathered is not None else 0,
                      "additivity_max_abs_err_sample": add_err, "timing": "CUDA events on the engine stream, max over ranks"}))
if world > 1:
    dist.destroy_process_group()
