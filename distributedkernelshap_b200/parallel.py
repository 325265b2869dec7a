"""Multi-GPU plumbing: one process per GPU with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in CPU
tests).  The KernelSHAP path shards over instances with no exchange during compute; the single collective is an
all-gather of the shap-value blocks (replaces ``DistributedExplainer.order_result`` + plasma gets,
explainers/distributed.py:152-179)."""
import os

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed():
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    try:
        dist = _dist()
    except ImportError:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return _dist().get_rank() if is_distributed() else 0


def world_size():
    return _dist().get_world_size() if is_distributed() else 1


def visible_gpus():
    """Number of CUDA devices this process can see (through libdks, without importing torch)."""
    import ctypes
    from . import _cabi
    n = ctypes.c_int(0)
    if _cabi.load().dks_device_count(ctypes.byref(n)) != 0:
        return 0
    return int(n.value)


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    import torch
    dist = _dist()
    if dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)


def shard_bounds(n, world):
    """Row range of every rank under the ``np.array_split`` rule: the first ``n % world`` ranks get one extra row."""
    base, extra = divmod(n, world)
    sizes = [base + 1 if r < extra else base for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(world)]


def allgather_rows(local, counts):
    """All-gather blocks ``local`` [C, n_r, G] (float64) whose row counts ``counts`` may differ by rank; returns the
    concatenation [C, sum(counts), G] on every rank.  Blocks are padded to the largest count because the collective
    needs equal sizes."""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    C, _, G = local.shape
    pad = max(counts)
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    send = torch.zeros((C, pad, G), dtype=torch.float64, device=device)
    if local.shape[1]:
        send[:, :local.shape[1]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    recv = torch.empty(world * send.numel(), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send.view(-1))
    recv = recv.view(world, C, pad, G).cpu().numpy()
    return np.concatenate([recv[r][:, :counts[r]] for r in range(world)], axis=1)


def allgather_rows_device(local_dev, counts):
    """Same collective for a block that is still on the GPU: ``local_dev`` [C, n_r, G] float64 CUDA tensor.  All-gathers on
    the device (NCCL), puts the rows in global order on the device, and makes ONE device-to-host copy into pinned memory
    (torch's caching host allocator: no cudaHostAlloc per call); returns a NumPy view [C, sum(counts), G] that owns it.
    Replaces the D2H -> H2D -> all-gather -> D2H x world round trip of ``allgather_rows`` on the SPMD host path."""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    C, nr, G = local_dev.shape
    pad = max(counts)
    if nr == pad:
        send = local_dev.contiguous()
    else:
        send = torch.zeros((C, pad, G), dtype=torch.float64, device=local_dev.device)
        send[:, :nr] = local_dev
    recv = torch.empty((world, C, pad, G), dtype=torch.float64, device=local_dev.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
    if all(c == pad for c in counts):
        ordered = recv.permute(1, 0, 2, 3).reshape(C, world * pad, G)
    else:
        ordered = torch.cat([recv[r][:, :counts[r]] for r in range(world)], dim=1)
    host = torch.empty(ordered.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(ordered, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


class PeerGather:
    """Gathered ``[world, C, n, G]`` float64 buffer in symmetric (peer-mapped) memory, one per rank.

    The engine stores its block of phi into slab ``rank`` of every peer's buffer with its own kernel over NVLink peer
    memory (``engine.set_peers`` / ``dks_set_peers``); ``barrier()`` is the cross-GPU signal exchange that makes the
    gathered buffer complete on every rank.  torch only allocates and maps the memory
    (``torch.distributed._symmetric_memory``).  Raises if the process group / driver cannot provide peer mappings: callers
    fall back to ``all_gather_into_tensor``."""

    def __init__(self, engine, C, n, G, device, own_sync=True):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        dist = _dist()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.buffer = symm_mem.empty((self.world, C, n, G), dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.buffer, dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.local = self.buffer[self.rank]                       # the solve writes this slab in place
        self.engine = engine
        engine.set_peers(self.world, self.rank, ptrs, C * n * G)
        # the engine's own completion protocol: one flag word per peer in peer-mapped memory (signal + wait at the end of
        # every explain call) instead of the symmetric-memory barrier
        self.own_sync = False
        if own_sync:
            self.flags = symm_mem.empty((max(self.world, 2),), dtype=torch.int64, device=device)
            self.flags.zero_()
            self.flag_handle = symm_mem.rendezvous(self.flags, dist.group.WORLD)
            torch.cuda.synchronize()
            dist.barrier()                                        # every rank has zeroed its flags before anyone signals
            engine.set_peer_flags([int(p) for p in self.flag_handle.buffer_ptrs])
            self.own_sync = True

    def barrier(self):
        if not self.own_sync:                                     # with flags the explain call itself ends with the exchange
            self.handle.barrier(channel=0)

    def close(self):
        self.engine.set_peers(0, 0, None, 0)
