"""Data-parallel plumbing: graphs are independent units, so rank r owns a contiguous shard of the
graphs (its own block-diagonal CSR and features) and the ONLY exchange per optimiser step is one
all-reduce(sum) of the flattened weight gradients [dW1|db1|dW2|db2] (502 003 floats = 2.0 MB at
F=1000,H=500,K=3) plus the loss scalar -- NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference has no distributed code at all; this is new capability layered on
the per-graph independence of its training loop (python/Training/TrainingNeural.py:371-388).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed when WORLD_SIZE > 1; one process per GPU."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` graphs for `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_max_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def flat_grad_layout(sizes) -> Tuple[list, int]:
    """Offsets of 16-byte aligned segments inside the flat gradient buffer (mirrors GCNEngine)."""
    offs, tot = [], 0
    for s in sizes:
        offs.append(tot)
        tot += (int(s) + 3) & ~3
    return offs, tot


# ---------------------------------------------------------------------------------------------------------------------
# Gradient exchange as one kernel over NVLink peer memory (csrc/peer.cu)
class _RawDeviceArray:
    """Lets torch wrap memory we allocated ourselves (cudaMalloc via gmc_peer_alloc) without copying."""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerAllReduce:
    """In-place all-reduce(sum) of one fp32 buffer that every rank of the node has mapped (cudaIpc): `tensor` is the
    local buffer (keep the gradients in it), `all_reduce_()` launches ONE kernel on the current stream that waits for the
    peers, reduces this rank's slice in rank order, writes it to all ranks and waits for theirs.  Construction is a
    collective (handles travel through torch.distributed); all ranks must construct their instances in the same order.
    Raises if the buffers cannot be shared (not one node, no peer access): callers fall back to NCCL."""

    def __init__(self, numel: int, device, group=None):
        import ctypes

        from . import _lib
        lib = _lib.lib()
        self._lib, self._check = lib, _lib.check
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("PeerAllReduce supports up to 16 ranks")
        self.numel = int(numel)
        self.padded = (self.numel + 3) // 4 * 4
        self.epoch = 0
        buf, flg = ctypes.c_void_p(), ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.gmc_peer_alloc(self.padded * 4, ctypes.byref(buf)), "gmc_peer_alloc")
            _lib.check(lib.gmc_peer_alloc(lib.gmc_peer_flag_bytes(), ctypes.byref(flg)), "gmc_peer_alloc")
            hb = lib.gmc_ipc_handle_bytes()
            h_buf, h_flg = ctypes.create_string_buffer(hb), ctypes.create_string_buffer(hb)
            _lib.check(lib.gmc_ipc_get_handle(buf, h_buf), "gmc_ipc_get_handle")
            _lib.check(lib.gmc_ipc_get_handle(flg, h_flg), "gmc_ipc_get_handle")
            mine = (os.uname().nodename, bytes(h_buf.raw), bytes(h_flg.raw))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            if any(e[0] != mine[0] for e in everyone):
                raise RuntimeError("PeerAllReduce needs all ranks on one node")
            self._local = (buf.value, flg.value)
            self._opened = []
            bufs, flags = (ctypes.c_void_p * self.world)(), (ctypes.c_void_p * self.world)()
            for q, (_, hb_q, hf_q) in enumerate(everyone):
                if q == self.rank:
                    bufs[q], flags[q] = buf.value, flg.value
                    continue
                pb, pf = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.check(lib.gmc_ipc_open_handle(hb_q, ctypes.byref(pb)), "gmc_ipc_open_handle")
                self._opened.append(pb.value)
                _lib.check(lib.gmc_ipc_open_handle(hf_q, ctypes.byref(pf)), "gmc_ipc_open_handle")
                self._opened.append(pf.value)
                bufs[q], flags[q] = pb.value, pf.value
            self._bufs, self._flags = bufs, flags
            self._holder = _RawDeviceArray(buf.value, self.padded)
            self.tensor = torch.as_tensor(self._holder, device=device)
        dist.barrier(group=group)                  # every rank has mapped every buffer before anyone launches

    def all_reduce_(self) -> torch.Tensor:
        self.epoch += 1
        self._check(self._lib.gmc_peer_allreduce_f32(self._bufs, self._flags, self.world, self.rank, self.padded,
                                                     self.epoch & 0xFFFFFFFF, torch.cuda.current_stream().cuda_stream),
                    "gmc_peer_allreduce_f32")
        return self.tensor


def peer_allreduce_wanted() -> bool:
    """Opt-in (GMC_PEER_ALLREDUCE=1).  Measured back to back on 8 B200s (scratch/ar_probe.py, 2 MB): 36.6 us per call
    against 32.4 us for NCCL 2.28 (NVLS) -- the exchange was never the cost of data parallelism here; what bench.py reports
    as the all-reduce interval (0.05-0.6 ms per rank) is the wait for the slowest rank of each step."""
    return os.environ.get("GMC_PEER_ALLREDUCE", "0") == "1"


def make_peer_allreduce(numel: int, device, group=None) -> Optional[PeerAllReduce]:
    """PeerAllReduce when every rank can build one, else None on every rank (the decision is itself a collective, so
    the ranks never disagree about which exchange they run)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    if not peer_allreduce_wanted() or dist.get_backend(group) != "nccl" or not torch.cuda.is_available():
        return None
    peer, ok = None, 1
    try:
        peer = PeerAllReduce(numel, device, group)
    except Exception as exc:       # noqa: BLE001 -- any failure means "use NCCL", reported once
        ok = 0
        if dist.get_rank(group) == 0:
            import warnings
            warnings.warn(f"peer-memory all-reduce unavailable ({exc}); using NCCL")
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return peer if int(flag.item()) == 1 else None
