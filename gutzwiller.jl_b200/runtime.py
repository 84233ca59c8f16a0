"""Device contexts, model handles and multi-GPU wiring for the host mirror.

One process drives one GPU (``LOCAL_RANK``); several processes are joined by one NCCL communicator
whose unique id is shipped between ranks with ``torch.distributed`` (plumbing only).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _capi
from ._capi import check, lib


class Context:
    """gwz_context: one CUDA device + stream (+ optional NCCL communicator)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().gwz_context_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self._finalizer = weakref.finalize(self, lib().gwz_context_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        dev, sm, clk, maj, mnr = (C.c_int() for _ in range(5))
        mem = C.c_size_t()
        check(lib().gwz_context_info(self._h, C.byref(dev), C.byref(sm), C.byref(clk), C.byref(maj), C.byref(mnr),
                                     C.byref(mem)))
        return dict(device=dev.value, sm_count=sm.value, sm_clock_khz=clk.value, cc=(maj.value, mnr.value),
                    global_mem_bytes=mem.value)

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().gwz_context_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self) -> None:
        check(lib().gwz_context_synchronize(self._h))

    def comm_info(self):
        w, r = C.c_int(), C.c_int()
        check(lib().gwz_context_comm_info(self._h, C.byref(w), C.byref(r)))
        return w.value, r.value

    def comm_init(self, world_size: int, rank: int, unique_id: bytes) -> None:
        buf = (C.c_uint8 * _capi.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        check(lib().gwz_context_comm_init(self._h, world_size, rank, buf))


def launch_count() -> int:
    """Kernels launched by the library so far (process-wide)."""
    n = C.c_uint64()
    check(lib().gwz_launch_count(C.byref(n)))
    return n.value


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * _capi.NCCL_ID_BYTES)()
    check(lib().gwz_nccl_unique_id(buf))
    return bytes(buf)


def comm_init_all(contexts) -> None:
    """Join several contexts of THIS process (one per device) in one communicator."""
    arr = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    check(lib().gwz_context_comm_init_all(arr, len(contexts)))


_default: dict[int, Context] = {}
_dist = dict(world=1, rank=0)


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _dist["world"] > 1 else 0
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [begin, end) of `total` global walker ids (or ranks) owned by `rank`.

    GPU r owns [r*total/G, (r+1)*total/G); the first total % world shards get one more.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(total), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def init_distributed(backend_for_id: str = "gloo") -> tuple[int, int]:
    """One process per GPU under torchrun: create the default context on LOCAL_RANK and join the
    NCCL communicator.  Returns (world_size, rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    _dist.update(world=world, rank=rank)
    if world == 1:
        default_context(0)
        return 1, 0
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group(backend=backend_for_id)
    ctx = default_context(int(os.environ.get("LOCAL_RANK", str(rank))))
    uid = torch.zeros(_capi.NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, src=0)
    ctx.comm_init(world, rank, bytes(uid.numpy().tobytes()))
    return world, rank


def distributed_state() -> tuple[int, int]:
    return _dist["world"], _dist["rank"]


class Model:
    """gwz_model: HubbardReal1D(BoseFS{N,M}; u, t) + GutzwillerAnsatz on one context."""

    def __init__(self, N: int, M: int, u: float, t: float, ctx: Context | None = None, v: float = 0.0):
        self.ctx = ctx or default_context()
        self._h = C.c_void_p()
        if v:  # ExtendedHubbardReal1D: served by the ansatz-generic engine only
            check(lib().gwz_model_create_ext_hubbard1d(self.ctx.handle, int(N), int(M), float(u), float(v), float(t),
                                                       C.byref(self._h)))
        else:
            check(lib().gwz_model_create_hubbard1d_gutzwiller(self.ctx.handle, int(N), int(M), float(u), float(t),
                                                              C.byref(self._h)))
        self.N, self.M, self.u, self.t, self.v = int(N), int(M), float(u), float(t), float(v)
        bits, words = C.c_int(), C.c_int()
        check(lib().gwz_model_info(self._h, None, None, None, None, C.byref(bits), C.byref(words)))
        self.bits, self.words = bits.value, words.value
        self._finalizer = weakref.finalize(self, lib().gwz_model_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def dimension(self) -> int:
        d, de = C.c_double(), C.c_uint64()
        check(lib().gwz_model_dimension(self._h, C.byref(d), C.byref(de)))
        return int(de.value) if de.value else int(d.value)

    def near_uniform(self) -> np.ndarray:
        a = np.zeros(self.words, dtype=np.uint64)
        check(lib().gwz_model_near_uniform(self._h, a.ctypes.data_as(C.POINTER(C.c_uint64))))
        return a

    def from_onr(self, onr) -> np.ndarray:
        o = np.ascontiguousarray(onr, dtype=np.int32)
        if o.shape != (self.M,):
            raise ValueError(f"onr must have {self.M} entries")
        a = np.zeros(self.words, dtype=np.uint64)
        check(lib().gwz_model_from_onr(self._h, o.ctypes.data_as(C.POINTER(C.c_int32)),
                                       a.ctypes.data_as(C.POINTER(C.c_uint64))))
        return a

    def to_onr(self, words) -> np.ndarray:
        a = np.ascontiguousarray(words, dtype=np.uint64)
        o = np.zeros(self.M, dtype=np.int32)
        check(lib().gwz_model_to_onr(self._h, a.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     o.ctypes.data_as(C.POINTER(C.c_int32))))
        return o

    # ---- per-address probe (kinetic_sample!, qmc.jl:14-44) ----
    def probe(self, addrs, params, u01=None, kernel=_capi.KERNEL_BITPARALLEL, want_rates=False) -> dict:
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(-1, self.words)
        n = addrs.shape[0]
        p = np.ascontiguousarray(np.atleast_1d(params), dtype=np.float64)
        if p.shape != (_capi.NUM_PARAMS,):
            raise ValueError("GutzwillerAnsatz takes exactly one parameter")
        u = None if u01 is None else np.ascontiguousarray(u01, dtype=np.float64)
        if u is not None and u.shape != (n,):
            raise ValueError("u01 must have one entry per address")
        out = dict(e_loc=np.zeros(n), residence_time=np.zeros(n), grad_ratio=np.zeros((n, 1)),
                   n_hops=np.zeros(n, dtype=np.int32), chosen=np.zeros(n, dtype=np.int32),
                   next_addr=np.zeros((n, self.words), dtype=np.uint64))
        rates = np.zeros((n, 2 * self.M)) if want_rates else None
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().gwz_probe_addresses(self._h, vp(addrs), n, p.ctypes.data_as(C.POINTER(C.c_double)), vp(u),
                                        int(kernel), vp(out["e_loc"]), vp(out["residence_time"]),
                                        vp(out["grad_ratio"]), vp(out["n_hops"]), vp(rates), vp(out["chosen"]),
                                        vp(out["next_addr"])))
        if want_rates:
            out["rates"] = rates
        return out

    def ansatz_val_and_grad(self, addrs, params):
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(-1, self.words)
        n = addrs.shape[0]
        p = np.ascontiguousarray(np.atleast_1d(params), dtype=np.float64)
        val, grad = np.zeros(n), np.zeros((n, 1))
        check(lib().gwz_ansatz_val_and_grad(self._h, addrs.ctypes.data_as(C.c_void_p), n,
                                            p.ctypes.data_as(C.POINTER(C.c_double)),
                                            val.ctypes.data_as(C.c_void_p), grad.ctypes.data_as(C.c_void_p)))
        return val, grad
