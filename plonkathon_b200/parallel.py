"""Multi-GPU: one process per GPU (torchrun), ``torch.distributed`` for the plumbing.

Two modes:

* replicas (bench.py's default at N > 1): proofs are independent units, every rank owns an SRS replica and
  proves its own instances -- no data-path collective.

* one proof across the GPUs of a box -- the MSM join of north_star: every commitment
  ``sum_i c_i [tau^i]G`` (setup.py:66-72) is sharded by **point range**; rank g computes the partial sum over
  powers [g*n/G, (g+1)*n/G) on its GPU, the 128-byte XYZZ partials are exchanged with ONE allgather per round
  (NCCL over NVLink on GPUs; gloo in the CPU tests), and every rank adds the G partials and converts to affine
  (NCCL has no elliptic-curve reduction, so the "reduce" is allgather + local add).  All ranks therefore see
  the same commitments, feed the same transcript and stay in lock step.  The transforms and element-wise
  kernels are replicated on every rank in this mode (slab-sharded NTT is not implemented yet)."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .prover import Prover, _pts


def shard_range(n: int, rank: int, world: int):
    """Contiguous point range [first, first+count) of rank `rank` out of `world` (the first n % world ranks
    get one extra point)."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def combine_partials(parts: bytes, count: int):
    """Sum `count` XYZZ partial sums (128 bytes each, as produced by pb200_prover_read_partials) into one
    affine point; returns (x||y little-endian bytes, is_identity).  Host arithmetic inside the library."""
    out = ctypes.create_string_buffer(64)
    ident = ctypes.c_int(0)
    _lib.check(_lib.lib().pb200_g1_combine_partials_host(parts, count, out, ctypes.byref(ident)))
    return out.raw, bool(ident.value)


def allgather_bytes(local: bytes, group=None, device=None) -> list:
    """All ranks contribute `local` (same length everywhere); returns the list of every rank's bytes.
    Uses a CUDA tensor (NCCL) when `device` is given, a CPU tensor (gloo) otherwise."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(local), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy().tobytes() for o in outs]


class ShardedProver(Prover):
    """``Prover`` whose commitments are point-sharded across the ranks of a process group."""

    @classmethod
    def from_arrays(cls, setup, group_order, pk_arrays, group=None):
        self = super().from_arrays(setup, group_order, pk_arrays)
        self._init_shard(group)
        return self

    def __init__(self, setup, program, group=None):
        super().__init__(setup, program)
        self._init_shard(group)

    def _init_shard(self, group):
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        first, count = shard_range(self.group_order, self.rank, self.world)
        _lib.check(_lib.lib().pb200_prover_set_shard(self._h, first, count, 1))
        self._device = torch.device("cuda", self.ctx.device)

    def _commitments(self, first_slot: int, count: int, raw: bytes):
        buf = ctypes.create_string_buffer(128 * count)
        _lib.check(_lib.lib().pb200_prover_read_partials(self._h, first_slot, count, buf))
        gathered = allgather_bytes(buf.raw, self.group, self._device)  # the one collective of this round
        xy = b""
        for k in range(count):
            parts = b"".join(g[128 * k:128 * (k + 1)] for g in gathered)
            pt, ident = combine_partials(parts, self.world)
            if ident:
                raise _lib.PlonkB200Error("commitment is the point at infinity (unsupported by the reference transcript)")
            xy += pt
        _lib.check(_lib.lib().pb200_prover_set_points(self._h, first_slot, count, xy))
        return _pts(xy, count)

    def prove_arrays(self, A, B, C, public) -> bytes:
        """Round-by-round (the commitments need the collective between rounds); returns the 768-byte proof."""
        from .curve import Scalar
        from .prover import _as_le_rows
        from .transcript import Transcript
        n = self.group_order
        a, b, c = (_as_le_rows(v, n) for v in (A, B, C))
        pub = _as_le_rows(public, len(public)) if len(public) else np.zeros((0, 32), dtype=np.uint8)
        vp = ctypes.c_void_p
        L = _lib.lib()
        tr = Transcript(b"plonk")
        out = ctypes.create_string_buffer(192)
        _lib.check(L.pb200_prover_round1(self._h, a.ctypes.data_as(vp), b.ctypes.data_as(vp), c.ctypes.data_as(vp),
                                         pub.ctypes.data_as(vp), pub.shape[0], out))
        from .transcript import Message1, Message2, Message3, Message5
        self.beta, self.gamma = tr.round_1(Message1(*self._commitments(0, 3, out.raw)))
        msg2 = self.round_2()
        self.alpha, self.fft_cofactor = tr.round_2(msg2)
        msg3 = self.round_3()
        self.zeta = tr.round_3(msg3)
        msg4 = self.round_4()
        self.v = tr.round_4(msg4)
        self.round_5()
        proof = ctypes.create_string_buffer(768)
        _lib.check(L.pb200_prover_serialize(self._h, proof))
        return proof.raw


# ------------------------------------------------------------------------------------------------
# slab-sharded NTT (north_star: "NTT by coefficient-slab across the GPUs with a single allgather at the join")
# ------------------------------------------------------------------------------------------------
def slab_ntt_plan(log_n: int, world: int):
    """N = 2^log_n over `world` = 2^log_g ranks -> (log_m, log_g): rank h owns the decimated input x[h::world]
    and produces the contiguous output slab [h*M, (h+1)*M), M = N / world."""
    log_g = world.bit_length() - 1
    assert world == 1 << log_g and 1 <= log_g <= 3 and log_n > log_g, "slab NTT: 2, 4 or 8 ranks"
    return log_n - log_g, log_g


def slab_ntt(x_full, log_n: int, inverse: bool = False, group=None, ctx: Optional[_lib.Context] = None):
    """Distributed NTT of the length-2^log_n vector `x_full` (a CUDA uint8/int32 tensor of N*32 bytes holding
    canonical or Montgomery Fr elements, present on every rank; only the rank's decimated part is read).
    Returns this rank's contiguous output slab as a CUDA tensor [M, 32] uint8.  One NCCL allgather."""
    import torch
    import torch.distributed as dist
    ctx = ctx or _lib.default_context()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    log_m, log_g = slab_ntt_plan(log_n, world)
    M = 1 << log_m
    L = _lib.lib()
    vp = ctypes.c_void_p
    local = torch.empty((M, 32), dtype=torch.uint8, device=x_full.device)
    _lib.check(L.pb200_fr_ntt_decimated(ctx.handle, vp(x_full.data_ptr()), vp(local.data_ptr()), log_m,
                                        1 if inverse else 0, world, rank))
    ctx.sync()  # the library stream is not torch's: make the sub-spectrum visible to the collective
    sub = torch.empty((world, M, 32), dtype=torch.uint8, device=x_full.device)
    dist.all_gather_into_tensor(sub, local, group=group)  # the one exchange step
    torch.cuda.current_stream().synchronize()
    out = torch.empty((M, 32), dtype=torch.uint8, device=x_full.device)
    _lib.check(L.pb200_fr_ntt_slab_combine(ctx.handle, vp(sub.data_ptr()), vp(out.data_ptr()), log_m, log_g, rank,
                                           1 if inverse else 0))
    ctx.sync()
    return out
