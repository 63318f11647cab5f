"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU MSM join (plonkathon_b200/parallel.py):
point-range sharding, one allgather of 128-byte XYZZ partial sums, local combination.  The per-rank partial
MSMs are produced by the oracle here (no GPU in this container); on a GPU box the same code path runs with
NCCL and the CUDA MSM (tests/test_gpu_multi.py)."""
import os
import random
import socket

import pytest
import torch.multiprocessing as mp

from oracle import plonk_oracle as O

R256 = 1 << 256


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _xyzz_bytes(pt):
    """affine oracle point (or None) -> 128-byte XYZZ in Montgomery limbs, as the library stores partials"""
    if pt is None:
        return bytes(128)
    m = lambda v: (v * R256 % O.Q_MOD).to_bytes(32, "little")  # noqa: E731
    return m(pt[0]) + m(pt[1]) + m(1) + m(1)


def _worker(rank, world, port, n, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plonkathon_b200 import parallel
    rng = random.Random(seed)
    pts = [O.g1_multiply(O.G1, rng.randrange(1, O.R_MOD)) for _ in range(n)]
    sc = [rng.randrange(O.R_MOD) for _ in range(n)]
    if seed == 2:  # make the total cancel to the identity: second half negates the first
        h = n // 2
        pts = pts[:h] + pts[:h]
        sc = sc[:h] + [(-s) % O.R_MOD for s in sc[:h]]
    first, count = parallel.shard_range(n, rank, world)
    part = O.ec_lincomb_naive(list(zip(pts[first:first + count], sc[first:first + count]))) if count else None
    gathered = parallel.allgather_bytes(_xyzz_bytes(part))
    xy, ident = parallel.combine_partials(b"".join(gathered), world)
    got = None if ident else (int.from_bytes(xy[:32], "little"), int.from_bytes(xy[32:], "little"))
    q.put((rank, got == O.ec_lincomb_naive(list(zip(pts, sc))), ident))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,seed", [(9, 1), (8, 2)])
def test_sharded_msm_join_world2(n, seed):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    if seed == 2:
        assert all(ident for _, _, ident in res)


def test_shard_range_partitions():
    from plonkathon_b200.parallel import shard_range
    for n in (1, 7, 8, 1 << 20):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                f, c = shard_range(n, r, world)
                assert f == pos
                pos += c
            assert pos == n


def _ntt_cpu_worker(rank, world, port, log_n, q):
    """the slab-NTT decomposition of parallel.slab_ntt with the oracle standing in for the CUDA kernels"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plonkathon_b200 import parallel
    n = 1 << log_n
    rng = random.Random(3)
    x = [rng.randrange(O.R_MOD) for _ in range(n)]
    log_m, log_g = parallel.slab_ntt_plan(log_n, world)
    M = 1 << log_m
    local = O.fft(x[rank::world])                      # step 1: local transform of the decimated sequence
    buf = b"".join(v.to_bytes(32, "little") for v in local)
    gathered = parallel.allgather_bytes(buf)           # step 2: the one allgather
    Y = [[int.from_bytes(g[32 * t:32 * t + 32], "little") for t in range(M)] for g in gathered]
    w = O.root_of_unity(n)
    slab = []
    for t in range(M):                                 # step 3: length-G DFT per element of the slab
        k = rank * M + t
        slab.append(sum(pow(w, h * k, O.R_MOD) * Y[h][t] for h in range(world)) % O.R_MOD)
    q.put((rank, slab == O.fft(x)[rank * M:(rank + 1) * M]))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_ntt_decomposition_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ntt_cpu_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
