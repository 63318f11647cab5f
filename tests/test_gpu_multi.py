"""GPU, 2 ranks (skipped on a single-GPU box): one proof across two GPUs with point-sharded commitments and
an NCCL allgather at every MSM join gives the same 768 bytes as the single-GPU prover."""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log_n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import plonkathon_b200 as pb
    from plonkathon_b200 import parallel, synthetic as syn
    n = 1 << log_n
    tau = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    c = syn.build_circuit(log_n, seed=11, n_public=2)
    pk, A, B, C, public = syn.circuit_arrays(c)
    setup = pb.Setup.generate(tau, n)
    single = pb.Prover.from_arrays(setup, n, pk).prove_arrays(A, B, C, public)
    sharded = parallel.ShardedProver.from_arrays(setup, n, pk).prove_arrays(A, B, C, public)
    q.put((rank, single == sharded))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("log_n", [10, 16])
def test_sharded_proof_equals_single_gpu(log_n):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, log_n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _ntt_worker(rank, world, port, log_n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import ctypes
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from plonkathon_b200 import _lib, parallel
    ctx = _lib.default_context()
    n = 1 << log_n
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randint(0, 2 ** 31 - 1, (n, 8), dtype=torch.int32, generator=g)
    x[:, 7] &= 0x0FFFFFFF
    x = x.cuda()
    ok = True
    for inverse in (0, 1):
        full = torch.empty_like(x)
        _lib.check(_lib.lib().pb200_fr_ntt(ctx.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(full.data_ptr()),
                                           log_n, inverse))
        ctx.sync()
        slab = parallel.slab_ntt(x.view(torch.uint8).reshape(n, 32), log_n, bool(inverse))
        m = n // world
        ref = full.view(torch.uint8).reshape(n, 32)[rank * m:(rank + 1) * m]
        ok = ok and bool(torch.equal(slab, ref))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("log_n", [12, 22])
def test_slab_ntt_equals_single_gpu(log_n):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ntt_worker, args=(r, 2, port, log_n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
