"""GPU, 2 / 4 / 8 ranks (skipped where the box has fewer GPUs): ONE proof across the GPUs -- coset slices, slab-sharded
interpolation, bucket-sharded commitments, the library's own NCCL allgathers at the joins (csrc/prover.cu with
world > 1) -- returns the same 768 bytes as the single-GPU prover; the sharded NTT and the sharded commitment equal
their single-GPU operators.  bench.py repeats the same checks at every N > 1 for the driver's scaling run."""
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


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return torch, dist


def _proof_worker(rank, world, port, log_n, n_public, q):
    torch, dist = _init(rank, world, port)
    import plonkathon_b200 as pb
    from plonkathon_b200 import parallel, synthetic as syn
    n = 1 << log_n
    tau = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    c = syn.build_circuit(log_n, seed=11, n_public=n_public)
    pk, A, B, C, public = syn.circuit_arrays(c)
    setup = pb.Setup.generate(tau, n)
    single = pb.Prover.from_arrays(setup, n, pk).prove_arrays(A, B, C, public)
    sp = parallel.ShardedProver.from_arrays(setup, n, pk)
    sharded = sp.prove_arrays(A, B, C, public)
    again = sp.prove_arrays(A, B, C, public)  # the prover object is reusable
    # a witness that breaks a gate fails on every rank alike (no rank is left waiting in a collective)
    bad = A.copy()
    bad[0, 0] ^= 1
    try:
        sp.prove_arrays(bad, B, C, public)
        rejected = False
    except AssertionError:
        rejected = True
    q.put((rank, single == sharded == again, rejected, parallel.comm_info(setup.ctx)))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(target, world, *args):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return res


@pytest.mark.parametrize("world,log_n,n_public", [(2, 10, 2), (2, 16, 2), (4, 12, 3), (8, 12, 2), (8, 16, 12), (2, 12, 12)])
def test_sharded_proof_equals_single_gpu(world, log_n, n_public):
    res = _spawn(_proof_worker, world, log_n, n_public)
    assert all(ok for _, ok, _, _ in res), res
    assert all(rej for _, _, rej, _ in res), res
    # per proof: 2 interpolation joins + the quotient join + 4 commitment joins (one more when PI is interpolated)
    assert all(info[1] == world and info[2] >= 2 * 7 for _, _, _, info in res), res


def _op_worker(rank, world, port, log_n, q):
    torch, dist = _init(rank, world, port)
    import ctypes
    import plonkathon_b200 as pb
    from plonkathon_b200 import _lib, parallel
    ctx = parallel.init_comm(_lib.default_context())
    n = 1 << log_n
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randint(0, 2 ** 31 - 1, (n, 8), dtype=torch.int32, generator=g)
    x[:, 7] &= 0x0FFFFFFF
    x = x.cuda()
    ok = True
    vp = ctypes.c_void_p
    for inverse in (0, 1):
        full = torch.empty_like(x)
        _lib.check(_lib.lib().pb200_fr_ntt(ctx.handle, vp(x.data_ptr()), vp(full.data_ptr()), log_n, inverse))
        ctx.sync()
        got = parallel.sharded_ntt(x.view(torch.uint8).reshape(n, 32), log_n, bool(inverse), ctx=ctx)
        ctx.sync()
        ok = ok and bool(torch.equal(got, full.view(torch.uint8).reshape(n, 32)))
    # sharded commitment == single-GPU commitment
    setup = pb.Setup.generate(0x1234567890ABCDEF, n, ctx=ctx)
    out = ctypes.create_string_buffer(64)
    ident = ctypes.c_int()
    _lib.check(_lib.lib().pb200_srs_commit_coeffs(ctx.handle, setup._srs, vp(x.data_ptr()), n, 0, out, ctypes.byref(ident)))
    ref = (int.from_bytes(out.raw[:32], "little"), int.from_bytes(out.raw[32:], "little"))
    ok = ok and parallel.sharded_commit(setup, x, n) == ref
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 12), (2, 22), (4, 14), (8, 14), (8, 22)])
def test_sharded_operators_equal_single_gpu(world, log_n):
    res = _spawn(_op_worker, world, log_n)
    assert all(ok for _, ok in res), res


def _golden_worker(rank, world, port, q):
    torch, dist = _init(rank, world, port)
    import json
    import plonkathon_b200 as pb
    from plonkathon_b200 import parallel, synthetic as syn
    from tests.golden_io import GOLDEN
    log_n = 22
    n = 1 << log_n
    c = syn.build_circuit(log_n, seed=7, n_public=2)
    pk, A, B, C, public = syn.circuit_arrays(c)
    setup = pb.Setup.generate(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF, n)
    raw = parallel.ShardedProver.from_arrays(setup, n, pk).prove_arrays(A, B, C, public)
    want = json.load(open(os.path.join(GOLDEN, "proof_2p22.json")))["proof_hex"]
    q.put((rank, raw.hex() == want))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [8, 4])
def test_sharded_2p22_gates_against_golden(world):
    """BASELINE.json configs[4]: a synthetic 2^22-gate circuit proved ONCE across the GPUs of the box (bucket-sharded
    MSM, slab-sharded NTT, NVLink allgathers), byte for byte against the oracle's golden proof
    (tests/golden/proof_2p22.json: the oracle prover over the C restatement, 2 CPU-hours)."""
    import torch
    if torch.cuda.device_count() != world:
        pytest.skip("runs on a box with exactly %d GPUs" % world)
    res = _spawn(_golden_worker, world)
    assert all(ok for _, ok in res), res
