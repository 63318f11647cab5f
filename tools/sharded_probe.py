"""Where the time of ONE proof across the ranks goes (torchrun, 2+ ranks): wall clock per round and the library's
per-category kernel timers on rank 0.  python -m torch.distributed.run --nproc-per-node G tools/sharded_probe.py [log_n]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import plonkathon_b200 as pb
from plonkathon_b200 import _lib, parallel, synthetic as syn
from plonkathon_b200.transcript import Transcript
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
L = _lib.lib(); ctx = _lib.default_context()
setup = pb.Setup.generate(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF, n, ctx=ctx)
c = syn.build_circuit(log_n, seed=20260924, n_public=2)
pk, A, B, C, public = syn.circuit_arrays(c)
mode = os.environ.get("PROBE_MODE", "")
if "pinned" in mode:
    A, B, C = (torch.from_numpy(x).pin_memory().numpy() for x in (A, B, C))
if "extra" in mode:  # what bench.py has alive when it times the sharded proof: single-GPU provers on two contexts
    extra = [pb.Prover.from_arrays(setup, n, pk), pb.Prover.from_arrays(setup, n, pk, ctx=_lib.Context(local))]
    for e in extra:
        e.prove_arrays(A, B, C, public)
if "vk" in mode and rank == 0:  # what bench.py's verification does on rank 0 before the sharded section
    vk = setup.verification_key_arrays(n, pk)
    if "verify" in mode:
        single = pb.Prover.from_arrays(setup, n, pk).prove_arrays(A, B, C, public)
        print("verified:", vk.verify_proof(n, pb.Proof.from_bytes(single), [int(x) for x in public]), flush=True)
sp = parallel.ShardedProver.from_arrays(setup, n, pk)
ref = sp.prove_arrays(A, B, C, public)
if "onecall" in mode:
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        sp.prove_arrays(A, B, C, public)
        torch.cuda.synchronize()
        if rank == 0:
            print("proof %d (one call): %.2f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
for it in range(3):
    _lib.check(L.pb200_ctx_timing(ctx.handle, 1))
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); marks = []
    tr = Transcript(b"plonk")
    m1 = sp.round_1_arrays(A, B, C, public); marks.append(time.perf_counter())
    sp.beta, sp.gamma = tr.round_1(m1)
    m2 = sp.round_2(); marks.append(time.perf_counter())
    sp.alpha, sp.fft_cofactor = tr.round_2(m2)
    m3 = sp.round_3(); marks.append(time.perf_counter())
    sp.zeta = tr.round_3(m3)
    m4 = sp.round_4(); marks.append(time.perf_counter())
    sp.v = tr.round_4(m4)
    m5 = sp.round_5(); marks.append(time.perf_counter())
    cats = []
    for cat in range(4):
        tot, cnt = ctypes.c_double(), ctypes.c_uint64()
        _lib.check(L.pb200_ctx_timing_read(ctx.handle, cat, ctypes.byref(tot), ctypes.byref(cnt)))
        cats.append("%.2f/%d" % (tot.value, cnt.value))
    _lib.check(L.pb200_ctx_timing(ctx.handle, 0))
    if rank == 0:
        rounds = [marks[0] - t0] + [marks[i] - marks[i - 1] for i in range(1, 5)]
        print("proof %d: %.2f ms; rounds %s; kernel timers acc/ntt/sort/reduce (ms/launches): %s"
              % (it, (marks[-1] - t0) * 1e3, " ".join("%.2f" % (r * 1e3) for r in rounds), " ".join(cats)), flush=True)
dist.barrier()
dist.destroy_process_group()
