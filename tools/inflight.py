"""Throughput with K proofs in flight on ONE GPU: K contexts (stream + scratch) and K provers sharing one SRS, each
driven by its own host thread (the C ABI releases the GIL).  Usage: python tools/inflight.py [K ...]"""
import ctypes, os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plonkathon_b200 as pb
from plonkathon_b200 import _lib, synthetic as syn
L = _lib.lib()
log_n = 20; n = 1 << log_n
ctx0 = _lib.Context(0)
setup = pb.Setup.generate(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF, n, ctx=ctx0)
circ = syn.build_circuit(log_n, seed=20260924, n_public=2)
pk, A, B, C, public = syn.circuit_arrays(circ)
pub = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in public), dtype=np.uint8).reshape(-1, 32).copy()
vp = ctypes.c_void_p
KMAX = max([int(a) for a in sys.argv[1:]] or [2])
lanes = []
for k in range(KMAX):
    ctx = ctx0 if k == 0 else _lib.Context(0)
    prover = pb.Prover.from_arrays(setup, n, pk, ctx=ctx)
    hA, hB, hC = (torch.from_numpy(x.copy()).pin_memory() for x in (A, B, C))
    dA, dB, dC = (x.cuda() for x in (hA, hB, hC))
    lanes.append(dict(ctx=ctx, prover=prover, h=(hA, hB, hC), d=(dA, dB, dC), proof=ctypes.create_string_buffer(768)))
def dev(l):
    _lib.check(L.pb200_prover_prove_device(l["prover"]._h, *[vp(t.data_ptr()) for t in l["d"]], pub.ctypes.data_as(vp), pub.shape[0], l["proof"]))
def host(l):
    _lib.check(L.pb200_prover_prove(l["prover"]._h, *[vp(t.data_ptr()) for t in l["h"]], pub.ctypes.data_as(vp), pub.shape[0], l["proof"]))
for l in lanes:
    dev(l); dev(l); host(l)
ref = lanes[0]["proof"].raw
assert all(l["proof"].raw == ref for l in lanes)
for K in [int(a) for a in sys.argv[1:]] or [1, 2]:
    pool = ThreadPoolExecutor(K)
    for name, fn in (("device", dev), ("host", host)):
        steps = 6
        def worker(l):
            for _ in range(steps): fn(l)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter()
        list(pool.map(worker, lanes[:K]))
        torch.cuda.synchronize(); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        print("K=%d %-6s %d proofs in %.1f ms (wall %.1f) -> %.2f proofs/s, %.2f ms/proof" % (K, name, K * steps, ms, (time.perf_counter() - t0) * 1e3, K * steps / ms * 1e3, ms / (K * steps)), flush=True)
    assert all(l["proof"].raw == ref for l in lanes[:K])
    pool.shutdown()
