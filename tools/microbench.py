"""Quick device-side timings (not the contract bench): modmul throughput, NTT and MSM at a few sizes."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from plonkathon_b200 import _lib

L = _lib.lib()
ctx = _lib.default_context()
res = {}
for field, name in ((0, "Fr"), (1, "Fq")):
    for threads in (148 * 1024, 148 * 2048):
        ms = ctypes.c_float()
        iters = 4096
        _lib.check(L.pb200_bench_modmul(ctx.handle, field, threads, iters, ctypes.byref(ms)))
        res["modmul_%s_t%d" % (name, threads)] = {"ms": ms.value, "Gmul_per_s": threads * iters / ms.value / 1e6}
stream = torch.cuda.ExternalStream(ctx.stream)
def time_it(fn, reps=5):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for logn in (16, 18, 20, 22, 24):
    n = 1 << logn
    x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
    x[:, 7] &= 0x0FFFFFFF
    y = torch.empty_like(x)
    f = lambda: _lib.check(L.pb200_fr_ntt(ctx.handle, x.data_ptr(), y.data_ptr(), logn, 0))
    ms = time_it(f)
    res["ntt_2^%d" % logn] = {"ms": ms, "Gelem_per_s": n / ms / 1e6, "hbm_frac": 64 * n / (ms * 1e-3) / 6569.6e9}
# MSM generic on tiled SRS points
from tests.golden_io import PTAU_HEAD
from oracle import plonk_oracle as O
osetup = O.Setup.from_file(PTAU_HEAD)
base = np.frombuffer(b"".join(p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little") for p in osetup.powers_of_x), dtype=np.uint8).reshape(2048, 64)
for logn in (16, 18, 20):
    n = 1 << logn
    pts = torch.from_numpy(np.ascontiguousarray(np.tile(base, (n // 2048, 1)))).cuda()
    sc = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda"); sc[:, 7] &= 0x0FFFFFFF
    out = ctypes.create_string_buffer(64); ident = ctypes.c_int()
    for c in ([None] if logn < 20 else [None, "14", "15", "16"]):
        if c: os.environ["PB200_MSM_C"] = c
        f = lambda: _lib.check(L.pb200_g1_msm(ctx.handle, pts.data_ptr(), sc.data_ptr(), n, out, ctypes.byref(ident)))
        t0 = time.time(); f(); t1 = time.time(); f(); t2 = time.time()
        res["msm_generic_2^%d_c%s" % (logn, c)] = {"ms_first": (t1 - t0) * 1e3, "ms": (t2 - t1) * 1e3, "Mpts_per_s": n / (t2 - t1) / 1e6}
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
