"""Three fixed-base KZG commitments of 2^k random coefficients (the prover's MSM shape) for ncu captures."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plonkathon_b200 as pb
from plonkathon_b200 import _lib
L = _lib.lib(); ctx = _lib.default_context()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
setup = pb.Setup.generate(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF, n, ctx=ctx)
x = torch.randint(0, 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda"); x[:, 7] &= 0x0FFFFFFF
out = ctypes.create_string_buffer(64); ident = ctypes.c_int()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(L.pb200_srs_commit_coeffs(ctx.handle, setup._srs, ctypes.c_void_p(x.data_ptr()), n, 0, out, ctypes.byref(ident)))
    print("commit 2^%d rep %d: %.2f ms" % (logn, rep, (time.time() - t0) * 1e3), flush=True)
