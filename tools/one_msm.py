import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plonkathon_b200 import _lib
from tests.golden_io import PTAU_HEAD
from oracle import plonk_oracle as O
L = _lib.lib(); ctx = _lib.default_context()
osetup = O.Setup.from_file(PTAU_HEAD)
base = np.frombuffer(b"".join(p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little") for p in osetup.powers_of_x), dtype=np.uint8).reshape(2048, 64)
for logn in [int(a) for a in sys.argv[1:]] or [20]:
    n = 1 << logn
    pts = torch.from_numpy(np.ascontiguousarray(np.tile(base, (n // 2048, 1)))).cuda()
    sc = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda"); sc[:, 7] &= 0x0FFFFFFF
    out = ctypes.create_string_buffer(64); ident = ctypes.c_int()
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        _lib.check(L.pb200_g1_msm(ctx.handle, pts.data_ptr(), sc.data_ptr(), n, out, ctypes.byref(ident)))
        print("msm 2^%d rep %d: %.2f ms" % (logn, rep, (time.time() - t0) * 1e3), flush=True)
