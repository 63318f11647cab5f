"""What ONE rank of a G-rank bucket-sharded commitment executes (pb200_srs_commit_partial with 1/G of the bucket
range), on a single GPU, for ncu launch lists: python tools/one_shard.py <log_n> <G>."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plonkathon_b200 as pb
from plonkathon_b200 import _lib
L = _lib.lib(); ctx = _lib.default_context()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = 1 << logn
setup = pb.Setup.generate(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF, n, ctx=ctx)
nb = ctypes.c_uint()
_lib.check(L.pb200_srs_bucket_count(setup._srs, ctypes.byref(nb)))
per = nb.value // G
x = torch.randint(0, 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda"); x[:, 7] &= 0x0FFFFFFF
out = ctypes.create_string_buffer(128)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(L.pb200_srs_commit_partial(ctx.handle, setup._srs, ctypes.c_void_p(x.data_ptr()), 0, n, 3 * per, 4 * per, 0, out))
    print("shard 3/%d of a 2^%d commitment, rep %d: %.2f ms" % (G, logn, rep, (time.time() - t0) * 1e3), flush=True)
