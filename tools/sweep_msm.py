"""sweep MSM tunables (segment length, window bits) on the fixed-base 2^20 commit, device-timed"""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
res = {}
for cfix in (os.environ.get("SWEEP_C", "19,20,21").split(",")):
    os.environ["PB200_MSM_C_FIXED"] = cfix
    import importlib
    import plonkathon_b200 as pb
    from plonkathon_b200 import _lib
    L = _lib.lib(); ctx = _lib.default_context()
    n = 1 << 20
    setup = pb.Setup.generate(0x1234567890ABCDEF, n)
    x = [torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda") for _ in range(3)]
    for t in x: t[:, 7] &= 0x0FFFFFFF
    stream = torch.cuda.ExternalStream(ctx.stream)
    out = ctypes.create_string_buffer(64); ident = ctypes.c_int()
    for seg in (16, 32, 64, 128):
        os.environ["PB200_MSM_SEG"] = str(seg)
        f = lambda: _lib.check(L.pb200_srs_commit_coeffs(ctx.handle, setup._srs, ctypes.c_void_p(x[0].data_ptr()), n, 0, out, ctypes.byref(ident)))
        f(); ctx.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): f()
        e1.record(stream); e1.synchronize()
        res["c%s_seg%d" % (cfix, seg)] = e0.elapsed_time(e1) / 5
    del setup
print(json.dumps(res, indent=1))
