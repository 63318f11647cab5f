import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import plonkathon_b200 as pb
from plonkathon_b200 import _lib, parallel, synthetic as syn
from plonkathon_b200.transcript import Transcript, Message1
log_n = 20; n = 1 << log_n
setup = pb.Setup.generate(0x1234567890ABCDEF, n)
c = syn.build_circuit(log_n, seed=1, n_public=2)
pk, A, B, C, public = syn.circuit_arrays(c)
sp = parallel.ShardedProver.from_arrays(setup, n, pk)
sp.prove_arrays(A, B, C, public)
dist.barrier(); torch.cuda.synchronize()
# instrument
orig_ag = parallel.allgather_bytes
T = {"ag": 0.0, "comb": 0.0}
def ag(*a, **k):
    t0 = time.perf_counter(); r = orig_ag(*a, **k); T["ag"] += time.perf_counter() - t0; return r
parallel.allgather_bytes = ag
orig_cb = parallel.combine_partials
def cb(*a, **k):
    t0 = time.perf_counter(); r = orig_cb(*a, **k); T["comb"] += time.perf_counter() - t0; return r
parallel.combine_partials = cb
for rep in range(3):
    T["ag"] = T["comb"] = 0.0
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sp.prove_arrays(A, B, C, public)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("rank", rank, "total %.1f ms allgather %.1f ms combine %.2f ms" % (tot * 1e3, T["ag"] * 1e3, T["comb"] * 1e3), flush=True)
dist.barrier(); dist.destroy_process_group()
