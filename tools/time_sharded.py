import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
import plonkathon_b200 as pb
from plonkathon_b200 import _lib, parallel, synthetic as syn
log_n = 20; n = 1 << log_n
setup = pb.Setup.generate(0x1234567890ABCDEF, n)
c = syn.build_circuit(log_n, seed=1, n_public=2)
pk, A, B, C, public = syn.circuit_arrays(c)
sp = parallel.ShardedProver.from_arrays(setup, n, pk)
single = pb.Prover.from_arrays(setup, n, pk)
ref = single.prove_arrays(A, B, C, public)
assert sp.prove_arrays(A, B, C, public) == ref
for name, fn in (("single prove_arrays (pageable host)", lambda: single.prove_arrays(A, B, C, public)),
                 ("sharded world=1", lambda: sp.prove_arrays(A, B, C, public))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 3 * 1e3, "ms")
# per-round breakdown of the sharded path
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); sp.prove_arrays(A, B, C, public); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
