"""CPU baseline table: the oracle port of the reference's Python path -- recursive
poly.fft, the multisubset ec_lincomb, Prover.prove -- timed on ONE host core at several sizes, with the fitted
2^20 extrapolation.  Writes profiles/<tag>_cpu_baseline.md."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random
from oracle import plonk_oracle as O
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rng = random.Random(1)
lines = ["# %s -- CPU baseline: oracle port of the reference's Python path, 1 core\n" % tag,
         "Host: %d cores visible; single-threaded like the reference. Real py_ecc is not installable here; the "
         "port's modular inverse is CPython's C `pow(a,-1,m)` (faster than py_ecc's Python Euclid), so these "
         "numbers are optimistic for the reference.\n" % (os.cpu_count() or 0)]
lines.append("| op | size | seconds | per-unit |\n|---|---:|---:|---:|")
fft_pts = []
for lg in (8, 10, 12, 14):
    v = [rng.randrange(O.R_MOD) for _ in range(1 << lg)]
    t0 = time.perf_counter(); O.fft(v); dt = time.perf_counter() - t0
    fft_pts.append((1 << lg, dt))
    lines.append("| poly.fft (poly.py:113-145) | 2^%d | %.4f | %.3g elems/s |" % (lg, dt, (1 << lg) / dt))
msm_pts = []
base = [O.g1_multiply(O.G1, rng.randrange(1, O.R_MOD)) for _ in range(64)]
for lg in (6, 8, 10):
    n = 1 << lg
    pts = [base[i % 64] for i in range(n)]
    sc = [rng.randrange(O.R_MOD) for _ in range(n)]
    t0 = time.perf_counter(); O.ec_lincomb(list(zip(pts, sc))); dt = time.perf_counter() - t0
    msm_pts.append((n, dt))
    lines.append("| ec_lincomb (curve.py:38-111) | 2^%d | %.3f | %.3g pts/s |" % (lg, dt, n / dt))
prove_pts = []
for lg in (4, 6, 8):
    n, times = bench.cpu_sample(lg, 1)
    prove_pts.append((n, times[0]))
    lines.append("| Prover.prove (prover.py:51-306) | 2^%d gates | %.2f | %.3g proofs/s |" % (lg, times[0], 1 / times[0]))
a = sum(dt / (n * math.log2(n)) for n, dt in fft_pts) / len(fft_pts)
b = sum(dt / n for n, dt in msm_pts) / len(msm_pts)
N = 1 << 20
lines.append("\nFits: t_fft = %.3g * n log2 n, t_msm = %.3g * n.  Extrapolated to 2^20: fft %.1f s, ec_lincomb %.0f s, "
             "prove (linear in gates from 2^8) %.0f s = %.2e proofs/s.\n" % (a, b, a * N * 20, b * N, prove_pts[-1][1] * N / prove_pts[-1][0],
                                                                         prove_pts[-1][0] / (prove_pts[-1][1] * N)))
os.makedirs("profiles", exist_ok=True)
open("profiles/%s_cpu_baseline.md" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
