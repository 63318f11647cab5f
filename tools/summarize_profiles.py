"""Turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py r02 [launches.csv] [capture.ncu-rep]

  * gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum of bench.py --inflight 1 --steps 1 --warmup 1)
      -> profiles/<tag>_launches_one_proof.md : one proof = the launches between two consecutive k_gate_check launches
  * gpurun_out/<tag>_full.ncu-rep (ncu --set full of the kernels of that proof)
      -> profiles/<tag>_ncu_kernels.md : the metrics B200_PROFILING.md asks for, one table per distinct kernel
         (the longest launch of each), and profiles/<tag>_dominant_kernel.json for bench.py's roofline.traffic."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/%s_launches.csv" % tag
capture = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/%s_full_raw.csv" % tag  # ncu -i <rep> --page raw --csv
os.makedirs("profiles", exist_ok=True)

# ---- launch list
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[ki] for r in rows[1:]]
    vals = [float(r[vi].replace(",", "")) for r in rows[1:]]
    gc = [i for i, n in enumerate(names) if "k_gate_check" in n]
    i0, i1 = gc[1], (gc[2] if len(gc) > 2 else len(names))
    agg, cnt, tot = {}, collections.Counter(), 0
    for n, v in zip(names[i0:i1], vals[i0:i1]):
        k = n.split("(")[0].replace("pb200::", "")
        agg[k] = agg.get(k, 0) + v
        cnt[k] += 1
        tot += v
    with open("profiles/%s_launches_one_proof.md" % tag, "w") as f:
        f.write("# %s -- ncu launch list of one 2^20-gate proof\n\n" % tag)
        f.write("Command (under gpurun): `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "
                "%s python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --no-verify`\n\n" % launches)
        f.write("One proof = the launches between two consecutive `k_gate_check` launches (%d kernels, sum of durations "
                "%.2f ms; per-launch times under ncu are cold-cache and serialised -- compare shares).\n\n" % (i1 - i0, tot / 1e6))
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, cnt[k], v / 1e6, 100 * v / tot))
    print(open("profiles/%s_launches_one_proof.md" % tag).read())

# ---- full capture
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
if os.path.exists(capture):
    if capture.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", capture, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        out = open(capture).read()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    best = {}
    for r in rows[2:]:
        name = r[kn].split("(")[0]
        t = float(r[ti].replace(",", ""))
        if name not in best or t > best[name][0]:
            best[name] = (t, r)
    with open("profiles/%s_ncu_kernels.md" % tag, "w") as f:
        f.write("# %s -- ncu --set full: the kernels of one proof (longest launch of each)\n\n" % tag)
        f.write("Command (under gpurun): `ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <skip> "
                "-c <n> -o %s python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --no-verify`; read here "
                "with `ncu -i ... --page raw --csv`.\n\n" % capture)
        for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
            f.write("## `%s`\n\n| metric | unit | value |\n|---|---|---:|\n" % name)
            for h, u, v in zip(hdr, units, r):
                if h in WANT:
                    f.write("| %s | %s | %s |\n" % (h, u, v))
            f.write("\n")
    # the dominant kernel's DRAM traffic per point for bench.py (the largest captured k_msm_seg_accumulate launch: a
    # batch of 1-3 commitments of 2^20 points; the point count follows from its grid)
    if "k_msm_seg_accumulate" in best:
        r = best["k_msm_seg_accumulate"][1]

        def val(metric):
            i = hdr.index(metric)
            x, u = float(r[i].replace(",", "")), units[i]
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        grid = float(r[hdr.index("launch__grid_size")].replace(",", ""))
        points = grid * 128 * 32 / 13.0  # 128 threads per block, 32 sorted entries per thread, 13 windows per point
        json.dump({"kernel": "k_msm_seg_accumulate", "log_n": 20, "points_in_captured_launch": points,
                   "grid_size": grid, "dram_bytes_in_captured_launch": dram, "dram_bytes_per_point": dram / points,
                   "source": capture}, open("profiles/%s_dominant_kernel.json" % tag, "w"), indent=1)
    print(open("profiles/%s_ncu_kernels.md" % tag).read()[:3000])
