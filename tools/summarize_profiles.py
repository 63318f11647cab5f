"""Turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/ (round tag argv[1])."""
import collections, csv, subprocess, sys, io, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(open("gpurun_out/r1_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
names = [r[ki] for r in rows[1:]]; vals = [float(r[vi].replace(',', '')) for r in rows[1:]]
gc = [i for i, n in enumerate(names) if 'k_gate_check' in n]
i0, i1 = gc[1], gc[2]
agg = {}; cnt = collections.Counter(); tot = 0
for n, v in zip(names[i0:i1], vals[i0:i1]):
    k = n.split('(')[0].replace('pb200::', '')
    agg[k] = agg.get(k, 0) + v; cnt[k] += 1; tot += v
with open("profiles/%s_launches_one_proof.md" % tag, "w") as f:
    f.write("# %s -- ncu launch list of one 2^20-gate proof\n\n" % tag)
    f.write("Command (under gpurun): `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "
            "gpurun_out/r1_launches.csv python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline`\n\n")
    f.write("One proof = the launches between two consecutive `k_gate_check` launches (%d kernels, sum of durations "
            "%.2f ms; per-launch times under ncu are cold-cache and serialised -- compare shares).\n\n" % (i1 - i0, tot / 1e6))
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, cnt[k], v / 1e6, 100 * v / tot))

# ---- full captures
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
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
for rep, title in (("r1_msm_acc", "k_msm_seg_accumulate (MSM bucket accumulation)"), ("r1_ntt", "k_ntt_pass (Fr NTT pass)")):
    out = subprocess.run(["ncu", "-i", "gpurun_out/%s.ncu-rep" % rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open("profiles/%s_%s.md" % (tag, rep.replace("r1_", "ncu_")), "w") as f:
        f.write("# %s -- ncu --set full: %s\n\n" % (tag, title))
        f.write("Command (under gpurun): `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> "
                "-o gpurun_out/%s python bench.py --steps 1 --warmup 1 --no-cpu-baseline`; read here with "
                "`ncu -i ... --page raw --csv`.\n\n" % rep)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write("## launch: `%s`\n\n| metric | unit | value |\n|---|---|---:|\n" % name.split("(")[0])
            for h, u, v in zip(hdr, units, r):
                if h in WANT:
                    f.write("| %s | %s | %s |\n" % (h, u, v))
            f.write("\n")
print(open("profiles/%s_launches_one_proof.md" % tag).read()[:1500])
print(open("profiles/%s_ncu_msm_acc.md" % tag).read())
