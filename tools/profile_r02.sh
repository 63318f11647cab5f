#!/bin/bash
# Round-2 profile captures on one GPU (under gpurun): launch lists of one proof and of one rank's share of a sharded
# commitment, and an ncu --set full capture of the kernels of a proof.  tools/summarize_profiles.py turns them into
# profiles/r02_*.
mkdir -p gpurun_out
for a in "20 8" "22 8"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_shard_$(echo $a | tr ' ' '_').csv python tools/one_shard.py $a > gpurun_out/r02_shard_$(echo $a | tr ' ' '_').log 2>&1
  tail -1 gpurun_out/r02_shard_$(echo $a | tr ' ' '_').log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_msm_seg_accumulate|k_ntt_pass|k_reduce_level0|k_reduce_block|k_msm_scatter|k_msm_histogram|k_quotient" -s 70 -c 16 -o /tmp/r02_full -f python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null   # the .ncu-rep itself (77 MB) stays on the box
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
grep "per proof" gpurun_out/r02_bench_1gpu.err | tail -1
du -sh gpurun_out
