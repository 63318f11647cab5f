#!/bin/bash
# Round-2 A/B of the bucket accumulation (batched affine rounds, default, vs PB200_MSM_ACC=xyzz) and of the chain
# length B; then launch list and ncu --set full captures of the new kernels.
# Usage (under gpurun): tools/ab_msm.sh [tests] [ncu] "acc B" ...
mkdir -p gpurun_out
if [ "$1" = tests ]; then shift
  PYTHONFAULTHANDLER=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1
  tail -4 gpurun_out/r2_tests.log
fi
NCU=0; if [ "$1" = ncu ]; then NCU=1; shift; fi
for mode in "$@"; do
  set -- $mode
  export PB200_MSM_ACC=$1 PB200_MSM_B=$2
  timeout 600 python bench.py --no-cpu-baseline --no-verify --steps 5 2>/tmp/b.err | tail -1 > /tmp/b.json
  grep "per proof" /tmp/b.err | tail -1 | tee -a gpurun_out/r2_ab_msm.log
  python -c "import json; d=json.load(open('/tmp/b.json')); print('acc=$1 B=$2 value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), 'msm2^20 ms', round(d['components']['g1_msm_fixed_base_2^20']['ms'],3), 'ntt pair ms', round(d['components']['fr_ntt_fwd_plus_inv_2^20']['ms'],3), 'golden', d['proof_matches_oracle_golden'])" 2>&1 | tail -1 | tee -a gpurun_out/r2_ab_msm.log
  [ -s /tmp/b.json ] || tail -5 /tmp/b.err | tee -a gpurun_out/r2_ab_msm.log
done
unset PB200_MSM_ACC PB200_MSM_B
if [ $NCU = 1 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r2_bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_aff_round|k_reduce_" -s 16 -c 10 -o gpurun_out/r2_msm -f python tools/one_commit.py 20 > gpurun_out/r2_ncu_msm.log 2>&1
fi
