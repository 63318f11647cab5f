#!/bin/bash
# A/B of the two bucket-accumulation paths (XYZZ segments vs batched-affine rounds): parity suite under the affine
# path, then bench value / e2e / fixed-base MSM component for each setting.
mkdir -p gpurun_out
PB200_MSM_AFFINE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/ab_affine_tests.log
for mode in "0 32 16" "1 32 16" "1 16 16" "1 64 16" "1 32 32" "1 64 8"; do
  set -- $mode
  export PB200_MSM_AFFINE=$1 PB200_MSM_AFFINE_B=$2 PB200_MSM_AFFINE_F=$3
  timeout 600 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | tail -1 > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('affine=$1 B=$2 F=$3', round(d['value'],3), round(d['ms_per_step'],3), round(d['e2e']['value'],3), round(d['components']['g1_msm_fixed_base_2^20']['ms'],3), round(d['roofline']['avg_launch_ms'],3))" | tee -a gpurun_out/ab_affine.log
done
