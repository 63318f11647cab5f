#!/bin/bash
# A/B of the two bucket-accumulation paths (XYZZ segments vs batched-affine rounds): bench value / e2e /
# fixed-base MSM component / average accumulation time per MSM call for each setting, then (optionally) an ncu
# capture of the affine round kernels.  Usage: tools/ab_affine.sh [tests] [ncu] "A B F" ...
mkdir -p gpurun_out
if [ "$1" = tests ]; then shift
  PB200_MSM_AFFINE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/ab_affine_tests.log
fi
NCU=0; if [ "$1" = ncu ]; then NCU=1; shift; fi
for mode in "$@"; do
  set -- $mode
  export PB200_MSM_AFFINE=$1 PB200_MSM_AFFINE_B=$2 PB200_MSM_AFFINE_F=$3
  timeout 600 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | tail -1 > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('affine=$1 B=$2 F=$3', round(d['value'],3), round(d['ms_per_step'],3), round(d['e2e']['value'],3), round(d['components']['g1_msm_fixed_base_2^20']['ms'],3), round(d['roofline']['avg_launch_ms'],3))" | tee -a gpurun_out/ab_affine.log
done
if [ $NCU = 1 ]; then
  export PB200_MSM_AFFINE=1 PB200_MSM_AFFINE_B=16 PB200_MSM_AFFINE_F=32
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aff_ -c 9 -o gpurun_out/r1_aff -f python tools/one_msm.py 20 > gpurun_out/ncu_aff.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_aff_launches.csv python tools/one_msm.py 20 > /dev/null 2>&1
fi
