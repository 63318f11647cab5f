#!/bin/bash
# A/B of alternative library builds placed in tools/variants/*.so (PB200_LIB override): bench value + e2e
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset PB200_LIB; else export PB200_LIB=$PWD/tools/variants/$lib; fi
  python bench.py --no-cpu-baseline --steps 8 2>/dev/null | tail -1 > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('$lib', round(d['value'],3), round(d['ms_per_step'],3), round(d['e2e']['value'],3), round(d['components']['g1_msm_fixed_base_2^20']['ms'],3))"
done
