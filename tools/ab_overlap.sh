#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prove" 2>&1 | tail -2
for o in 0 1 0 1; do
  PB200_OVERLAP=$o python bench.py --no-cpu-baseline --steps 8 2>/dev/null | tail -1 > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('overlap=$o', round(d['value'],3), round(d['ms_per_step'],3), round(d['e2e']['value'],3))"
done
