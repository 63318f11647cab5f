#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_probe.py 20 2>&1 | grep "^proof" ; }
run A=1
run PB200_SHARD_CONTIGUOUS=1
run PB200_MSM_SEG=32 PB200_MSM_G=16
run PB200_OVERLAP=0
