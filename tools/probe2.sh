#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_probe.py 20 2>&1 | grep "^proof\|^verified" ; }
run PROBE_MODE=pinned,vk
run PROBE_MODE=pinned,vk,verify
