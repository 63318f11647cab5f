#!/bin/bash
# Round-2 validation on an 8-GPU box (under gpurun --gpus 8): the 8-rank cases of tests/test_gpu_multi.py incl. the
# 2^22-gate golden proof (BASELINE.json configs[4]), then bench.py at 8 GPUs for 2^20 and 2^22 gates.
mkdir -p gpurun_out
T=tests/test_gpu_multi.py
PYTHONFAULTHANDLER=1 timeout 900 python -m pytest -x -q -m gpu \
  "$T::test_sharded_proof_equals_single_gpu[8-12-2]" "$T::test_sharded_proof_equals_single_gpu[8-16-12]" \
  "$T::test_sharded_proof_equals_single_gpu[4-12-3]" \
  "$T::test_sharded_operators_equal_single_gpu[8-14]" "$T::test_sharded_operators_equal_single_gpu[8-22]" \
  "$T::test_sharded_2p22_gates_against_golden[8]" > gpurun_out/r2_tests_8gpu.log 2>&1
tail -4 gpurun_out/r2_tests_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
grep "rank 0" gpurun_out/r2_bench_8gpu.err | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 8 --log-n 22 --seed 7 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2_bench_8gpu_2p22.json 2> gpurun_out/r2_bench_8gpu_2p22.err
grep "rank 0" gpurun_out/r2_bench_8gpu_2p22.err | tail -6
