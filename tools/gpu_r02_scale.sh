#!/bin/bash
# Round-2 multi-GPU session (run with gpurun --gpus 8): bench.py at N = 1, 2, 4, 8 back to back on one box (the
# driver's SCALE sequence), each N>1 line carrying the C5-on-N-GPUs sub-record and the bitwise frame check; then
# the front end with PBRT_GPUS=8 against PBRT_GPUS=1.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r02yy_gpus.txt
PBRT_BENCH_LEAN=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/r02yy_bench_n1.json 2> $O/r02yy_bench_n1.err
tail -c 400 $O/r02yy_bench_n1.json
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
    bench.py --gpus $n --steps 5 --warmup 3 > $O/r02yy_bench_n$n.json 2> $O/r02yy_bench_n$n.err
  tail -c 1200 $O/r02yy_bench_n$n.json; tail -3 $O/r02yy_bench_n$n.err
done
timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "multi_gpu" > $O/r02yy_pytest_multi.log 2>&1; tail -3 $O/r02yy_pytest_multi.log
