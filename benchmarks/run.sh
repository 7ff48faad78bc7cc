#!/bin/bash
# Benchmark matrix (counterpart of reference benchmarks/run.sh:8-48): {1, 4 GPUs} x {dp, fsdp+gc} x {bf16, fp16}.
set -e
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for prec in --bf16 --fp16; do
  python benchmarks/transformer.py $prec
  $TR --nproc-per-node 4 benchmarks/transformer.py $prec --dp_size 4
  $TR --nproc-per-node 4 benchmarks/transformer.py $prec --fsdp_size 4 --gc
done
