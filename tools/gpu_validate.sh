#!/usr/bin/env bash
# One-call hardware validation (run under gpurun; N = number of GPUs of the box):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_validate.sh 2'
#   gpurun          --timeout  900 -- 'bash tools/gpu_validate.sh 1'
# Every section has its own timeout and log under gpurun_out/validate/; the summary at the end is what to read.
set -u
N=${1:-1}
OUT=gpurun_out/validate; mkdir -p "$OUT"
PORT=29600
run() {   # run <name> <timeout_s> <cmd...>
  local name=$1 t=$2; shift 2
  ( timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" >> "$OUT/$name.log" )
  printf "%-28s %s\n" "$name" "$(grep -E 'passed|failed|^\{|exit' "$OUT/$name.log" | tail -2 | tr '\n' ' ' | cut -c1-260)"
}
T=""   # torchrun prefix with a fresh rendezvous port (set by `nextport`; must run in THIS shell, not in $(...))
nextport() { PORT=$((PORT+1)); T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"; }

if [ "$N" -eq 1 ]; then
  run gpu_tests      600 python -m pytest tests -q -m gpu
  run smoke          200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"
  run gemm_ab        200 build/gemm_test 2 1 1
  run bench_n1       300 python bench.py --steps 5 --warmup 3
else
  nextport; run multigpu_tests 400 $T -m pytest tests/test_multigpu.py -m multigpu -q
  nextport; run pp_tour        120 $T examples/parallelism_tour.py --mode pp
  nextport; run fused_tp       200 $T benchmarks/fused_tp_bench.py
  nextport; run overlap        200 $T benchmarks/overlap_bench.py
  nextport; run bench          300 $T bench.py --gpus "$N" --steps 5 --warmup 3 --trace "$OUT/trace_n$N.json"
  nextport; TORCHACC_B200_SPLIT_HEAD=1 run bench_split_head 300 $T bench.py --gpus "$N" --steps 5 --warmup 3 --no-e2e
  gzip -f "$OUT"/trace_*.json 2>/dev/null
fi
echo "logs: $OUT/"
