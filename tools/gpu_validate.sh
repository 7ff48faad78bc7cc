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
  run gpu_tests      700 python -m pytest tests -q -m gpu
  run smoke          200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"
  run bench_n1       300 python bench.py --steps 5 --warmup 3
  # ncu --set full on the round-2 kernels (one launch each after a warm-up launch): summaries go to profiles/
  [ "${VALIDATE_NCU:-0}" = "1" ] && run ncu_r2         400 ncu --set full --clock-control none --import-source on \
      -k "regex:gemm_mxfp8_kernel|gemm_bf16_kernel|flash_fwd_kernel" --launch-skip 7 --launch-count 4 \
      -f -o "$OUT/ncu_r2" python tools/ncu_targets.py
else
  nextport; run multigpu_tests 400 $T -m pytest tests/test_multigpu.py -m multigpu -q
  nextport; run bench          300 $T bench.py --gpus "$N" --steps 5 --warmup 3
  # compute-sanitizer on the peer-memory kernels (rank-local tools: synccheck = barrier misuse, racecheck = smem hazards)
  [ "${VALIDATE_SANITIZER:-0}" = "1" ] && nextport && run synccheck      150 $T --no-python compute-sanitizer --tool synccheck --log-file "$OUT/synccheck.%p.txt" \
      python -m pytest tests/test_multigpu.py -m multigpu -q -k "symm_collectives or carried or fused_tp"
  if [ "${VALIDATE_EXTRA:-0}" = "1" ]; then
    nextport; run pp_tour        120 $T examples/parallelism_tour.py --mode pp
    nextport; run fused_tp       200 $T benchmarks/fused_tp_bench.py
    nextport; run bench_trace    300 $T bench.py --gpus "$N" --steps 5 --warmup 3 --trace "$OUT/trace_n$N.json"
    gzip -f "$OUT"/trace_*.json 2>/dev/null
  fi
fi
echo "logs: $OUT/"
