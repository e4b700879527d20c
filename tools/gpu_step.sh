#!/bin/bash
# One GPU slot: run the named stages with their own timeouts; logs land in gpurun_out/<tag>/ (merged back by gpurun).
#   gpurun --timeout 900 -- 'bash tools/gpu_step.sh TAG stage1 stage2 ...'
set -u
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$out/smi.txt" 2>&1
for stage in "$@"; do
  echo "== $stage"
  case $stage in
    pcg)     timeout 300 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_dist.py -x -q -rs > "$out/pcg.log" 2>&1; tail -15 "$out/pcg.log" ;;
    suite)   timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > "$out/suite.log" 2>&1; tail -25 "$out/suite.log" ;;
    smoke)   timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -3 "$out/smoke.log" ;;
    quick)   timeout 600 python bench.py --steps 3 --warmup 2 --hex8 c4s --record "$out/record_quick.json" > "$out/bench_quick.json" 2> "$out/bench_quick.err"; tail -c 3000 "$out/bench_quick.json"; tail -5 "$out/bench_quick.err" ;;
    pcg0)    PF2_PCG=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-hex8 --no-extra-legs --no-cpu-baseline > "$out/bench_pcg0.json" 2> "$out/bench_pcg0.err"; tail -c 1500 "$out/bench_pcg0.json" ;;
    full)    timeout 1500 python bench.py --steps 20 --warmup 5 --record "$out/record_full.json" > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 4000 "$out/bench_full.json"; tail -5 "$out/bench_full.err" ;;
    ref)     timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; tail -c 2500 "$out/bench_ref.json" ;;
    *)       echo "unknown stage $stage" ;;
  esac
done
