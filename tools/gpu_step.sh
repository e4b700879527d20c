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
    pcg)     timeout 600 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_dist.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "not two_gpus" > "$out/pcg.log" 2>&1; tail -15 "$out/pcg.log" ;;
    l2p)     for w in 2m c2; do for e in 0 1; do PF2_L2_PERSIST=$e PF2_PCG=0 timeout 120 python tools/pcg_tune.py $w 2 2>&1 | tail -1 | cut -c1-260 | tee -a "$out/l2p.jsonl"; done; done ;;
    suite)   timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > "$out/suite.log" 2>&1; tail -25 "$out/suite.log" ;;
    smoke)   timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -3 "$out/smoke.log" ;;
    quick)   timeout 600 python bench.py --steps 3 --warmup 2 --hex8 c4s --record "$out/record_quick.json" > "$out/bench_quick.json" 2> "$out/bench_quick.err"; tail -c 3000 "$out/bench_quick.json"; tail -5 "$out/bench_quick.err" ;;
    pcg0)    PF2_PCG=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-hex8 --no-extra-legs --no-cpu-baseline > "$out/bench_pcg0.json" 2> "$out/bench_pcg0.err"; tail -c 1500 "$out/bench_pcg0.json" ;;
    full)    timeout 1500 python bench.py --steps 20 --warmup 5 --record "$out/record_full.json" > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 4000 "$out/bench_full.json"; tail -5 "$out/bench_full.err" ;;
    ref)     timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; tail -c 2500 "$out/bench_ref.json" ;;
    san)     timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pcg.py -x -q -k "three_kernel or iteration_cap or one_rank" > "$out/san_memcheck.log" 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" "$out/san_memcheck.log" | tail -3
             timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pcg.py -x -q -k "three_kernel and lambda3" > "$out/san_racecheck.log" 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" "$out/san_racecheck.log" | tail -3 ;;
    tune)    for w in ${TUNE_W:-2m c2 c4s}; do
               PF2_PCG=0 timeout 120 python tools/pcg_tune.py $w 2>&1 | tail -1 | tee -a "$out/tune.jsonl"
               for lib in pansfem2_b200/bin/libpf2_*.so; do PF2_LIB=$PWD/$lib timeout 120 python tools/pcg_tune.py $w 2>&1 | tail -1 | tee -a "$out/tune.jsonl"; done
             done ;;
    tune2)   for w in ${TUNE_W:-2m}; do
               for lib in pansfem2_b200/bin/libpf2_*.so; do for g in 0 592 444 296; do
                 PF2_PCG_GRID=$g PF2_LIB=$PWD/$lib timeout 120 python tools/pcg_tune.py $w 2>&1 | tail -1 | tee -a "$out/tune2.jsonl"; done; done
             done ;;
    dist2)   timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -rs > "$out/dist2.log" 2>&1; tail -12 "$out/dist2.log" ;;
    small)   for w in c1 pair c4s; do for m in 0 1; do PF2_PCG=$m timeout 120 python tools/pcg_tune.py $w 2 2>&1 | tail -1 | tee -a "$out/small.jsonl"; done; done ;;
    b2)      for m in 1 0; do PF2_PCG=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-hex8 --no-extra-legs --no-headline-2m > "$out/bench_n2_pcg$m.json" 2> "$out/bench_n2_pcg$m.err"; tail -c 1800 "$out/bench_n2_pcg$m.json"; tail -3 "$out/bench_n2_pcg$m.err"; done ;;
    b8)      for m in 1 0; do PF2_PCG=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus ${NG:-8} --steps 3 --warmup 2 --hex8 c4 --no-extra-legs --no-headline-2m > "$out/bench_n${NG:-8}_pcg$m.json" 2> "$out/bench_n${NG:-8}_pcg$m.err"; python tools/bench_brief.py "$out/bench_n${NG:-8}_pcg$m.json"; tail -3 "$out/bench_n${NG:-8}_pcg$m.err" | cut -c1-300; done ;;
    probe)   for w in ${PROBE_W:-c2 c4}; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29531 tools/dist_pcg_probe.py $w ${PROBE_IT:-1500} 2>"$out/probe_$w.err" | grep "^{" | tee -a "$out/probe.jsonl"; tail -2 "$out/probe_$w.err" | cut -c1-200; done ;;
    asm3)    ASM_ONLY=h8 timeout 300 python tools/assemble_probe.py "$out/assemble_probe_h8.json" > "$out/assemble_probe_h8.log" 2>&1; tail -25 "$out/assemble_probe_h8.log" ;;
    new)     timeout 900 python -m pytest tests/test_gpu_loadvec.py tests/test_gpu_pcg.py tests/test_gpu_cpp_dropin.py tests/test_gpu_matrix_free.py -x -q -k "loadvec or load_vec or selection or assembled_loads or warm or batched or hex8_gather" > "$out/new.log" 2>&1; tail -8 "$out/new.log"
             timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "hex8_gather or reproducible" >> "$out/new.log" 2>&1; tail -4 "$out/new.log" ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file "$out/launches_c2.csv" python bench.py --steps 1 --warmup 1 --no-hex8 --no-extra-legs --no-headline-2m --no-cpu-baseline > "$out/launches_bench.log" 2>&1; tail -2 "$out/launches_bench.log" | cut -c1-200; wc -l "$out/launches_c2.csv" ;;
    ncufull) PF2_PCG=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_kernel|cg_update_kernel|cg_pupdate_kernel" -s 30 -c 3 -o "$out/cg_kernels_c2" python tools/ncu_target.py 2d 2000 1000 --itr 40 2>&1 | tail -3 ;;
    san2)    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_loadvec.py tests/test_gpu_dist.py -x -q -k "not two_gpus and not matches_the_oracle" > "$out/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" "$out/sanitizer_memcheck.log" | tail -3
             timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pcg.py -x -q -k "three_kernel_loop_and_oracle and lambda3 or one_launch_sweeps" > "$out/sanitizer_racecheck.log" 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" "$out/sanitizer_racecheck.log" | tail -3 ;;
    cg1)     [ -n "${SKIP_TESTS:-}" ] || timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -rs -k "cg1 or env18 or env19 or env20 or env21 or env22" > "$out/cg1.log" 2>&1; tail -12 "$out/cg1.log"
             [ -n "${SKIP_BENCH:-}" ] || timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus ${NG:-2} --steps 2 --warmup 1 --no-hex8 --no-headline-2m > "$out/bench_cg1_n${NG:-2}.json" 2> "$out/bench_cg1_n${NG:-2}.err"; python tools/bench_brief.py "$out/bench_cg1_n${NG:-2}.json"; tail -3 "$out/bench_cg1_n${NG:-2}.err" | cut -c1-300 ;;
    hex20)   timeout 400 python -m pytest tests/test_gpu_families.py tests/test_gpu_dist.py -x -q -k "hex20 or one_rank" > "$out/hex20.log" 2>&1; tail -5 "$out/hex20.log"
             timeout 200 python tools/hex20_probe.py "$out/hex20_probe.json" 2>&1 | tail -2 ;;
    *)       echo "unknown stage $stage" ;;
  esac
done
