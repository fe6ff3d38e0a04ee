#!/usr/bin/env bash
# Round-2 single-GPU check:  gpurun -- 'bash tools/gpu_check_r2.sh [stage ...]'
set -u
mkdir -p gpurun_out
t() { local secs=$1; shift; timeout -k 10 "$secs" "$@"; }
for stage in "${@:-marlin rope smoke mbench bench}"; do
for s in $stage; do
case "$s" in
  marlin) t 420 python -m pytest tests/test_gpu_marlin.py tests/test_gpu_marlin_moe.py -x -q --tb=short 2>&1 | tail -12 | tee gpurun_out/r02_marlin_tests.log ;;
  attn)   t 600 python -m pytest tests/test_gpu_attention.py -q --tb=short -x 2>&1 | tail -8 | tee gpurun_out/r02_attn_tests.log
          for kv in auto fp8; do t 100 python tools/bench_attn.py --kv-dtype $kv 2>&1 | tail -1; done | tee gpurun_out/r02_attn_bench.jsonl
          t 100 python tools/bench_attn.py --kv-dtype fp8 --bs 1024 --ctx 8192 --heads 4 --kv-heads 1 --layers 1 2>&1 | tail -1 | tee -a gpurun_out/r02_attn_bench.jsonl
          t 100 python tools/bench_attn.py --kv-dtype auto --bs 256 --ctx 4096 --heads 4 --kv-heads 1 --layers 8 2>&1 | tail -1 | tee -a gpurun_out/r02_attn_bench.jsonl ;;
  moe)    t 300 python -m pytest tests/test_gpu_mixtral_moe.py tests/test_gpu_moe.py -q --tb=short 2>&1 | tail -8 | tee gpurun_out/r02_moe_tests.log ;;
  rope)   t 200 python -m pytest tests/test_gpu_fused_rope_cache.py tests/test_gpu_norm_rope_act.py tests/test_gpu_cache.py -q --tb=short 2>&1 | tail -8 | tee gpurun_out/r02_rope_tests.log ;;
  smoke)  t 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log ;;
  mbench) t 300 python tools/bench_marlin.py --m 256 64 --sustained 2>&1 | tee gpurun_out/r02_bench_marlin.jsonl | cut -c1-260 ;;
  vsref)  t 500 python tests/bench_vs_ref_cuda.py --ms 256 64 32 16 1 > gpurun_out/r02_vs_ref.jsonl 2> gpurun_out/r02_vs_ref.err; cut -c1-220 gpurun_out/r02_vs_ref.jsonl ;;
  bench)  t 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 2500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err ;;
  gptq)   t 400 python bench.py --quant gptq --no-cpu-baseline > gpurun_out/r02_bench_n1_gptq.json 2> gpurun_out/r02_bench_n1_gptq.err; tail -c 1500 gpurun_out/r02_bench_n1_gptq.json; tail -3 gpurun_out/r02_bench_n1_gptq.err ;;
  ncu)    # evidence captures (one GPU): launch list of a step, full sets of the fp8 attention, both Marlin kernels
          t 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_step_ncu.csv \
             python bench.py --steps 1 --warmup 1 --no-secondary --no-ref-cuda --no-cpu-baseline --no-graph > gpurun_out/r02_ncu_step.log 2>&1; tail -1 gpurun_out/r02_ncu_step.log | cut -c1-200
          t 240 ncu --set full --clock-control none --import-source on -k regex:paged_attention_tc -s 3 -c 1 -o gpurun_out/r02_attn_fp8_cfg4 -f \
             python tools/bench_attn.py --kv-dtype fp8 --bs 1024 --ctx 8192 --heads 4 --kv-heads 1 --layers 1 --iters 3 > gpurun_out/r02_ncu_attn.log 2>&1; tail -1 gpurun_out/r02_ncu_attn.log | cut -c1-200
          t 240 ncu --set full --clock-control none --import-source on -k regex:marlin_w4a16_tc5 -s 3 -c 1 -o gpurun_out/r02_marlin_tc5_m256 -f \
             python tools/bench_marlin.py --m 256 --iters 2 --kn 4096 28672 > gpurun_out/r02_ncu_marlin.log 2>&1; tail -1 gpurun_out/r02_ncu_marlin.log | cut -c1-200
          t 240 ncu --set full --clock-control none --import-source on -k regex:marlin_w4a16_small -s 3 -c 1 -o gpurun_out/r02_marlin_small_m16 -f \
             python tools/bench_marlin.py --m 16 --iters 2 --kn 4096 28672 > gpurun_out/r02_ncu_marlin_small.log 2>&1; tail -1 gpurun_out/r02_ncu_marlin_small.log | cut -c1-200 ;;
  full)   t 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu.log ;;
esac
done
done
