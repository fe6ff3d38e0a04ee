#!/usr/bin/env bash
# One gpurun call = several minutes of fixed cost (box acquisition + push), so batch work per call:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh full'
# Every step runs under its own `timeout -k` so that a hung kernel cannot hold the box; outputs land in gpurun_out/.
#   quick   marlin + moe + attention-vs-reference tests (~40 s)        marlin   marlin tests + GEMM bench + cycle counters
#   full    whole -m gpu suite, smoke(), bench.py (bf16 and gptq)      attn     attention tests + fp8 / bf16 micro-bench
#   vsref   per-kernel table beside the reference's CUDA kernels       ncu-marlin / ncu-attn   one --set full capture
set -u
mkdir -p gpurun_out
t() { local secs=$1; shift; timeout -k 10 "$secs" "$@"; }
case "${1:-quick}" in
  quick)
    t 400 python -m pytest tests/test_gpu_marlin.py tests/test_gpu_marlin_moe.py tests/test_gpu_vs_ref_cuda.py -q --tb=short 2>&1 | tail -15 | tee gpurun_out/check_quick.log ;;
  full)
    t 700 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee gpurun_out/check_pytest.log
    t 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/check_smoke.log
    t 300 python bench.py > gpurun_out/check_bench_n1.json 2> gpurun_out/check_bench_n1.err; cut -c1-220 gpurun_out/check_bench_n1.json
    t 300 python bench.py --quant gptq > gpurun_out/check_bench_gptq.json 2> gpurun_out/check_bench_gptq.err; cut -c1-220 gpurun_out/check_bench_gptq.json ;;
  marlin)
    t 300 python -m pytest tests/test_gpu_marlin.py tests/test_gpu_marlin_moe.py -q --tb=short 2>&1 | tail -6 | tee gpurun_out/check_marlin.log
    t 120 python tools/prof_marlin.py 2>&1 | tee gpurun_out/check_marlin_prof.log
    t 300 python tools/bench_marlin.py --m 256 64 16 --sustained 2>&1 | tee gpurun_out/check_marlin_bench.jsonl | cut -c1-200 ;;
  attn)
    t 500 python -m pytest tests/test_gpu_attention.py -q --tb=short 2>&1 | tail -5 | tee gpurun_out/check_attn.log
    for kv in auto fp8; do t 100 python tools/bench_attn.py --kv-dtype $kv 2>&1 | tail -1; done | tee gpurun_out/check_attn_bench.jsonl
    t 100 python tools/bench_attn.py --kv-dtype fp8 --bs 1024 --ctx 8192 --heads 4 --kv-heads 1 --layers 1 2>&1 | tail -1 | tee -a gpurun_out/check_attn_bench.jsonl ;;
  vsref)
    t 500 python tests/bench_vs_ref_cuda.py --ms 256 64 16 > gpurun_out/check_vs_ref.jsonl 2> gpurun_out/check_vs_ref.err; cut -c1-200 gpurun_out/check_vs_ref.jsonl ;;
  ncu-marlin)
    t 300 ncu --set full --clock-control none --import-source on -k regex:marlin_w4a16_tc5 -s 3 -c 1 -o gpurun_out/check_marlin_full -f \
      python tools/bench_marlin.py --m 256 --iters 2 --kn 4096 28672 > gpurun_out/check_ncu_marlin.log 2>&1; tail -2 gpurun_out/check_ncu_marlin.log ;;
  ncu-attn)
    t 300 ncu --set full --clock-control none --import-source on -k regex:paged_attention_tc -s 2 -c 1 -o gpurun_out/check_attn_full -f \
      python tools/bench_attn.py --kv-dtype "${2:-auto}" --iters 2 --layers 1 > gpurun_out/check_ncu_attn.log 2>&1; tail -2 gpurun_out/check_ncu_attn.log ;;
  *) echo "usage: $0 quick|full|marlin|attn|vsref|ncu-marlin|ncu-attn [kv-dtype]"; exit 2 ;;
esac
