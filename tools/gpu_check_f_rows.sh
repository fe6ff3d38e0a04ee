#!/usr/bin/env bash
# SURVEY §8(f) rows on one GPU:  gpurun -- 'bash tools/gpu_check_f_rows.sh [stage ...]'
# Stages are ordered from the least to the most hang-prone kernel and each is time-boxed, so one stuck launch cannot
# cost the others their results.
set -u
mkdir -p gpurun_out
t() { local secs=$1; shift; timeout -k 10 "$secs" "$@"; }
for stage in "${@:-prefill sampling quant mm smoke bench}"; do
for s in $stage; do
case "$s" in
  prefill)  t 240 python -m pytest tests/test_gpu_f_rows.py -q --tb=short -k "prefill" 2>&1 | tail -40 | tee gpurun_out/r02_f1_prefill_tests.log ;;
  sampling) t 300 python -m pytest tests/test_gpu_f_rows.py -q --tb=short -k "sampl or renorm" 2>&1 | tail -40 | tee gpurun_out/r02_f2_sampling_tests.log ;;
  quant)    t 200 python -m pytest tests/test_gpu_f_rows.py -q --tb=short -k "fp8_quant" 2>&1 | tail -30 | tee gpurun_out/r02_f4_quant_tests.log ;;
  mm)       t 300 python -m pytest tests/test_gpu_f_rows.py -q --tb=short -k "scaled_mm" 2>&1 | tail -40 | tee gpurun_out/r02_f4_mm_tests.log ;;
  smoke)    t 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02_smoke.log ;;
  bench)    t 300 python bench_f_rows.py > gpurun_out/r02_bench_f_rows.json 2> gpurun_out/r02_bench_f_rows.err; tail -c 6000 gpurun_out/r02_bench_f_rows.json; tail -5 gpurun_out/r02_bench_f_rows.err ;;
  full)     t 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log ;;
  headline) t 500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 3000 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err ;;
  ncu)      # one full-set capture per new kernel + SASS-level evidence comes from tools/sass_opcodes.py on the CPU box
            t 200 ncu --set full --clock-control none --import-source on -k regex:scaled_mm_tc5 -s 4 -c 1 -o gpurun_out/r02_scaled_mm_fp8_m256 -f \
               python tools/f_rows_driver.py mm > gpurun_out/r02_ncu_mm.log 2>&1; tail -2 gpurun_out/r02_ncu_mm.log | cut -c1-200
            t 200 ncu --set full --clock-control none --import-source on -k regex:prefill_attention -s 1 -c 1 -o gpurun_out/r02_prefill_attention -f \
               python tools/f_rows_driver.py prefill > gpurun_out/r02_ncu_prefill.log 2>&1; tail -2 gpurun_out/r02_ncu_prefill.log | cut -c1-200
            t 200 ncu --set full --clock-control none --import-source on -k regex:rejection_sampling -s 1 -c 1 -o gpurun_out/r02_sampling_topk -f \
               python tools/f_rows_driver.py sampling > gpurun_out/r02_ncu_sampling.log 2>&1; tail -2 gpurun_out/r02_ncu_sampling.log | cut -c1-200
            t 200 ncu --set full --clock-control none --import-source on -k regex:fp8_quant_token -s 1 -c 1 -o gpurun_out/r02_fp8_quant_token -f \
               python tools/f_rows_driver.py quant > gpurun_out/r02_ncu_quant.log 2>&1; tail -2 gpurun_out/r02_ncu_quant.log | cut -c1-200 ;;
esac
done
done
