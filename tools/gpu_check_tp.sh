#!/usr/bin/env bash
# Multi-GPU checks in one gpurun call:  gpurun --gpus N -- 'bash tools/gpu_check_tp.sh N [quick]'
set -u
N=${1:-2}
mkdir -p gpurun_out
t() { local secs=$1; shift; timeout -k 10 "$secs" "$@"; }
run() { t "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 "${@:2}"; }
run 100 tools/debug_tp_fused.py > gpurun_out/r02_debug_tp_n$N.log 2>&1; grep "us/call\|us/pair\|equal\|Error\|error" gpurun_out/r02_debug_tp_n$N.log | sed 's/\[r[1-9][^]]*\][^[]*//g' | head -30
t 400 python -m pytest tests/test_gpu_tp_fused.py tests/test_gpu_tp_decoder.py tests/test_gpu_custom_all_reduce.py -q --tb=short -s -k "[$N]" 2>&1 | tail -25 | tee gpurun_out/r02_tp_tests_n$N.log
if [ "${2:-}" != "quick" ]; then
  run 500 bench.py --gpus "$N" --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 3000 gpurun_out/r02_bench_n$N.json; tail -5 gpurun_out/r02_bench_n$N.err
fi
