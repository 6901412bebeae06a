#!/usr/bin/env bash
# Dev helper (run under gpurun): one box lease, several measurements.  usage: tools/gpu_session.sh <step> [<step> ...]
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for step in "$@"; do
  echo "=== $step" 
  case "$step" in
    chains)   timeout 300 ./tools/microbench/chain_r02 2>&1 | tee gpurun_out/chain_r02.txt ;;
    bench)    timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json ;;
    benchfull) timeout 1500 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json ;;
    patchncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 50 -c 1 \
                 -o gpurun_out/prof_patch -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_patch.log 2>&1; tail -3 gpurun_out/ncu_patch.log ;;
    projncu)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 40 -c 1 \
                 -o gpurun_out/prof_proj -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_proj.log 2>&1; tail -3 gpurun_out/ncu_proj.log ;;
    gemmncu)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 51 -c 4 \
                 -o gpurun_out/prof_gemms -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_gemms.log 2>&1; tail -3 gpurun_out/ncu_gemms.log ;;
    qkvncu)   timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 51 -c 1 \
                 -o gpurun_out/prof_qkv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_qkv.log 2>&1; tail -3 gpurun_out/ncu_qkv.log ;;
    attnncu)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 1 \
                 -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log ;;
    attntrace) VITB200_ATTN_TRACE=gpurun_out/attn_trace.txt timeout 600 python tools/attn_trace.py 2>&1 | tee gpurun_out/attn_trace_print.txt | tail -40 ;;
    tests)    timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15 ;;
    *)        echo "running: $step"; timeout 1200 bash -c "$step" ;;
  esac
done
