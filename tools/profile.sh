#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list of one bench step + full ncu captures of the dominant kernels.
# Outputs land in gpurun_out/ (scratch); tools/summarize_ncu.py turns them into the tracked profiles/*.md|csv.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}
# every launch of one warm step with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 100 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
# full sections for the GEMM family (one of each epilogue in the middle of the network) and the attention kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 21 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 5 -c 1 \
    -o gpurun_out/prof_attn_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 5 -c 1 \
    -o gpurun_out/prof_ln_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ln_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the two-sweep attention kernel (ViT-L/16-384 geometry: 577 tokens, 16 heads) on its own
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_long -c 1 \
    -o gpurun_out/prof_attn_long_${TAG} -f python tools/run_attn_long.py > gpurun_out/ncu_attn_long_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
