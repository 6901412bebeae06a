#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list of one bench step + full ncu captures of the dominant kernels.
# Outputs land in gpurun_out/ (scratch); tools/summarize_ncu.py turns them into the tracked profiles/*.md|json.
# Launch order of one forward (batch 256, ViT-B): patchify, cls_rows, patch GEMM, LN, then per block qkv GEMM, attention, proj GEMM, LN,
# fc1 GEMM, fc2 GEMM (+ LN of the next block), final LN, head GEMM, soft-max: 89 launches, 50 of them GEMMs.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --sustained-seconds 0.05"
# every launch of one warm step with its device time (cold-cache, serialised: compare SHARES, not absolutes); bench runs 3 warm-up forwards first
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 267 -c 89 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
# full sections: the four block GEMMs of the second forward's first block (qkv, proj, fc1, fc2) ...
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 51 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f $BENCH > gpurun_out/ncu_gemm_${TAG}.log 2>&1
# ... the patch-embedding GEMM and the patchify kernel in front of it ...
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 50 -c 1 \
    -o gpurun_out/prof_patch_${TAG} -f $BENCH > gpurun_out/ncu_patch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patchify -s 1 -c 1 \
    -o gpurun_out/prof_patchify_${TAG} -f $BENCH > gpurun_out/ncu_patchify_${TAG}.log 2>&1
# ... the opt-in single-kernel gather variant of the patch embedding, for the record ...
VITB200_PATCH_GATHER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 50 -c 1 \
    -o gpurun_out/prof_patchgather_${TAG} -f $BENCH > gpurun_out/ncu_patchgather_${TAG}.log 2>&1
# ... attention (split-precision operands), LayerNorm ...
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 1 \
    -o gpurun_out/prof_attn_${TAG} -f $BENCH > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 27 -c 1 \
    -o gpurun_out/prof_ln_${TAG} -f $BENCH > gpurun_out/ncu_ln_${TAG}.log 2>&1
# ... and the two-sweep attention kernel (ViT-L/16-384 geometry: 577 tokens, 16 heads) on its own
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_long -c 1 \
    -o gpurun_out/prof_attn_long_${TAG} -f python tools/run_attn_long.py > gpurun_out/ncu_attn_long_${TAG}.log 2>&1
ls -la gpurun_out/*_${TAG}.ncu-rep
