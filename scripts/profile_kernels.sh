#!/usr/bin/env bash
# ncu captures of the dominant kernels inside one bench step (one GPU; run under gpurun), as in round 1:
#   bash scripts/profile_kernels.sh [tag]          -> gpurun_out/<tag>_<kernel>.ncu-rep + <tag>_launches.csv
# Environment switches (MER_ATT_F16_VER=2 ...) are inherited, so a variant is profiled by exporting its switch.
# Read the reports back in the build container:
#   ncu -i gpurun_out/<tag>_attention_f16_kernel.ncu-rep --page raw --csv
#   ncu -i ... --page source --csv --print-source sass        (per-instruction executed counts / stall samples)
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p "$out"
# launch list of a window that contains at least one full step (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 900 --csv --log-file "$out/${tag}_launches.csv" \
    python bench.py --steps 1 --warmup 1 --cpu-clips 2 --no-extras > "$out/${tag}_launches.log" 2>&1
# kernel regex, matching launches to skip, launches to capture, name of the report.  GEMM launch order inside a step:
# [0] ViT patch embedding (TF32), [1..48] ViT layers (F16: qkv, out-proj, fc1, fc2 per layer), [49..54] HuBERT conv1-6
# (BF16X3), [55] feature projection, [56] positional conv (F16), then the HuBERT and BERT layers (F16)
for spec in "gemm_kernel 21 4 gemm_f16_vit_layer" "gemm_kernel 49 2 gemm_bf16x3_conv" "attention_f16_kernel 2 1 attention_f16" \
            "attention_tc_kernel 2 1 attention_tc" "layernorm_kernel 8 1 layernorm" "conv0_apply 0 1 conv0_apply" \
            "conv0_stats 0 1 conv0_stats" "fus_rows_fast_kernel 0 1 fus_rows" "fus_wgrad_kernel 0 1 fus_wgrad"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -f \
      -o "$out/${tag}_$4" python bench.py --steps 1 --warmup 1 --cpu-clips 2 --no-extras > "$out/${tag}_$4.log" 2>&1
  echo "$4: exit $?"
done
