# one launch of each non-GEMM encoder kernel + one BF16X3 GEMM, inside the bench (second step)
for spec in "layernorm_kernel 80" "attention_f16_kernel 13" "attention_tc_kernel 26" "conv0_apply_kernel 1" "conv0_stats_kernel 1" "cast_f16_kernel 1"; do
  set -- $spec
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_r1_$1 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-150
done
timeout 300 python -m pytest tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
