#!/bin/bash
# sweep the GEMM variants: 1-CTA vs CTA-pair kernel, direct vs smem-staged epilogue
mkdir -p gpurun_out
for cta in 1 2; do for epi in 0 1 2 3; do
  if [ $cta = 1 ]; then export LN3_GEMM_1CTA=1; else unset LN3_GEMM_1CTA; fi
  export LN3_GEMM_EPI=$epi
  echo "=== cta=$cta epi=$epi"
  timeout 100 python tools/gpu_check_gemm.py 2>&1 | grep -E "tflops|OK|FAIL|ERROR|rel_l2': 0\.[1-9]" | cut -c1-140
done; done
