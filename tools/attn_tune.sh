#!/bin/bash
# attention kernel tuning sweep (run under gpurun)
for p in 0 1 2 3 4; do
  echo "== SA_ATTN_POLY=$p"
  SA_ATTN_POLY=$p python tools/gpu_kernel_check.py attn_small attn_tail attn_big 2>&1 | grep -E "rel|TFLOP|exit [1-9]"
done
