#!/bin/bash
# attention kernel tuning sweep (run under gpurun): SA_ATTN_IMPL x SA_ATTN_POLY
for impl in ${IMPLS:-4}; do for p in ${POLYS:-0 2 4}; do
  echo "== SA_ATTN_IMPL=$impl SA_ATTN_POLY=$p"
  SA_ATTN_IMPL=$impl SA_ATTN_POLY=$p python tools/gpu_kernel_check.py attn_small attn_tail attn_cross attn_big 2>&1 | grep -E "rel|TFLOP|exit [1-9]|timed out|Error|error" | head -12
done; done
