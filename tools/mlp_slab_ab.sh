#!/bin/bash
# A/B of the forward kernel's two forms (W1 resident / W1 streamed in slabs) over the feature width
for cfg in "" "2,4" "1,8" "4,2" "3,3"; do
  echo "== LTR_MLP_SLAB=$cfg"
  LTR_MLP_SLAB=$cfg timeout 300 python tools/mlp_width_sweep.py 48 64 136 160 220 288 320 700 2>&1 | grep "^F="
done
