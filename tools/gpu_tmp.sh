#!/bin/bash
for lib in "" build/variants/libspn_a_unroll2.so ""; do
  echo "== ${lib:-product (unroll 4)}"
  SPN_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tools/step_profile.py 2>&1 | grep -i "leaf_tiled\|total" | head -4
done
timeout 300 python -m pytest tests/test_gpu_leaf.py -q 2>&1 | tail -2
