#!/bin/bash
for b in 64 128 256 512; do
  for l in 0 1; do
    echo "== B=$b LDP_LOOP=$l: $(LDP_B=$b LDP_LOOP=$l timeout 300 python scripts/ops_profile.py small_${b}_$l 2>&1 | head -1)"
  done
done
