#!/bin/bash
# usage (GPU box): tools/substep_block_sweep.sh  — substep phase time vs CTA size of the substep kernel
for b in 32 64 128 256; do
  echo "block $b"; TI5_SUBSTEP_BLOCK=$b python tools/substep_ablate.py 8192 2>&1 | grep "all features"
done
