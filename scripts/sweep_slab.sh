#!/bin/bash
# per-launch conv_slab_kernel times of one overlap step (512 clips) for a few shared-memory budgets / tile counts
for cfg in "MMLA_CONV_SLAB_KB=113" "MMLA_CONV_SLAB_KB=75" "MMLA_CONV_SLAB_KB=56" "MMLA_CONV_SLAB_TILES=2" "MMLA_CONV_SLAB_TILES=1"; do
  echo "== $cfg"
  env $cfg timeout 120 python scripts/trace_overlap.py | awk '/conv_slab/{printf "%s ", $2; s+=$2} /^total/{t=$2} END{printf "\n slab %.3f total %.3f\n", s, t}'
done
