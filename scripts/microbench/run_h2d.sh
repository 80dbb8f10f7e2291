#!/bin/bash
# Per-GPU pinned host->device bandwidth with 1, 2, 4, 8 concurrent uploaders (one process per GPU).
# usage: scripts/microbench/run_h2d.sh [max_gpus] > gpurun_out/h2d.jsonl
set -e
cd "$(dirname "$0")"
[ -x h2d_bw.bin ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o h2d_bw.bin h2d_bw.cu
MAXG=${1:-$(nvidia-smi -L | wc -l)}
for N in 1 2 4 8; do
  [ "$N" -le "$MAXG" ] || break
  START=$(python3 -c "import time; print(time.time() + 3)")
  echo "{\"concurrent_uploaders\": $N}"
  for ((g = 0; g < N; g++)); do ./h2d_bw.bin $g 94 2 20 $START & done
  wait
done
