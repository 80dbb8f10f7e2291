timeout 300 python -m pytest tests/test_conv_slab_gpu.py -x -q 2>&1 | tail -5
MMLA_CONV_SLAB_VERBOSE=1 timeout 120 python scripts/prof_conv_slab.py 2>&1 | awk '!seen[$0]++' | tee gpurun_out/prof_conv_slab_v6.txt
for cfg in "X=1" "MMLA_CONV_SLAB_EPI=0" "MMLA_CONV_SLAB_KB=226" "MMLA_CONV_SLAB_TILES=2"; do
  echo "== $cfg"
  env $cfg timeout 120 python scripts/trace_overlap.py | awk '/conv_slab/{printf "%s ", $2; s+=$2} /^total/{t=$2} END{printf "\n slab %.3f total %.3f\n", s, t}'
done 2>&1 | tee gpurun_out/sweep_slab_v6.txt
