echo "== default"; MMLA_RB_VERBOSE=1 python scripts/prof_resblock2d.py 2>&1 | sort -u | grep -v "^$"
for kb in 56 75 113 226; do for t in 1 2 3 4; do echo "== KB=$kb T=$t"; STAMPS=0 MMLA_RB_KB=$kb MMLA_RB_TILES=$t python scripts/prof_resblock2d.py 2>&1 | tail -2; done; done
