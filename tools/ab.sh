#!/bin/bash
# usage (GPU box): tools/ab.sh name1 name2 ...  — bench3 (c2, c3, c1; 50 steps) for every variants/libuvt_<name>.so, twice, interleaved
mkdir -p gpurun_out
for rep in 1 2; do for v in "$@"; do UVT_NO_REBUILD=1 tools/bench3.sh ab_${v}_$rep UVT_LIB_PATH=variants/libuvt_$v.so UVT_NO_REBUILD=1; done; done
