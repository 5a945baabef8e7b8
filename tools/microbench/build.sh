#!/bin/sh
# Builds the stand-alone microbenchmarks next to their sources (sm_100a; binaries are git-ignored
# and travel to the GPU box with the working tree).   tools/microbench/build.sh [name ...]
set -e
cd "$(dirname "$0")"
names="$*"
[ -n "$names" ] || names="solve_bench tc_bench dist_solve_bench red_bench bulk_issue_bench dmma_bench lat_bench tile_bench"
for n in $names; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -o "$n" "$n.cu"
  echo "built $n"
done
