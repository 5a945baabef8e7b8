#!/bin/sh
# compute-sanitizer recipe for the CUDA path on the small parity fixtures (run on a GPU box):
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# memcheck runs with torch's caching allocator OFF, so that an access past the end of a torch tensor
# is an access past a real cudaMalloc allocation (with the caching allocator such reads land inside
# torch's pool and go unnoticed -- this is how the prefetch past the last chunk of the tile-streamed
# back-substitution was found).  racecheck / synccheck cover the shared-memory protocols of the
# elimination kernel and the solver (named barriers, staging slots, bulk-copy sources).
# The tcgen05 trailing update / blocked solve have their own harness (memcheck, racecheck and synccheck
# clean: profiles/r2t_sanitizer.txt):
#   compute-sanitizer --tool memcheck tools/microbench/tc_bench syrk 704 2 6 64
#   compute-sanitizer --tool memcheck tools/microbench/tc_bench solve 199 8 6 64 1 1
set -e
cd "$(dirname "$0")/.."
tool="${1:-memcheck}"
expr="${2:-stages or update_matches or ragged or subset or param_mask or rank_deficient}"
PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool "$tool" --print-limit 5 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$expr"
