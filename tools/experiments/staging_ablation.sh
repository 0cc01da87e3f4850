#!/bin/bash
# Builds three copies of the library with the producers' operand staging ablated (D2T_ABLATE_STAGING = 1, 2, 3; see the
# headers of csrc/corr_umma_fwd.cu / corr_umma_bwd.cu) into tools/_build/.  Results of those builds are garbage; only
# their kernel times mean something: they bound what a TMA-fed variant of the tensor-core correlation kernels (raw tiles
# land in shared memory without passing the LSU; threads only derive the lo parts) could reach on a 16-byte-aligned
# (W-padded) copy of the maps.  Timed by tools/time_staging_ablation.py.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
for m in 1 2 3; do
  make -s -j 16 -C "$ROOT/detect-to-track_b200/csrc" EXTRA=-DD2T_ABLATE_STAGING=$m \
       OUT="$ROOT/tools/_build/libd2t_ablate$m.so" OBJD="$ROOT/tools/_build/ablate$m"
done
