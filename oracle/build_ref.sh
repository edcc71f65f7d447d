#!/usr/bin/env bash
# Builds the UNMODIFIED reference (niclasberg/biblbm) library objects into oracle/_ref/ from the
# sources where they lie under /root/reference. TEST INFRASTRUCTURE ONLY.
# No reference source is copied into this repo; only compiled objects/binaries land in oracle/_ref/
# (git-ignored). Usage: oracle/build_ref.sh [harness ...]   (default: all harnesses in oracle/harness)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${BIBLBM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF (fine on the GPU box: prebuilt oracle/_ref is used)"; exit 0; fi
mkdir -p "$OUT/obj"
export FLAGS="-std=gnu++98 -fpermissive -w -O2 -DPLB_MPI_PARALLEL -I$HERE/mpi_shim -I$REF/src -I$REF/externalLibraries"
export OUT
if [ ! -f "$OUT/libplb_ref.a" ]; then
  ls "$REF"/src/*/*.cpp "$REF"/externalLibraries/tinyxml/*.cpp | \
    xargs -P "$JOBS" -I{} sh -c 'o="$OUT/obj/$(echo {} | tr / _).o"; [ -f "$o" ] || g++ $FLAGS -c {} -o "$o"'
  rm -f "$OUT/libplb_ref.a"; ar rcs "$OUT/libplb_ref.a" "$OUT"/obj/*.o
fi
HARNESSES=("$@")
if [ ${#HARNESSES[@]} -eq 0 ]; then HARNESSES=($(ls "$HERE"/harness/*.cpp 2>/dev/null | xargs -n1 basename | sed 's/\.cpp$//')); fi
for h in "${HARNESSES[@]}"; do
  src="$HERE/harness/$h.cpp"
  if [ ! -f "$OUT/$h" ] || [ "$src" -nt "$OUT/$h" ]; then
    echo "building oracle/_ref/$h"
    g++ $FLAGS -O3 -I"$REF/progs" -I"$HERE/harness" "$src" "$OUT/libplb_ref.a" -o "$OUT/$h"
  fi
done
