#!/usr/bin/env bash
# Compiles the reference's own drivers (progs/*) against the B200 binding: the overlay include directory comes before the
# reference's src/ so that "fsi/headers.h" / "fsi/headers.hh" resolve to overlay/fsi/*, which include the reference headers
# unchanged (#include_next) and add the bindings of overlay/biblbm_overlay/. Nothing of the reference is copied into this
# repository: sources are compiled where they lie under /root/reference; binaries land in overlay/_build/ (git-ignored).
#
#   overlay/build_drivers.sh                 # every supported driver, unchanged source  -> overlay/_build/<driver>
#   plus, for the parity tests, a parameter-patched copy (step count / mesh file / output cadence read from the environment;
#   the time loop and all calls are untouched) built twice: against the binding (-> overlay/_build/<driver>_short_b200) and
#   against the pure reference (-> oracle/_ref/<driver>_short_cpu).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${BIBLBM_REFERENCE:-/root/reference}"
OUT="$HERE/_build"
REFOUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF (prebuilt binaries in overlay/_build are used)"; exit 0; fi
[ -f "$REFOUT/libplb_ref.a" ] || "$ROOT/oracle/build_ref.sh" >/dev/null
[ -f "$ROOT/biblbm_b200/libbiblbm_b200.so" ] || python "$ROOT/biblbm_b200/build.py" >/dev/null
mkdir -p "$OUT/src"
BASE="-std=gnu++98 -fpermissive -w -O2 -DPLB_MPI_PARALLEL -I$ROOT/oracle/mpi_shim"
REFINC="-I$REF/src -I$REF/externalLibraries -I$REF/progs"
B200="-I$HERE -I$ROOT/include"
LINK_B200="$REFOUT/libplb_ref.a -L$ROOT/biblbm_b200 -lbiblbm_b200 -Wl,-rpath,\$ORIGIN/../../biblbm_b200"

build() {   # build <out> <src> <driver dir> <b200|cpu> [extra flags]
  local out="$1" src="$2" dir="$3" kind="$4" extra="${5:-}"
  if [ -f "$out" ] && [ "$out" -nt "$src" ] && [ "$out" -nt "$HERE/biblbm_overlay/impl.hh" ] && [ "$out" -nt "$HERE/biblbm_overlay/decl.h" ] \
     && [ "$out" -nt "$HERE/biblbm_overlay/impl_desc.inc" ] && [ "$out" -nt "$HERE/biblbm_overlay/decl_desc.inc" ] \
     && [ "$out" -nt "$ROOT/include/biblbm_b200.h" ]; then return; fi
  echo "building $out"
  if [ "$kind" = b200 ]; then g++ $BASE $extra $B200 $REFINC -I"$dir" "$src" $LINK_B200 -o "$out"
  else g++ $BASE $REFINC -I"$dir" "$src" "$REFOUT/libplb_ref.a" -o "$out"; fi
}

# parameter patch: literal values -> environment (defaults = the driver's own values)
patch_params() {  # patch_params <driver.cpp> <out.cpp>
  sed -E \
    -e 's|^(\s*)plint Nt = ([0-9]+);|\1plint Nt = getenv("BIB_NT") ? atol(getenv("BIB_NT")) : \2;|' \
    -e 's|^(\s*)plint output_interval = ([0-9]+);|\1plint output_interval = getenv("BIB_OUTPUT_INTERVAL") ? atol(getenv("BIB_OUTPUT_INTERVAL")) : \2;|' \
    -e 's|for\(int i = 0; i < 13; \+\+i\) \{|for(int i = 0; i < (getenv("BIB_NCASES") ? atoi(getenv("BIB_NCASES")) : 13); ++i) {|' \
    -e 's|^(\s*)const char \* mesh_file = "sphere_dx0_33.msh";|\1const char * mesh_file = getenv("BIB_MESH") ? getenv("BIB_MESH") : "sphere_dx0_33.msh";|' \
    -e 's|^(\s*)const char \* mesh_file = "sphere_r8.gmsh";|\1const char * mesh_file = getenv("BIB_MESH") ? getenv("BIB_MESH") : "sphere_r8.gmsh";|' \
    -e 's|^(\s*)plint geoNx = 96, geoNy = 96, geoNz = 96;|\1plint geoNx = getenv("BIB_GEO_N") ? atol(getenv("BIB_GEO_N")) : 96, geoNy = geoNx, geoNz = geoNx;|' \
    -e 's|^(\s*)plint Nt = std::ceil\(20 / gamma_lb\);|\1plint Nt = getenv("BIB_NT") ? atol(getenv("BIB_NT")) : (plint) std::ceil(20 / gamma_lb);|' \
    -e 's|^(\s*)plint output_interval = std::ceil\(1 /gamma_lb / 20\);|\1plint output_interval = getenv("BIB_OUTPUT_INTERVAL") ? atol(getenv("BIB_OUTPUT_INTERVAL")) : (plint) std::ceil(1 /gamma_lb / 20);|' \
    -e 's|get_volume\(\), 10000\);|get_volume(), getenv("BIB_SHRINK_ITERS") ? atol(getenv("BIB_SHRINK_ITERS")) : 10000);|' \
    -e 's|get_volume\(\), 100000\);|get_volume(), getenv("BIB_SHRINK_ITERS") ? atol(getenv("BIB_SHRINK_ITERS")) : 100000);|' \
    -e 's|^(\s*)plint Nit = 40000;|\1plint Nit = getenv("BIB_GROW_ITERS") ? atol(getenv("BIB_GROW_ITERS")) : 40000;|' \
    -e 's|shrink_volume\(0.59, 40000\)|shrink_volume(0.59, getenv("BIB_SHRINK_ITERS") ? atol(getenv("BIB_SHRINK_ITERS")) : 40000)|' \
    "$1" > "$2"
}

DRIVERS="${@:-rbc_in_shear capsule_in_shear rbcs_in_shear rbc_in_pipe rigid_particle_in_shear sphere_in_ext_flow deformation_test}"
JOBS="${BIB_BUILD_JOBS:-$(( $(nproc) > 6 ? 6 : $(nproc) ))}"      # one g++ per translation unit, ~2 GB each
TASKS="$(mktemp)"
trap 'rm -f "$TASKS"' EXIT
for drv in $DRIVERS; do
  dir="$REF/progs/$drv"
  src="$(ls "$dir"/*.cpp | head -1)"
  extra="-"
  # sphere_in_ext_flow defines its own data-processing functional with a private iteration count that the binding has to read
  # (overlay/biblbm_overlay/decl_desc.inc): access control is switched off for this driver's build instead of editing its source
  [ "$drv" = sphere_in_ext_flow ] && extra="-fno-access-control"
  patch_params "$src" "$OUT/src/${drv}_short.cpp.new"
  if ! cmp -s "$OUT/src/${drv}_short.cpp.new" "$OUT/src/${drv}_short.cpp" 2>/dev/null; then mv "$OUT/src/${drv}_short.cpp.new" "$OUT/src/${drv}_short.cpp"; else rm "$OUT/src/${drv}_short.cpp.new"; fi
  echo "$OUT/$drv $src $dir b200 $extra" >> "$TASKS"                                   # the unchanged driver
  echo "$OUT/${drv}_short_b200 $OUT/src/${drv}_short.cpp $dir b200 $extra" >> "$TASKS"
  echo "$REFOUT/${drv}_short_cpu $OUT/src/${drv}_short.cpp $dir cpu -" >> "$TASKS"
done
export -f build
export HERE ROOT REF OUT REFOUT BASE REFINC B200 LINK_B200
xargs -P "$JOBS" -L 1 bash -c 'e="$5"; [ "$e" = "-" ] && e=""; build "$1" "$2" "$3" "$4" "$e"' _ < "$TASKS"
