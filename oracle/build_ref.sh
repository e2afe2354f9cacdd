#!/usr/bin/env bash
# Build the reference's own daily/routing path into a CPU harness (TEST INFRASTRUCTURE).
#
#   oracle/build_ref.sh [NG]          (default NG = 67420, the reference's def.h:10 value)
#
# * compiles the reference sources WHERE THEY LIE under $WG_REF_SRC (default
#   /root/reference/source) - nothing is copied into the repository;
# * outputs only into oracle/_ref/ (git-ignored): objects + ref_harness_<NG>;
# * two shims that do not touch physics:
#     - oracle/stub/netcdf.h on the include path (netcdf-c is not installed; only ".nc" file
#       branches use it),
#     - json11.cpp is compiled from a sed-patched TEMPORARY copy (deleted afterwards):
#       GCC >= 7 rejects `nullptr < nullptr` in Value<NUL, std::nullptr_t>::less
#       (json11.cpp:157); the upstream json11 fix replaces std::nullptr_t by a NullStruct;
# * NG != 67420: the grid size is injected with `-include oracle/ref_def_override.h`, which
#   pre-defines def.h's include guard (def.h:1) and its constants with a different `ng`.
#   The 720x360 raster (geo.h:16) is unchanged, so small worlds are sub-masks of it.
# * flags: -O2 -fopenmp -ffp-contract=off, no -march=native (no FMA): the parity build.
set -euo pipefail
NG="${1:-67420}"
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${WG_REF_SRC:-/root/reference/source}"
OUT="$HERE/_ref"
OBJ="$OUT/obj_$NG"
if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: reference sources not found at $SRC (prebuilt oracle/_ref is used as is)" >&2
  exit 0
fi
mkdir -p "$OBJ"
CXX="${WGK_CXX:-/usr/bin/g++}"  # NB: $CXX in this image points at /opt/gcc, which ships no libgomp
FLAGS="-std=c++14 -O2 -fopenmp -ffp-contract=off -w -I$HERE/stub -I$SRC"
if [ "$NG" != "67420" ]; then
  FLAGS="$FLAGS -DWGK_REF_NG=$NG -include $HERE/ref_def_override.h"
fi
FILES="additionalOutputInputFile calcWaterTemp calib_basins calib_param calibration clcl climate
 climateYear configFile daily geo globals glacierYear gw_frac initializeWGHM integrateWGHM lai land
 option permafrost random rout_prepare routing s_max snowInElevationFile timestring
 upstream_stations wghmStateFile enKF2wghmState extractsub parameterJsonFile matrix"
pids=()
for f in $FILES; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$SRC/$f.cpp" -nt "$OBJ/$f.o" ]; then
    $CXX $FLAGS -c "$SRC/$f.cpp" -o "$OBJ/$f.o" &
    pids+=($!)
    if [ "${#pids[@]}" -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
if [ ! -f "$OBJ/json11.o" ]; then
  TMP="$(mktemp -d)"
  sed -e 's/std::nullptr_t/NullStruct/g' \
      -e '0,/^static void dump(NullStruct/s//struct NullStruct { bool operator==(NullStruct) const { return true; } bool operator<(NullStruct) const { return false; } };\nstatic void dump(NullStruct/' \
      -e 's/JsonNull() : Value(nullptr) {}/JsonNull() : Value({}) {}/' \
      -e 's/Json::Json(NullStruct) noexcept/Json::Json(std::nullptr_t) noexcept/' \
      "$SRC/json11.cpp" > "$TMP/json11.cpp"
  $CXX $FLAGS -c "$TMP/json11.cpp" -o "$OBJ/json11.o"
  rm -rf "$TMP"
fi
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$OBJ/ref_harness.o"
$CXX -fopenmp "$OBJ"/*.o -o "$OUT/ref_harness_$NG"
echo "built $OUT/ref_harness_$NG"
