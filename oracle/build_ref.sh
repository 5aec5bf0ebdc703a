#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the reference's own QUERY_MODE=HOST implementation of the
# AQL batch-execution hot path (query/*.cu compiled as plain C++ with Thrust's CPP backend)
# into oracle/_ref/{libmem_ref.so,libalgorithm.so}.
#
# - Sources are read where they lie under $ARESDB_REFERENCE (default /root/reference).
# - They need a small mechanical patch for Thrust 2.8 API drift; the patch
#   is applied with sed while streaming each file into a throw-away mktemp build dir, so
#   no reference source is ever written into this repository.  Only the two .so files and
#   a build stamp land in oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# - We do NOT run the reference's CMake build; this is g++ on 14 files + gcc on malloc.c.
#
# Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
# may load what this script produces.
set -euo pipefail

REF="${ARESDB_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
OPT="${ARESDB_REF_OPT:--O3}"      # reference uses -O3 for host objects (CMakeLists.txt:165)
JOBS="${ARESDB_REF_JOBS:-8}"

if [[ ! -d "$REF/query" ]]; then
  echo "build_ref: $REF/query not present; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
if [[ -f "$OUT/libalgorithm.so" && -f "$OUT/libmem_ref.so" && "${1:-}" != "--force" ]]; then
  echo "build_ref: oracle/_ref already built (use --force to rebuild)"
  exit 0
fi

CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
TMP="$(mktemp -d /tmp/aresdb_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/query" "$TMP/cgoutils/memory" "$OUT"

# --- stream sources through the Thrust-2.8 patch -----------------------------------------
patch_stream() {
  # 1. thrust::make_tuple no longer accepts explicit (lvalue) template arguments
  # 2. thrust::tuple is cuda::std::tuple now: no ::head_type
  sed -E \
    -e 's/thrust::make_tuple<[^()]*>\(/thrust::make_tuple(/g' \
    -e 's/typename ([A-Za-z0-9_]+)::value_type::head_type/typename thrust::tuple_element<0, typename \1::value_type>::type/g'
}
for f in "$REF"/query/*.hpp "$REF"/query/*.h "$REF"/query/*.cu; do
  b="$(basename "$f")"
  case "$b" in *_unittest.cu|unittest_utils.hpp|geo_intersects.cu) continue;; esac
  patch_stream < "$f" > "$TMP/query/$b"
done
cp "$REF"/cgoutils/*.h "$TMP/cgoutils/"
cp "$REF"/cgoutils/memory/malloc.c "$TMP/cgoutils/memory/"

# multi-line make_tuple<\n UUIDT,\n bool>( in hash_lookup.cu
perl -0pi -e 's/thrust::make_tuple<\s*UUIDT,\s*bool>\(/thrust::make_tuple(/g' "$TMP/query/hash_lookup.cu"
# 3. GeoPoint/UUID BinaryFunctor default arm returns `false` where a tuple is expected
perl -0pi -e 's/(\/\/ should not came here, GeoPoint only support equal function\s*\n\s*)return false;/$1return result_type();/g' "$TMP/query/functor.hpp"
# 4. algorithm headers are no longer transitively included
perl -0pi -e 's/(#include <thrust\/tuple.h>\n)/$1#include <thrust\/sort.h>\n#include <thrust\/remove.h>\n#include <thrust\/sequence.h>\n#include <thrust\/transform.h>\n#include <thrust\/copy.h>\n#include <thrust\/reduce.h>\n#include <thrust\/scan.h>\n#include <thrust\/scatter.h>\n#include <thrust\/merge.h>\n#include <thrust\/for_each.h>\n#include <thrust\/functional.h>\n#include <thrust\/iterator\/discard_iterator.h>\n#include <thrust\/iterator\/counting_iterator.h>\n#include <thrust\/iterator\/transform_iterator.h>\n/' "$TMP/query/utils.hpp"
# 5. unqualified min(uint32_t,int)
sed -i -E 's/int outputLen = min\(totalCount,/int outputLen = std::min<int>(totalCount,/' "$TMP/query/sort_reduce.cu"

cat > "$TMP/shim.h" <<'EOF'
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
using std::min;
using std::max;
EOF

cd "$TMP"
# Distinct file names / sonames: a process that also loads the B200 engine (tests) must not let the
# dynamic loader alias this malloc-backed libmem with the engine's CUDA libmem.
gcc $OPT -fPIC -shared cgoutils/memory/malloc.c -Wl,-soname,libmem_ref.so -o "$OUT/libmem_ref.so"

SRCS=(sort_reduce filter transform dimension_transform measure_transform scratch_space_transform
      hash_reduction hll hash_lookup functor utils iterator algorithm memory)
printf '%s\n' "${SRCS[@]}" | xargs -P "$JOBS" -I{} \
  g++ -x c++ -std=c++17 $OPT -fPIC -w -include shim.h \
      -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP -DSUPPORT_HASH_REDUCTION=1 \
      -I. -I"$CUDA_INC" -c query/{}.cu -o {}.o

g++ -shared -Wl,-soname,libalgorithm_ref.so -o "$OUT/libalgorithm.so" ./*.o -L"$OUT" -lmem_ref \
    -L"${CUDA_HOME:-/usr/local/cuda}/lib64" -lcudart -Wl,-rpath,'$ORIGIN'
echo "built from $REF with g++ $(g++ -dumpversion) $OPT on $(date -u +%FT%TZ)" > "$OUT/BUILD_STAMP"
echo "build_ref: wrote $OUT/libmem_ref.so $OUT/libalgorithm.so"
