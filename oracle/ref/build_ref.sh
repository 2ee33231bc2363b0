#!/bin/bash
# Builds the reference's own host framework IN PLACE from /root/reference (read-only) into oracle/_ref/ (git-ignored,
# travels to the GPU box).  TEST INFRASTRUCTURE: nothing under cuda_sdr_b200/ uses these artefacts.
#
# The reference's CMake build cannot run here (clang, GNU Radio, libhackrf, FFmpeg, remez, gsdr are all absent), so
# this recipe compiles the hot path's ~55 host sources directly with g++.  No reference source is copied into the
# repo; the only transformations happen in a scratch directory under /tmp at build time:
#   * include/gpusdrpipeline/Result.h: `#pragma pack` sits between `template<...>` and `struct`, which only clang
#     accepts -> the pragma lines are moved outside the templates;
#   * src/Factories.cpp: the three #include lines of factories that need libhackrf / FFmpeg / remez are replaced
#     by oracle/ref/unavailable_factories.h (same class names, every create*() returns Status_NotFound).
# `gsdr` (the kernel library, un-vendored) is supplied at LINK time, two ways: oracle/ref/gsdr_naive.cu (a
# straightforward restatement = baseline B1) and this repo's libb200sdr.so (the product, as a gsdr drop-in).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF=${REFERENCE_ROOT:-/root/reference}
OUT="$ROOT/oracle/_ref"
CXX=/usr/bin/g++
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CUDA_HOME=${CUDA_HOME:-/usr/local/cuda}
JOBS=${JOBS:-8}

if [ ! -d "$REF/src/filters" ]; then
  echo "build_ref.sh: $REF is not mounted; keeping whatever is already in $OUT" >&2
  exit 0
fi
JSON_INC=$(python3 - <<'EOF'
import os, sys
for p in sys.path:
    c = os.path.join(p, "include", "cudnn_frontend", "thirdparty")
    if os.path.exists(os.path.join(c, "nlohmann", "json.hpp")):
        print(c); break
EOF
)
[ -n "$JSON_INC" ] || { echo "nlohmann/json.hpp not found" >&2; exit 1; }

SCRATCH=$(mktemp -d /tmp/refbuild.XXXXXX)
trap 'rm -rf "$SCRATCH"' EXIT
mkdir -p "$OUT" "$SCRATCH/obj" "$SCRATCH/include"

# ---- scratch include tree: symlinks to every reference header, a fixed Result.h, the generated export header ----
(cd "$REF/include" && find gpusdrpipeline -type d) | while read -r d; do mkdir -p "$SCRATCH/include/$d"; done
(cd "$REF/include" && find gpusdrpipeline -type f) | while read -r f; do ln -s "$REF/include/$f" "$SCRATCH/include/$f"; done
rm "$SCRATCH/include/gpusdrpipeline/Result.h"
python3 - "$REF/include/gpusdrpipeline/Result.h" "$SCRATCH/include/gpusdrpipeline/Result.h" <<'EOF'
import re, sys
src = open(sys.argv[1]).read()
# move each `#pragma pack(push, 8)` that directly follows a `template <...>` line to just before that line
fixed, n = re.subn(r"(template <[^\n]*>\n)(#pragma pack\(push, 8\)\n)", r"\2\1", src)
assert n >= 2, "Result.h no longer matches the expected shape"
open(sys.argv[2], "w").write(fixed)
EOF
cat > "$SCRATCH/include/gpusdrpipeline/gpusdrpipeline_export.h" <<'EOF'
#pragma once
#define GS_PUBLIC __attribute__((visibility("default")))
#define GS_PRIVATE __attribute__((visibility("hidden")))
EOF
sed -e 's|#include "filters/factories/AacFileWriterFactory.h"|#include "unavailable_factories.h"|' \
    -e '/#include "filters\/factories\/HackrfSourceFactory.h"/d' \
    -e '/#include "filters\/factories\/RfToPcmAudioFactory.h"/d' \
    "$REF/src/Factories.cpp" > "$SCRATCH/Factories.cpp"

INC=(-I"$SCRATCH/include" -I"$SCRATCH/include/gpusdrpipeline" -I"$REF/src" -I"$HERE" -I"$ROOT/include" -I"$JSON_INC"
     -I"$CUDA_HOME/include")
FLAGS=(-std=c++20 -O2 -fPIC -w -DNDEBUG "${INC[@]}")

SOURCES=("$SCRATCH/Factories.cpp" "$REF"/src/GSLog.cpp "$REF"/src/FileLogger.cpp "$REF"/src/ParseJson.cpp
         "$REF"/src/buffers/*.cpp "$REF"/src/commandqueue/*.cpp "$REF"/src/driver/*.cpp
         "$REF"/src/util/util.cpp "$REF"/src/util/CudaUtil.cpp)
for f in BaseFilter BaseSink BaseSource Fir Int8ToFloat CosineSource ComplexCosineSource Multiply QuadAmDemod QuadFmDemod \
         Magnitude AddConst AddConstToVectorLength CudaMemcpyFilter PortRemappingSink PortRemappingSource \
         ReadByteCountMonitor FilterFactories FileReader; do
  SOURCES+=("$REF/src/filters/$f.cpp")
done

echo "build_ref: compiling ${#SOURCES[@]} reference sources in place"
printf '%s\n' "${SOURCES[@]}" | xargs -P "$JOBS" -I{} bash -c '
  src="$1"; shift
  obj="'"$SCRATCH"'/obj/$(echo "$src" | md5sum | cut -c1-8)_$(basename "$src").o"
  '"$CXX"' "$@" -c "$src" -o "$obj"' _ {} "${FLAGS[@]}"

"$CXX" -shared -o "$OUT/libgpusdrpipeline_ref.so" "$SCRATCH"/obj/*.o -L"$CUDA_HOME/lib64" -lcudart -lpthread -ldl
echo "build_ref: nvcc gsdr_naive.cu"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC,-fvisibility=hidden -ccbin "$CXX" \
  -o "$OUT/libgsdr_naive.so" "$HERE/gsdr_naive.cu"

# ---- executables: the reference's own tests (unmodified, gtest shim) and the chain driver, per gsdr library ----
B200LIB="$ROOT/cuda_sdr_b200"
link() {  # link <output> <gsdr lib dir> <gsdr lib name> <sources...>
  local out=$1 dir=$2 lib=$3; shift 3
  "$CXX" "${FLAGS[@]}" -o "$out" "$@" -L"$OUT" -lgpusdrpipeline_ref -L"$dir" -l"$lib" -L"$CUDA_HOME/lib64" -lcudart \
    -lpthread -ldl '-Wl,-rpath,$ORIGIN' '-Wl,-rpath,$ORIGIN/../../cuda_sdr_b200' -Wl,-rpath,"$CUDA_HOME/lib64"
}
TESTS=("$REF/tests/FirTests.cpp" "$REF/tests/CosineSourceTests.cpp" "$HERE/gtest_main.cpp")
link "$OUT/ref_tests_naive" "$OUT" gsdr_naive "${TESTS[@]}"
link "$OUT/ref_chain_naive" "$OUT" gsdr_naive "$HERE/ref_chain.cpp"
# the graph driver (tests/graph_probe.cpp: SteppingDriver + nested FilterDrivers, am_test.cpp style) on the reference's own
# host framework: what tests/test_gpu_graph.py compares this repo's drivers with
link "$OUT/graph_probe_naive" "$OUT" gsdr_naive "$ROOT/tests/graph_probe.cpp"
if [ -f "$B200LIB/libb200sdr.so" ]; then
  link "$OUT/ref_tests_b200" "$B200LIB" b200sdr "${TESTS[@]}"
  link "$OUT/ref_chain_b200" "$B200LIB" b200sdr "$HERE/ref_chain.cpp"
fi
# ---- THIS repo's host framework (libgpusdrpipeline.so) behind the same sources --------------------------------
#   *_ours     : compiled against the REFERENCE's headers, linked against our library  -> binary (vtable/ABI) compatibility
#   *_ours_hdr : compiled against OUR re-authored headers                               -> source compatibility
if [ -f "$B200LIB/libgpusdrpipeline.so" ]; then
  OURS=(-L"$B200LIB" -lgpusdrpipeline -lb200sdr -L"$CUDA_HOME/lib64" -lcudart -lpthread -ldl '-Wl,-rpath,$ORIGIN/../../cuda_sdr_b200'
        -Wl,-rpath,"$CUDA_HOME/lib64")
  "$CXX" "${FLAGS[@]}" -o "$OUT/ref_tests_ours" "${TESTS[@]}" "${OURS[@]}"
  "$CXX" "${FLAGS[@]}" -o "$OUT/ref_chain_ours" "$HERE/ref_chain.cpp" "${OURS[@]}"
  "$CXX" "${FLAGS[@]}" -o "$OUT/graph_probe_ours" "$ROOT/tests/graph_probe.cpp" "${OURS[@]}"
  HDR=(-std=c++20 -O2 -fPIC -w -DNDEBUG -I"$ROOT/include" -I"$HERE" -I"$CUDA_HOME/include")
  "$CXX" "${HDR[@]}" -o "$OUT/ref_tests_ours_hdr" "${TESTS[@]}" "${OURS[@]}"
  "$CXX" "${HDR[@]}" -DREF_CHAIN_HAS_FUSED -o "$OUT/ref_chain_ours_hdr" "$HERE/ref_chain.cpp" "${OURS[@]}"
fi
ls -la "$OUT"
echo "build_ref: done"
