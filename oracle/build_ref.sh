#!/usr/bin/env bash
# oracle/_ref: the UNMODIFIED reference (`MindTheGap` + gatb-core) compiled from the sources where they lie under
# /root/reference, by the recipe of SURVEY.md §0.5 / §8(c). Test infrastructure only: it validates the oracle
# restatement (tests/), produces the full-size golden fixtures (tests/golden/make_fullsize_fixtures.py) and is the
# CPU arm of `bench.py --impl reference`. Nothing of the product links or executes it.
#
# Outputs (all under oracle/_ref/, git-ignored, shipped to the GPU box by gpurun like our own .so files):
#   bin/MindTheGap        the reference CLI (stock `find` / `fill`)
#   bin/gatb-h5dump       HDF5 dataset dumper used to pin Bloom/debloom bytes
#   lib/libgatbcore.a, lib/libhdf5.a, include/   what tools/h5_handoff needs to link gatb-core's own HDF5 writer
# The cmake build tree lives in oracle/_ref/build (listed in .gpurunignore: only the products travel).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MTG_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF absent (GPU box?) - using prebuilt oracle/_ref if present" >&2
  exit 0
fi
mkdir -p "$OUT/build" "$OUT/bin" "$OUT/lib" "$OUT/include"
cd "$OUT/build"
if [ -x "$OUT/bin/MindTheGap" ] && [ -x "$OUT/bin/gatb-h5dump" ] && [ -f "$OUT/lib/libgatbcore.a" ] && [ -z "${FORCE:-}" ]; then
  echo "build_ref: reference binaries up to date"
else
cmake "$REF" -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 > cmake.log 2>&1
make -j"$JOBS" MindTheGap > make.log 2>&1
make -j"$JOBS" gatb-h5dump >> make.log 2>&1
cp -f bin/MindTheGap "$OUT/bin/"
cp -f "$(find . -name gatb-h5dump -type f -perm -u+x | head -1)" "$OUT/bin/"
cp -f "$(find . -name libgatbcore.a | head -1)" "$OUT/lib/"
cp -f "$(find . -name libhdf5.a | head -1)" "$OUT/lib/"
# generated headers (config.hpp, hdf5 public headers) for linking a dumper / the .h5 hand-off tool
find . -name config.hpp -path '*gatb*' -exec sh -c 'mkdir -p "$0/include/gatb/system/api" && cp -f "$1" "$0/include/gatb/system/api/"' "$OUT" {} \;
H5INC="$(dirname "$(find . -name H5pubconf.h | head -1)")"
[ -n "$H5INC" ] && mkdir -p "$OUT/include/hdf5" && cp -f "$H5INC"/*.h "$OUT/include/hdf5/" 2>/dev/null || true
fi
# gatb-linked helpers (rebuilt when their source is newer)
# gatb-linked helper that reads dsk/solid back from a reference .h5 (oracle/ref_tools/h5solid.cpp)
INC="-I ext/gatb-core/include -I ext/gatb-core/include/Release -I$REF/thirdparty/gatb-core/gatb-core/src -I$REF/thirdparty/gatb-core/gatb-core/thirdparty -I ext/gatb-core/thirdparty/hdf5/src -I$REF/thirdparty/gatb-core/gatb-core/thirdparty/hdf5/src"
for tool in "$HERE"/ref_tools/*.cpp; do
  exe="$OUT/bin/$(basename "$tool" .cpp)"
  [ -x "$exe" ] && [ "$exe" -nt "$tool" ] && continue
  g++ -O2 -std=c++11 -DNDEBUG -D_FILE_OFFSET_BITS=64 -w $INC "$tool" -o "$OUT/bin/$(basename "$tool" .cpp)" \
      ext/gatb-core/lib/Release/libgatbcore.a ext/gatb-core/lib/Release/libhdf5.a -ldl -lpthread -lz -lm
done
# the product's .h5 hand-off host tool (mindthegap_b200/csrc/h5_handoff.cpp) links the reference's gatb-core / HDF5 for the file I/O
H5TOOL="$HERE/../mindthegap_b200/csrc/h5_handoff.cpp"
H5EXE="$HERE/../mindthegap_b200/_build/mtg_h5"
if [ -f "$H5TOOL" ] && { [ ! -x "$H5EXE" ] || [ "$H5TOOL" -nt "$H5EXE" ]; }; then
  mkdir -p "$(dirname "$H5EXE")"
  g++ -O2 -std=c++11 -DNDEBUG -D_FILE_OFFSET_BITS=64 -w $INC "$H5TOOL" -o "$H5EXE" \
      ext/gatb-core/lib/Release/libgatbcore.a ext/gatb-core/lib/Release/libhdf5.a -ldl -lpthread -lz -lm
fi
# the reference-side binding, COMPILED: the reference's own src/ with three statements of Finder.cpp replaced by calls into the C ABI
# (integration/finder_shim.hpp, integration/make_shim.py) -> oracle/_ref/bin/MindTheGap_mtg, linked against libmtg_b200.so
SHIM_EXE="$OUT/bin/MindTheGap_mtg"
LIBMTG="$HERE/../mindthegap_b200/_build/libmtg_b200.so"
if [ -f "$LIBMTG" ] && { [ ! -x "$SHIM_EXE" ] || [ "$HERE/../integration/finder_shim.hpp" -nt "$SHIM_EXE" ] || [ "$HERE/../integration/make_shim.py" -nt "$SHIM_EXE" ] || [ "$HERE/../include/mtg_b200.h" -nt "$SHIM_EXE" ]; }; then
  python "$HERE/../integration/make_shim.py" "$REF/src" "$OUT/shim_src" > /dev/null
  g++ -O2 -std=c++11 -DNDEBUG -D_FILE_OFFSET_BITS=64 -w $INC -I"$OUT/shim_src" -I"$HERE/../include" "$OUT"/shim_src/*.cpp -o "$SHIM_EXE" \
      ext/gatb-core/lib/Release/libgatbcore.a ext/gatb-core/lib/Release/libhdf5.a -L"$(dirname "$LIBMTG")" -lmtg_b200 \
      -Wl,-rpath,'$ORIGIN/../../../mindthegap_b200/_build' -ldl -lpthread -lz -lm
fi
echo "build_ref: done -> $OUT/bin/MindTheGap"
