#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the REAL reference oracle + CPU baseline.
#
# Compiles the reference's own megdnn CPU backends (dnn/src/{common,naive,fallback,x86}) from the
# sources where they lie under /root/reference into oracle/_ref/libmegdnn_ref.so, plus our small
# driver (oracle/ref_driver.cpp) into oracle/_ref/libref_driver.so.  No reference source is copied
# into the repo; only binaries land in oracle/_ref/ (git-ignored, NOT gpurun-ignored).
#
#   Handle::make(comp, 2) -> naive::HandleImpl   (the ORACLE,      dnn/src/common/handle.cpp:78)
#   Handle::make(comp, 0) -> x86::HandleImpl     (the CPU BASELINE, dnn/src/common/handle.cpp:61)
#
# The reference's own build system (cmake + generated cutlass/cudnn/MKL) is NOT run: only the CPU
# translation units are compiled directly with g++ (recipe verified in SURVEY.md Appendix C).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/dnn/src" ]; then
    echo "build_ref.sh: $REF not present (GPU box?) -- using prebuilt oracle/_ref if any"; exit 0
fi
mkdir -p "$OUT/gen/megdnn" "$OUT/gen/src/common" "$OUT/genfiles" "$OUT/obj"

STAMP="$OUT/.built"
if [ -f "$OUT/libmegdnn_ref.so" ] && [ -f "$STAMP" ] && [ "$STAMP" -nt "$HERE/build_ref.sh" ]; then
    echo "build_ref.sh: libmegdnn_ref.so up to date"
else
    # 1. generated param headers (dnn/CMakeLists.txt:1-34)
    ( cd "$REF/dnn/scripts"
      python3 -W ignore gen_param_defs.py opr_param_defs.py "$OUT/gen/megdnn/opr_param_defs.h"
      python3 -W ignore gen_param_defs.py opr_param_defs.py /dev/null --write-cppjson "$OUT/gen/megdnn/opr_param_json.h"
      python3 -W ignore gen_param_defs.py --enumv opr_param_defs.py "$OUT/gen/src/common/opr_param_defs_enumv.cuh" )
    # 2. megbrain_build_config.h from the reference template: every "#cmakedefine01 X" -> 0|1
    ON="MGB_ENABLE_EXCEPTION MGB_HAVE_THREAD MEGDNN_ENABLE_RTTI MEGDNN_ENABLE_LOGGING MEGDNN_ENABLE_EXCEPTIONS MEGDNN_64_BIT MGB_ENABLE_LOGGING MGB_ENABLE_JSON MEGDNN_X86 MEGDNN_X86_64 MEGDNN_ENABLE_MULTI_THREADS"
    python3 - "$REF/src/megbrain_build_config.h.in" "$OUT/genfiles/megbrain_build_config.h" $ON <<'EOF'
import re, sys
src, dst, on = sys.argv[1], sys.argv[2], set(sys.argv[3:])
out = []
for line in open(src):
    m = re.match(r"\s*#cmakedefine01\s+(\w+)", line)
    if m:
        out.append("#define %s %d\n" % (m.group(1), 1 if m.group(1) in on else 0))
    elif re.match(r"\s*#cmakedefine\s+(\w+)", line):
        out.append("/* %s */\n" % line.strip())
    else:
        out.append(line)
open(dst, "w").writelines(out)
EOF
    INC="-I$REF/dnn -I$REF/dnn/include -I$OUT/gen -I$OUT/genfiles -I$REF/third_party/midout/src"
    CXXFLAGS="-std=c++17 -O2 -fPIC -w -msse4.2 -mfpmath=sse"
    cd "$REF/dnn/src"
    find common naive fallback x86 -name '*.cpp' | sort > "$OUT/obj/tus.txt"
    echo "build_ref.sh: compiling $(wc -l < "$OUT/obj/tus.txt") reference TUs with $JOBS jobs"
    compile_one() {
        f="$1"; o="$OUT/obj/$(echo "$f" | tr '/' '_').o"
        [ -f "$o" ] && [ "$o" -nt "$f" ] && return 0
        extra=""
        [ "$f" = "x86/utils.cpp" ] && extra="-include pmmintrin.h"
        g++ $CXXFLAGS $extra $INC -c "$f" -o "$o" || { echo "FAILED: $f"; return 1; }
    }
    export -f compile_one; export OUT INC CXXFLAGS
    xargs -P "$JOBS" -I{} bash -c 'compile_one {}' < "$OUT/obj/tus.txt"
    g++ -shared -o "$OUT/libmegdnn_ref.so" "$OUT"/obj/*.o -lpthread
    touch "$STAMP"
    echo "build_ref.sh: linked $OUT/libmegdnn_ref.so"
fi

# 3. our driver: a C-ABI shim over the reference's megdnn::ConvBiasForward (naive or x86 handle)
if [ ! -f "$OUT/libref_driver.so" ] || [ "$HERE/ref_driver.cpp" -nt "$OUT/libref_driver.so" ] || [ "$OUT/libmegdnn_ref.so" -nt "$OUT/libref_driver.so" ]; then
    g++ -std=c++17 -O2 -fPIC -w -shared -o "$OUT/libref_driver.so" "$HERE/ref_driver.cpp" \
        -I"$REF/dnn/include" -I"$OUT/gen" -I"$OUT/genfiles" -I"$HERE/../include" \
        -L"$OUT" -lmegdnn_ref -Wl,-rpath,'$ORIGIN' -lpthread
    echo "build_ref.sh: built $OUT/libref_driver.so"
fi
