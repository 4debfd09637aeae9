#!/usr/bin/env bash
# Builds the C++ host binding (megdnn::b200::ConvBiasForwardImpl, a subclass of the reference's abstract
# megdnn::ConvBiasForward) against the reference's own public headers, and the drop-in test driver.
# Needs /root/reference (headers) and oracle/_ref (generated param headers + libmegdnn_ref.so = the
# reference's megdnn common layer: Handle, TensorLayout, ConvBiasForward::check_exec / deduce_layout).
# On the GPU box the prebuilt binaries travel with the snapshot; this script then does nothing.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
GEN="$ROOT/oracle/_ref"
LIB="$ROOT/megengine_b200/lib"
if [ ! -d "$REF/dnn/include" ] || [ ! -f "$GEN/libmegdnn_ref.so" ]; then
    echo "build_host.sh: reference headers / libmegdnn_ref.so not present -- using prebuilt binding if any"; exit 0
fi
CUDA=${CUDA_HOME:-/usr/local/cuda}
INC="-I$REF/dnn/include -I$GEN/gen -I$GEN/genfiles -I$ROOT/include -I$HERE -I$CUDA/include"
if [ ! -f "$LIB/libmegdnn_b200.so" ] || [ "$HERE/conv_bias_b200.cpp" -nt "$LIB/libmegdnn_b200.so" ] || \
   [ "$HERE/conv_bias_b200.h" -nt "$LIB/libmegdnn_b200.so" ] || [ "$ROOT/include/b200cb.h" -nt "$LIB/libmegdnn_b200.so" ]; then
    g++ -std=c++17 -O2 -fPIC -w -shared -o "$LIB/libmegdnn_b200.so" "$HERE/conv_bias_b200.cpp" $INC \
        -L"$LIB" -lb200cb -L"$GEN" -lmegdnn_ref -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../oracle/_ref'
    echo "build_host.sh: built $LIB/libmegdnn_b200.so"
fi
DRV="$GEN/dropin_driver"
if [ ! -f "$DRV" ] || [ "$ROOT/tests/cpp/dropin_driver.cpp" -nt "$DRV" ] || [ "$LIB/libmegdnn_b200.so" -nt "$DRV" ]; then
    g++ -std=c++17 -O2 -w -o "$DRV" "$ROOT/tests/cpp/dropin_driver.cpp" $INC \
        -L"$LIB" -lmegdnn_b200 -lb200cb -L"$GEN" -lmegdnn_ref -L"$CUDA/lib64" -lcudart -lpthread \
        -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../megengine_b200/lib' -Wl,-rpath,"$CUDA/lib64"
    echo "build_host.sh: built $DRV"
fi
