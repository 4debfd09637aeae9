#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds the C++ host binding (megengine_b200/host: megdnn::b200::ConvBiasForwardImpl, a
# subclass of the reference's abstract megdnn::ConvBiasForward) against the reference's own public headers,
# and the drop-in test driver.  The binding is a megdnn plug-in: it needs megdnn's common layer (Handle,
# TensorLayout, ConvBiasForward::check_exec / deduce_layout), which in this repo exists only as the
# compiled reference under oracle/_ref -- so the built binding lives there too (oracle/_ref/
# libmegdnn_b200.so), NOT in the product package: nothing under megengine_b200/ links the reference.
# Needs /root/reference (headers) and oracle/_ref (generated param headers + libmegdnn_ref.so).
# On the GPU box the prebuilt binaries travel with the snapshot; this script then does nothing.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SRC="$ROOT/megengine_b200/host"
GEN="$ROOT/oracle/_ref"
LIB="$ROOT/megengine_b200/lib"
if [ ! -d "$REF/dnn/include" ] || [ ! -f "$GEN/libmegdnn_ref.so" ]; then
    echo "build_host.sh: reference headers / libmegdnn_ref.so not present -- using prebuilt binding if any"; exit 0
fi
CUDA=${CUDA_HOME:-/usr/local/cuda}
INC="-I$REF/dnn/include -I$GEN/gen -I$GEN/genfiles -I$ROOT/include -I$SRC -I$CUDA/include"
OUT="$GEN/libmegdnn_b200.so"
if [ ! -f "$OUT" ] || [ "$SRC/conv_bias_b200.cpp" -nt "$OUT" ] || [ "$SRC/nn_b200.cpp" -nt "$OUT" ] || \
   [ "$SRC/conv_bias_b200.h" -nt "$OUT" ] || [ "$SRC/nn_b200.h" -nt "$OUT" ] || [ "$ROOT/include/b200cb.h" -nt "$OUT" ] || \
   [ "$ROOT/include/b200cb_nn.h" -nt "$OUT" ]; then
    g++ -std=c++17 -O2 -fPIC -w -shared -o "$OUT" "$SRC/conv_bias_b200.cpp" "$SRC/nn_b200.cpp" $INC \
        -L"$LIB" -lb200cb -L"$GEN" -lmegdnn_ref -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../megengine_b200/lib'
    echo "build_host.sh: built $OUT"
fi
DRV="$GEN/dropin_driver"
if [ ! -f "$DRV" ] || [ "$HERE/dropin_driver.cpp" -nt "$DRV" ] || [ "$OUT" -nt "$DRV" ]; then
    g++ -std=c++17 -O2 -w -o "$DRV" "$HERE/dropin_driver.cpp" $INC \
        -L"$GEN" -lmegdnn_b200 -lmegdnn_ref -L"$LIB" -lb200cb -L"$CUDA/lib64" -lcudart -lpthread \
        -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../megengine_b200/lib' -Wl,-rpath,"$CUDA/lib64"
    echo "build_host.sh: built $DRV"
fi
