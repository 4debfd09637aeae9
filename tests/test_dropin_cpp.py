"""The C++ host binding (megengine_b200/host: a subclass of the reference's abstract
megdnn::ConvBiasForward built against the reference's own headers) driven like the reference's
Checker drives a backend: expected values from the reference's naive operator, every algorithm
forced through execution_policy(), QuantizedS8 bit-exact, bf16 / fp16 at the reference tests' eps.
The binaries are built where /root/reference exists (tests/cpp/build_host.sh, into oracle/_ref: the
binding links the reference's megdnn common layer, so it is not part of the product package) and travel
to the GPU box with the snapshot."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "oracle", "_ref", "libmegdnn_b200.so")
DRIVER = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")


def test_host_binding_exports_factory():
    if not os.path.exists(HOST_LIB):
        pytest.skip("C++ host binding not built (needs the reference headers: build_host.sh)")
    lib = ctypes.CDLL(HOST_LIB)
    # megdnn::b200::create_conv_bias_forward(megdnn::Handle*, CUstream_st*)
    assert hasattr(lib, "_ZN6megdnn4b20024create_conv_bias_forwardEPNS_6HandleEP11CUstream_st")


@pytest.mark.gpu
def test_dropin_driver_all_cases_pass():
    if not os.path.exists(DRIVER):
        pytest.skip("drop-in driver not built")
    r = subprocess.run([DRIVER], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "PASS nchw4_qs8_3x3_z_relu" in r.stdout and "PASS nhwc_bf16_3x3_z_relu" in r.stdout
    # the operators either side of ConvBias through the reference's abstract TypeCvt / Elemwise / Pooling interfaces
    for name in ("typecvt_qs32_qs8", "elemwise_add_relu_f16", "pool_max3s2_bf16_nhwc", "pool_avg7_qs8_nhwc",
                 "pool_max3s2_qs8_nchw4"):
        assert "PASS " + name in r.stdout, name
    assert "0 failed" in r.stdout
