"""ctypes binding of the C-ABI in include/b200cb.h (libb200cb.so).

The library is built in-tree by ``__graft_entry__.build()`` (``megengine_b200/csrc/Makefile``).
There is NO fallback: if the shared object is missing, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200CB_LIB") or os.path.join(_HERE, "lib", "libb200cb.so")

# enums (values == megdnn's, see include/b200cb.h)
FORMAT_NCHW, FORMAT_NHWC, FORMAT_NCHW4 = 0, 1, 3
NL_IDENTITY, NL_RELU, NL_SIGMOID, NL_H_SWISH = 0, 1, 2, 3
MODE_CROSS_CORRELATION, MODE_CONVOLUTION = 0, 1
SPARSE_DENSE, SPARSE_GROUP = 0, 1
COMPUTE_DEFAULT, COMPUTE_FLOAT32 = 0, 1
BIAS_NONE, BIAS_CHANNEL, BIAS_FULL = 0, 1, 2
DT_F32, DT_F16, DT_BF16, DT_QS8, DT_QS32, DT_NONE = 0, 1, 2, 3, 4, 255

RELAYOUT_NCHW_NHWC, RELAYOUT_NHWC_NCHW, RELAYOUT_NCHW_NCHW4, RELAYOUT_NCHW4_NCHW = 0, 1, 2, 3
RELAYOUT_NCHW4_NHWC, RELAYOUT_NHWC_NCHW4, RELAYOUT_NHWC_PAD_C = 4, 5, 6

DT_SIZE = {DT_F32: 4, DT_F16: 2, DT_BF16: 2, DT_QS8: 1, DT_QS32: 4}
DT_NAME = {DT_F32: "Float32", DT_F16: "Float16", DT_BF16: "BFloat16", DT_QS8: "QuantizedS8",
           DT_QS32: "QuantizedS32", DT_NONE: "none"}


class Desc(ctypes.Structure):
    """b200cb_desc"""
    _fields_ = [
        ("format", ctypes.c_uint32), ("sparse", ctypes.c_uint32), ("mode", ctypes.c_uint32),
        ("nonline_mode", ctypes.c_uint32), ("compute_mode", ctypes.c_uint32),
        ("bias_mode", ctypes.c_uint32), ("has_z", ctypes.c_uint32),
        ("src_dtype", ctypes.c_uint32), ("filter_dtype", ctypes.c_uint32),
        ("bias_dtype", ctypes.c_uint32), ("z_dtype", ctypes.c_uint32), ("dst_dtype", ctypes.c_uint32),
        ("src_scale", ctypes.c_float), ("filter_scale", ctypes.c_float),
        ("bias_scale", ctypes.c_float), ("z_scale", ctypes.c_float), ("dst_scale", ctypes.c_float),
        ("n", ctypes.c_uint32), ("ic", ctypes.c_uint32), ("ih", ctypes.c_uint32), ("iw", ctypes.c_uint32),
        ("oc", ctypes.c_uint32), ("oh", ctypes.c_uint32), ("ow", ctypes.c_uint32),
        ("fh", ctypes.c_uint32), ("fw", ctypes.c_uint32), ("group", ctypes.c_uint32),
        ("pad_h", ctypes.c_uint32), ("pad_w", ctypes.c_uint32),
        ("stride_h", ctypes.c_uint32), ("stride_w", ctypes.c_uint32),
        ("dilate_h", ctypes.c_uint32), ("dilate_w", ctypes.c_uint32),
        ("src_batch_stride", ctypes.c_uint64), ("z_batch_stride", ctypes.c_uint64),
        ("dst_batch_stride", ctypes.c_uint64),
    ]

    def copy(self):
        d = Desc()
        ctypes.memmove(ctypes.byref(d), ctypes.byref(self), ctypes.sizeof(Desc))
        return d


class Plan(ctypes.Structure):
    """b200cb_plan"""
    _fields_ = [
        ("magic", ctypes.c_uint32), ("algo", ctypes.c_int32), ("kernel", ctypes.c_uint32),
        ("block_n", ctypes.c_uint32), ("stages", ctypes.c_uint32), ("grid", ctypes.c_uint32),
        ("smem_bytes", ctypes.c_uint32), ("a_mode", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32 * 8),
        ("workspace_bytes", ctypes.c_uint64), ("flops", ctypes.c_uint64), ("bytes", ctypes.c_uint64),
        ("desc", Desc),
    ]


# every symbol include/b200cb.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
SYMBOLS = [
    ("b200cb_algo_count", ctypes.c_int, []),
    ("b200cb_algo_name", ctypes.c_char_p, [ctypes.c_int]),
    ("b200cb_query", ctypes.c_int, [ctypes.POINTER(Desc), ctypes.POINTER(Plan)]),
    ("b200cb_query_algo", ctypes.c_int, [ctypes.POINTER(Desc), ctypes.c_int, ctypes.POINTER(Plan)]),
    ("b200cb_launch", ctypes.c_int,
     [ctypes.POINTER(Plan), _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    ("b200cb_preprocessed_filter_bytes", ctypes.c_size_t, [ctypes.POINTER(Plan)]),
    ("b200cb_preprocess_filter", ctypes.c_int, [ctypes.POINTER(Plan), _P, _P, _P]),
    ("b200cb_plan_use_preprocessed_filter", ctypes.c_int, [ctypes.POINTER(Plan)]),
    ("b200cb_plan_num_launches", ctypes.c_int, [ctypes.POINTER(Plan)]),
    ("b200cb_selftest", ctypes.c_int, [ctypes.c_int, _P, ctypes.POINTER(ctypes.c_uint64)]),
    ("b200cb_host_scratch_bytes", ctypes.c_size_t, [ctypes.POINTER(Plan)]),
    ("b200cb_exec_host", ctypes.c_int,
     [ctypes.POINTER(Plan), _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    ("b200cb_relayout", ctypes.c_int,
     [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
      ctypes.c_uint32, _P, _P, _P]),
    ("b200cb_strerror", ctypes.c_char_p, [ctypes.c_int]),
    ("b200cb_last_error_detail", ctypes.c_char_p, []),
]



class PoolDesc(ctypes.Structure):
    """b200cb_pool_desc (include/b200cb_nn.h)"""
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "format", "mode", "dtype", "n", "c", "ih", "iw", "oh", "ow", "window_h", "window_w", "stride_h", "stride_w",
        "pad_h", "pad_w")]


POOL_MAX, POOL_AVERAGE, POOL_AVERAGE_COUNT_EXCLUDE_PADDING = 0, 1, 2
ELEM_RELU, ELEM_ADD, ELEM_MAX, ELEM_MUL, ELEM_FUSE_ADD_RELU, ELEM_H_SWISH = 0, 16, 18, 21, 37, 49

# every symbol include/b200cb_nn.h declares
_U32, _U64, _F = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float
NN_SYMBOLS = [
    ("b200cb_pooling", ctypes.c_int, [ctypes.POINTER(PoolDesc), _P, _P, _P]),
    ("b200cb_elemwise", ctypes.c_int, [_U32, _U32, _U64, _P, _P, _P, _P]),
    ("b200cb_elemwise_q", ctypes.c_int, [_U32, _U64, _P, _F, _P, _F, _P, _F, _P]),
    ("b200cb_typecvt", ctypes.c_int, [_U32, _F, _U32, _F, _U64, _P, _P, _P]),
    ("b200cb_convolution_query", ctypes.c_int, [ctypes.POINTER(Desc), ctypes.POINTER(Plan)]),
    ("b200cb_convolution_launch", ctypes.c_int, [ctypes.POINTER(Plan), _P, _P, _P, _P, ctypes.c_size_t, _P]),
    ("b200cb_deconv_query", ctypes.c_int, [ctypes.POINTER(Desc), ctypes.POINTER(Plan)]),
    ("b200cb_deconv_launch", ctypes.c_int, [ctypes.POINTER(Plan), _P, _P, _P, _P, ctypes.c_size_t, _P]),
    ("b200cb_batch_convbias_query", ctypes.c_int, [ctypes.POINTER(Desc), ctypes.POINTER(Plan)]),
    ("b200cb_batch_convbias_launch", ctypes.c_int,
     [ctypes.POINTER(Plan), _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
]

_lib = None


def load():
    """dlopen libb200cb.so and type every entry point. Fails loudly when the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb200cb.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "-- megengine_b200 has no CPU/PyTorch fallback path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS + NN_SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200CBError(RuntimeError):
    """Mirror of megdnn's MegDNNError for this path: C-ABI codes < 0 or unexpected > 0."""

    def __init__(self, code, where=""):
        lib = load()
        msg = lib.b200cb_strerror(code).decode()
        detail = lib.b200cb_last_error_detail().decode()
        super().__init__("%s: [%d] %s%s" % (where, code, msg, (" (" + detail + ")") if detail and code < 0 else ""))
        self.code = code


def strerror(code):
    return load().b200cb_strerror(code).decode()
