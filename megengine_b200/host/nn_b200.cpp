// See nn_b200.h.  Reference counterparts: dnn/src/cuda/type_cvt/opr_impl.cpp, dnn/src/cuda/elemwise/opr_impl.cpp,
// dnn/src/cuda/pooling/opr_impl.cpp + algo.h (operator / AlgoBase / AlgoPack).
#include "nn_b200.h"

#include <cstring>
#include <stdexcept>

namespace megdnn {
namespace b200 {
namespace {

uint32_t dtype_code(const DType& dt) {
    switch (dt.enumv()) {
        case DTypeEnum::Float32: return B200CB_DT_F32;
        case DTypeEnum::Float16: return B200CB_DT_F16;
        case DTypeEnum::BFloat16: return B200CB_DT_BF16;
        case DTypeEnum::QuantizedS8: return B200CB_DT_QS8;
        case DTypeEnum::QuantizedS32: return B200CB_DT_QS32;
        default: return B200CB_DT_NONE;
    }
}
float dtype_scale(const DType& dt) {
    if (dt.enumv() == DTypeEnum::QuantizedS8)
        return dt.param<dtype::QuantizedS8>().scale;
    if (dt.enumv() == DTypeEnum::QuantizedS32)
        return dt.param<dtype::QuantizedS32>().scale;
    return 0.f;
}
[[noreturn]] void b200_throw(const std::string& msg) {
    megdnn::ErrorHandler::on_megdnn_error(msg);   // what megdnn_throw ends in (basic_types.h:21-58)
    throw std::runtime_error(msg);                // not reached
}
void check_rc(const char* what, int rc) {
    if (rc != B200CB_OK)
        b200_throw(std::string("b200cb ") + what + ": " + b200cb_strerror(rc) + " (" + b200cb_last_error_detail() + ")");
}

// the single algorithm of the pooling operator (an AlgoPack of one)
class AlgoPoolingB200 final : public detail::Algorithm {
public:
    AlgoPoolingB200() { m_handle_type = Handle::HandleType::CUDA; }
    Attribute attribute() const override { return Attribute::REPRODUCIBLE; }
    const char* name() const override { return "B200_POOLING_DIRECT"; }
    uint32_t type() const override { return 0xB201u; }
    std::string param() const override { return {}; }
};
AlgoPoolingB200& pooling_algo() {
    static AlgoPoolingB200 a;
    return a;
}

}  // namespace

// ---------------------------------------------------------------- TypeCvt
void TypeCvtImpl::exec(_megdnn_tensor_in src, _megdnn_tensor_out dst) {
    check_exec(src.layout, dst.layout);   // the reference's own checks (dnn/src/common/type_cvt.cpp)
    if (!src.layout.is_contiguous() || !dst.layout.is_contiguous())
        b200_throw("b200 TypeCvt: non-contiguous tensors are not supported");
    check_rc("typecvt", b200cb_typecvt(
                                dtype_code(src.layout.dtype), dtype_scale(src.layout.dtype), dtype_code(dst.layout.dtype),
                                dtype_scale(dst.layout.dtype), src.layout.total_nr_elems(), src.raw_ptr(), dst.raw_ptr(),
                                m_stream));
}

// ---------------------------------------------------------------- Elemwise
void ElemwiseForwardImpl::exec(_megdnn_in const TensorNDArray& src, _megdnn_tensor_out dst) {
    {   // ElemwiseForward's own protected check (arity, dtypes, broadcastability: dnn/src/common/elemwise/opr_impl.cpp)
        std::vector<TensorLayout> lays;
        for (auto& s : src)
            lays.push_back(s.layout);
        TensorLayoutPtrArray ptrs;
        for (auto& l : lays)
            ptrs.push_back(&l);
        check_layout_and_broadcast(ptrs, dst.layout);
    }
    for (auto& s : src)
        if (!s.layout.eq_layout(dst.layout) || !s.layout.is_contiguous())
            b200_throw("b200 Elemwise: operands must match dst (same shape, contiguous, no broadcasting)");
    const void* b = src.size() > 1 ? src[1].raw_ptr() : nullptr;
    check_rc("elemwise", b200cb_elemwise(
                                 static_cast<uint32_t>(param().mode), dtype_code(dst.layout.dtype),
                                 dst.layout.total_nr_elems(), src[0].raw_ptr(), b, dst.raw_ptr(), m_stream));
}

// ---------------------------------------------------------------- Pooling
b200cb_pool_desc PoolingForwardImpl::make_desc(const TensorLayout& src, const TensorLayout& dst) const {
    b200cb_pool_desc d;
    memset(&d, 0, sizeof(d));
    const auto& p = param();
    d.format = static_cast<uint32_t>(p.format);
    d.mode = static_cast<uint32_t>(p.mode);
    d.dtype = dtype_code(src.dtype);
    if (p.format == param::Pooling::Format::NHWC) {
        d.n = src[0], d.ih = src[1], d.iw = src[2], d.c = src[3], d.oh = dst[1], d.ow = dst[2];
    } else if (p.format == param::Pooling::Format::NCHW) {
        d.n = src[0], d.c = src[1], d.ih = src[2], d.iw = src[3], d.oh = dst[2], d.ow = dst[3];
    } else if (p.format == param::Pooling::Format::NCHW4) {
        d.n = src[0], d.c = src[1] * 4, d.ih = src[2], d.iw = src[3], d.oh = dst[2], d.ow = dst[3];
    } else {
        b200_throw("b200 Pooling: format not supported");
    }
    d.window_h = p.window_h, d.window_w = p.window_w, d.stride_h = p.stride_h, d.stride_w = p.stride_w;
    d.pad_h = p.pad_h, d.pad_w = p.pad_w;
    return d;
}

void PoolingForwardImpl::exec(_megdnn_tensor_in src, _megdnn_tensor_out dst, _megdnn_workspace workspace) {
    check_exec(src.layout, dst.layout, workspace.size);   // PoolingForward::check_exec (dnn/src/common/pooling.cpp)
    if (!src.layout.is_contiguous() || !dst.layout.is_contiguous())
        b200_throw("b200 Pooling: non-contiguous tensors are not supported");
    const b200cb_pool_desc d = make_desc(src.layout, dst.layout);
    check_rc("pooling", b200cb_pooling(&d, src.raw_ptr(), dst.raw_ptr(), m_stream));
}

std::vector<PoolingForward::Algorithm*> PoolingForwardImpl::get_all_algorithms(
        const TensorLayout& src, const TensorLayout& dst) {
    (void)make_desc(src, dst);
    return {&pooling_algo()};
}
std::vector<PoolingForward::Algorithm*> PoolingForwardImpl::get_all_algorithms_safe(
        const TensorLayout& src, const TensorLayout& dst) {
    return get_all_algorithms(src, dst);
}
PoolingForward::Algorithm* PoolingForwardImpl::get_algorithm_heuristic(
        const TensorLayout& src, const TensorLayout& dst, size_t, const AlgoAttribute&, const AlgoAttribute&) {
    (void)make_desc(src, dst);
    return &pooling_algo();
}
PoolingForward::Algorithm* PoolingForwardImpl::get_algorithm_from_desc(const AlgorithmDesc& desc) {
    if (desc.name != pooling_algo().name())
        b200_throw("b200 Pooling: unknown algorithm " + desc.name);
    return &pooling_algo();
}

std::unique_ptr<TypeCvtForward> create_type_cvt(Handle* handle, cudaStream_t stream) {
    return std::unique_ptr<TypeCvtForward>(new TypeCvtImpl(handle, stream));
}
std::unique_ptr<ElemwiseForward> create_elemwise(Handle* handle, cudaStream_t stream) {
    return std::unique_ptr<ElemwiseForward>(new ElemwiseForwardImpl(handle, stream));
}
std::unique_ptr<PoolingForward> create_pooling(Handle* handle, cudaStream_t stream) {
    return std::unique_ptr<PoolingForward>(new PoolingForwardImpl(handle, stream));
}

}  // namespace b200
}  // namespace megdnn
