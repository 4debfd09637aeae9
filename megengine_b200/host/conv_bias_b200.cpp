// See conv_bias_b200.h.  Reference counterparts: dnn/src/cuda/conv_bias/opr_impl.cpp:16-288 (operator),
// dnn/src/cuda/conv_bias/algo.h:38-165 + algo.cpp:7-150 (AlgoBase / AlgoPack / SizeArgs).
#include "conv_bias_b200.h"

#include <cstring>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <unordered_map>

namespace megdnn {
namespace b200 {
namespace {

// one megdnn Algorithm object per entry of the library's algorithm table -- the AlgoPack
// (dnn/src/cuda/conv_bias/algo.cpp:7-108): immutable after static init, shared by all operators
class AlgoB200 final : public detail::Algorithm {
public:
    explicit AlgoB200(int idx) : m_idx(idx), m_name(b200cb_algo_name(idx)) {
        m_handle_type = Handle::HandleType::CUDA;
    }
    Attribute attribute() const override { return Attribute::REPRODUCIBLE; }
    const char* name() const override { return m_name.c_str(); }
    //! appended after the reference's last AlgoType value (dnn/src/cuda/conv_bias/algo.h:43-82)
    uint32_t type() const override { return 0xB200u; }
    std::string param() const override {
        std::string r;
        serialize_write_pod(m_idx, r);
        return r;
    }
    int index() const { return m_idx; }

private:
    int m_idx;
    std::string m_name;
};

struct AlgoPack {
    std::vector<std::unique_ptr<AlgoB200>> all;
    AlgoPack() {
        for (int i = 0; i < b200cb_algo_count(); ++i)
            all.emplace_back(new AlgoB200(i));
    }
};
const AlgoPack& algo_pack() {
    static AlgoPack pack;
    return pack;
}

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
// true when only stride[0] differs from the contiguous layout; *batch_stride = stride[0] then (0 when
// the tensor is fully contiguous, the C-ABI's "dense" value)
bool batch_stride_only(const TensorLayout& l, uint64_t* batch_stride) {
    *batch_stride = 0;
    if (l.ndim == 0 || l.is_contiguous())
        return true;
    ptrdiff_t expect = 1;
    for (size_t i = l.ndim; i-- > 1;) {
        if (l.shape[i] != 1 && l.stride[i] != expect)
            return false;
        expect *= static_cast<ptrdiff_t>(l.shape[i]);
    }
    if (l.stride[0] < expect)
        return false;
    *batch_stride = static_cast<uint64_t>(l.stride[0]);
    return true;
}

// megdnn_throw / megdnn_assert live in the reference's private dnn/src/common/utils.h; the public
// hook they end in is ErrorHandler::on_megdnn_error (dnn/include/megdnn/basic_types.h:21-58)
[[noreturn]] void b200_throw(const std::string& msg) {
    megdnn::ErrorHandler::on_megdnn_error(msg);
    throw std::runtime_error(msg);  // not reached: the handler must not return
}
#define b200_assert(cond, msg)                                         \
    do {                                                               \
        if (!(cond))                                                   \
            b200_throw(std::string("assertion `" #cond "' failed: ") + (msg)); \
    } while (0)
[[noreturn]] void throw_rc(const char* what, int rc) {
    // megdnn_throw (dnn/include/megdnn/basic_types.h:21-58): the default handler raises
    megdnn::ErrorHandler::on_megdnn_error(
            std::string("b200cb ") + what + ": " + b200cb_strerror(rc) + " (" +
            b200cb_last_error_detail() + ")");
    throw std::runtime_error("unreachable");
}

}  // namespace

ConvBiasForwardImpl::ConvBiasForwardImpl(Handle* handle, cudaStream_t stream)
        : ConvBiasForward(handle), m_stream(stream) {}

std::unique_ptr<ConvBiasForward> create_conv_bias_forward(Handle* handle, cudaStream_t stream) {
    return std::unique_ptr<ConvBiasForward>(new ConvBiasForwardImpl(handle, stream));
}

// SizeArgs -> POD (dnn/src/cuda/conv_bias/algo.cpp:114-150, helper.h:27-36)
b200cb_desc ConvBiasForwardImpl::make_desc(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst) const {
    b200cb_desc d;
    memset(&d, 0, sizeof(d));
    const auto& p = param();
    const auto fm = make_canonized_filter_meta(src.ndim, filter);
    d.format = static_cast<uint32_t>(p.format);
    d.sparse = static_cast<uint32_t>(p.sparse);
    d.mode = static_cast<uint32_t>(p.mode);
    d.nonline_mode = static_cast<uint32_t>(p.nonlineMode);
    d.compute_mode = static_cast<uint32_t>(p.compute_mode);
    d.group = fm.group;
    d.ic = fm.icpg * fm.group;
    d.oc = fm.ocpg * fm.group;
    d.fh = fm.spatial[0], d.fw = fm.spatial[1];
    d.pad_h = fm.padding[0], d.pad_w = fm.padding[1];
    d.stride_h = fm.stride[0], d.stride_w = fm.stride[1];
    d.dilate_h = fm.dilation[0], d.dilate_w = fm.dilation[1];
    size_t h_axis;
    switch (p.format) {
        case param::ConvBias::Format::NHWC: h_axis = 1; break;
        case param::ConvBias::Format::NCHW:
        case param::ConvBias::Format::NCHW4: h_axis = 2; break;
        default: h_axis = 0; d.format = 0xffu; break;  // rejected by b200cb_query
    }
    d.n = static_cast<uint32_t>(src[0]);
    if (h_axis) {
        d.ih = static_cast<uint32_t>(src[h_axis]), d.iw = static_cast<uint32_t>(src[h_axis + 1]);
        d.oh = static_cast<uint32_t>(dst[h_axis]), d.ow = static_cast<uint32_t>(dst[h_axis + 1]);
    }
    d.src_dtype = dtype_code(src.dtype), d.src_scale = dtype_scale(src.dtype);
    d.filter_dtype = dtype_code(filter.dtype), d.filter_scale = dtype_scale(filter.dtype);
    d.dst_dtype = dtype_code(dst.dtype), d.dst_scale = dtype_scale(dst.dtype);
    d.bias_dtype = B200CB_DT_NONE, d.z_dtype = B200CB_DT_NONE;
    // BiasMode (nn.h:450-454); "absent" == ndim 0 (dnn/src/naive/convolution/helper.h:730-732)
    if (bias.ndim == 0) {
        d.bias_mode = B200CB_BIAS_NONE;
    } else {
        d.bias_mode = bias.eq_shape(dst) ? B200CB_BIAS_FULL : B200CB_BIAS_CHANNEL;
        d.bias_dtype = dtype_code(bias.dtype), d.bias_scale = dtype_scale(bias.dtype);
    }
    // A non-contiguous BATCH stride must be accepted (dnn/src/cuda/conv_bias/algo.h:32-37); any other
    // non-contiguity (channel- or W-sliced views forwarded by Subtensor / Concat) passes the reference's
    // check_exec_allow_noncontiguous and has to be rejected per algorithm, as the reference's CUDA algos
    // do (`!src_layout->is_contiguous()` -> unavailable): the descriptor is then marked unsupported.
    bool ok = filter.is_contiguous() && (bias.ndim == 0 || bias.is_contiguous());
    ok = ok && batch_stride_only(src, &d.src_batch_stride) && batch_stride_only(dst, &d.dst_batch_stride);
    if (z.ndim != 0) {
        d.has_z = 1;
        d.z_dtype = dtype_code(z.dtype), d.z_scale = dtype_scale(z.dtype);
        ok = ok && batch_stride_only(z, &d.z_batch_stride);
    }
    if (!ok)
        d.format = 0xffu;  // b200cb_query: B200CB_UNSUPPORTED_FORMAT -> is_available() == false
    return d;
}

std::vector<detail::Algorithm*> ConvBiasForwardImpl::get_all_algorithms(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst) {
    const b200cb_desc d = make_desc(src, filter, bias, z, dst);
    std::vector<Algorithm*> ret;
    for (auto& a : algo_pack().all) {
        b200cb_plan plan;
        if (b200cb_query_algo(&d, a->index(), &plan) == B200CB_OK)  // AlgoBase::is_available
            ret.push_back(a.get());
    }
    return ret;
}

std::vector<detail::Algorithm*> ConvBiasForwardImpl::get_all_algorithms_safe(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst) {
    auto ret = get_all_algorithms(src, filter, bias, z, dst);
    b200_assert(!ret.empty(), "no usable conv_bias fwd algorithm");  // algo_chooser.h:71-89
    return ret;
}

detail::Algorithm* ConvBiasForwardImpl::get_algorithm_heuristic(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst, size_t workspace_limit_in_bytes,
        const AlgoAttribute& positive_attr, const AlgoAttribute& negative_attr) {
    const b200cb_desc d = make_desc(src, filter, bias, z, dst);
    b200cb_plan plan;
    int rc = b200cb_query(&d, &plan);
    if (rc == B200CB_OK && plan.workspace_bytes <= workspace_limit_in_bytes) {
        Algorithm* a = algo_pack().all[plan.algo].get();
        if (a->contain_attribute_all(positive_attr) && !a->contain_attribute_any(negative_attr))
            return a;
    }
    // fall back to the first available one that fits (reference: opr_impl.cpp:226-238)
    for (auto* a : get_all_algorithms(src, filter, bias, z, dst)) {
        b200cb_plan p2;
        if (b200cb_query_algo(&d, static_cast<AlgoB200*>(a)->index(), &p2) == B200CB_OK &&
            p2.workspace_bytes <= workspace_limit_in_bytes && a->contain_attribute_all(positive_attr) &&
            !a->contain_attribute_any(negative_attr))
            return a;
    }
    b200_throw("no usable conv_bias fwd algorithm for B200");
}

detail::Algorithm* ConvBiasForwardImpl::get_algorithm_from_desc(const AlgorithmDesc& desc) {
    for (auto& a : algo_pack().all)
        if (a->desc() == desc)
            return a.get();
    b200_throw("can not find the B200 conv_bias algorithm of this desc");
}

// The reference resolves an algorithm in three steps -- execution_policy().algo, the process-wide
// AlgorithmCache (keyed by the serialised layouts + param, dnn/include/megdnn/algorithm_cache.h:12-80,
// dnn/src/common/algo_chooser.h:24-50), then the heuristic.  Same here: the cache maps the POD descriptor
// (which carries every field of the layouts and the param that the algorithms look at) + the forced
// algorithm to the finished plan, so that exec() / get_workspace_in_bytes() of a steady-state network
// cost one hash lookup instead of a query over the whole algorithm table.
namespace {
struct PlanCache {
    struct Key {
        b200cb_desc desc;
        int algo;  // -1: heuristic
        bool operator==(const Key& o) const { return algo == o.algo && !memcmp(&desc, &o.desc, sizeof(desc)); }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            // FNV-1a over the POD (memset-initialised, so padding bytes are deterministic)
            const unsigned char* p = reinterpret_cast<const unsigned char*>(&k.desc);
            uint64_t h = 1469598103934665603ull ^ static_cast<uint64_t>(k.algo + 1);
            for (size_t i = 0; i < sizeof(k.desc); ++i)
                h = (h ^ p[i]) * 1099511628211ull;
            return static_cast<size_t>(h);
        }
    };
    std::mutex mtx;
    std::unordered_map<Key, b200cb_plan, Hash> map;
    size_t hits = 0, misses = 0;
};
PlanCache& plan_cache() {
    static PlanCache c;
    return c;
}
}  // namespace

void ConvBiasForwardImpl::resolve_plan(const b200cb_desc& desc, b200cb_plan* plan) {
    PlanCache::Key key;
    key.desc = desc;
    key.algo = -1;
    if (execution_policy().algo.valid())
        key.algo = static_cast<AlgoB200*>(get_algorithm_from_desc(execution_policy().algo))->index();
    PlanCache& cache = plan_cache();
    {
        std::lock_guard<std::mutex> lk(cache.mtx);
        auto it = cache.map.find(key);
        if (it != cache.map.end()) {
            ++cache.hits;
            *plan = it->second;
            return;
        }
        ++cache.misses;
    }
    const int rc = key.algo >= 0 ? b200cb_query_algo(&desc, key.algo, plan) : b200cb_query(&desc, plan);
    if (rc != B200CB_OK)
        throw_rc("query", rc);
    std::lock_guard<std::mutex> lk(cache.mtx);
    if (cache.map.size() >= 4096)
        cache.map.clear();  // bounded like the reference's per-opr caches
    cache.map.emplace(key, *plan);
}

void plan_cache_stats(size_t* hits, size_t* misses) {
    PlanCache& cache = plan_cache();
    std::lock_guard<std::mutex> lk(cache.mtx);
    *hits = cache.hits, *misses = cache.misses;
}

size_t ConvBiasForwardImpl::get_workspace_in_bytes(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst,
        const PreprocessedFilter* preprocessed_filter) {
    b200cb_plan plan;
    resolve_plan(make_desc(src, filter, bias, z, dst), &plan);
    if (preprocessed_filter && !preprocessed_filter->tensors.empty())
        b200cb_plan_use_preprocessed_filter(&plan);
    return plan.workspace_bytes;
}

SmallVector<TensorLayout> ConvBiasForwardImpl::deduce_preprocessed_filter_layout(
        const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
        const TensorLayout& z, const TensorLayout& dst) {
    b200cb_plan plan;
    resolve_plan(make_desc(src, filter, bias, z, dst), &plan);
    const size_t bytes = b200cb_preprocessed_filter_bytes(&plan);
    if (!bytes)
        return {};  // empty == "no preprocessing" (nn.h:436-447)
    return {TensorLayout({bytes}, dtype::Byte())};
}

void ConvBiasForwardImpl::exec_preprocess(
        const TensorLayout& src_layout, _megdnn_tensor_in filter, _megdnn_tensor_in bias,
        const TensorLayout& z_layout, const TensorLayout& dst_layout,
        PreprocessedFilter* preprocessed_filter, _megdnn_workspace) {
    b200_assert(preprocessed_filter && !preprocessed_filter->tensors.empty(), "exec_preprocess needs storage");
    b200cb_plan plan;
    resolve_plan(make_desc(src_layout, filter.layout, bias.layout, z_layout, dst_layout), &plan);
    int rc = b200cb_preprocess_filter(
            &plan, filter.raw_ptr(), preprocessed_filter->tensors[0].raw_ptr(), m_stream);
    if (rc != B200CB_OK)
        throw_rc("preprocess_filter", rc);
}

void ConvBiasForwardImpl::exec(
        _megdnn_tensor_in src, _megdnn_tensor_in filter, _megdnn_tensor_in bias,
        _megdnn_tensor_in z, _megdnn_tensor_out dst, const PreprocessedFilter* preprocessed_filter,
        _megdnn_workspace workspace) {
    // the reference's own argument validation (dnn/src/common/conv_bias.cpp:8-133)
    check_exec_allow_noncontiguous(
            src.layout, filter.layout, bias.layout, z.layout, dst.layout, workspace.size,
            preprocessed_filter);
    b200cb_plan plan;
    resolve_plan(make_desc(src.layout, filter.layout, bias.layout, z.layout, dst.layout), &plan);
    const void* flt = filter.raw_ptr();
    if (preprocessed_filter && !preprocessed_filter->tensors.empty()) {
        b200cb_plan_use_preprocessed_filter(&plan);
        flt = preprocessed_filter->tensors[0].raw_ptr();
    }
    int rc = b200cb_launch(
            &plan, src.raw_ptr(), flt, bias.layout.ndim ? bias.raw_ptr() : nullptr,
            z.layout.ndim ? z.raw_ptr() : nullptr, dst.raw_ptr(), workspace.raw_ptr, workspace.size,
            m_stream);
    if (rc != B200CB_OK)
        throw_rc("launch", rc);
}

}  // namespace b200
}  // namespace megdnn
