// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product (libb200cb.so).
//
// C-ABI shim over the REAL reference operator: megdnn::ConvBiasForward from
// oracle/_ref/libmegdnn_ref.so (the reference's own dnn/src/{common,naive,fallback,x86}
// compiled by oracle/build_ref.sh).  Used by tests/ to pin the C restatement
// (oracle/convbias_oracle.c), by tests/golden/make_golden.py to generate fixtures, and by
// bench.py's cpu_baseline / --impl reference legs.
//
//   handle_kind 2 -> naive::HandleImpl  (oracle:   dnn/src/naive/conv_bias/opr_impl.cpp:211-292)
//   handle_kind 0 -> x86::HandleImpl    (baseline: dnn/src/fallback/conv_bias/opr_impl.cpp:306-323)
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200cb.h"
#include "b200cb_nn.h"
#include "megcore.h"
#include "megdnn/basic_types.h"
#include "megdnn/dtype.h"
#include "megdnn/handle.h"
#include "megdnn/oprs.h"

using namespace megdnn;

namespace {

thread_local std::string g_err;

// A MegcoreCPUDispatcher backed by a persistent thread pool (the reference ships one only in its
// test tree: dnn/test/common/task_executor.cpp).  Thread 0 is the calling thread.
class PoolDispatcher final : public MegcoreCPUDispatcher {
public:
    explicit PoolDispatcher(size_t n) : m_n(n ? n : 1) {
        for (size_t t = 1; t < m_n; ++t)
            m_workers.emplace_back([this, t] { worker(t); });
    }
    ~PoolDispatcher() noexcept override {
        {
            std::lock_guard<std::mutex> lk(m_mtx);
            m_stop = true;
            ++m_epoch;
        }
        m_cv.notify_all();
        for (auto& w : m_workers)
            w.join();
    }
    void dispatch(Task&& task) override { task(); }
    void dispatch(MultiThreadingTask&& task, size_t parallelism) override {
        if (m_n == 1 || parallelism <= 1) {
            for (size_t i = 0; i < parallelism; ++i)
                task(i, 0);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_mtx);
            m_task = &task;
            m_par = parallelism;
            m_next.store(0);
            m_pending.store(m_n - 1);
            ++m_epoch;
        }
        m_cv.notify_all();
        run(0);
        while (m_pending.load(std::memory_order_acquire) != 0)
            std::this_thread::yield();
        m_task = nullptr;
    }
    void sync() override {}
    size_t nr_threads() override { return m_n; }

private:
    void run(size_t tid) {
        for (;;) {
            size_t i = m_next.fetch_add(1);
            if (i >= m_par)
                break;
            (*m_task)(i, tid);
        }
    }
    void worker(size_t tid) {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_mtx);
                m_cv.wait(lk, [&] { return m_epoch != seen; });
                seen = m_epoch;
                if (m_stop)
                    return;
            }
            run(tid);
            m_pending.fetch_sub(1, std::memory_order_release);
        }
    }
    size_t m_n;
    std::vector<std::thread> m_workers;
    std::mutex m_mtx;
    std::condition_variable m_cv;
    size_t m_epoch = 0;
    bool m_stop = false;
    MultiThreadingTask* m_task = nullptr;
    size_t m_par = 0;
    std::atomic<size_t> m_next{0}, m_pending{0};
};

struct HandleBox {
    megcoreDeviceHandle_t dev = nullptr;
    megcoreComputingHandle_t comp = nullptr;
    std::unique_ptr<Handle> handle;
};

HandleBox* get_handle(int kind, int threads) {
    static std::mutex mtx;
    static std::map<std::pair<int, int>, std::unique_ptr<HandleBox>> cache;
    std::lock_guard<std::mutex> lk(mtx);
    auto key = std::make_pair(kind, threads);
    auto it = cache.find(key);
    if (it != cache.end())
        return it->second.get();
    auto box = std::make_unique<HandleBox>();
    megcoreCreateDeviceHandle(&box->dev, megcorePlatformCPU);
    if (threads > 1) {
        megcoreCreateComputingHandleWithCPUDispatcher(
                &box->comp, box->dev, std::make_shared<PoolDispatcher>(threads));
    } else {
        megcoreCreateComputingHandle(&box->comp, box->dev);
    }
    box->handle = Handle::make(box->comp, kind);
    auto* raw = box.get();
    cache[key] = std::move(box);
    return raw;
}

DType make_dtype(uint32_t dt, float scale) {
    switch (dt) {
        case B200CB_DT_F32: return dtype::Float32();
        case B200CB_DT_F16: return dtype::Float16();
        case B200CB_DT_BF16: return dtype::BFloat16();
        case B200CB_DT_QS8: return dtype::QuantizedS8(scale);
        case B200CB_DT_QS32: return dtype::QuantizedS32(scale);
        default: throw std::runtime_error("ref_driver: bad dtype");
    }
}

struct Layouts {
    TensorLayout src, flt, bias, z, dst;
};

Layouts make_layouts(const b200cb_desc& d) {
    Layouts l;
    const size_t G = d.group, icpg = d.ic / G, ocpg = d.oc / G;
    const bool grp = d.sparse == B200CB_SPARSE_GROUP;
    DType sd = make_dtype(d.src_dtype, d.src_scale), fd = make_dtype(d.filter_dtype, d.filter_scale);
    DType dd = make_dtype(d.dst_dtype, d.dst_scale);
    DType bd = d.bias_dtype == B200CB_DT_NONE ? dd : make_dtype(d.bias_dtype, d.bias_scale);
    DType zd = d.has_z ? make_dtype(d.z_dtype, d.z_scale) : dd;
    TensorShape bias_shape;
    if (d.format == B200CB_FORMAT_NHWC) {
        l.src = TensorLayout({d.n, d.ih, d.iw, d.ic}, sd);
        l.flt = grp ? TensorLayout({G, ocpg, d.fh, d.fw, icpg}, fd)
                    : TensorLayout({d.oc, d.fh, d.fw, d.ic}, fd);
        l.dst = TensorLayout({d.n, d.oh, d.ow, d.oc}, dd);
        bias_shape = {1, 1, 1, d.oc};
    } else if (d.format == B200CB_FORMAT_NCHW) {
        l.src = TensorLayout({d.n, d.ic, d.ih, d.iw}, sd);
        l.flt = grp ? TensorLayout({G, ocpg, icpg, d.fh, d.fw}, fd)
                    : TensorLayout({d.oc, d.ic, d.fh, d.fw}, fd);
        l.dst = TensorLayout({d.n, d.oc, d.oh, d.ow}, dd);
        bias_shape = {1, d.oc, 1, 1};
    } else if (d.format == B200CB_FORMAT_NCHW4) {
        l.src = TensorLayout({d.n, d.ic / 4, d.ih, d.iw, 4}, sd);
        l.flt = grp ? TensorLayout({G, ocpg, icpg / 4, d.fh, d.fw, 4}, fd)
                    : TensorLayout({d.oc, d.ic / 4, d.fh, d.fw, 4}, fd);
        l.dst = TensorLayout({d.n, d.oc / 4, d.oh, d.ow, 4}, dd);
        bias_shape = {1, d.oc / 4, 1, 1, 4};
    } else {
        throw std::runtime_error("ref_driver: bad format");
    }
    if (d.src_batch_stride)
        l.src.stride[0] = d.src_batch_stride;
    if (d.dst_batch_stride)
        l.dst.stride[0] = d.dst_batch_stride;
    if (d.bias_mode == B200CB_BIAS_CHANNEL)
        l.bias = TensorLayout(bias_shape, bd);
    else if (d.bias_mode == B200CB_BIAS_FULL)
        l.bias = TensorLayout(TensorShape(l.dst), bd);
    else
        l.bias = TensorLayout(bd);  // ndim == 0: "absent", dtype still set by callers
    if (d.has_z) {
        l.z = TensorLayout(TensorShape(l.dst), zd);
        if (d.z_batch_stride)
            l.z.stride[0] = d.z_batch_stride;
    } else {
        l.z = TensorLayout(zd);
    }
    return l;
}

}  // namespace

extern "C" {

const char* ref_last_error(void) {
    return g_err.c_str();
}

// Runs the reference ConvBiasForward `repeat` times; seconds[i] receives the wall time of run i
// (may be NULL).  algo_name receives Algorithm::name() chosen by the reference's own heuristic.
int ref_convbias(
        const b200cb_desc* d, int handle_kind, int threads, const void* src, const void* filter,
        const void* bias, const void* z, void* dst, int repeat, double* seconds, char* algo_name,
        int algo_name_cap) {
    try {
        HandleBox* hb = get_handle(handle_kind, threads);
        auto opr = hb->handle->create_operator<ConvBiasForward>();
        auto& p = opr->param();
        p.nonlineMode = static_cast<param::ConvBias::NonlineMode>(d->nonline_mode);
        p.mode = static_cast<param::ConvBias::Mode>(d->mode);
        p.sparse = static_cast<param::ConvBias::Sparse>(d->sparse);
        p.format = static_cast<param::ConvBias::Format>(d->format);
        p.pad_h = d->pad_h, p.pad_w = d->pad_w;
        p.stride_h = d->stride_h, p.stride_w = d->stride_w;
        p.dilate_h = d->dilate_h, p.dilate_w = d->dilate_w;
        p.compute_mode = static_cast<param::ConvBias::ComputeMode>(d->compute_mode);
        Layouts l = make_layouts(*d);
        // cross-check our shape deduction against the reference's (common/convolution.cpp:633-1049)
        TensorLayout deduced;
        deduced.dtype = l.dst.dtype;
        opr->deduce_layout(l.src, l.flt, l.bias, l.z, deduced);
        if (!deduced.eq_shape(l.dst))
            throw std::runtime_error(
                    "ref_driver: dst shape mismatch, reference deduces " + deduced.to_string());
        size_t ws_size = opr->get_workspace_in_bytes(l.src, l.flt, l.bias, l.z, l.dst, nullptr);
        std::vector<dt_byte> ws(ws_size + 64);
        if (algo_name && algo_name_cap > 0) {
            auto info = opr->get_algorithm_info_heuristic(l.src, l.flt, l.bias, l.z, l.dst);
            strncpy(algo_name, info.desc.name.c_str(), algo_name_cap - 1);
            algo_name[algo_name_cap - 1] = 0;
        }
        static char dummy[64];
        TensorND tsrc(const_cast<void*>(src), l.src), tflt(const_cast<void*>(filter), l.flt);
        TensorND tbias(bias ? const_cast<void*>(bias) : dummy, l.bias);
        TensorND tz(z ? const_cast<void*>(z) : dummy, l.z), tdst(dst, l.dst);
        for (int r = 0; r < (repeat > 0 ? repeat : 1); ++r) {
            auto t0 = std::chrono::steady_clock::now();
            opr->exec(tsrc, tflt, tbias, tz, tdst, nullptr, Workspace(ws.data(), ws_size));
            megcoreSynchronize(hb->comp);
            auto t1 = std::chrono::steady_clock::now();
            if (seconds)
                seconds[r] = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown exception";
        return 2;
    }
}

// ---- operators either side of ConvBias (include/b200cb_nn.h), reference's naive implementations
#define REF_TRY try {
#define REF_CATCH                          \
    return 0;                              \
    }                                      \
    catch (const std::exception& e) {      \
        g_err = e.what();                  \
        return 1;                          \
    }                                      \
    catch (...) {                          \
        g_err = "unknown exception";       \
        return 2;                          \
    }

int ref_pooling(const b200cb_pool_desc* d, const void* src, void* dst) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<PoolingForward>();
    auto& p = opr->param();
    p.mode = static_cast<param::Pooling::Mode>(d->mode);
    p.format = static_cast<param::Pooling::Format>(d->format);
    p.pad_h = d->pad_h, p.pad_w = d->pad_w, p.stride_h = d->stride_h, p.stride_w = d->stride_w;
    p.window_h = d->window_h, p.window_w = d->window_w;
    DType dt = make_dtype(d->dtype, 0.5f);
    TensorLayout ls, ld;
    if (d->format == B200CB_FORMAT_NHWC)
        ls = TensorLayout({d->n, d->ih, d->iw, d->c}, dt);
    else if (d->format == B200CB_FORMAT_NCHW)
        ls = TensorLayout({d->n, d->c, d->ih, d->iw}, dt);
    else
        ls = TensorLayout({d->n, d->c / 4, d->ih, d->iw, 4}, dt);
    ld.dtype = dt;
    opr->deduce_layout(ls, ld);
    const size_t oh = d->format == B200CB_FORMAT_NHWC ? ld[1] : ld[2], ow = d->format == B200CB_FORMAT_NHWC ? ld[2] : ld[3];
    if (oh != d->oh || ow != d->ow)
        throw std::runtime_error("ref_driver: pooling dst shape mismatch, reference deduces " + ld.to_string());
    size_t ws_size = opr->get_workspace_in_bytes(ls, ld);
    std::vector<dt_byte> ws(ws_size + 64);
    opr->exec(TensorND(const_cast<void*>(src), ls), TensorND(dst, ld), Workspace(ws.data(), ws_size));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

int ref_elemwise(uint32_t mode, uint32_t dtype, uint64_t n, const void* a, const void* b, void* dst) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<ElemwiseForward>();
    opr->param().mode = static_cast<param::Elemwise::Mode>(mode);
    TensorLayout l({(size_t)n}, make_dtype(dtype, 1.f));
    TensorNDArray in;
    in.emplace_back(const_cast<void*>(a), l);
    if (b)
        in.emplace_back(const_cast<void*>(b), l);
    opr->exec(in, TensorND(dst, l));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

int ref_elemwise_q(
        uint32_t mode, uint64_t n, const void* a, float sa, const void* b, float sb, void* dst, float sd) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<ElemwiseMultiType>();
    using M = param::ElemwiseMultiType::Mode;
    M m;
    switch (mode) {
        case B200CB_ELEM_RELU: m = M::QRELU; break;
        case B200CB_ELEM_ADD: m = M::QADD; break;
        case B200CB_ELEM_FUSE_ADD_RELU: m = M::QFUSE_ADD_RELU; break;
        case B200CB_ELEM_H_SWISH: m = M::QH_SWISH; break;
        default: throw std::runtime_error("ref_driver: bad quantized elemwise mode");
    }
    opr->param().mode = m;
    TensorNDArray in;
    in.emplace_back(const_cast<void*>(a), TensorLayout({(size_t)n}, dtype::QuantizedS8(sa)));
    if (b)
        in.emplace_back(const_cast<void*>(b), TensorLayout({(size_t)n}, dtype::QuantizedS8(sb)));
    opr->exec(in, TensorND(dst, TensorLayout({(size_t)n}, dtype::QuantizedS8(sd))));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

int ref_typecvt(uint32_t sdt, float ss, uint32_t ddt, float sd, uint64_t n, const void* src, void* dst) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<TypeCvt>();
    opr->exec(
            TensorND(const_cast<void*>(src), TensorLayout({(size_t)n}, make_dtype(sdt, ss))),
            TensorND(dst, TensorLayout({(size_t)n}, make_dtype(ddt, sd))));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

static void fill_conv_param(param::Convolution& p, const b200cb_desc* d) {
    p.mode = static_cast<param::Convolution::Mode>(d->mode);
    p.sparse = static_cast<param::Convolution::Sparse>(d->sparse);
    p.format = static_cast<param::Convolution::Format>(d->format);
    p.pad_h = d->pad_h, p.pad_w = d->pad_w;
    p.stride_h = d->stride_h, p.stride_w = d->stride_w;
    p.dilate_h = d->dilate_h, p.dilate_w = d->dilate_w;
    p.compute_mode = static_cast<param::Convolution::ComputeMode>(d->compute_mode);
}

// ConvolutionForward (bias / z / activation fields of the descriptor are ignored)
int ref_convolution(const b200cb_desc* d, const void* src, const void* filter, void* dst) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<ConvolutionForward>();
    fill_conv_param(opr->param(), d);
    b200cb_desc dd = *d;
    dd.bias_mode = B200CB_BIAS_NONE, dd.has_z = 0;
    Layouts l = make_layouts(dd);
    TensorLayout deduced;
    deduced.dtype = l.dst.dtype;
    opr->deduce_layout(l.src, l.flt, deduced);
    if (!deduced.eq_shape(l.dst))
        throw std::runtime_error("ref_driver: conv dst shape mismatch, reference deduces " + deduced.to_string());
    size_t ws_size = opr->get_workspace_in_bytes(l.src, l.flt, l.dst, nullptr);
    std::vector<dt_byte> ws(ws_size + 64);
    opr->exec(
            TensorND(const_cast<void*>(src), l.src), TensorND(const_cast<void*>(filter), l.flt), TensorND(dst, l.dst),
            nullptr, Workspace(ws.data(), ws_size));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

// ConvolutionBackwardData: `d` describes the forward convolution (src = grad, dst = diff)
int ref_deconv(const b200cb_desc* d, const void* filter, const void* diff, void* grad) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<ConvolutionBackwardData>();
    fill_conv_param(opr->param(), d);
    b200cb_desc dd = *d;
    dd.bias_mode = B200CB_BIAS_NONE, dd.has_z = 0;
    Layouts l = make_layouts(dd);
    TensorLayout deduced;
    deduced.dtype = l.src.dtype;
    opr->deduce_layout(l.flt, l.dst, deduced);
    // (deduce_layout gives the smallest grad extent; the forward descriptor may name a larger one)
    size_t ws_size = opr->get_workspace_in_bytes(l.flt, l.dst, l.src);
    std::vector<dt_byte> ws(ws_size + 64);
    opr->exec(
            TensorND(const_cast<void*>(filter), l.flt), TensorND(const_cast<void*>(diff), l.dst), TensorND(grad, l.src),
            Workspace(ws.data(), ws_size));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

// BatchConvBiasForward: filter carries a leading batch dimension
int ref_batch_convbias(
        const b200cb_desc* d, const void* src, const void* filter, const void* bias, const void* z, void* dst) {
    REF_TRY
    HandleBox* hb = get_handle(2, 1);
    auto opr = hb->handle->create_operator<BatchConvBiasForward>();
    auto& p = opr->param();
    p.nonlineMode = static_cast<param::BatchConvBias::NonlineMode>(d->nonline_mode);
    p.mode = static_cast<param::BatchConvBias::Mode>(d->mode);
    p.sparse = static_cast<param::BatchConvBias::Sparse>(d->sparse);
    p.format = static_cast<param::BatchConvBias::Format>(d->format);
    p.pad_h = d->pad_h, p.pad_w = d->pad_w, p.stride_h = d->stride_h, p.stride_w = d->stride_w;
    p.dilate_h = d->dilate_h, p.dilate_w = d->dilate_w;
    p.compute_mode = static_cast<param::BatchConvBias::ComputeMode>(d->compute_mode);
    Layouts l = make_layouts(*d);
    TensorShape fs;
    fs.ndim = l.flt.ndim + 1;
    fs[0] = d->n;
    for (size_t i = 0; i < l.flt.ndim; ++i)
        fs[i + 1] = l.flt[i];
    TensorLayout lf(fs, l.flt.dtype);
    size_t ws_size = opr->get_workspace_in_bytes(l.src, lf, l.bias, l.z, l.dst);
    std::vector<dt_byte> ws(ws_size + 64);
    static char dummy[64];
    opr->exec(
            TensorND(const_cast<void*>(src), l.src), TensorND(const_cast<void*>(filter), lf),
            TensorND(bias ? const_cast<void*>(bias) : dummy, l.bias), TensorND(z ? const_cast<void*>(z) : dummy, l.z),
            TensorND(dst, l.dst), Workspace(ws.data(), ws_size));
    megcoreSynchronize(hb->comp);
    REF_CATCH
}

int ref_hardware_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}
}
