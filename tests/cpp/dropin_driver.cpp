// TEST INFRASTRUCTURE -- drop-in check of megdnn::b200::ConvBiasForwardImpl (megengine_b200/host/).
//
// Drives the B200 operator exactly as the reference's Checker<ConvBiasForward> drives a backend
// (dnn/test/common/checker.h, dnn/test/cuda/conv_bias.cpp:83-124, conv_bias_int8.cpp:798-835): the
// reference's own naive operator (Handle::make(comp, 2), dnn/src/naive/conv_bias/opr_impl.cpp:211-292)
// computes the expected dst on the host from the same TensorNDs, param and dtypes; the B200 operator
// runs from device copies through the inherited public interface -- deduce_layout,
// get_all_algorithms_info_safe, execution_policy (every algorithm, as fastrun does),
// get_workspace_in_bytes, exec -- and the results are compared: QuantizedS8 bit for bit, BFloat16 /
// Float16 with the reference test's own eps (2e-2 / 2e-2, metric dnn/test/common/utils.h:189-193).
#include <cuda_runtime.h>

#include <time.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "conv_bias_b200.h"
#include "nn_b200.h"
#include "megcore.h"
#include "megdnn/dtype.h"
#include "megdnn/handle.h"
#include "megdnn/oprs.h"

using namespace megdnn;

namespace {

#define CUDA_OK(x)                                                                  \
    do {                                                                            \
        cudaError_t e__ = (x);                                                      \
        if (e__ != cudaSuccess) {                                                   \
            fprintf(stderr, "CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e__)); \
            exit(3);                                                                \
        }                                                                           \
    } while (0)

struct Case {
    const char* name;
    param::ConvBias::Format format;
    DTypeEnum kind;
    size_t n, ic, h, w, oc, f, stride, pad, group;
    param::ConvBias::NonlineMode nl;
    bool has_z;
};

DType mk(DTypeEnum k, float scale) {
    switch (k) {
        case DTypeEnum::QuantizedS8: return dtype::QuantizedS8(scale);
        case DTypeEnum::BFloat16: return dtype::BFloat16();
        case DTypeEnum::Float16: return dtype::Float16();
        default: return dtype::Float32();
    }
}

void fill(std::vector<uint8_t>& buf, const DType& dt, size_t elems, std::mt19937& rng, float fscale) {
    buf.resize(elems * dt.size());
    if (dt.enumv() == DTypeEnum::QuantizedS8) {
        std::uniform_int_distribution<int> u(-127, 127);
        for (size_t i = 0; i < elems; ++i)
            reinterpret_cast<int8_t*>(buf.data())[i] = static_cast<int8_t>(u(rng));
    } else if (dt.enumv() == DTypeEnum::QuantizedS32) {
        std::uniform_int_distribution<int> u(-5000, 5000);
        for (size_t i = 0; i < elems; ++i)
            reinterpret_cast<int32_t*>(buf.data())[i] = u(rng);
    } else {
        std::normal_distribution<float> g(0.f, fscale);
        for (size_t i = 0; i < elems; ++i) {
            float v = g(rng);
            if (dt.enumv() == DTypeEnum::BFloat16)
                reinterpret_cast<dt_bfloat16*>(buf.data())[i] = dt_bfloat16(v);
            else if (dt.enumv() == DTypeEnum::Float32)
                reinterpret_cast<float*>(buf.data())[i] = v;
            else
                reinterpret_cast<dt_float16*>(buf.data())[i] = dt_float16(v);
        }
    }
}

float to_f32(const uint8_t* p, const DType& dt, size_t i) {
    switch (dt.enumv()) {
        case DTypeEnum::QuantizedS8: return reinterpret_cast<const int8_t*>(p)[i];
        case DTypeEnum::BFloat16: return static_cast<float>(reinterpret_cast<const dt_bfloat16*>(p)[i]);
        case DTypeEnum::Float16: return static_cast<float>(reinterpret_cast<const dt_float16*>(p)[i]);
        default: return reinterpret_cast<const float*>(p)[i];
    }
}

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { CUDA_OK(cudaMalloc(&p, bytes ? bytes : 16)); }
    ~DevBuf() { cudaFree(p); }
};

bool run_case(Handle* host, cudaStream_t stream, const Case& c) {
    std::mt19937 rng(1234 + c.ic * 7 + c.oc);
    const bool q = c.kind == DTypeEnum::QuantizedS8;
    const float s_src = 0.37f, s_flt = 0.71f, s_z = 1.9f, s_dst = 23.0f;
    DType sd = mk(c.kind, s_src), fd = mk(c.kind, s_flt), dd = mk(c.kind, s_dst), zd = mk(c.kind, s_z);
    DType bd = q ? DType(dtype::QuantizedS32(s_src * s_flt)) : sd;
    const size_t oh = (c.h + 2 * c.pad - c.f) / c.stride + 1, ow = (c.w + 2 * c.pad - c.f) / c.stride + 1;
    TensorLayout lsrc, lflt, lbias, ldst;
    const bool grp = c.group > 1;
    const size_t icpg = c.ic / c.group, ocpg = c.oc / c.group;
    if (c.format == param::ConvBias::Format::NHWC) {
        lsrc = TensorLayout({c.n, c.h, c.w, c.ic}, sd);
        lflt = grp ? TensorLayout({c.group, ocpg, c.f, c.f, icpg}, fd) : TensorLayout({c.oc, c.f, c.f, c.ic}, fd);
        lbias = TensorLayout({1, 1, 1, c.oc}, bd);
    } else if (c.format == param::ConvBias::Format::NCHW) {
        lsrc = TensorLayout({c.n, c.ic, c.h, c.w}, sd);
        lflt = grp ? TensorLayout({c.group, ocpg, icpg, c.f, c.f}, fd) : TensorLayout({c.oc, c.ic, c.f, c.f}, fd);
        lbias = TensorLayout({1, c.oc, 1, 1}, bd);
    } else {  // NCHW4
        lsrc = TensorLayout({c.n, c.ic / 4, c.h, c.w, 4}, sd);
        lflt = TensorLayout({c.oc, c.ic / 4, c.f, c.f, 4}, fd);
        lbias = TensorLayout({1, c.oc / 4, 1, 1, 4}, bd);
    }
    auto set_param = [&](ConvBiasForward* opr) {
        auto& p = opr->param();
        p.format = c.format;
        p.nonlineMode = c.nl;
        p.sparse = grp ? param::ConvBias::Sparse::GROUP : param::ConvBias::Sparse::DENSE;
        p.pad_h = p.pad_w = c.pad;
        p.stride_h = p.stride_w = c.stride;
        p.compute_mode = (q || c.kind == DTypeEnum::Float32) ? param::ConvBias::ComputeMode::DEFAULT
                                                             : param::ConvBias::ComputeMode::FLOAT32;
    };
    auto ref = host->create_operator<ConvBiasForward>();
    auto dut = b200::create_conv_bias_forward(host, stream);
    set_param(ref.get());
    set_param(dut.get());
    // shape deduction must agree between the reference operator and ours (same inherited code path)
    ldst.dtype = dd;
    TensorLayout lz_absent(zd);
    dut->deduce_layout(lsrc, lflt, lbias, lz_absent, ldst);
    TensorLayout lz = c.has_z ? TensorLayout(TensorShape(ldst), zd) : lz_absent;
    if (ldst.total_nr_elems() != c.n * oh * ow * c.oc) {
        printf("FAIL %-28s deduce_layout -> %s\n", c.name, ldst.to_string().c_str());
        return false;
    }
    std::vector<uint8_t> hsrc, hflt, hbias, hz, href(ldst.span().dist_byte()), hgot(ldst.span().dist_byte());
    fill(hsrc, sd, lsrc.total_nr_elems(), rng, 1.f);
    fill(hflt, fd, lflt.total_nr_elems(), rng, std::sqrt(2.f / (icpg * c.f * c.f)));
    fill(hbias, bd, lbias.total_nr_elems(), rng, 0.3f);
    if (c.has_z)
        fill(hz, zd, lz.total_nr_elems(), rng, 1.f);
    static char dummy[64];
    // expected: the reference's naive operator on the host
    {
        size_t ws = ref->get_workspace_in_bytes(lsrc, lflt, lbias, lz, ldst, nullptr);
        std::vector<dt_byte> wbuf(ws + 64);
        ref->exec(
                TensorND(hsrc.data(), lsrc), TensorND(hflt.data(), lflt), TensorND(hbias.data(), lbias),
                TensorND(c.has_z ? (void*)hz.data() : (void*)dummy, lz), TensorND(href.data(), ldst), nullptr,
                Workspace(wbuf.data(), ws));
    }
    DevBuf dsrc(hsrc.size()), dflt(hflt.size()), dbias(hbias.size()), dz(hz.size()), ddst(hgot.size());
    CUDA_OK(cudaMemcpyAsync(dsrc.p, hsrc.data(), hsrc.size(), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dflt.p, hflt.data(), hflt.size(), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dbias.p, hbias.data(), hbias.size(), cudaMemcpyHostToDevice, stream));
    if (c.has_z)
        CUDA_OK(cudaMemcpyAsync(dz.p, hz.data(), hz.size(), cudaMemcpyHostToDevice, stream));
    // fastrun protocol: every algorithm the operator reports, forced through execution_policy()
    auto infos = dut->get_all_algorithms_info_safe(lsrc, lflt, lbias, lz, ldst);
    auto heur = dut->get_algorithm_info_heuristic(lsrc, lflt, lbias, lz, ldst);
    bool all_ok = true;
    std::string tried;
    for (auto& info : infos) {
        dut->execution_policy().algo = info.desc;
        size_t ws = dut->get_workspace_in_bytes(lsrc, lflt, lbias, lz, ldst, nullptr);
        DevBuf dws(ws);
        CUDA_OK(cudaMemsetAsync(ddst.p, 0x55, hgot.size(), stream));
        dut->exec(
                TensorND(dsrc.p, lsrc), TensorND(dflt.p, lflt), TensorND(dbias.p, lbias),
                TensorND(c.has_z ? dz.p : nullptr, lz), TensorND(ddst.p, ldst), nullptr, Workspace((dt_byte*)dws.p, ws));
        CUDA_OK(cudaMemcpyAsync(hgot.data(), ddst.p, hgot.size(), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        size_t bad = 0;
        double max_err = 0;
        const size_t ne = ldst.total_nr_elems();
        for (size_t i = 0; i < ne; ++i) {
            if (q) {
                bad += hgot[i] != href[i];
            } else {
                float a = to_f32(hgot.data(), dd, i), b = to_f32(href.data(), dd, i);
                double e = std::fabs(a - b) / std::fmax(std::fmax(std::fabs(a), std::fabs(b)), 1.0);
                // eps: dnn/test/cuda/conv_bias.cpp:83-100 (Float32 1e-3), :102-124,234-265 (16-bit 2e-2)
                if (!(e <= (c.kind == DTypeEnum::Float32 ? 1e-3 : 2e-2)))
                    ++bad;
                max_err = std::fmax(max_err, e);
            }
        }
        tried += std::string(" ") + info.desc.name + (info.desc == heur.desc ? "*" : "") +
                 (bad ? "[BAD]" : "[ok]");
        if (bad) {
            all_ok = false;
            printf("     %s: %zu / %zu elements differ (max rel err %.3g)\n", info.desc.name.c_str(), bad, ne, max_err);
        }
    }
    printf("%s %-28s algos:%s\n", all_ok ? "PASS" : "FAIL", c.name, tried.c_str());
    return all_ok;
}

// ---- the operators either side of ConvBias (megengine_b200/host/nn_b200.h): each B200 operator against the
// reference's naive operator of the same class on the same TensorNDs, bit for bit
bool same_bytes(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, const char* name) {
    size_t bad = 0;
    for (size_t i = 0; i < a.size(); ++i)
        bad += a[i] != b[i];
    printf("%s %-28s %zu / %zu bytes differ\n", bad ? "FAIL" : "PASS", name, bad, a.size());
    return bad == 0;
}

int run_nn_cases(Handle* host, cudaStream_t stream) {
    int failed = 0;
    std::mt19937 rng(77);
    // TypeCvt: QuantizedS32 -> QuantizedS8 (the requantize between int8 blocks), Float32 -> BFloat16, QuantizedS8 -> Float16
    struct Cvt { const char* name; DType s, d; };
    const Cvt cvts[] = {{"typecvt_qs32_qs8", dtype::QuantizedS32(0.0013f), dtype::QuantizedS8(0.9f)},
                        {"typecvt_f32_bf16", dtype::Float32(), dtype::BFloat16()},
                        {"typecvt_qs8_f16", dtype::QuantizedS8(0.3f), dtype::Float16()}};
    for (auto& c : cvts) {
        const size_t n = 100003;
        TensorLayout ls({n}, c.s), ld({n}, c.d);
        std::vector<uint8_t> hs, want(n * c.d.size()), got(n * c.d.size());
        fill(hs, c.s, n, rng, 3.f);
        host->create_operator<TypeCvt>()->exec(TensorND(hs.data(), ls), TensorND(want.data(), ld));
        DevBuf ds(hs.size()), dd(got.size());
        CUDA_OK(cudaMemcpyAsync(ds.p, hs.data(), hs.size(), cudaMemcpyHostToDevice, stream));
        b200::create_type_cvt(host, stream)->exec(TensorND(ds.p, ls), TensorND(dd.p, ld));
        CUDA_OK(cudaMemcpyAsync(got.data(), dd.p, got.size(), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        failed += !same_bytes(got, want, c.name);
    }
    // Elemwise: FUSE_ADD_RELU (the residual add gopt leaves when z-fusion is not legal), bf16 and fp16; H_SWISH fp32
    struct El { const char* name; param::Elemwise::Mode mode; DType dt; int arity; };
    const El els[] = {{"elemwise_add_relu_bf16", param::Elemwise::Mode::FUSE_ADD_RELU, dtype::BFloat16(), 2},
                      {"elemwise_add_relu_f16", param::Elemwise::Mode::FUSE_ADD_RELU, dtype::Float16(), 2},
                      {"elemwise_hswish_f32", param::Elemwise::Mode::H_SWISH, dtype::Float32(), 1}};
    for (auto& c : els) {
        TensorLayout l({3, 14, 14, 40}, c.dt);
        const size_t n = l.total_nr_elems();
        std::vector<uint8_t> ha, hb, want(n * c.dt.size()), got(n * c.dt.size());
        fill(ha, c.dt, n, rng, 2.f);
        fill(hb, c.dt, n, rng, 2.f);
        auto ref = host->create_operator<ElemwiseForward>();
        ref->param().mode = c.mode;
        TensorNDArray in;
        in.emplace_back(ha.data(), l);
        if (c.arity > 1)
            in.emplace_back(hb.data(), l);
        ref->exec(in, TensorND(want.data(), l));
        DevBuf da(ha.size()), db(hb.size()), dd(got.size());
        CUDA_OK(cudaMemcpyAsync(da.p, ha.data(), ha.size(), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(db.p, hb.data(), hb.size(), cudaMemcpyHostToDevice, stream));
        auto dut = b200::create_elemwise(host, stream);
        dut->param().mode = c.mode;
        TensorNDArray din;
        din.emplace_back(da.p, l);
        if (c.arity > 1)
            din.emplace_back(db.p, l);
        dut->exec(din, TensorND(dd.p, l));
        CUDA_OK(cudaMemcpyAsync(got.data(), dd.p, got.size(), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        failed += !same_bytes(got, want, c.name);
    }
    // Pooling: the ResNet stem max pool (bf16 NHWC), global average (QuantizedS8 NHWC), max pool NCHW4 int8
    struct Po { const char* name; param::Pooling::Format fmt; param::Pooling::Mode mode; DType dt; size_t n, c, h, w, win, st, pad; };
    const Po pos[] = {
            {"pool_max3s2_bf16_nhwc", param::Pooling::Format::NHWC, param::Pooling::Mode::MAX, dtype::BFloat16(), 2, 64, 28, 28, 3, 2, 1},
            {"pool_avg7_qs8_nhwc", param::Pooling::Format::NHWC, param::Pooling::Mode::AVERAGE, dtype::QuantizedS8(0.5f), 3, 128, 7, 7, 7, 7, 0},
            {"pool_avgx3s2_f16_nhwc", param::Pooling::Format::NHWC, param::Pooling::Mode::AVERAGE_COUNT_EXCLUDE_PADDING, dtype::Float16(), 2, 24, 9, 11, 3, 2, 1},
            {"pool_max3s2_qs8_nchw4", param::Pooling::Format::NCHW4, param::Pooling::Mode::MAX, dtype::QuantizedS8(0.5f), 2, 16, 15, 15, 3, 2, 1}};
    for (auto& c : pos) {
        auto ref = host->create_operator<PoolingForward>();
        auto dut = b200::create_pooling(host, stream);
        for (PoolingForward* o : {ref.get(), dut.get()}) {
            auto& p = o->param();
            p.format = c.fmt, p.mode = c.mode;
            p.window_h = p.window_w = c.win, p.stride_h = p.stride_w = c.st, p.pad_h = p.pad_w = c.pad;
        }
        TensorLayout ls = c.fmt == param::Pooling::Format::NHWC ? TensorLayout({c.n, c.h, c.w, c.c}, c.dt)
                                                                 : TensorLayout({c.n, c.c / 4, c.h, c.w, 4}, c.dt);
        TensorLayout ld, ld2;
        ld.dtype = ld2.dtype = c.dt;
        ref->deduce_layout(ls, ld);
        dut->deduce_layout(ls, ld2);
        if (!ld.eq_layout(ld2)) {
            printf("FAIL %-28s deduce_layout differs\n", c.name);
            ++failed;
            continue;
        }
        const size_t ne = ls.total_nr_elems(), no = ld.total_nr_elems();
        std::vector<uint8_t> hs, want(no * c.dt.size()), got(no * c.dt.size());
        fill(hs, c.dt, ne, rng, 2.f);
        std::vector<dt_byte> ws(ref->get_workspace_in_bytes(ls, ld) + 64);
        ref->exec(TensorND(hs.data(), ls), TensorND(want.data(), ld), Workspace(ws.data(), ws.size() - 64));
        DevBuf ds(hs.size()), dd(got.size());
        CUDA_OK(cudaMemcpyAsync(ds.p, hs.data(), hs.size(), cudaMemcpyHostToDevice, stream));
        auto algos = dut->get_all_algorithms_info(ls, ld);
        dut->exec(TensorND(ds.p, ls), TensorND(dd.p, ld), Workspace(nullptr, 0));
        CUDA_OK(cudaMemcpyAsync(got.data(), dd.p, got.size(), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        failed += !(same_bytes(got, want, c.name) && algos.size() == 1);
    }
    return failed;
}

}  // namespace

int main() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "dropin_driver: no CUDA device\n");
        return 2;
    }
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));
    megcoreDeviceHandle_t dev;
    megcoreComputingHandle_t comp;
    megcoreCreateDeviceHandle(&dev, megcorePlatformCPU);
    megcoreCreateComputingHandle(&comp, dev);
    auto host = Handle::make(comp, 2);  // naive: the oracle
    using F = param::ConvBias::Format;
    using NL = param::ConvBias::NonlineMode;
    const Case cases[] = {
            {"nhwc_qs8_3x3_z_relu", F::NHWC, DTypeEnum::QuantizedS8, 2, 64, 14, 14, 128, 3, 1, 1, 1, NL::RELU, true},
            {"nhwc_qs8_1x1_hswish", F::NHWC, DTypeEnum::QuantizedS8, 3, 32, 9, 9, 64, 1, 1, 0, 1, NL::H_SWISH, false},
            {"nhwc_qs8_7x7s2_stem", F::NHWC, DTypeEnum::QuantizedS8, 2, 16, 30, 30, 64, 7, 2, 3, 1, NL::RELU, false},
            {"nchw4_qs8_3x3_z_relu", F::NCHW4, DTypeEnum::QuantizedS8, 2, 32, 9, 11, 64, 3, 1, 1, 1, NL::RELU, true},
            {"nchw4_qs8_ic4_5x5s2", F::NCHW4, DTypeEnum::QuantizedS8, 2, 4, 17, 17, 32, 5, 2, 2, 1, NL::IDENTITY, false},
            {"nhwc_bf16_3x3_z_relu", F::NHWC, DTypeEnum::BFloat16, 2, 64, 14, 14, 256, 3, 1, 1, 1, NL::RELU, true},
            {"nhwc_f16_1x1s2_identity", F::NHWC, DTypeEnum::Float16, 2, 64, 14, 14, 128, 1, 2, 0, 1, NL::IDENTITY, false},
            {"nhwc_bf16_depthwise3x3", F::NHWC, DTypeEnum::BFloat16, 2, 32, 14, 14, 32, 3, 1, 1, 32, NL::RELU, false},
            {"nchw_bf16_3x3_z_relu", F::NCHW, DTypeEnum::BFloat16, 2, 24, 12, 12, 64, 3, 1, 1, 1, NL::RELU, true},
            {"nchw_f16_7x7s2_ic3_stem", F::NCHW, DTypeEnum::Float16, 2, 3, 30, 30, 64, 7, 2, 3, 1, NL::RELU, false},
            {"nchw_f16_depthwise3x3s2", F::NCHW, DTypeEnum::Float16, 2, 48, 15, 15, 48, 3, 2, 1, 48, NL::RELU, false},
            {"nhwc_qs8_depthwise3x3s2", F::NHWC, DTypeEnum::QuantizedS8, 2, 32, 15, 15, 32, 3, 2, 1, 32, NL::RELU, false},
            {"nchw_f32_3x3_z_relu", F::NCHW, DTypeEnum::Float32, 2, 20, 12, 12, 64, 3, 1, 1, 1, NL::RELU, true},
            {"nhwc_f32_7x7s2_ic3", F::NHWC, DTypeEnum::Float32, 2, 3, 30, 30, 64, 7, 2, 3, 1, NL::RELU, false},
            {"nhwc_bf16_group4_3x3_z", F::NHWC, DTypeEnum::BFloat16, 2, 64, 14, 14, 128, 3, 1, 1, 4, NL::RELU, true},
            {"nhwc_qs8_group2_1x1", F::NHWC, DTypeEnum::QuantizedS8, 2, 64, 9, 9, 64, 1, 1, 0, 2, NL::IDENTITY, false},
    };
    int failed = 0;
    for (auto& c : cases) {
        try {
            failed += !run_case(host.get(), stream, c);
        } catch (const std::exception& e) {
            printf("FAIL %-28s exception: %s\n", c.name, e.what());
            ++failed;
        }
    }
    // error behaviour: an unsupported cell must throw through megdnn's error handler, not crash
    {
        auto dut = b200::create_conv_bias_forward(host.get(), stream);
        dut->param().format = F::NCHW;
        // (Float32 with an output channel count that is not a multiple of 4: outside the supported cells)
        TensorLayout s({2, 8, 10, 10}, dtype::Float32()), f({6, 8, 3, 3}, dtype::Float32());
        TensorLayout b({1, 6, 1, 1}, dtype::Float32()), z{DType(dtype::Float32())}, d({2, 6, 8, 8}, dtype::Float32());
        bool threw = false;
        try {
            dut->get_all_algorithms_info_safe(s, f, b, z, d);
        } catch (const std::exception&) {
            threw = true;
        }
        printf("%s %-28s\n", threw ? "PASS" : "FAIL", "unsupported_cell_throws");
        failed += !threw;
    }
    // layouts that are non-contiguous in anything but the batch stride pass the reference's
    // check_exec_allow_noncontiguous and must be reported as "no algorithm available" (the reference's CUDA
    // algos: `!src_layout->is_contiguous()` -> unavailable), never run as if they were dense.
    {
        auto dut = b200::create_conv_bias_forward(host.get(), stream);
        auto& p = dut->param();
        p.format = F::NHWC;
        p.pad_h = p.pad_w = 1;
        p.compute_mode = param::ConvBias::ComputeMode::FLOAT32;
        DType bf = dtype::BFloat16();
        TensorLayout s({2, 14, 14, 64}, bf), f({64, 3, 3, 64}, bf), b({1, 1, 1, 64}, bf), z{bf}, d({2, 14, 14, 64}, bf);
        const size_t n_dense = dut->get_all_algorithms_info(s, f, b, z, d).size();
        TensorLayout s_batch = s;   // batch-strided view: must stay available
        s_batch.stride[0] = s.stride[0] + 4096;
        const size_t n_batch = dut->get_all_algorithms_info(s_batch, f, b, z, d).size();
        TensorLayout s_chan = s;    // channel-sliced view of a 128-channel tensor: pixel stride 128
        s_chan.stride[2] = 128, s_chan.stride[1] = 128 * 14, s_chan.stride[0] = 128 * 14 * 14;
        const size_t n_chan = dut->get_all_algorithms_info(s_chan, f, b, z, d).size();
        TensorLayout d_w = d;       // W-sliced dst view
        d_w.stride[1] = 64 * 20, d_w.stride[0] = 64 * 20 * 14;
        const size_t n_w = dut->get_all_algorithms_info(s, f, b, z, d_w).size();
        const bool ok = n_dense >= 1 && n_batch == n_dense && n_chan == 0 && n_w == 0;
        printf("%s %-28s dense %zu batch-strided %zu channel-sliced %zu w-sliced %zu\n", ok ? "PASS" : "FAIL",
               "noncontiguous_views", n_dense, n_batch, n_chan, n_w);
        failed += !ok;
    }
    // a14: the plan cache (AlgorithmCache equivalent) must serve the repeated exec / get_workspace_in_bytes
    // calls of the cases above
    {
        size_t hits = 0, misses = 0;
        b200::plan_cache_stats(&hits, &misses);
        const bool ok = hits > 0 && misses > 0;
        printf("%s %-28s hits %zu misses %zu\n", ok ? "PASS" : "FAIL", "plan_cache", hits, misses);
        failed += !ok;
    }
    // host cost of one exec() through the C++ binding (check_exec + descriptor + cached plan + launch),
    // measured on an enqueue-only loop: the stream is synchronised outside the timed region
    {
        auto dut = b200::create_conv_bias_forward(host.get(), stream);
        auto& p = dut->param();
        p.format = F::NHWC;
        p.nonlineMode = NL::RELU;
        p.compute_mode = param::ConvBias::ComputeMode::FLOAT32;
        DType bf = dtype::BFloat16();
        TensorLayout s({8, 14, 14, 256}, bf), f({256, 1, 1, 256}, bf), b({1, 1, 1, 256}, bf), z{bf}, d({8, 14, 14, 256}, bf);
        DevBuf ds(s.span().dist_byte()), df(f.span().dist_byte()), db(b.span().dist_byte()), dd(d.span().dist_byte());
        CUDA_OK(cudaMemsetAsync(ds.p, 0, s.span().dist_byte(), stream));
        CUDA_OK(cudaMemsetAsync(df.p, 0, f.span().dist_byte(), stream));
        CUDA_OK(cudaMemsetAsync(db.p, 0, b.span().dist_byte(), stream));
        const size_t ws = dut->get_workspace_in_bytes(s, f, b, z, d, nullptr);
        DevBuf dws(ws);
        auto run = [&](int n) {
            for (int i = 0; i < n; ++i)
                dut->exec(TensorND(ds.p, s), TensorND(df.p, f), TensorND(db.p, b), TensorND(nullptr, z),
                          TensorND(dd.p, d), nullptr, Workspace((dt_byte*)dws.p, ws));
        };
        run(20);
        CUDA_OK(cudaStreamSynchronize(stream));
        const int reps = 2000;
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        run(reps);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        CUDA_OK(cudaStreamSynchronize(stream));
        const double us = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 1e3 / reps;
        printf("INFO host_us_per_exec %.2f (C++ binding, enqueue only, %d execs)\n", us, reps);
    }
    try {
        failed += run_nn_cases(host.get(), stream);
    } catch (const std::exception& e) {
        printf("FAIL %-28s exception: %s\n", "nn_operators", e.what());
        ++failed;
    }
    printf("dropin_driver: %d failed\n", failed);
    return failed ? 1 : 0;
}
