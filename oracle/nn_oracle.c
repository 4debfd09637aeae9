/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product (libb200cb.so).
 *
 * CPU restatement of the reference's naive operators either side of ConvBias, the
 * checker of include/b200cb_nn.h.  Plain C, one function per operator, each citing what it restates:
 *   oracle_pooling    dnn/src/naive/pooling/opr_impl.cpp:17-110,196-292 (MaxPooler, MeanIncludePooler,
 *                     MeanExcludePooler, MeanIncludeRoundedPooler / MeanExcludeRoundedPooler for QuantizedS8,
 *                     pooling_forward_impl loop order n, c, oh, ow, fh, fw)
 *   oracle_elemwise   dnn/src/common/elemwise/kern_defs.cuh:169-270 (ctype arithmetic)
 *   oracle_elemwise_q dnn/src/naive/elemwise_multi_type/opr_impl.h:45-120 + dnn/include/megdnn/dtype.h:845-854
 *   oracle_typecvt    dnn/src/naive/type_cvt/opr_impl.cpp:24-57, dtype.h:845-897
 *   oracle_deconv     dnn/src/naive/convolution/opr_impl.cpp:92-167 + helper.h:669-681 (StrategyBwdData scatter,
 *                     loop order n, oc, oh, ow, fh, fw, ic; fp32 / int32 accumulator, final TypeCvt)
 * Pinned by tests/test_nn_oracle.py against the reference's own compiled operators (oracle/_ref, live) and the
 * committed golden outputs of those operators (tests/golden/nn_golden.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "b200cb_nn.h"

uint16_t oracle_f2h(float v);
uint16_t oracle_f2bf(float v);
float oracle_h2f(uint16_t h);
float oracle_bf2f(uint16_t h);

static float ld(const void* p, uint64_t i, uint32_t dt) {
    switch (dt) {
        case B200CB_DT_F32: return ((const float*)p)[i];
        case B200CB_DT_F16: return oracle_h2f(((const uint16_t*)p)[i]);
        case B200CB_DT_BF16: return oracle_bf2f(((const uint16_t*)p)[i]);
        case B200CB_DT_QS8: return (float)((const int8_t*)p)[i];
        default: return (float)((const int32_t*)p)[i];
    }
}
/* v rounded to dtype dt and widened again */
static float rnd(float v, uint32_t dt) {
    if (dt == B200CB_DT_F16)
        return oracle_h2f(oracle_f2h(v));
    if (dt == B200CB_DT_BF16)
        return oracle_bf2f(oracle_f2bf(v));
    return v;
}
static void st(void* p, uint64_t i, uint32_t dt, float v) {
    switch (dt) {
        case B200CB_DT_F32: ((float*)p)[i] = v; break;
        case B200CB_DT_F16: ((uint16_t*)p)[i] = oracle_f2h(v); break;
        case B200CB_DT_BF16: ((uint16_t*)p)[i] = oracle_f2bf(v); break;
        case B200CB_DT_QS8: ((int8_t*)p)[i] = (int8_t)v; break;
        default: ((int32_t*)p)[i] = (int32_t)v; break;
    }
}
/* DTypeParamImpl<dt_qint8>::quantize (dtype.h:845-853) */
static int8_t quant_s8(float in, float scale) {
    float v = in / scale;
    v = roundf(v);
    v = fminf(fmaxf(-128.f, v), 127.f);
    return (int8_t)v;
}
/* DTypeParamImpl<dt_qint32>::quantize (dtype.h:888-896) */
static int32_t quant_s32(float in, float scale) {
    float v = in / scale;
    v = roundf(v);
    v = fminf(fmaxf(-2147483648.f, v), 2147483520.f);
    return (int32_t)v;
}

static uint64_t pool_idx(uint32_t fmt, uint64_t n, uint64_t c, uint64_t h, uint64_t w, uint64_t C, uint64_t H, uint64_t W) {
    if (fmt == B200CB_FORMAT_NHWC)
        return ((n * H + h) * W + w) * C + c;
    if (fmt == B200CB_FORMAT_NCHW)
        return ((n * C + c) * H + h) * W + w;
    return (((n * (C >> 2) + (c >> 2)) * H + h) * W + w) * 4 + (c & 3);   /* NCHW4IdxGetter */
}

int oracle_pooling(const b200cb_pool_desc* d, const void* src, void* dst) {
    const uint32_t dt = d->dtype;
    for (uint32_t n = 0; n < d->n; ++n)
        for (uint32_t c = 0; c < d->c; ++c)
            for (uint32_t oh = 0; oh < d->oh; ++oh)
                for (uint32_t ow = 0; ow < d->ow; ++ow) {
                    float fsum = 0.f, best = 0.f;
                    int32_t isum = 0;
                    uint32_t fed = 0;
                    if (dt == B200CB_DT_F32)
                        best = -3.402823466e+38f;
                    else if (dt == B200CB_DT_F16)
                        best = -65504.f;
                    else if (dt == B200CB_DT_BF16)
                        best = oracle_bf2f(0xff7f);
                    else
                        best = -128.f;
                    for (uint32_t fh = 0; fh < d->window_h; ++fh)
                        for (uint32_t fw = 0; fw < d->window_w; ++fw) {
                            const uint32_t ih = oh * d->stride_h + fh - d->pad_h;   /* size_t wrap in the reference */
                            const uint32_t iw = ow * d->stride_w + fw - d->pad_w;
                            if (ih < d->ih && iw < d->iw) {
                                const float x = ld(src, pool_idx(d->format, n, c, ih, iw, d->c, d->ih, d->iw), dt);
                                ++fed;
                                best = best > x ? best : x;                /* MaxPooler::feed */
                                if (dt == B200CB_DT_QS8)
                                    isum += (int32_t)x;                    /* Mean*RoundedPooler: int32 sum */
                                else
                                    fsum = rnd(fsum + x, dt);              /* sum += x in ctype */
                            }
                        }
                    if (!fed)
                        return 1;   /* "The pooling window lies outside completely" */
                    float out;
                    if (d->mode == B200CB_POOL_MAX) {
                        out = best;
                    } else if (dt == B200CB_DT_QS8) {
                        const float cnt = d->mode == B200CB_POOL_AVERAGE ? (float)(int32_t)(d->window_h * d->window_w)
                                                                         : (float)fed;
                        out = roundf((float)isum / cnt);
                    } else {
                        const float cnt = d->mode == B200CB_POOL_AVERAGE ? rnd((float)(d->window_h * d->window_w), dt)
                                                                         : rnd((float)fed, dt);
                        out = fsum / cnt;   /* ctype division: rounded to ctype by st() */
                    }
                    st(dst, pool_idx(d->format, n, c, oh, ow, d->c, d->oh, d->ow), dt, out);
                }
    return 0;
}

static float hswish(float x) {
    return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
}

int oracle_elemwise(uint32_t mode, uint32_t dt, uint64_t n, const void* a, const void* b, void* dst) {
    for (uint64_t i = 0; i < n; ++i) {
        const float x = ld(a, i, dt);
        const float y = b ? ld(b, i, dt) : 0.f;
        float r;
        switch (mode) {
            case B200CB_ELEM_RELU: r = x <= 0.f ? 0.f : x; break;
            case B200CB_ELEM_ADD: r = x + y; break;
            case B200CB_ELEM_MAX: r = x > y ? x : y; break;
            case B200CB_ELEM_MUL: r = x * y; break;
            case B200CB_ELEM_FUSE_ADD_RELU: {
                const float s = rnd(x + y, dt);   /* (x + y) is a ctype value before the comparison */
                r = s <= 0.f ? 0.f : s;
                break;
            }
            case B200CB_ELEM_H_SWISH: r = hswish(x); break;
            default: return 1;
        }
        st(dst, i, dt, r);
    }
    return 0;
}

int oracle_elemwise_q(
        uint32_t mode, uint64_t n, const int8_t* a, float sa, const int8_t* b, float sb, int8_t* dst, float sd) {
    for (uint64_t i = 0; i < n; ++i) {
        const float x = a[i] * sa;                 /* dequantize: in.as_int8() * scale */
        const float y = b ? b[i] * sb : 0.f;
        float r;
        switch (mode) {
            case B200CB_ELEM_RELU: r = x <= 0.f ? 0.f : x; break;
            case B200CB_ELEM_ADD: r = x + y; break;
            case B200CB_ELEM_FUSE_ADD_RELU: r = (x + y) <= 0.f ? 0.f : (x + y); break;
            case B200CB_ELEM_H_SWISH: r = hswish(x); break;
            default: return 1;
        }
        dst[i] = quant_s8(r, sd);
    }
    return 0;
}

int oracle_typecvt(uint32_t sdt, float ss, uint32_t ddt, float sd, uint64_t n, const void* src, void* dst) {
    for (uint64_t i = 0; i < n; ++i) {
        float v = ld(src, i, sdt);
        if (sdt == B200CB_DT_QS8 || sdt == B200CB_DT_QS32)
            v = v * ss;   /* (float)int * scale; int32 -> float rounds to nearest */
        if (ddt == B200CB_DT_QS8)
            ((int8_t*)dst)[i] = quant_s8(v, sd);
        else if (ddt == B200CB_DT_QS32)
            ((int32_t*)dst)[i] = quant_s32(v, sd);
        else
            st(dst, i, ddt, v);
    }
    return 0;
}

/* ConvolutionBackwardData, NHWC dense.  `f` describes the forward convolution (b200cb_nn.h). */
int oracle_deconv(const b200cb_desc* f, const void* filter, const void* diff, void* grad) {
    if (f->format != B200CB_FORMAT_NHWC || f->group != 1)
        return 1;
    const uint64_t ge = (uint64_t)f->n * f->ih * f->iw * f->ic;
    const int is_q = f->filter_dtype == B200CB_DT_QS8;
    float* facc = is_q ? NULL : (float*)calloc(ge, sizeof(float));
    int32_t* iacc = is_q ? (int32_t*)calloc(ge, sizeof(int32_t)) : NULL;
    if (!facc && !iacc)
        return 2;
    const int flip = f->mode == B200CB_MODE_CONVOLUTION;
    for (uint32_t n = 0; n < f->n; ++n)
        for (uint32_t oc = 0; oc < f->oc; ++oc)
            for (uint32_t oh = 0; oh < f->oh; ++oh)
                for (uint32_t ow = 0; ow < f->ow; ++ow) {
                    const float dv = ld(diff, (((uint64_t)n * f->oh + oh) * f->ow + ow) * f->oc + oc, f->dst_dtype);
                    for (uint32_t fh = 0; fh < f->fh; ++fh)
                        for (uint32_t fw = 0; fw < f->fw; ++fw) {
                            const uint32_t ih = oh * f->stride_h + fh * f->dilate_h - f->pad_h;
                            const uint32_t iw = ow * f->stride_w + fw * f->dilate_w - f->pad_w;
                            if (ih >= f->ih || iw >= f->iw)
                                continue;
                            const uint32_t kh = flip ? f->fh - 1 - fh : fh, kw = flip ? f->fw - 1 - fw : fw;
                            for (uint32_t ic = 0; ic < f->ic; ++ic) {
                                const float wv = ld(
                                        filter, (((uint64_t)oc * f->fh + kh) * f->fw + kw) * f->ic + ic, f->filter_dtype);
                                const uint64_t gi = (((uint64_t)n * f->ih + ih) * f->iw + iw) * f->ic + ic;
                                if (is_q)
                                    iacc[gi] += (int32_t)dv * (int32_t)wv;
                                else
                                    facc[gi] += dv * wv;
                            }
                        }
                }
    for (uint64_t i = 0; i < ge; ++i) {
        if (is_q) {
            const float s = f->filter_scale * f->dst_scale;   /* QuantizedS32(filter.scale * diff.scale) */
            if (f->src_dtype == B200CB_DT_QS32)
                ((int32_t*)grad)[i] = iacc[i];
            else
                ((int8_t*)grad)[i] = quant_s8((float)iacc[i] * s, f->src_scale);
        } else {
            st(grad, i, f->src_dtype, facc[i]);
        }
    }
    free(facc);
    free(iacc);
    return 0;
}
