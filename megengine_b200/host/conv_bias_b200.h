// megdnn::ConvBiasForward on B200 -- the C++ host side above the C-ABI (include/b200cb.h).
//
// A concrete subclass of the REFERENCE's abstract operator (dnn/include/megdnn/oprs/nn.h:358-536),
// compiled against the reference's own public headers, so every caller of megdnn::ConvBiasForward --
// src/opr ConvBiasForward::scn_do_execute (src/opr/impl/dnn/convolution.cpp:804-877), fastrun
// (src/rdnn/impl/algo_chooser.cpp), DnnOprCaller (imperative/src/impl/dnn_op_helper.h:254-272), the
// test Checker (dnn/test/common/checker.h) -- drives it through the interface it already uses:
// param(), deduce_layout, check_exec, get_workspace_in_bytes, exec, the weight-preprocess trio,
// get_all_algorithms_info(_safe), get_algorithm_info_heuristic, execution_policy().
//
// It is what dnn/src/cuda/conv_bias/opr_impl.{h,cpp} + the AlgoB200 class of INTEGRATION.md reduce to
// when only this algorithm family is registered: layout checks are the reference's (inherited from
// dnn/src/common/conv_bias.cpp), algorithm objects wrap the library's algorithm table, exec enqueues
// on the caller's CUDA stream and returns (no synchronisation, no allocation, errors -> megdnn_throw).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "b200cb.h"
#include "megdnn/handle.h"
#include "megdnn/oprs.h"

namespace megdnn {
namespace b200 {

class ConvBiasForwardImpl final : public ConvBiasForward {
public:
    //! `handle` only parents the operator (OperatorBase keeps a Handle*); kernels run on `stream`
    ConvBiasForwardImpl(Handle* handle, cudaStream_t stream);

    void exec(
            _megdnn_tensor_in src, _megdnn_tensor_in filter, _megdnn_tensor_in bias,
            _megdnn_tensor_in z, _megdnn_tensor_out dst,
            const PreprocessedFilter* preprocessed_filter, _megdnn_workspace workspace) override;
    void exec_preprocess(
            const TensorLayout& src_layout, _megdnn_tensor_in filter, _megdnn_tensor_in bias,
            const TensorLayout& z_layout, const TensorLayout& dst_layout,
            PreprocessedFilter* preprocessed_filter, _megdnn_workspace workspace) override;
    size_t get_workspace_in_bytes(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst,
            const PreprocessedFilter* preprocessed_filter) override;
    size_t get_preprocess_workspace_in_bytes(
            const TensorLayout&, const TensorLayout&, const TensorLayout&, const TensorLayout&,
            const TensorLayout&) override {
        return 0;
    }
    SmallVector<TensorLayout> deduce_preprocessed_filter_layout(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst) override;

    const char* get_algorithm_set_name() const override { return "CONV_BIAS_B200"; }
    Algorithm* get_algorithm_from_desc(const AlgorithmDesc& desc) override;

    //! the POD handed to the C-ABI; public so that tests can compare it with their own
    b200cb_desc make_desc(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst) const;

protected:
    std::vector<Algorithm*> get_all_algorithms(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst) override;
    std::vector<Algorithm*> get_all_algorithms_safe(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst) override;
    Algorithm* get_algorithm_heuristic(
            const TensorLayout& src, const TensorLayout& filter, const TensorLayout& bias,
            const TensorLayout& z, const TensorLayout& dst, size_t workspace_limit_in_bytes,
            const AlgoAttribute& positive_attr, const AlgoAttribute& negative_attr) override;

private:
    //! execution_policy().algo if set (fastrun / user), else the library heuristic
    void resolve_plan(const b200cb_desc& desc, b200cb_plan* plan);
    cudaStream_t m_stream;
};

//! hit / miss counters of the process-wide (descriptor, forced algorithm) -> plan cache (the
//! AlgorithmCache equivalent, dnn/include/megdnn/algorithm_cache.h:12-80)
void plan_cache_stats(size_t* hits, size_t* misses);

//! what handle->create_operator<ConvBiasForward>() returns for this backend
std::unique_ptr<ConvBiasForward> create_conv_bias_forward(Handle* handle, cudaStream_t stream);

}  // namespace b200
}  // namespace megdnn
