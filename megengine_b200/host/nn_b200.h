// The operators either side of ConvBias on B200 -- C++ host side above include/b200cb_nn.h.
//
// Concrete subclasses of the REFERENCE's abstract operators, compiled against the reference's own public headers
// (same recipe as conv_bias_b200.h): megdnn::TypeCvtForward and megdnn::ElemwiseForward
// (dnn/include/megdnn/oprs/general.h:19-72,890-900) and megdnn::PoolingForward (dnn/include/megdnn/oprs/nn.h:900-948),
// so that their callers -- src/opr TypeCvt / Elemwise / Pooling nodes, DnnOprCaller, the test Checker -- drive them
// through the interfaces they already use.  Layout checks are the reference's (check_exec / deduce_layout inherited
// from dnn/src/common); exec enqueues on the caller's CUDA stream and returns; what the C-ABI reports as
// unsupported (broadcast operands, non-contiguous views, other dtypes / modes) is thrown like megdnn_throw.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "b200cb_nn.h"
#include "megdnn/handle.h"
#include "megdnn/oprs.h"

namespace megdnn {
namespace b200 {

class TypeCvtImpl final : public TypeCvtForward {
public:
    TypeCvtImpl(Handle* handle, cudaStream_t stream) : TypeCvtForward(handle), m_stream(stream) {}
    void exec(_megdnn_tensor_in src, _megdnn_tensor_out dst) override;

private:
    cudaStream_t m_stream;
};

class ElemwiseForwardImpl final : public ElemwiseForward {
public:
    ElemwiseForwardImpl(Handle* handle, cudaStream_t stream) : ElemwiseForward(handle), m_stream(stream) {}
    void exec(_megdnn_in const TensorNDArray& src, _megdnn_tensor_out dst) override;

private:
    cudaStream_t m_stream;
};

class PoolingForwardImpl final : public PoolingForward {
public:
    PoolingForwardImpl(Handle* handle, cudaStream_t stream) : PoolingForward(handle), m_stream(stream) {}
    void exec(_megdnn_tensor_in src, _megdnn_tensor_out dst, _megdnn_workspace workspace) override;
    size_t get_workspace_in_bytes(const TensorLayout&, const TensorLayout&) override { return 0; }
    const char* get_algorithm_set_name() const override { return "POOLING_B200"; }
    Algorithm* get_algorithm_from_desc(const AlgorithmDesc& desc) override;
    //! the POD handed to the C-ABI (throws when the layouts are not a cell of b200cb_pooling)
    b200cb_pool_desc make_desc(const TensorLayout& src, const TensorLayout& dst) const;

protected:
    std::vector<Algorithm*> get_all_algorithms(const TensorLayout& src, const TensorLayout& dst) override;
    std::vector<Algorithm*> get_all_algorithms_safe(const TensorLayout& src, const TensorLayout& dst) override;
    Algorithm* get_algorithm_heuristic(
            const TensorLayout& src, const TensorLayout& dst, size_t workspace_limit_in_bytes,
            const AlgoAttribute& positive_attr, const AlgoAttribute& negative_attr) override;

private:
    cudaStream_t m_stream;
};

std::unique_ptr<TypeCvtForward> create_type_cvt(Handle* handle, cudaStream_t stream);
std::unique_ptr<ElemwiseForward> create_elemwise(Handle* handle, cudaStream_t stream);
std::unique_ptr<PoolingForward> create_pooling(Handle* handle, cudaStream_t stream);

}  // namespace b200
}  // namespace megdnn
