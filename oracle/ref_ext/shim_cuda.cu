// TEST INFRASTRUCTURE ONLY -- compiles the reference's own DCNv3 CUDA translation unit, unmodified and from where it
// lies (/root/reference/network/ops_dcnv3/src/cuda/dcnv3_cuda.cu, passed as -DGP_REF_CUDA_TU), against torch 2.11.
//
// The reference dispatches with AT_DISPATCH_FLOATING_TYPES_AND_HALF(input.type(), ...) (dcnv3_cuda.cu:69-70,147-148);
// torch >= 2.x no longer converts at::DeprecatedTypeProperties to c10::ScalarType, so the stock macro does not compile.
// Nothing of the reference is copied or edited: the torch headers are included first (their include guards make the
// reference's own #includes no-ops), the one dispatch macro is re-pointed at an overload that accepts both types, and
// then the reference file is #included verbatim.
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <ATen/cuda/CUDAContext.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <torch/torch.h>

namespace gp_ref_shim {
inline c10::ScalarType scalar_type_of(c10::ScalarType t) { return t; }
inline c10::ScalarType scalar_type_of(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}   // namespace gp_ref_shim

#undef AT_DISPATCH_FLOATING_TYPES_AND_HALF
#define AT_DISPATCH_FLOATING_TYPES_AND_HALF(TYPE, NAME, ...) \
    AT_DISPATCH_SWITCH(::gp_ref_shim::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES_AND_HALF(__VA_ARGS__))

#include GP_REF_CUDA_TU
