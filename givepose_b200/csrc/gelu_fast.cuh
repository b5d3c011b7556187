// givepose_b200 -- GELU for 16-bit storage, shared by gn_apply_kernel (posenet_kernels.cuh) and the operand transform of the
// tcgen05 convolution (conv3x3_tc.cu): both must produce the same bits.
//
// 0.5 x (1 + tanh(x (a + b x^2 + c x^4))) with (a, b, c) fitted to the exact erf GELU (max abs deviation 2.5e-5 on [-8, 8]; the
// argument is clamped to +-6 where tanh has saturated) and the hardware tanh.approx.f32 (rel. error 2^-11): total error
// < 2.5e-4 |x|, an order of magnitude below the bf16 rounding of the stored result.
#pragma once

namespace gp {

__device__ __forceinline__ float gelu_fast16(float x) {
    const float xc = fminf(fmaxf(x, -6.f), 6.f);
    const float x2 = xc * xc;
    const float u = xc * fmaf(x2, fmaf(x2, -3.51523083e-4f, 3.70056758e-2f), 7.97507859e-1f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

}  // namespace gp
