// tc_common.cuh -- helpers shared by the tcgen05 kernels (gemm_tc2.cuh, gemm_tc2h.cuh): shared-memory addresses, mbarrier
// primitives with a bounded wait, and the bf16 x 3 split of FP32 values.
//
// There is no FP32 tcgen05 kind, so FP32-accurate products are built from BF16 splits:
//     a = a1 + a2 + a3 (each bf16, 3 x 8 = 24 significand bits),
//     a*b ~= a1b1 + a1b2 + a2b1 + a1b3 + a3b1 + a2b2      (6 kind::f16 MMAs, FP32 accumulation in TMEM)
// (A first engine that split the operands inside the consuming CTA and accumulated everything in TMEM lived here in
// round 1; it missed the 1e-5 wavefield bar -- the tensor core truncates when it adds into its FP32 accumulator,
// profiles/exp_tc_accum_r01.txt -- and was removed from the product surface.  gemm_tc2.cuh is the engine.)
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ust {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug becomes a launch failure instead of a hang
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}

struct Split3 { uint32_t w[3]; };  // bf16x2 words of the three split planes for two consecutive k
__device__ __forceinline__ Split3 split2(float a0, float a1) {
    Split3 s;
    __nv_bfloat162 h1 = __floats2bfloat162_rn(a0, a1);
    float r0 = a0 - __low2float(h1), r1 = a1 - __high2float(h1);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(r0, r1);
    r0 -= __low2float(h2); r1 -= __high2float(h2);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(r0, r1);
    s.w[0] = *reinterpret_cast<uint32_t*>(&h1);
    s.w[1] = *reinterpret_cast<uint32_t*>(&h2);
    s.w[2] = *reinterpret_cast<uint32_t*>(&h3);
    return s;
}

}  // namespace tc
}  // namespace ust
