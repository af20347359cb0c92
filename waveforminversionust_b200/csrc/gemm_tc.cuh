// gemm_tc.cuh -- complex64 GEMM tile engine on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//     Cout = (Cin ? Cin : 0) + sgn * op(A) * B,   op(A) = A or conj(A)^T,   complex64 in / complex64 out
//
// There is no FP32 tcgen05 kind, so FP32-accurate products are built from BF16 splits:
//     a = a1 + a2 + a3 (each bf16, 3 x 8 = 24 significand bits),
//     a*b ~= a1b1 + a1b2 + a2b1 + a1b3 + a3b1 + a2b2      (6 kind::f16 MMAs, FP32 accumulation in TMEM)
// The tensor core truncates when it adds into the FP32 accumulator (measured: error grows linearly with
// the number of accumulating MMAs), so the leading a1b1 terms and the five small correction terms go to
// two separate TMEM accumulators that are summed in the epilogue: the big accumulator then sees 2 instead
// of 12 truncating additions per 16 k.
// and the complex product from real ones with the accumulator laid out as [Cr | Ci] (256 fp32 columns):
//     [Cr | Ci] += Ar * [Br | Bi]  +  Ai * [-Bi | Br].
// One CTA = one 128 x 128 complex tile.  Warps 0-7 stream fp32 operands from global memory, split them
// and write K-major, non-swizzled UMMA operand planes into a 3-stage shared-memory ring (16 complex k per
// stage); warp 8 issues tcgen05.mma (one elected thread) and releases stages with tcgen05.commit; the
// accumulator (128 lanes x 256 columns) lives in TMEM and is read back with tcgen05.ld for the epilogue.
// The operands are converted in flight (global -> registers -> smem), which is why they are not fed by
// TMA: the tensor map cannot split a float into three bf16 planes.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace ust {
namespace tc {

constexpr int TM = 128;     // tile rows (complex)
constexpr int TN = 128;     // tile columns (complex); accumulator columns = 2*TN
constexpr int KC = 16;      // complex k per stage = one UMMA K (16 bf16)
constexpr int STAGES = 3;
constexpr int A_PLANE = TM * KC * 2;          // bytes of one bf16 plane [128 x 16]
constexpr int B_PLANE = 2 * TN * KC * 2;      // bytes of one bf16 plane [256 x 16]
constexpr int STAGE_BYTES = 6 * A_PLANE + 6 * B_PLANE;
constexpr int NUM_THREADS = 288;              // 8 producer/epilogue warps + 1 MMA warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;  // + alignment slack + barriers
constexpr uint32_t TMEM_COLS = 512;  // D1 = main a1*b1 terms (cols 0..255), D2 = the five correction terms (256..511)
// K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units:
// element (row r, k) of a plane sits at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2
constexpr uint32_t LBO = 128, SBO = 256;

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

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(LBO >> 4) << 16;
    d |= (uint64_t)(SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

// kind::f16, BF16 x BF16 -> F32, M=128, N=256, both operands K-major
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// one task = 8 consecutive complex k of one operand line (row of op(A) or column of B), held in registers
struct Line8 { float re[8], im[8]; };

// split a line and store the six planes' 16-byte chunks.  plane p of the A block: Ar1,Ar2,Ar3,Ai1,Ai2,Ai3.
__device__ __forceinline__ void store_a_line(uint32_t stage_base, int row, int c, const Line8& v) {
    uint32_t wr[3][4], wi[3][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        Split3 sr = split2(v.re[2 * q], v.re[2 * q + 1]);
        Split3 si = split2(v.im[2 * q], v.im[2 * q + 1]);
#pragma unroll
        for (int s = 0; s < 3; ++s) { wr[s][q] = sr.w[s]; wi[s][q] = si.w[s]; }
    }
    const uint32_t off = (uint32_t)(row >> 3) * SBO + (uint32_t)c * LBO + (uint32_t)(row & 7) * 16;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        st_shared_v4(stage_base + s * A_PLANE + off, wr[s][0], wr[s][1], wr[s][2], wr[s][3]);
        st_shared_v4(stage_base + (3 + s) * A_PLANE + off, wi[s][0], wi[s][1], wi[s][2], wi[s][3]);
    }
}
// B block planes: B1_s = [Br_s | Bi_s] (rows n and 128+n), B2_s = [-Bi_s | Br_s]
__device__ __forceinline__ void store_b_line(uint32_t b_base, int n, int c, const Line8& v) {
    uint32_t wr[3][4], wi[3][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        Split3 sr = split2(v.re[2 * q], v.re[2 * q + 1]);
        Split3 si = split2(v.im[2 * q], v.im[2 * q + 1]);
#pragma unroll
        for (int s = 0; s < 3; ++s) { wr[s][q] = sr.w[s]; wi[s][q] = si.w[s]; }
    }
    const uint32_t off_lo = (uint32_t)(n >> 3) * SBO + (uint32_t)c * LBO + (uint32_t)(n & 7) * 16;
    const uint32_t off_hi = off_lo + (uint32_t)(TN >> 3) * SBO;
    const uint32_t NEG = 0x80008000u;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        st_shared_v4(b_base + s * B_PLANE + off_lo, wr[s][0], wr[s][1], wr[s][2], wr[s][3]);
        st_shared_v4(b_base + s * B_PLANE + off_hi, wi[s][0], wi[s][1], wi[s][2], wi[s][3]);
        st_shared_v4(b_base + (3 + s) * B_PLANE + off_lo, wi[s][0] ^ NEG, wi[s][1] ^ NEG, wi[s][2] ^ NEG, wi[s][3] ^ NEG);
        st_shared_v4(b_base + (3 + s) * B_PLANE + off_hi, wr[s][0], wr[s][1], wr[s][2], wr[s][3]);
    }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// row range [skip_lo, skip_hi) of the output is left untouched (Gauss-Jordan pivot block row)
struct TcExtra { int skip_lo, skip_hi; };

template <bool TA>
__device__ __forceinline__ void cgemm_tile(const GemmTile<float>& t, const TcExtra& ex, unsigned char* smem_raw) {
    typedef cx<float> C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // carve shared memory: [stages][A planes | B planes] then barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 8); mbar_init(empty_bar(s), 1); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot_ptr;

    const int nk = (t.K + KC - 1) / KC;

    if (warp < 8) {
        // ---------------- producers: global fp32 -> bf16 split planes in smem ----------------
        const int line = tid & 127;  // A row / B column within the tile
        const int c = tid >> 7;      // which half (8 k) of the 16-k stage
        Line8 va, vb, na, nb;
        auto load_lines = [&](int kt, Line8& a, Line8& b) {
            const int k0 = kt * KC + c * 8;
            // ---- A line ----
            const int gm = t.m0 + line;
            if (!TA) {
                const C* p = t.A + (size_t)gm * t.lda + k0;
                const bool rowok = gm < t.M;
                if (rowok && k0 + 8 <= t.K && ((t.lda & 1) == 0) && ((((uintptr_t)t.A) & 15) == 0)) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 f = *reinterpret_cast<const float4*>(p + 2 * q);
                        a.re[2 * q] = f.x; a.im[2 * q] = f.y; a.re[2 * q + 1] = f.z; a.im[2 * q + 1] = f.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        C z = (rowok && k0 + q < t.K) ? p[q] : cxzero<float>();
                        a.re[q] = z.re; a.im[q] = z.im;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int gk = k0 + q;
                    C z = (gk < t.K && gm < t.M) ? t.A[(size_t)gk * t.lda + gm] : cxzero<float>();
                    a.re[q] = z.re; a.im[q] = -z.im;  // conj(A)^T
                }
            }
            // ---- B line ----
            const int gn = t.n0 + line;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int gk = k0 + q;
                C z = (gk < t.K && gn < t.N) ? t.B[(size_t)gk * t.ldb + gn] : cxzero<float>();
                b.re[q] = z.re; b.im[q] = z.im;
            }
        };
        load_lines(0, va, vb);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % STAGES;
            const uint32_t use = (uint32_t)(kt / STAGES);
            if (kt + 1 < nk) load_lines(kt + 1, na, nb);  // next stage's loads fly while this one is converted
            mbar_wait(empty_bar(s), (use & 1u) ^ 1u);
            const uint32_t sb = smem_base + s * STAGE_BYTES;
            store_a_line(sb, line, c, va);
            store_b_line(sb + 6 * A_PLANE, line, c, vb);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
            va = na; vb = nb;
        }
    } else {
        // ---------------- MMA issuer: one elected thread ----------------
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % STAGES;
            const uint32_t use = (uint32_t)(kt / STAGES);
            mbar_wait(full_bar(s), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sb = smem_base + s * STAGE_BYTES;
                const uint32_t bb = sb + 6 * A_PLANE;
                // split pairs (i,j): a_i * b_j for i+j <= 4
                const int pi[6] = {0, 0, 1, 0, 2, 1};
                const int pj[6] = {0, 1, 0, 2, 0, 1};
                const uint32_t acc = kt > 0 ? 1u : 0u;
                const uint32_t d1 = tmem_acc, d2 = tmem_acc + 2 * TN;
                umma(d1, make_desc(sb), make_desc(bb), acc);                             // Ar1 * [Br1|Bi1]
                umma(d1, make_desc(sb + 3 * A_PLANE), make_desc(bb + 3 * B_PLANE), 1u);  // Ai1 * [-Bi1|Br1]
#pragma unroll
                for (int q = 1; q < 6; ++q) {
                    umma(d2, make_desc(sb + pi[q] * A_PLANE), make_desc(bb + pj[q] * B_PLANE), (q == 1) ? acc : 1u);
                    umma(d2, make_desc(sb + (3 + pi[q]) * A_PLANE), make_desc(bb + (3 + pj[q]) * B_PLANE), 1u);
                }
                umma_commit(empty_bar(s));                   // frees the stage when these MMAs retire
                if (kt == nk - 1) umma_commit(accum_bar);    // accumulator complete
            }
            __syncwarp();
        }
    }

    // ---------------- epilogue: TMEM -> registers -> global ----------------
    if (warp < 8) {
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_base = (warp & 3) * 32;   // TMEM lanes this warp may read
        const int nhalf = (warp >> 2) * 64;      // columns [nhalf, nhalf+64) of the complex tile
        const int m = t.m0 + lane_base + lane;
        const bool row_ok = m < t.Mstore && !(m >= ex.skip_lo && m < ex.skip_hi);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const int ncol = nhalf + ch * 16;
            uint32_t rr[16], ri[16], cr[16], ci[16];
            const uint32_t ta_r = tmem_acc + ((uint32_t)lane_base << 16) + (uint32_t)ncol;
            tmem_ld16(ta_r, rr);
            tmem_ld16(ta_r + TN, ri);
            tmem_ld16(ta_r + 2 * TN, cr);
            tmem_ld16(ta_r + 3 * TN, ci);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __uint_as_float(cr[j]));
                ri[j] = __float_as_uint(__uint_as_float(ri[j]) + __uint_as_float(ci[j]));
            }
            if (row_ok) {
                const bool vec_ok = ((t.ldc & 1) == 0) && ((((uintptr_t)t.Cout) & 15) == 0) &&
                                    (!t.Cin || (((t.ldcin & 1) == 0) && ((((uintptr_t)t.Cin) & 15) == 0)));
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const int n = t.n0 + ncol + j;
                    if (n >= t.N) continue;
                    const bool pair = vec_ok && (n + 1 < t.N) && ((n & 1) == 0);
                    C c0 = cxzero<float>(), c1 = cxzero<float>();
                    if (t.Cin) {
                        const C* ci = t.Cin + (size_t)m * t.ldcin + n;
                        if (pair) {
                            float4 f = *reinterpret_cast<const float4*>(ci);
                            c0 = C(f.x, f.y); c1 = C(f.z, f.w);
                        } else {
                            c0 = ci[0];
                            if (n + 1 < t.N) c1 = ci[1];
                        }
                        if (n >= t.mask_lo && n < t.mask_hi) c0 = cxzero<float>();
                        if (n + 1 >= t.mask_lo && n + 1 < t.mask_hi) c1 = cxzero<float>();
                    }
                    c0.re += t.sgn * __uint_as_float(rr[j]);     c0.im += t.sgn * __uint_as_float(ri[j]);
                    c1.re += t.sgn * __uint_as_float(rr[j + 1]); c1.im += t.sgn * __uint_as_float(ri[j + 1]);
                    C* co = t.Cout + (size_t)m * t.ldc + n;
                    if (pair) {
                        *reinterpret_cast<float4*>(co) = make_float4(c0.re, c0.im, c1.re, c1.im);
                    } else {
                        co[0] = c0;
                        if (n + 1 < t.N) co[1] = c1;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace ust
