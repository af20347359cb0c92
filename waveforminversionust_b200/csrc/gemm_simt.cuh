// gemm_simt.cuh -- batched complex GEMM tile engine on the CUDA cores (float / double).
//
// This is the engine for complex128 (there is no FP64 tcgen05 kind) and the always-available
// engine for complex64.  For double the inner product runs on the FP64 tensor-core instruction
// (mma.sync.m8n8k4.f64, "DMMA"): same FP64 FMA arithmetic per product, but a warp reads 12 doubles per
// lane and k4 step from shared memory for 32 MMAs instead of 64 for the same work on the FMA pipe --
// the FMA form was shared-memory bound at ~15 TFLOP/s.  One CTA (256 threads) computes a BM x BN complex tile
//     Cout = (Cin ? Cin : 0) + sgn * op(A) * B,   op(A) = A  or  conj(A)^T
// with a register-staged double-buffered smem pipeline over K.  Thread tile (BM/16) x (BN/16)
// complex, interleaved in 2-element chunks so smem reads are conflict-free 16-byte vectors.
#pragma once
#include "common.cuh"

namespace ust {

template <typename R> struct GemmCfg;
template <> struct GemmCfg<float>  { static constexpr int BK = 16; static constexpr int VEC = 2; };
template <> struct GemmCfg<double> { static constexpr int BK = 8;  static constexpr int VEC = 1; };

// 16-byte pack of VEC complex numbers (the unit of every vectorised global/shared access)
template <typename R> struct Pack;
template <> struct alignas(16) Pack<float>  { cx<float> v[2]; };
template <> struct alignas(16) Pack<double> { cx<double> v[1]; };

template <typename R, int BM, int BN>
struct alignas(16) GemmSmem {
    static constexpr int BK = GemmCfg<R>::BK;
    cx<R> As[2][BK][BM];
    cx<R> Bs[2][BK][BN];
};

// double: real and imaginary parts in separate planes (the DMMA fragments are scalars), rows padded by 8 doubles so that the
// four k-rows of a fragment load fall into two 128-byte bank windows = the two wavefronts an 8-byte warp load needs anyway
template <int BM, int BN>
struct alignas(16) GemmSmem<double, BM, BN> {
    static constexpr int BK = GemmCfg<double>::BK;
    double Ar[2][BK][BM + 8], Ai[2][BK][BM + 8];
    double Br[2][BK][BN + 8], Bi[2][BK][BN + 8];
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Problem description for one tile.
template <typename R>
struct GemmTile {
    const cx<R>* A; int lda;      // !TA: A[m*lda+k]       TA: op(A)[m][k] = conj(A[k*lda+m])
    const cx<R>* B; int ldb;      // B[k*ldb+n]
    const cx<R>* Cin; int ldcin;  // nullable
    cx<R>* Cout; int ldc;
    int M, N, K;                  // logical bounds: m<M, n<N, k<K are valid
    int Mstore;                   // rows m<Mstore are written
    int m0, n0;                   // tile origin
    int mask_lo, mask_hi;         // Cin columns n in [mask_lo, mask_hi) read as zero
    R sgn;
};

template <typename R, int BM, int BN, bool TA>
__device__ __forceinline__ void cgemm_tile(const GemmTile<R>& t, GemmSmem<R, BM, BN>& sm) {
    constexpr int BK = GemmCfg<R>::BK;
    constexpr int VEC = GemmCfg<R>::VEC;
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int CM = (TM >= 2) ? TM / 2 : 1;  // chunks of 2 along m
    constexpr int CN = (TN >= 2) ? TN / 2 : 1;
    static_assert(TM == 2 || TM == 4, "BM must be 32 or 64");
    static_assert(TN == 2 || TN == 4, "BN must be 32 or 64");
    constexpr int A_LOADS = BM * BK / (256 * VEC);
    constexpr int B_LOADS = BN * BK / (256 * VEC);
    static_assert(A_LOADS >= 1 && B_LOADS >= 1, "tile too small");

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    const bool a_vec = ((t.lda % VEC) == 0) && ((((uintptr_t)t.A) & 15) == 0);
    const bool b_vec = ((t.ldb % VEC) == 0) && ((((uintptr_t)t.B) & 15) == 0) && ((t.n0 % VEC) == 0);

    Pack<R> ra[A_LOADS], rb[B_LOADS];

    auto load_a = [&](int k0) {
#pragma unroll
        for (int j = 0; j < A_LOADS; ++j) {
            int idx = tid + j * 256;
            if (!TA) {
                constexpr int KV = BK / VEC;
                int m = idx / KV, kv = idx % KV;
                int gm = t.m0 + m, gk = k0 + kv * VEC;
                const cx<R>* p = t.A + (size_t)gm * t.lda + gk;
                if (a_vec && gm < t.M && gk + VEC <= t.K) {
                    ra[j] = *reinterpret_cast<const Pack<R>*>(p);
                } else {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) ra[j].v[q] = (gm < t.M && gk + q < t.K) ? p[q] : cxzero<R>();
                }
            } else {
                constexpr int MV = BM / VEC;
                int k = idx / MV, mv = idx % MV;
                int gk = k0 + k, gm = t.m0 + mv * VEC;
                const cx<R>* p = t.A + (size_t)gk * t.lda + gm;
                if (a_vec && gk < t.K && gm + VEC <= t.M) {
                    ra[j] = *reinterpret_cast<const Pack<R>*>(p);
                } else {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) ra[j].v[q] = (gk < t.K && gm + q < t.M) ? p[q] : cxzero<R>();
                }
#pragma unroll
                for (int q = 0; q < VEC; ++q) ra[j].v[q] = conj(ra[j].v[q]);
            }
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int j = 0; j < B_LOADS; ++j) {
            int idx = tid + j * 256;
            constexpr int NV = BN / VEC;
            int k = idx / NV, nv = idx % NV;
            int gk = k0 + k, gn = t.n0 + nv * VEC;
            const cx<R>* p = t.B + (size_t)gk * t.ldb + gn;
            if (b_vec && gk < t.K && gn + VEC <= t.N) {
                rb[j] = *reinterpret_cast<const Pack<R>*>(p);
            } else {
#pragma unroll
                for (int q = 0; q < VEC; ++q) rb[j].v[q] = (gk < t.K && gn + q < t.N) ? p[q] : cxzero<R>();
            }
        }
    };
    auto store_ab = [&](int buf) {
        if constexpr (sizeof(R) == 8) {
#pragma unroll
            for (int j = 0; j < A_LOADS; ++j) {
                const int idx = tid + j * 256;
                const int m = !TA ? idx / BK : idx % BM, k = !TA ? idx % BK : idx / BM;
                sm.Ar[buf][k][m] = ra[j].v[0].re; sm.Ai[buf][k][m] = ra[j].v[0].im;
            }
#pragma unroll
            for (int j = 0; j < B_LOADS; ++j) {
                const int idx = tid + j * 256;
                const int k = idx / BN, n = idx % BN;
                sm.Br[buf][k][n] = rb[j].v[0].re; sm.Bi[buf][k][n] = rb[j].v[0].im;
            }
        } else {
#pragma unroll
        for (int j = 0; j < A_LOADS; ++j) {
            int idx = tid + j * 256;
            if (!TA) {
                constexpr int KV = BK / VEC;
                int m = idx / KV, kv = idx % KV;
#pragma unroll
                for (int q = 0; q < VEC; ++q) sm.As[buf][kv * VEC + q][m] = ra[j].v[q];
            } else {
                constexpr int MV = BM / VEC;
                int k = idx / MV, mv = idx % MV;
                *reinterpret_cast<Pack<R>*>(&sm.As[buf][k][mv * VEC]) = ra[j];
            }
        }
#pragma unroll
        for (int j = 0; j < B_LOADS; ++j) {
            int idx = tid + j * 256;
            constexpr int NV = BN / VEC;
            int k = idx / NV, nv = idx % NV;
            *reinterpret_cast<Pack<R>*>(&sm.Bs[buf][k][nv * VEC]) = rb[j];
        }
        }
    };

    const int nk = (t.K + BK - 1) / BK;
    if constexpr (sizeof(R) == 8) {
        // ---- FP64 tensor-core path: warp grid 2 (m) x 4 (n); a warp owns MT x NT tiles of 8 x 8 ----
        constexpr int MT = BM / 16, NT = BN / 32;
        const int lane = tid & 31, warp = tid >> 5;
        const int wm = (warp & 1) * (BM / 2), wn = (warp >> 1) * (BN / 4);
        const int fr = lane >> 2, fk = lane & 3;  // fragment row (A) / column (B), fragment k
        double cr[MT][NT][2], ci[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) { cr[i][j][0] = cr[i][j][1] = 0.0; ci[i][j][0] = ci[i][j][1] = 0.0; }
        load_a(0);
        load_b(0);
        store_ab(0);
        __syncthreads();
        for (int kt = 0; kt < nk; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < nk) {
                load_a((kt + 1) * BK);
                load_b((kt + 1) * BK);
            }
#pragma unroll
            for (int k4 = 0; k4 < BK; k4 += 4) {
                double ar[MT], ai[MT], br[NT], bi[NT], nbi[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) { ar[i] = sm.Ar[buf][k4 + fk][wm + 8 * i + fr]; ai[i] = sm.Ai[buf][k4 + fk][wm + 8 * i + fr]; }
#pragma unroll
                for (int j = 0; j < NT; ++j) { br[j] = sm.Br[buf][k4 + fk][wn + 8 * j + fr]; bi[j] = sm.Bi[buf][k4 + fk][wn + 8 * j + fr]; nbi[j] = -bi[j]; }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                        dmma884(cr[i][j][0], cr[i][j][1], ai[i], nbi[j]);
                        dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
                        dmma884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
                    }
            }
            if (kt + 1 < nk) {
                store_ab(buf ^ 1);
                __syncthreads();
            }
        }
        // epilogue: lane holds C[row = lane / 4][columns 2 (lane % 4), + 1] of every 8 x 8 tile
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int m = t.m0 + wm + 8 * i + fr;
            if (m >= t.Mstore) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int n = t.n0 + wn + 8 * j + 2 * fk + q;
                    if (n >= t.N) continue;
                    cx<R> c = cxzero<R>();
                    if (t.Cin && !(n >= t.mask_lo && n < t.mask_hi)) c = t.Cin[(size_t)m * t.ldcin + n];
                    c.re += t.sgn * cr[i][j][q];
                    c.im += t.sgn * ci[i][j][q];
                    t.Cout[(size_t)m * t.ldc + n] = c;
                }
        }
    } else {
    cx<R> acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = cxzero<R>();

    load_a(0);
    load_b(0);
    store_ab(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_a((kt + 1) * BK);
            load_b((kt + 1) * BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            cx<R> a[TM], b[TN];
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                a[2 * c] = sm.As[buf][kk][c * (BM / CM) + ty * 2];
                a[2 * c + 1] = sm.As[buf][kk][c * (BM / CM) + ty * 2 + 1];
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                b[2 * c] = sm.Bs[buf][kk][c * (BN / CN) + tx * 2];
                b[2 * c + 1] = sm.Bs[buf][kk][c * (BN / CN) + tx * 2 + 1];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) cmac(acc[i][j], a[i], b[j]);
        }
        if (kt + 1 < nk) {
            store_ab(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = t.m0 + (i >> 1) * (BM / CM) + ty * 2 + (i & 1);
        if (m >= t.Mstore) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = t.n0 + (j >> 1) * (BN / CN) + tx * 2 + (j & 1);
            if (n >= t.N) continue;
            cx<R> c = cxzero<R>();
            if (t.Cin && !(n >= t.mask_lo && n < t.mask_hi)) c = t.Cin[(size_t)m * t.ldcin + n];
            c.re += t.sgn * acc[i][j].re;
            c.im += t.sgn * acc[i][j].im;
            t.Cout[(size_t)m * t.ldc + n] = c;
        }
    }
    }  // FMA path
}

}  // namespace ust
