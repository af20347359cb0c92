// factor.cuh -- two-sided block-tridiagonal factorisation with explicit block inverses.
//
//   down chain: S_i = D_i - L_i T_{i-1} U_{i-1}          T_i = S_i^{-1}
//   up   chain: S_i = D_i - U_i T_{i+1} L_{i+1}
//   middle    : S_m = D_m - L_m T_{m-1} U_{m-1} - U_m T_{m+1} L_{m+1}
// (replaces the LU half of SuperLU gssv behind solve_helmholtz.py:15-18.)
//
// Each T_i is formed by an in-place-equivalent blocked Gauss-Jordan inversion (block GJ_NB=64, no
// inter-block pivoting: the Schur complements are well conditioned, cond ~ 9, first PML row ~ 7e2),
// ping-ponging between the T slot and a scratch block so that no kernel reads what it writes:
//   step k:  P  = inv(X_kk)
//            R  = P * Xtilde_k,:          (Xtilde = X with block column k replaced by e_k blocks)
//            X' = Xtilde - X_:,k * R      (rows i != k);    X'_k,: = R
// schur_kernel     : builds S_i from T_prev and the coefficient planes (O(n^2), 3x3 stencil on T_prev)
// gj_pivot_kernel   : pivot-block inverse in shared memory (one CTA per chain)
// gj_rowpanel_kernel: row panel R = P * Xtilde_k,:
// gj_update_kernel : rank-64 update of every other block row (the GEMM-shaped part)
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"

namespace ust {

template <typename R>
struct FactorArgs {
    Geom g;
    int phase, step, nbatch;
    const cx<R>* planes;  // [nfreq][9][Ny][Nx]
    cx<R>* T;             // [nfreq][M][nP*nP]
    cx<R>* scratch;       // [2*nfreq][nP*nP]
    cx<R>* pbuf;          // [2*nfreq][64*64] pivot-block inverses (transposed)
    int* status;
    // TMA-fed tensor-core update (gemm_tc2.cuh, complex64 only; null otherwise)
    uint16_t* Rp;         // [2*nfreq] B planes of the row panel R (64 x nP), bplanes layout with 4 k-chunks
    uint16_t* Cp;         // [2*nfreq][6][nP/8][8][8][8] A planes of the column panel X_:,k (nP x 64)
    size_t rp_stride;     // elements per batch entry of Rp
};

// buffer holding X^{(k)} for batch entry z working on block row `row`
template <typename R>
__device__ __forceinline__ cx<R>* gj_buffer(const FactorArgs<R>& a, int z, int freq, int row, int k) {
    const size_t bs = (size_t)a.g.nP * a.g.nP;
    cx<R>* slot = a.T + ((size_t)freq * a.g.M + row) * bs;
    cx<R>* scr = a.scratch + (size_t)z * bs;
    const int nblk = a.g.nP / GJ_NB;
    const bool k_even = (k & 1) == 0;
    const bool slot_holds_even = (nblk & 1) == 0;  // X^{(nblk)} must land in the T slot
    return (k_even == slot_holds_even) ? slot : scr;
}

template <typename R>
__device__ __forceinline__ cx<R> schur_term(const FactorArgs<R>& a, const cx<R>* __restrict__ planes_f,
                                            const cx<R>* __restrict__ Tp, int lkind, int ly, int rkind, int ry, int ai, int bi) {
    const int nI = a.g.nI, nP = a.g.nP;
    cx<R> l[3], r[3];
    tri3<R>(planes_f, a.g, lkind, false, ly, ai, l[0], l[1], l[2]);  // L[a, a-1..a+1]
    tri3<R>(planes_f, a.g, rkind, false, ry, bi, r[0], r[1], r[2]);  // U[b-1..b+1, b]
    cx<R> s = cxzero<R>();
#pragma unroll
    for (int dp = 0; dp < 3; ++dp) {
        int p = ai - 1 + dp;
        if (p < 0 || p >= nI) continue;
        cx<R> rowacc = cxzero<R>();
#pragma unroll
        for (int dq = 0; dq < 3; ++dq) {
            int q = bi - 1 + dq;
            if (q < 0 || q >= nI) continue;
            cmac(rowacc, Tp[(size_t)p * nP + q], r[dq]);
        }
        cmac(s, l[dp], rowacc);
    }
    return s;
}

template <typename R>
__global__ void __launch_bounds__(256) schur_kernel(FactorArgs<R> a) {
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z), dir = chain_dir(a.phase, z);
    const int bi = blockIdx.x * 16 + threadIdx.x;  // column (fast)
    const int ai = blockIdx.y * 16 + threadIdx.y;  // row
    const int nI = a.g.nI, nP = a.g.nP, M = a.g.M;
    if (ai >= nP || bi >= nP) return;
    cx<R>* X0 = gj_buffer(a, z, freq, row, 0);
    cx<R> v;
    if (ai < nI && bi < nI) {
        const size_t pl = (size_t)a.g.Nx * a.g.Ny;
        const cx<R>* planes_f = a.planes + (size_t)freq * 9 * pl;
        const int y = row + 1;
        const size_t o = (size_t)y * a.g.Nx + (ai + 1);
        v = cxzero<R>();
        if (bi == ai) v = planes_f[PL_C * pl + o];
        else if (bi == ai - 1) v = planes_f[PL_L * pl + o];
        else if (bi == ai + 1) v = planes_f[PL_R * pl + o];
        const size_t bs = (size_t)nP * nP;
        if ((dir == 0 || dir == 2) && row > 0) {
            const cx<R>* Tp = a.T + ((size_t)freq * M + (row - 1)) * bs;
            v = v - schur_term(a, planes_f, Tp, TRI_L, y, TRI_UC, y - 1, ai, bi);
        }
        if ((dir == 1 || dir == 2) && row < M - 1) {
            const cx<R>* Tp = a.T + ((size_t)freq * M + (row + 1)) * bs;
            v = v - schur_term(a, planes_f, Tp, TRI_U, y, TRI_LC, y + 1, ai, bi);
        }
    } else {
        v = (ai == bi) ? cxone<R>() : cxzero<R>();
    }
    X0[(size_t)ai * nP + bi] = v;
}

// ---------------------------------------------------------------------------------------------
// Pivot: P = inv(X_kk) by unpivoted Gauss-Jordan, one CTA per chain (the latency-bound part of the
// factorisation: 64 dependent steps).  The block lives in REGISTERS: thread (i = tid/4, q = tid%4) holds
// columns 16q..16q+15 of row i of G = X_kk^T (the inverse of a transpose is the transpose of the inverse;
// the row-panel kernel wants P[r][kk] = G[kk][r] with unit stride).  Per step p the four owners of row p
// scale and publish it through a double-buffered shared row (ONE barrier per step); every thread takes its
// multiplier G[i][p] (the pivot itself for the owners) from its row partner with a shuffle and updates 16 entries.
// grid = (1, 1, nbatch), 256 threads, dynamic smem = gj_pivot_smem<R>().
// ---------------------------------------------------------------------------------------------
template <typename R>
constexpr int gj_pivot_qs() { return sizeof(R) == 4 ? 18 : 17; }  // padded quarter stride: the four quarters of a row land in different banks
template <typename R>
constexpr size_t gj_pivot_smem() { return sizeof(cx<R>) * 2 * 4 * gj_pivot_qs<R>(); }

template <typename R>
__global__ void __launch_bounds__(256, 1) gj_pivot_kernel(FactorArgs<R> a, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int QS = sizeof(R) == 4 ? 18 : 17;
    cx<R>(*rowbuf)[4 * QS] = reinterpret_cast<cx<R>(*)[4 * QS]>(smem_raw);  // [2][4 quarters][QS] scaled pivot rows, double buffered
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const cx<R>* __restrict__ Xc = gj_buffer(a, z, freq, row, k);
    const int k0 = k * GJ_NB;
    const int tid = threadIdx.x;
    const int i = tid >> 2, q = tid & 3, lane = tid & 31;
    cx<R> g[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = Xc[(size_t)(k0 + 16 * q + c) * nP + k0 + i];  // G[i][16q+c] = X_kk[16q+c][i]
    bool bad = false;
#pragma unroll 1
    for (int pq = 0; pq < 4; ++pq) {
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) {
            const int p = 16 * pq + pp;
            cx<R>* rb = rowbuf[p & 1];
            // G[i][p] lives in register pp of the row partner that owns column quarter pq: the multiplier of row i,
            // and for the owners of row p the pivot itself
            cx<R> m;
            m.re = __shfl_sync(0xffffffffu, g[pp].re, (lane & ~3) | pq);
            m.im = __shfl_sync(0xffffffffu, g[pp].im, (lane & ~3) | pq);
            const bool own = (q == pq);
            if (i == p) {  // scale the pivot row, publish it
                const R mag = m.re * m.re + m.im * m.im;
                if (!(mag > R(0)) || isinf(mag)) bad = true;  // zero, NaN or overflowing pivot
                const cx<R> ip = crecip(m);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    g[c] = (own && c == pp) ? ip : g[c] * ip;
                    rb[QS * q + c] = g[c];
                }
            }
            __syncthreads();
            if (i != p) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const cx<R> rj = rb[QS * q + c];
                    const cx<R> old = (own && c == pp) ? cxzero<R>() : g[c];
                    g[c] = old - m * rj;
                }
            }
        }
    }
    if (bad) atomicOr(a.status, 1);
    cx<R>* Pg = a.pbuf + (size_t)z * GJ_NB * GJ_NB;
#pragma unroll
    for (int c = 0; c < 16; ++c) Pg[i * GJ_NB + 16 * q + c] = g[c];
}

// Row panel: R_j = P * Xtilde_kj written into block row k of X'.  grid = (nblk, 1, nbatch), 256 threads,
// dynamic smem = 2 * 64x64 complex.
template <typename R>
__global__ void __launch_bounds__(256) gj_rowpanel_kernel(FactorArgs<R> a, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<R>(*G)[GJ_NB] = reinterpret_cast<cx<R>(*)[GJ_NB]>(smem_raw);                                   // G[kk][r] = P[r][kk]
    cx<R>(*Tl)[GJ_NB] = reinterpret_cast<cx<R>(*)[GJ_NB]>(smem_raw + sizeof(cx<R>) * GJ_NB * GJ_NB);  // Xtilde_kj tile
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const cx<R>* __restrict__ Xc = gj_buffer(a, z, freq, row, k);
    cx<R>* __restrict__ Xn = gj_buffer(a, z, freq, row, k + 1);
    const cx<R>* __restrict__ Pg = a.pbuf + (size_t)z * GJ_NB * GJ_NB;
    const int k0 = k * GJ_NB, j0 = blockIdx.x * GJ_NB;
    const int tid = threadIdx.x;
    for (int e = tid; e < GJ_NB * GJ_NB; e += 256) {
        int r = e / GJ_NB, c = e % GJ_NB;
        G[r][c] = Pg[e];
        cx<R> tv;
        if (j0 == k0) tv = (r == c) ? cxone<R>() : cxzero<R>();
        else tv = Xc[(size_t)(k0 + r) * nP + j0 + c];
        Tl[r][c] = tv;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    cx<R> acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = cxzero<R>();
#pragma unroll 8
    for (int kk = 0; kk < GJ_NB; ++kk) {
        cx<R> av[4], bv[4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            av[2 * c] = G[kk][c * 32 + ty * 2];
            av[2 * c + 1] = G[kk][c * 32 + ty * 2 + 1];
            bv[2 * c] = Tl[kk][c * 32 + tx * 2];
            bv[2 * c + 1] = Tl[kk][c * 32 + tx * 2 + 1];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) cmac(acc[i][jj], av[i], bv[jj]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = (i >> 1) * 32 + ty * 2 + (i & 1);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            int c = (jj >> 1) * 32 + tx * 2 + (jj & 1);
            Xn[(size_t)(k0 + r) * nP + j0 + c] = acc[i][jj];
        }
    }
    if constexpr (sizeof(R) == 4) {
        if (a.Rp) {
            // the update GEMM's B operand: R as bf16 x 3 planes.  Stage the tile in shared memory (Tl is free once
            // every thread has left the k loop) so that a thread owns 8 consecutive k of one column.
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int r = (i >> 1) * 32 + ty * 2 + (i & 1);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    int c = (jj >> 1) * 32 + tx * 2 + (jj & 1);
                    Tl[r][c] = acc[i][jj];
                }
            }
            __syncthreads();
            const int nloc = tid & 63;
            const int n = j0 + nloc;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kg = (tid >> 6) + 4 * h;
                float re[8], im[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { cx<float> v = Tl[kg * 8 + c][nloc]; re[c] = v.re; im[c] = v.im; }
                uint16_t* chunk = a.Rp + (size_t)z * a.rp_stride +
                                  ((size_t)(n / tc2::TN) * (GJ_NB / tc2::KC) + (kg >> 1)) * (tc2::B_STAGE / 2);
                tc2::store_b8(chunk, n % tc2::TN, kg & 1, re, im);
            }
        }
    }
}

// X'_ij = Xtilde_ij - X_ik R_j for block rows i != k.  grid = (nblk, nblk-1, nbatch).
template <typename R>
__global__ void __launch_bounds__(256) gj_update_kernel(FactorArgs<R> a, int k) {
    __shared__ GemmSmem<R, GJ_NB, GJ_NB> sm;
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const cx<R>* Xc = gj_buffer(a, z, freq, row, k);
    cx<R>* Xn = gj_buffer(a, z, freq, row, k + 1);
    const int ib = blockIdx.y + (blockIdx.y >= k ? 1 : 0);
    GemmTile<R> t;
    t.A = Xc + k * GJ_NB; t.lda = nP;
    t.B = Xn + (size_t)k * GJ_NB * nP; t.ldb = nP;
    t.Cin = Xc; t.ldcin = nP;
    t.Cout = Xn; t.ldc = nP;
    t.M = nP; t.N = nP; t.K = GJ_NB; t.Mstore = nP;
    t.m0 = ib * GJ_NB; t.n0 = blockIdx.x * GJ_NB;
    t.mask_lo = k * GJ_NB; t.mask_hi = (k + 1) * GJ_NB;
    t.sgn = R(-1);
    cgemm_tile<R, GJ_NB, GJ_NB, false>(t, sm);
}

// tensor-core variant (complex64): 128x128 tiles over the whole matrix; the pivot block row (which the
// panel kernel already wrote) is skipped in the epilogue.  grid = (ceil(nP/128), ceil(nP/128), nbatch)
__global__ void __launch_bounds__(tc::NUM_THREADS, 1) tc_gj_update_kernel(FactorArgs<float> a, int k) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const cx<float>* Xc = gj_buffer(a, z, freq, row, k);
    cx<float>* Xn = gj_buffer(a, z, freq, row, k + 1);
    GemmTile<float> t;
    t.A = Xc + k * GJ_NB; t.lda = nP;
    t.B = Xn + (size_t)k * GJ_NB * nP; t.ldb = nP;
    t.Cin = Xc; t.ldcin = nP;
    t.Cout = Xn; t.ldc = nP;
    t.M = nP; t.N = nP; t.K = GJ_NB; t.Mstore = nP;
    t.m0 = blockIdx.y * tc::TM; t.n0 = blockIdx.x * tc::TN;
    t.mask_lo = k * GJ_NB; t.mask_hi = (k + 1) * GJ_NB;
    t.sgn = -1.f;
    tc::TcExtra ex; ex.skip_lo = k * GJ_NB; ex.skip_hi = (k + 1) * GJ_NB;
    tc::cgemm_tile<false>(t, ex, tc_smem);
}

// Column panel X_:,k (nP x 64, FP32) -> bf16 x 3 A planes of the TMA-fed update.  grid = (nP/32, 1, nbatch), 256 threads:
// thread = (block row I, block column J, row r in the block), a warp writes 512 contiguous bytes per plane.
__global__ void __launch_bounds__(256) gj_colsplit_kernel(FactorArgs<float> a, int k) {
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const cx<float>* __restrict__ Xc = gj_buffer(a, z, freq, row, k);
    const int tid = threadIdx.x;
    const int I = blockIdx.x * 4 + (tid >> 6), J = (tid >> 3) & 7, r = tid & 7;
    const int rr = I * 8 + r;
    if (rr >= nP) return;
    const cx<float>* src = Xc + (size_t)rr * nP + k * GJ_NB + J * 8;
    float re[8], im[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { cx<float> v = src[c]; re[c] = v.re; im[c] = v.im; }
    uint32_t wr[3][4], wi[3][4];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        tc::Split3 sr = tc::split2(re[2 * qd], re[2 * qd + 1]);
        tc::Split3 si = tc::split2(im[2 * qd], im[2 * qd + 1]);
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) { wr[sp][qd] = sr.w[sp]; wi[sp][qd] = si.w[sp]; }
    }
    const size_t plane = (size_t)nP * GJ_NB;  // elements per plane
    uint16_t* dst = a.Cp + (size_t)z * tc2::NPL_A * plane + ((size_t)I * 8 + J) * 64 + r * 8;
#pragma unroll
    for (int sp = 0; sp < 3; ++sp) {
        *reinterpret_cast<uint4*>(dst + sp * plane) = make_uint4(wr[sp][0], wr[sp][1], wr[sp][2], wr[sp][3]);
        *reinterpret_cast<uint4*>(dst + (3 + sp) * plane) = make_uint4(wi[sp][0], wi[sp][1], wi[sp][2], wi[sp][3]);
    }
}

// TMA-fed tensor-core update: X' = Xtilde - X_:,k R over the whole matrix except the pivot block row.
// grid = (ceil(nP/128), ceil(nP/128), nbatch), 576 threads.
__global__ void __launch_bounds__(tc2::NUM_THREADS, 1) tc2_gj_update_kernel(FactorArgs<float> a, int k, float bias_fix,
                                                                             const __grid_constant__ CUtensorMap cmap) {
    extern __shared__ __align__(1024) unsigned char tc2_smem[];
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    tc2::Tc2Tile t;
    t.bplanes = a.Rp + (size_t)z * a.rp_stride;
    t.amat = z;
    t.Cin = gj_buffer(a, z, freq, row, k); t.ldcin = nP;
    t.Cout = gj_buffer(a, z, freq, row, k + 1); t.ldc = nP;
    t.M = nP; t.N = nP; t.K = GJ_NB; t.Mstore = nP;
    t.m0 = blockIdx.y * tc2::TM; t.n0 = blockIdx.x * tc2::TN;
    t.mask_lo = k * GJ_NB; t.mask_hi = (k + 1) * GJ_NB;
    t.skip_lo = k * GJ_NB; t.skip_hi = (k + 1) * GJ_NB;
    t.sgn = -1.f;
    t.bias_fix = bias_fix;
    tc2::cgemm_tile<false>(t, &cmap, tc2_smem);
}

// Split the finished block inverse T_row (FP32) into the bf16 x 3 operand planes of the TMA-fed engine
// (gemm_tc2.cuh); padding rows/columns (>= nI) are written as zero.  grid = (nP/256 rounded up, nP/8, nbatch).
__global__ void __launch_bounds__(256) t_split_kernel(FactorArgs<float> a, uint16_t* __restrict__ Tp) {
    const int z = blockIdx.z;
    const int row = chain_row(a.g, a.phase, z, a.step);
    if (row < 0) return;
    const int freq = chain_freq(a.phase, z);
    const int nP = a.g.nP;
    const size_t mat = (size_t)freq * a.g.M + row;
    tc2::a_split_body(a.T + mat * (size_t)nP * nP, nP, a.g.nI, a.g.nI, Tp + mat * (size_t)tc2::NPL_A * nP * nP, nP);
}

}  // namespace ust
