// sweep.cuh -- multi-RHS block sweeps on the stored block inverses (forward and adjoint systems).
//
// Two-sided elimination / back-substitution, in place on the (N, nrhs) array:
//   eliminate   z_i = op(T_i) (b_i - K_i z_prev)            down chain: prev=i-1, up chain: prev=i+1
//   middle      x_m = op(T_m) (b_m - K_lo z_{m-1} - K_hi z_{m+1})
//   back-subst  x_i = z_i - op(T_i) (K'_i x_next)           outwards from the middle
// with K = L_i / U_i for the forward system and U_{i-1}^H / L_{i+1}^H for the adjoint system
// (conj(H)^T, solve_helmholtz.py:66-73) and op(T) = T or T^H.
// Per block row: tri_apply_kernel (HBM/L2-bound, builds the GEMM's B operand W) + sweep_gemm_kernel.
// Replaces the triangular-solve half of SuperLU gssv behind solve_helmholtz.py:15-18,85-93.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc2.cuh"

namespace ust {

enum SweepMode { SW_ELIM = 0, SW_BACK = 1 };

template <typename R>
struct SweepArgs {
    Geom g;
    int phase, step, nbatch, mode, adjoint, nrhs;
    const cx<R>* planes;  // [nfreq][9][Ny][Nx]   (pre-offset to the first frequency of the batch)
    const cx<R>* T;       // [nfreq][M][nP*nP]
    cx<R>* W;             // [nbatch][nP*nrhs] scratch
    cx<R>* X;             // solution arrays, X + freq*x_stride, each (N, nrhs)
    size_t x_stride;
    int f0;               // plan frequency index of the batch's first frequency (operand planes are indexed by plan frequency)
    // One-hot right-hand sides (forward solves of the FWI loop): during elimination a 128-column tile whose sources all lie
    // further along the chain is still identically zero (X was zero-filled), so its products are skipped.
    // first_row[t] / last_row[t] = smallest / largest interior block row holding a source of column tile t; onehot = 0 disables.
    int onehot;
    short first_row[8], last_row[8];
};

template <typename R>
__device__ __forceinline__ bool sweep_tile_is_zero(const SweepArgs<R>& s, int dir, int row, int tn) {
    if (!s.onehot || s.mode != 0 /* SW_ELIM */ || s.phase != PH_CHAIN || tn >= 8) return false;
    return dir == 0 ? row < s.first_row[tn] : row > s.last_row[tn];
}

// which coupling terms a block row needs
struct Coupling { int kind, y, src_row; bool on; };

__device__ __forceinline__ void sweep_couplings(const Geom& g, int mode, int adjoint, int dir, int row, Coupling& lo, Coupling& hi) {
    // lo: term that involves block row row-1 ; hi: term that involves block row row+1
    lo.on = false; hi.on = false;
    lo.src_row = row - 1; hi.src_row = row + 1;
    if (!adjoint) { lo.kind = TRI_L; lo.y = row + 1; hi.kind = TRI_U; hi.y = row + 1; }
    else          { lo.kind = TRI_UC; lo.y = row;    hi.kind = TRI_LC; hi.y = row + 2; }
    const bool has_lo = row > 0, has_hi = row < g.M - 1;
    if (mode == SW_ELIM) {
        if (dir == 0) lo.on = has_lo;
        else if (dir == 1) hi.on = has_hi;
        else { lo.on = has_lo; hi.on = has_hi; }
    } else {  // back-substitution: rows above the middle look down (row+1), rows below look up
        if (dir == 0) hi.on = has_hi;
        else lo.on = has_lo;
    }
}

// W[a,t] = (ELIM ? b[a,t] : 0) -/+ sum_terms (K v)[a,t]   (ELIM: minus, BACK: plus)
template <typename R>
__global__ void __launch_bounds__(256) tri_apply_kernel(SweepArgs<R> s) {
    const int z = blockIdx.z;
    const int row = chain_row(s.g, s.phase, z, s.step);
    if (row < 0) return;
    const int freq = chain_freq(s.phase, z), dir = chain_dir(s.phase, z);
    const int nI = s.g.nI, nrhs = s.nrhs, Nx = s.g.Nx;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)nI * nrhs) return;
    const int a = (int)(idx / nrhs), t = (int)(idx % nrhs);
    const size_t pl = (size_t)s.g.Nx * s.g.Ny;
    const cx<R>* planes_f = s.planes + (size_t)freq * 9 * pl;
    cx<R>* Xf = s.X + (size_t)freq * s.x_stride;
    Coupling lo, hi;
    sweep_couplings(s.g, s.mode, s.adjoint, dir, row, lo, hi);
    cx<R> acc = cxzero<R>();
    Coupling cs[2] = {lo, hi};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (!cs[c].on) continue;
        cx<R> k0, k1, k2;
        tri3<R>(planes_f, s.g, cs[c].kind, s.adjoint != 0, cs[c].y, a, k0, k1, k2);
        const cx<R>* v = Xf + ((size_t)(cs[c].src_row + 1) * Nx + 1) * nrhs;  // interior slab of that grid row
        if (a > 0) cmac(acc, k0, v[(size_t)(a - 1) * nrhs + t]);
        cmac(acc, k1, v[(size_t)a * nrhs + t]);
        if (a < nI - 1) cmac(acc, k2, v[(size_t)(a + 1) * nrhs + t]);
    }
    cx<R> w;
    if (s.mode == SW_ELIM) {
        cx<R> b = Xf[((size_t)(row + 1) * Nx + 1 + a) * nrhs + t];
        w = b - acc;
    } else {
        w = acc;
    }
    s.W[(size_t)z * s.g.nP * nrhs + (size_t)a * nrhs + t] = w;
}

// X_row = (BACK ? X_row : 0) +/- op(T_row) W.   grid = (ceil(nrhs/BN), nP/BM, nbatch)
template <typename R, int BM, int BN, bool TA>
__global__ void __launch_bounds__(256) sweep_gemm_kernel(SweepArgs<R> s) {
    __shared__ GemmSmem<R, BM, BN> sm;
    const int z = blockIdx.z;
    const int row = chain_row(s.g, s.phase, z, s.step);
    if (row < 0) return;
    const int freq = chain_freq(s.phase, z);
    const int nP = s.g.nP, nI = s.g.nI, nrhs = s.nrhs;
    if ((int)blockIdx.y * BM >= nI) return;
    GemmTile<R> t;
    t.A = s.T + ((size_t)freq * s.g.M + row) * (size_t)nP * nP; t.lda = nP;
    t.B = s.W + (size_t)z * nP * nrhs; t.ldb = nrhs;
    cx<R>* out = s.X + (size_t)freq * s.x_stride + ((size_t)(row + 1) * s.g.Nx + 1) * nrhs;
    t.Cin = (s.mode == SW_BACK) ? out : nullptr; t.ldcin = nrhs;
    t.Cout = out; t.ldc = nrhs;
    t.M = nP; t.N = nrhs; t.K = nI; t.Mstore = nI;
    t.m0 = blockIdx.y * BM; t.n0 = blockIdx.x * BN;
    t.mask_lo = 0; t.mask_hi = 0;
    t.sgn = (s.mode == SW_BACK) ? R(-1) : R(1);
    cgemm_tile<R, BM, BN, TA>(t, sm);
}

// ---------------------------------------------------------------------------------------------
// TMA-fed tensor-core sweep (gemm_tc2.cuh).  tri_apply2_kernel is tri_apply_kernel with the output written
// as the GEMM's pre-split B planes (bf16 x 3, core-matrix layout) instead of FP32: one thread = one column n,
// eight consecutive rows a.  grid = (kpad/8, ceil(nrhs/128), nbatch), 128 threads.
// ---------------------------------------------------------------------------------------------
struct Tc2SweepExtra {
    uint16_t* Wp;        // [nbatch][bplanes_elems(kpad, nrhs)]
    size_t wp_stride;    // elements per batch entry
    int kpad;            // nI rounded up to 16
    float bias_fix;
    int drain_every;     // drain period (chunks) of the leading accumulator
    int prefetch_cin;    // back-substitution: L2 prefetch of the tile's Cin rows when the CTA starts
    int ksplit;          // 2 / 4 = launched as clusters of that many CTAs which split the k loop of a tile (few-tile launches), else 1
};

__global__ void __launch_bounds__(128) tri_apply2_kernel(SweepArgs<float> s, Tc2SweepExtra x) {
    pdl_trigger();
    pdl_wait();
    typedef float R;
    const int z = blockIdx.z;
    const int row = chain_row(s.g, s.phase, z, s.step);
    if (row < 0) return;
    const int freq = chain_freq(s.phase, z), dir = chain_dir(s.phase, z);
    const int nI = s.g.nI, nrhs = s.nrhs, Nx = s.g.Nx;
    const int kg = blockIdx.x, tn = blockIdx.y, r = threadIdx.x;
    if (sweep_tile_is_zero(s, dir, row, tn)) return;  // the matching GEMM tiles are skipped as well
    const int n = tn * tc2::TN + r;
    const int a0 = kg * 8;
    // the tridiagonal coefficients of the block's 8 rows are the same for all 128 columns: stage them once.  The wavefield
    // loads do not depend on them and are issued first, so that the two global round trips overlap.
    __shared__ cx<R> coef[2][8][3];
    const size_t pl = (size_t)s.g.Nx * s.g.Ny;
    const cx<R>* planes_f = s.planes + (size_t)freq * 9 * pl;
    Coupling lo, hi;
    sweep_couplings(s.g, s.mode, s.adjoint, dir, row, lo, hi);
    const Coupling cs[2] = {lo, hi};
    const bool active = n < nrhs && a0 < nI;
    const cx<R>* Xf = s.X + (size_t)freq * s.x_stride;
    cx<R> vv[2][10], bb[8];
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        const cx<R>* v = Xf + ((size_t)(cs[ci].src_row + 1) * Nx + 1) * nrhs + n;  // interior slab of that grid row, column n
#pragma unroll
        for (int c = 0; c < 10; ++c) {
            const int a = a0 - 1 + c;
            vv[ci][c] = (active && cs[ci].on && a >= 0 && a < nI) ? v[(size_t)a * nrhs] : cxzero<R>();
        }
    }
    {
        const cx<R>* b = Xf + ((size_t)(row + 1) * Nx + 1) * nrhs + n;
#pragma unroll
        for (int c = 0; c < 8; ++c) bb[c] = (active && s.mode == SW_ELIM && a0 + c < nI) ? b[(size_t)(a0 + c) * nrhs] : cxzero<R>();
    }
    if (r < 16) {
        const int ci = r >> 3, c = r & 7;
        cx<R> k0 = cxzero<R>(), k1 = cxzero<R>(), k2 = cxzero<R>();
        if (cs[ci].on && a0 + c < nI) tri3<R>(planes_f, s.g, cs[ci].kind, s.adjoint != 0, cs[ci].y, a0 + c, k0, k1, k2);
        coef[ci][c][0] = k0; coef[ci][c][1] = k1; coef[ci][c][2] = k2;
    }
    __syncthreads();
    float re[8], im[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { re[c] = 0.f; im[c] = 0.f; }
    if (active) {
        cx<R> acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = cxzero<R>();
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
            if (!cs[ci].on) continue;
#pragma unroll
            for (int c = 0; c < 8; ++c) {  // rows a >= nI carry zero coefficients
                cmac(acc[c], coef[ci][c][0], vv[ci][c]);
                cmac(acc[c], coef[ci][c][1], vv[ci][c + 1]);
                cmac(acc[c], coef[ci][c][2], vv[ci][c + 2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int a = a0 + c;
            if (a >= nI) continue;
            cx<R> w = acc[c];
            if (s.mode == SW_ELIM) w = bb[c] - acc[c];
            re[c] = w.re; im[c] = w.im;
        }
    }
    uint16_t* chunk = x.Wp + (size_t)z * x.wp_stride + ((size_t)tn * (x.kpad / tc2::KC) + (kg >> 1)) * (tc2::B_STAGE / 2);
    tc2::store_b8(chunk, r, kg & 1, re, im);
}

// X_row = (BACK ? X_row : 0) +/- op(T_row) W on the TMA/tcgen05 engine.  grid = (ceil(nrhs/128), ceil(nI/128), nbatch), 576 threads.
template <bool TA>
__global__ void __launch_bounds__(tc2::NUM_THREADS, 1) tc2_sweep_gemm_kernel(SweepArgs<float> s, Tc2SweepExtra x,
                                                                              const __grid_constant__ CUtensorMap amap) {
    extern __shared__ __align__(1024) unsigned char tc2_smem[];
    pdl_trigger();
    const int z = blockIdx.z;
    const int row = chain_row(s.g, s.phase, z, s.step);
    if (row < 0) { pdl_wait(); return; }  // an early exit must not let the grid complete before its predecessor
    const int freq = chain_freq(s.phase, z);
    const int nI = s.g.nI, nrhs = s.nrhs;
    const int tn = x.ksplit > 1 ? (int)(blockIdx.x / x.ksplit) : (int)blockIdx.x;  // the CTAs of a split-K cluster work on the same tile
    if (sweep_tile_is_zero(s, chain_dir(s.phase, z), row, tn)) { pdl_wait(); return; }  // the output rows stay zero
    tc2::Tc2Tile t;
    tc2::tile_no_emit(t);
    t.bplanes = x.Wp + (size_t)z * x.wp_stride;
    t.amat = (s.f0 + freq) * s.g.M + row;
    cx<float>* out = s.X + (size_t)freq * s.x_stride + ((size_t)(row + 1) * s.g.Nx + 1) * nrhs;
    t.Cin = (s.mode == SW_BACK) ? out : nullptr; t.ldcin = nrhs;
    t.Cout = out; t.ldc = nrhs;
    t.M = nI; t.N = nrhs; t.K = nI; t.Mstore = nI;
    t.m0 = blockIdx.y * tc2::TM; t.n0 = tn * tc2::TN;
    t.mask_lo = 0; t.mask_hi = 0; t.skip_lo = 0; t.skip_hi = 0;
    t.sgn = (s.mode == SW_BACK) ? -1.f : 1.f;
    t.bias_fix = x.bias_fix;
    t.drain_every = x.drain_every;
    t.prefetch_cin = x.prefetch_cin;
    t.ksplit = x.ksplit;
    tc2::cgemm_tile<TA>(t, &amap, tc2_smem);
}

// ---------------------------------------------------------------------------------------------
// Dirichlet ring (identity rows, solve_helmholtz.py:266-276).  The ring unknowns are eliminated from
// the factorisation; these two kernels restore the reference's semantics for general right-hand sides:
//   forward:  u_ring = b_ring,   b_int -= H[int,ring] b_ring          (ring_pre_kernel, before the sweeps)
//   adjoint:  x_ring = b_ring - H[int,ring]^H x_int                   (ring_post_adj_kernel, after)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int plane_of_offset(int dy, int dx) {
    // plane holding H[(x,y) -> (x+dx, y+dy)]
    if (dy == 0) return dx == 0 ? PL_C : (dx < 0 ? PL_L : PL_R);
    if (dy < 0) return dx == 0 ? PL_D : (dx < 0 ? PL_DL : PL_DR);
    return dx == 0 ? PL_U : (dx < 0 ? PL_UL : PL_UR);
}

__device__ __forceinline__ bool ring_node(const Geom& g, int x, int y) { return x == 0 || y == 0 || x == g.Nx - 1 || y == g.Ny - 1; }

// enumerate the "inner ring" (interior nodes adjacent to the Dirichlet ring): count = 2*nI + 2*(M-2)
__device__ __forceinline__ void inner_ring_node(const Geom& g, int i, int& x, int& y) {
    const int nI = g.nI, M = g.M;
    if (i < nI) { x = 1 + i; y = 1; return; }
    i -= nI;
    if (i < nI) { x = 1 + i; y = M; return; }
    i -= nI;
    if (i < M - 2) { x = 1; y = 2 + i; return; }
    i -= M - 2;
    x = nI; y = 2 + i;
}
// enumerate the Dirichlet ring: count = 2*Nx + 2*(Ny-2)
__device__ __forceinline__ void outer_ring_node(const Geom& g, int i, int& x, int& y) {
    if (i < g.Nx) { x = i; y = 0; return; }
    i -= g.Nx;
    if (i < g.Nx) { x = i; y = g.Ny - 1; return; }
    i -= g.Nx;
    if (i < g.Ny - 2) { x = 0; y = 1 + i; return; }
    i -= g.Ny - 2;
    x = g.Nx - 1; y = 1 + i;
}

template <typename R>
__global__ void __launch_bounds__(256) ring_pre_kernel(Geom g, const cx<R>* __restrict__ planes_f, cx<R>* __restrict__ X, int nrhs, int count) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)count * nrhs) return;
    const int i = (int)(idx / nrhs), t = (int)(idx % nrhs);
    int x, y;
    inner_ring_node(g, i, x, y);
    const size_t pl = (size_t)g.Nx * g.Ny;
    cx<R> acc = cxzero<R>();
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            int xx = x + dx, yy = y + dy;
            if (!ring_node(g, xx, yy)) continue;
            cx<R> c = planes_f[plane_of_offset(dy, dx) * pl + (size_t)y * g.Nx + x];
            cmac(acc, c, X[((size_t)yy * g.Nx + xx) * nrhs + t]);
        }
    size_t o = ((size_t)y * g.Nx + x) * nrhs + t;
    X[o] = X[o] - acc;
}

template <typename R>
__global__ void __launch_bounds__(256) ring_post_adj_kernel(Geom g, const cx<R>* __restrict__ planes_f, cx<R>* __restrict__ X, int nrhs, int count) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)count * nrhs) return;
    const int i = (int)(idx / nrhs), t = (int)(idx % nrhs);
    int x, y;
    outer_ring_node(g, i, x, y);
    const size_t pl = (size_t)g.Nx * g.Ny;
    cx<R> acc = cxzero<R>();
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            int xq = x + dx, yq = y + dy;  // interior neighbour q whose row references this ring node
            if (xq < 1 || yq < 1 || xq > g.Nx - 2 || yq > g.Ny - 2) continue;
            cx<R> c = conj(planes_f[plane_of_offset(-dy, -dx) * pl + (size_t)yq * g.Nx + xq]);
            cmac(acc, c, X[((size_t)yq * g.Nx + xq) * nrhs + t]);
        }
    size_t o = ((size_t)y * g.Nx + x) * nrhs + t;
    X[o] = X[o] - acc;
}

}  // namespace ust
