// common.cuh -- complex arithmetic, error plumbing and indexing shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>

namespace ust {

// ---------------------------------------------------------------------------------------------
// complex number with the alignment of a 2-vector of R (float2 / double2 compatible layout)
// ---------------------------------------------------------------------------------------------
template <typename R>
struct alignas(2 * sizeof(R)) cx {
    R re, im;
    cx() = default;
    __host__ __device__ cx(R r, R i) : re(r), im(i) {}
};

template <typename R> __host__ __device__ __forceinline__ cx<R> operator+(cx<R> a, cx<R> b) { return cx<R>(a.re + b.re, a.im + b.im); }
template <typename R> __host__ __device__ __forceinline__ cx<R> operator-(cx<R> a, cx<R> b) { return cx<R>(a.re - b.re, a.im - b.im); }
template <typename R> __host__ __device__ __forceinline__ cx<R> operator-(cx<R> a) { return cx<R>(-a.re, -a.im); }
template <typename R> __host__ __device__ __forceinline__ cx<R> operator*(cx<R> a, cx<R> b) {
    return cx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename R> __host__ __device__ __forceinline__ cx<R> operator*(R s, cx<R> a) { return cx<R>(s * a.re, s * a.im); }
template <typename R> __host__ __device__ __forceinline__ cx<R> conj(cx<R> a) { return cx<R>(a.re, -a.im); }
template <typename R> __host__ __device__ __forceinline__ cx<R> cxzero() { return cx<R>(R(0), R(0)); }
template <typename R> __host__ __device__ __forceinline__ cx<R> cxone() { return cx<R>(R(1), R(0)); }
// acc += a*b  (4 FMAs)
template <typename R> __device__ __forceinline__ void cmac(cx<R>& acc, cx<R> a, cx<R> b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(a.im, b.re, acc.im);
}
template <typename R> __host__ __device__ __forceinline__ cx<R> crecip(cx<R> a) {
    R d = a.re * a.re + a.im * a.im;
    R s = R(1) / d;
    return cx<R>(a.re * s, -a.im * s);
}
template <typename R> __host__ __device__ __forceinline__ cx<R> cdiv(cx<R> a, cx<R> b) { return a * crecip(b); }

// ---------------------------------------------------------------------------------------------
// Geometry of one factorisation problem, shared by every kernel.
//   nI = Nx-2 interior columns = block size;  M = Ny-2 interior rows = number of block rows;
//   nP = nI rounded up to the Gauss-Jordan block (64) = leading dimension of every dense block.
//   Two elimination chains per frequency: chain z = 2*freq + dir, dir 0 walks rows 0..mid-1 downwards,
//   dir 1 walks rows M-1..mid+1 upwards; block row `mid` joins them (twisted factorisation).
// ---------------------------------------------------------------------------------------------
struct Geom {
    int Nx, Ny, nI, M, nP, mid;
    long long N;  // Nx*Ny
};

constexpr int GJ_NB = 64;

enum Phase { PH_CHAIN = 0, PH_MID = 1 };

// block row handled by batch entry z at chain step `step`; -1 when the chain has already ended.
__host__ __device__ __forceinline__ int chain_row(const Geom& g, int phase, int z, int step) {
    if (phase == PH_MID) return g.mid;
    int dir = z & 1;
    int row = dir ? (g.M - 1 - step) : step;
    bool ok = dir ? (row > g.mid) : (row < g.mid);
    return ok ? row : -1;
}
__host__ __device__ __forceinline__ int chain_freq(int phase, int z) { return phase == PH_MID ? z : (z >> 1); }
__host__ __device__ __forceinline__ int chain_dir(int phase, int z) { return phase == PH_MID ? 2 : (z & 1); }

// plane order (matches the reference's column order, solve_helmholtz.py:198-200)
enum Plane { PL_C = 0, PL_L, PL_R, PL_D, PL_U, PL_DL, PL_DR, PL_UL, PL_UR };

// The three coefficients (multiplying v[a-1], v[a], v[a+1]) of a tridiagonal coupling block applied to a
// vector, for interior column a (grid x = a+1) -- SURVEY.md Appendix A.5.
//   TRI_L  : L_i   (row form; planes of grid row y=i+1): dl, d, dr at (y, x)
//   TRI_U  : U_i   (row form): ul, u, ur at (y, x)
//   TRI_UC : column form of U_j  (U[a-1,a], U[a,a], U[a+1,a]) = ur(y,x-1), u(y,x), ul(y,x+1)
//   TRI_LC : column form of L_j  = dr(y,x-1), d(y,x), dl(y,x+1)
// With conj=true the column forms are the row forms of U_j^H / L_j^H (adjoint sweeps).
enum TriKind { TRI_L = 0, TRI_U = 1, TRI_UC = 2, TRI_LC = 3 };

template <typename R>
__device__ __forceinline__ void tri3(const cx<R>* __restrict__ planes, const Geom& g, int kind, bool cj, int y, int a,
                                     cx<R>& c0, cx<R>& c1, cx<R>& c2) {
    const size_t pl = (size_t)g.Nx * g.Ny;
    const size_t o = (size_t)y * g.Nx + (a + 1);
    switch (kind) {
        case TRI_L: c0 = planes[PL_DL * pl + o]; c1 = planes[PL_D * pl + o]; c2 = planes[PL_DR * pl + o]; break;
        case TRI_U: c0 = planes[PL_UL * pl + o]; c1 = planes[PL_U * pl + o]; c2 = planes[PL_UR * pl + o]; break;
        case TRI_UC: c0 = planes[PL_UR * pl + o - 1]; c1 = planes[PL_U * pl + o]; c2 = planes[PL_UL * pl + o + 1]; break;
        default: c0 = planes[PL_DR * pl + o - 1]; c1 = planes[PL_D * pl + o]; c2 = planes[PL_DL * pl + o + 1]; break;
    }
    if (cj) { c0 = conj(c0); c1 = conj(c1); c2 = conj(c2); }
}

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& s);
extern thread_local long long g_launches;

#define UST_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            ust::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                           std::to_string(__LINE__));                                                   \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

#define UST_LAUNCH_CHECK()                                                                              \
    do {                                                                                                \
        ++ust::g_launches;                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess) {                                                                       \
            ust::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                           std::to_string(__LINE__));                                                   \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

#define UST_TRY(call)            \
    do {                         \
        int r__ = (call);        \
        if (r__) return r__;     \
    } while (0)

inline int cdiv_i(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): the kernels of the factor / sweep chains are short (5-100 us) and strictly
// dependent, so the launch latency and prologue of kernel i+1 are overlapped with the tail wave of kernel i.  Every kernel
// launched through launch_pdl() calls pdl_trigger() first (its dependents may then be scheduled as soon as all of its own
// CTAs are resident or done) and pdl_wait() before its first global-memory access (blocks until the preceding grid has
// completed and its writes are visible).  Both are no-ops for a kernel launched without the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
extern bool g_use_pdl;
// Set by the launch schedules per batch: with <= 4 chains per launch (one or two frequencies on the GPU) every launch is a
// single partial wave, and the early-resident CTAs of the dependent launch cost more than the overlap buys (measured:
// 2 frequencies at 512^2 148 vs 159 ms, cfg4 619 vs 641 ms without the attribute; 16 frequencies 308 vs 295 ms with it).
extern thread_local bool g_pdl_batch_ok;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (g_use_pdl && g_pdl_batch_ok) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// same, as clusters of `cx` CTAs along x (grid.x must be a multiple of cx)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cx_, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cx_; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (g_use_pdl && g_pdl_batch_ok) ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace ust
