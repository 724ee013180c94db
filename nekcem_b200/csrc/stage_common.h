// Device helpers shared by the fused RK-stage kernels (internal; not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "stage_args.h"

namespace nkb {

// read-only data (everything except the RK registers and the auxiliary ODE fields) goes through
// ld.global.nc so that the compiler may hoist the loads above earlier stores
__device__ __forceinline__ double ldg(const double *p) { return __ldg(p); }
__device__ __forceinline__ int ldg(const int *p) { return __ldg(p); }

// pull a 128-byte line into L2 without occupying a register or shared memory
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// one warp prefetches `bytes` starting at base (any alignment)
__device__ __forceinline__ void prefetch_chunk(const void *base, int bytes, int lane)
{
    const char *b = (const char *)base;
    for (int off = lane * 128; off < bytes; off += 32 * 128) prefetch_l2(b + off);
    if (lane == 0) prefetch_l2(b + bytes - 8);
}

// neighbour trace of a 3D face point from its vmapP code: base pointer of component 0 and the
// stride between components.  vp >= 0 volume node; vp <= -3: halo slot or x-face mirror entry;
// otherwise (PEC / unpaired: value unused) the node `dflt`.
__device__ __forceinline__ const double *nbr_trace(const StageArgs &a, int vp, long long dflt,
                                                   long long &stride)
{
    if (vp >= 0) { stride = a.ld; return a.u_in + vp; }
    if (vp <= -3) {
        const int s = -(vp + 3);
        if (s >= XTR_BIAS) { stride = a.ldx; return a.xtr_in + (s - XTR_BIAS); }
        stride = 1;
        return a.halo + 6ll * (long long)s;
    }
    stride = a.ld;
    return a.u_in + dflt;
}

// ---- shared-memory layout of one field component of an element (or of a k-slab of it) --------
__host__ __device__ constexpr int pad_j(int n) { return (n == 6 || n == 14) ? 3 : (n == 12 ? 1 : 0); }
__host__ __device__ constexpr int pad_k(int n)
{
    return (n == 3 || n == 4 || n == 7) ? 3 : (n == 10 ? 7 : 0);
}
// n = 8 and n = 16 use an XOR swizzle (no padding) that makes the r-, s-, t-pencil and the
// linear access patterns all free of 64-bit bank conflicts; other orders use the padding found
// by scripts/smem_banks.py.
template <int N>
struct Lay {
    static constexpr bool SWZ = (N == 8 || N == 16);
    static constexpr int SJ = SWZ ? N : N + pad_j(N);
    static constexpr int SK = SWZ ? N * N : SJ * N + pad_k(N);
    static constexpr int SC = SK * N;
    __device__ __forceinline__ static int at(int i, int j, int k)
    {
        if constexpr (N == 8) return (i ^ ((j >> 1) + 4 * (k & 1))) + 8 * (j ^ (k & 1)) + 64 * k;
        else if constexpr (N == 16) return (i ^ j) + 16 * j + 256 * k;
        else return i + SJ * j + SK * k;
    }
};

// kernel parameter block: the launch arguments plus dxm1 by value, so that D(i,m) with
// compile-time indices is a constant-bank operand of the FMA (no load instruction)
template <int N>
struct StageParams {
    StageArgs a;
    double D[N * N]; // dxm1, column-major: D(i,m) at i + N*m
};

// one direction's share of the curl: (d3*my - d2*mz, d1*mz - d3*mx, d2*mx - d1*my)
__device__ __forceinline__ void curl_part(const double (&d)[3], double mx, double my, double mz,
                                          double (&c)[3])
{
    c[0] = d[2] * my - d[1] * mz;
    c[1] = d[0] * mz - d[2] * mx;
    c[2] = d[1] * mx - d[0] * my;
}

} // namespace nkb
