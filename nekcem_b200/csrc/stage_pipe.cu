// Fused Maxwell RK-stage kernel for sm_100a, "pipelined slab" formulation (3D, the general
// per-node-metric path and the constant-metric path).
//
// Same decomposition and arithmetic as stage_slab.cu (one work item = one k-slab of one element,
// all six components, phases P1..P5 of that file), but the data movement is Blackwell's:
//
//   * persistent CTAs (grid = SMs x resident CTAs), work item q = blockIdx.x + it*gridDim.x;
//   * every streaming operand of an item arrives in shared memory through 1-D bulk asynchronous
//     copies (cp.async.bulk.shared.global, completion on an mbarrier with expect_tx), issued by
//     warp 0 half an item to one item AHEAD of the phase that consumes it, so no phase waits for
//     HBM or L2 and none of these loads occupies a register or the LSU pipe:
//
//        region  holds                          filled after    consumed in
//        Y       fields H,E of the slab (6)     flux of it-1    P0 of it  (re-laid out into U)
//        Y       face geometry/impedances (8)   P0 of it        P3 flux of it
//        X       rx..sz (6)                     P5 of it-1      P2 s-pencils of it
//        X       RK registers kH,kE (6)         P2 of it        P5 epilogue of it
//        Z       tx..tz (3)                     P4 of it-1      P4 t-pencils of it
//
//     Each region is single-buffered: it is refilled right after the phase that read it, which
//     still leaves two to three phases of lead.  U (fields, bank-conflict-free layout Lay<N>) and R
//     (residuals) are as in stage_slab.cu.
//   * what is not contiguous stays on the LSU: the neighbour traces (gather through vmapP, L2
//     prefetched one phase ahead), the two mass arrays (loaded one phase ahead into registers),
//     the t-lines through the other slabs of the element when KS > 1, and the results (streaming
//     stores).
//
// A bulk copy needs 16-byte aligned addresses and sizes; slabs of odd order start on odd
// multiples of 8 bytes, so every copy fetches the enclosing aligned range and the readers add the
// parity of the source address (0 or 1 doubles).
//
// Reference semantics: SURVEY.md 8a rows a4-a18; citations at the phases in stage_slab.cu.
#include <cstdint>

#include "stage_aux.h"

namespace nkb {
namespace {

#ifndef PIPE_SPLIT_OVERRIDE
#define PIPE_SPLIT_OVERRIDE 0
#endif
#ifndef PIPE_FLUX_LAST
#define PIPE_FLUX_LAST 1 // 1: t-pencils before the flux phase (gathers in flight meanwhile)
#endif
#ifndef PIPE_IG_T
#define PIPE_IG_T 0 // 1: the t-pencil lanes interleave the two groups as well
#endif
#ifndef PIPE_NBR_PREFETCH
#define PIPE_NBR_PREFETCH 1 // 1: L2 prefetch of the neighbour traces at P0
#endif
#ifndef PIPE_L2_AHEAD
#define PIPE_L2_AHEAD 0 // 1: L2 prefetch of the next item's fields and rx..sz one item ahead
#endif
#ifndef PIPE_PML_L2
#define PIPE_PML_L2 1 // 1: L2 prefetch of a PML element's auxiliary arrays at the start of the item
#endif
#ifndef PIPE_XOUT
// 1: the x-face mirror is written coalesced from shared memory at the top of the next item;
// 0: straight from the epilogue.  With one CTA per SM (nx1 = 9, 10) the extra pass is serial time
// (measured 0.68 -> 0.695), with two it is free.
#define PIPE_XOUT (PIPE_ONLY_N <= 8)
#endif
#ifndef PIPE_SKIP
#define PIPE_SKIP 0 // timing experiments only (wrong results): 1 P1, 2 P2, 4 flux, 8 P4, 16 relayout
#endif
#ifndef PIPE_REGS
#define PIPE_REGS 128
#endif
#ifndef PIPE_W3S_MAX
#define PIPE_W3S_MAX 1000 // w3mn of one element in shared memory up to this many nodes
#endif
#ifndef PIPE_IG
#define PIPE_IG 1 // 1: r/s-pencil lanes interleave the H and E groups (cofactor reads broadcast)
#endif

__host__ __device__ constexpr int p_round32(int x) { return ((x + 31) / 32) * 32; }
__host__ __device__ constexpr int p_even(int x) { return (x + 1) / 2 * 2; }
__host__ __device__ constexpr int p_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int p_min(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int p_nt_for(int items)
{
    if (items <= 320) return p_round32(items);
    int passes = (items + 255) / 256;
    return p_round32((items + passes - 1) / passes);
}
// slabs per element.  Whole elements (KS = 1) wherever the regions fit one SM: two CTAs per SM up
// to nx1 = 8, one CTA of 512 threads for nx1 = 9, 10 (measured: 0.69 / 0.67 of the HBM roofline
// against 0.52 / 0.42 with two half-element slabs per element, whose t-lines through the other
// slab come from L2).  Above that only thin slabs fit next to the staging regions.
__host__ __device__ constexpr int pipe_ks_for(int n)
{
#ifdef PIPE_KS_OVERRIDE_N
    if (n == PIPE_KS_OVERRIDE_N) return PIPE_KS_OVERRIDE;
#endif
    return n <= 10 ? 1 : (n <= 11 ? 3 : (n <= 12 ? 4 : (n <= 13 ? 5 : 8)));
}

// Layout of one component of U / R.  As Lay<N> (stage_common.h) except where the group-interleaved
// s-pencil lanes of this kernel prefer another padding (scripts/smem_banks_pipe.py: weighted
// wavefronts per access over all phases): nx1 = 10 without the k padding.
template <int N>
struct PLay : Lay<N> {};
template <>
struct PLay<10> {
    static constexpr int SJ = 10, SK = 100, SC = 1000;
    __device__ __forceinline__ static int at(int i, int j, int k) { return i + 10 * j + 100 * k; }
};

// skew of the E components in U and R for the group-interleaved s-pencil lanes
// (scripts/smem_banks_pipe.py, for the slab thickness pipe_ks_for gives)
__host__ __device__ constexpr int pipe_he_for(int n)
{
    return n == 6 ? 14 : n == 7 ? 3 : n == 8 ? 8 : n == 9 ? 14 : n == 10 ? 2 : n == 11 ? 15
         : n == 13 ? 12 : n == 14 ? 10 : n == 15 ? 9 : 0;
}

template <int N, int KS>
struct PT {
    static constexpr int N2 = N * N, N3 = N2 * N, NF = 6 * N2;
    static constexpr int KB = (N + KS - 1) / KS; // thickest slab
    static constexpr int SPLIT = PIPE_SPLIT_OVERRIDE ? PIPE_SPLIT_OVERRIDE : (N <= 5 ? 4 : 2);
    static constexpr int RS_BLK = p_round32(2 * N * KB);
    static constexpr int RS_ITEMS = SPLIT * RS_BLK;
    static constexpr int TSPLIT = KS == 1 ? SPLIT : 1;
    static constexpr int T_BLK = p_round32(2 * N2);
    static constexpr int T_ITEMS = TSPLIT * T_BLK;
#ifdef PIPE_NT
    static constexpr int NT = PIPE_NT;
#else
    static constexpr int NT = (KS == 1 && N == 10) ? 640 : (KS == 1 && N == 9) ? 512
                              : p_max(p_nt_for(RS_ITEMS), p_round32(N2));
#endif
    static constexpr int SC = PLay<N>::SK * KB; // component stride in U and R
    static constexpr int XL = p_even(N2 * KB + 2); // linear slab of one array + alignment slack
    // face staging: KS == 1: the 6*N2 points of the element are contiguous per array;
    // KS > 1: the slab's strips of the four x/y faces and one z face
    static constexpr int SEGXY = p_even(N * KB + 2), SEGZ = p_even(N2 + 2);
    static constexpr int FL = KS == 1 ? p_even(6 * N2 + 2) : 4 * SEGXY + SEGZ;
    static constexpr int YL = p_max(6 * XL, 8 * FL);
    static constexpr int FXY = 4 * N * KB;
    static constexpr int FZ = KS == 1 ? 2 * N2 : N2;
    static constexpr int F_ITEMS = FXY + FZ;
    static constexpr int FPT = (F_ITEMS + NT - 1) / NT;
    static constexpr int NBAR = 5;
    // w3mn of one element in shared memory wherever it fits next to the regions (whole elements:
    // nx1 <= 10), else read through L1.  nx1 = 9: 0.691 -> 0.704 of the roofline against L1 reads
    // (the streaming gathers keep evicting the table)
    static constexpr bool W3S = N3 <= PIPE_W3S_MAX;
    // the E components of U and R start HE doubles later than a multiple of the component stride
    // (PIPE_IG: the s-pencil lanes interleave the H and E groups, so that both read one cofactor
    // address; the skew puts their field and residual accesses on different banks)
    static constexpr int HE = PIPE_IG ? pipe_he_for(N) : 0;
    static constexpr int OFF_U = 0, OFF_R = 6 * SC + HE, OFF_X = 12 * SC + 2 * HE,
                         OFF_Y = OFF_X + 6 * XL, OFF_Z = OFF_Y + YL, OFF_W = OFF_Z + 3 * XL,
                         OFF_BAR = OFF_W + (W3S ? p_even(N3) : 0);
    static constexpr size_t SMEM = sizeof(double) * (OFF_BAR + NBAR + 1);
    // work item w of an r/s pencil phase -> output share h, group g (0: curl H, 1: -curl E),
    // pencil (pa, pb).  h is warp-uniform (it selects code).
    template <bool IG>
    __device__ __forceinline__ static void rs_item(int w, int &h, int &g, int &pa, int &pb)
    {
        h = w / RS_BLK;
        const int r = w - h * RS_BLK;
        if (IG) { // lanes: pa (N) x g (2) x pb
            const int q = r / N;
            pa = r - q * N;
            g = 2 * N * KB <= r ? 2 : (q & 1);
            pb = q >> 1;
        } else {
            g = r / (N * KB);
            const int p = r - g * (N * KB);
            pa = p % N;
            pb = p / N;
        }
    }
    __host__ __device__ static constexpr int k0(int s) { return s * N / KS; }
    __host__ __device__ static constexpr int kb(int s) { return (s + 1) * N / KS - s * N / KS; }
    static constexpr int MINB_SMEM = (227 * 1024) / ((int)SMEM + 1024);
    static constexpr int MINB_THR = 2048 / NT;
    static constexpr int REGS = NT > 512 ? 96 : PIPE_REGS;
    static constexpr int MINB_REG = 65536 / (NT * REGS);
    static constexpr int MINB0 = p_min(p_min(MINB_SMEM, MINB_THR), MINB_REG);
    static constexpr int MINB = MINB0 < 1 ? 1 : (MINB0 > 4 ? 4 : MINB0);
};

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// global -> shared bulk copy of the 16-byte aligned range enclosing [src, src+cnt doubles);
// the first double lands at dst + (src & 15)/8.  Returns the bytes in flight.
__device__ __forceinline__ uint32_t bulk_load(double *dst, const double *src, int cnt, uint64_t *bar)
{
    const uintptr_t s = (uintptr_t)src;
    const uint32_t head = (uint32_t)(s & 15);
    const uint32_t bytes = (head + (uint32_t)cnt * 8u + 15u) & ~15u;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(s - head), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
    return bytes;
}
// L2 prefetch of the same aligned range (no shared-memory destination, no completion)
__device__ __forceinline__ void bulk_prefetch(const double *src, int cnt)
{
    const uintptr_t s = (uintptr_t)src;
    const uint32_t head = (uint32_t)(s & 15);
    const uint32_t bytes = (head + (uint32_t)cnt * 8u + 15u) & ~15u;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s - head), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t bulk_bytes(const double *src, int cnt)
{
    const uint32_t head = (uint32_t)((uintptr_t)src & 15);
    return (head + (uint32_t)cnt * 8u + 15u) & ~15u;
}
__device__ __forceinline__ int par_of(const double *p) { return (int)(((uintptr_t)p >> 3) & 1); }

// ---- pencil phases ---------------------------------------------------------------------------------
template <int N, int DIR, int KOFF>
__device__ __forceinline__ int ps_at(int m, int pa, int pb)
{
    return DIR == 0 ? PLay<N>::at(m, pa, pb)
                    : (DIR == 1 ? PLay<N>::at(pa, m, pb) : PLay<N>::at(pa, pb, m - KOFF));
}
// slab-local linear node of position m along the pencil (DIR 2: m counts from the slab's k0)
template <int N, int DIR>
__device__ __forceinline__ int pn_at(int m, int pa, int pb)
{
    return DIR == 0 ? m + N * pa + N * N * pb : (DIR == 1 ? pa + N * m + N * N * pb : pa + N * pb + N * N * m);
}

// One thread, outputs O0..O1-1 of pencil (pa,pb):
//   d_c = sum_m D(o,m) u_c(m)                               (mxfK order, left to right)
//   DIR 0: R = d
//   DIR 1: R = (curl_part(R; rx,ry,rz) + curl_part(d; sx,sy,sz)) * (sg*w3)
//   DIR 2: R = R + (sg*w3) * curl_part(d; tx,ty,tz)
// cofs: shared-memory copy of the slab's cofactor arrays (array q at cofs + q*XL, slab-local linear
// node index); CM: the element's constant cofactors come from global memory (cm[]).
template <int N, int DIR, int KOFF, int O0, int O1, int SC, int XL, bool CM>
__device__ __forceinline__ void pipe_pencil(const double (&D)[N * N], const double *U, double *R,
                                            const double *cofs, const double (&cm)[6],
                                            const double *w3, int pa, int pb, double sg)
{
    constexpr int NO = O1 - O0;
    if constexpr (NO > 0) {
        double acc[3][NO];
#pragma unroll
        for (int m = 0; m < N; m++) {
            const int so = ps_at<N, DIR, KOFF>(m, pa, pb);
            const double u0 = U[so], u1 = U[SC + so], u2 = U[2 * SC + so];
#pragma unroll
            for (int o = 0; o < NO; o++) {
                const double dv = D[(O0 + o) + N * m];
                if (m == 0) {
                    acc[0][o] = dv * u0; acc[1][o] = dv * u1; acc[2][o] = dv * u2;
                } else {
                    acc[0][o] = acc[0][o] + dv * u0;
                    acc[1][o] = acc[1][o] + dv * u1;
                    acc[2][o] = acc[2][o] + dv * u2;
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; o++) {
            const double d[3] = {acc[0][o], acc[1][o], acc[2][o]};
            double *Ro = R + ps_at<N, DIR, KOFF>(O0 + o, pa, pb);
            if constexpr (DIR == 0) {
                Ro[0] = d[0]; Ro[SC] = d[1]; Ro[2 * SC] = d[2];
            } else {
                const int nd = pn_at<N, DIR>(DIR == 2 ? O0 + o - KOFF : O0 + o, pa, pb);
                const double wv = sg * w3[nd];
                double c[3];
                if constexpr (DIR == 1) {
                    const double dr[3] = {Ro[0], Ro[SC], Ro[2 * SC]};
                    double cr[3];
                    if constexpr (CM) {
                        curl_part(dr, cm[0], cm[1], cm[2], cr);
                        curl_part(d, cm[3], cm[4], cm[5], c);
                    } else {
                        curl_part(dr, cofs[nd], cofs[XL + nd], cofs[2 * XL + nd], cr);
                        curl_part(d, cofs[3 * XL + nd], cofs[4 * XL + nd], cofs[5 * XL + nd], c);
                    }
                    Ro[0] = (cr[0] + c[0]) * wv;
                    Ro[SC] = (cr[1] + c[1]) * wv;
                    Ro[2 * SC] = (cr[2] + c[2]) * wv;
                } else {
                    if constexpr (CM) curl_part(d, cm[0], cm[1], cm[2], c);
                    else curl_part(d, cofs[nd], cofs[XL + nd], cofs[2 * XL + nd], c);
                    Ro[0] = Ro[0] + wv * c[0];
                    Ro[SC] = Ro[SC] + wv * c[1];
                    Ro[2 * SC] = Ro[2 * SC] + wv * c[2];
                }
            }
        }
    }
}

// dispatch on the thread's share h of the outputs LO..HI-1 (compile-time ranges)
template <int N, int DIR, int KOFF, int LO, int HI, int SPLIT, int SC, int XL, bool CM>
__device__ __forceinline__ void pipe_split(const double (&D)[N * N], const double *U, double *R,
                                           const double *cofs, const double (&cm)[6],
                                           const double *w3, int pa, int pb, double sg, int h)
{
    constexpr int L = HI - LO, HN = (L + SPLIT - 1) / SPLIT;
    constexpr int E1 = LO + (HN < L ? HN : L), E2 = LO + (2 * HN < L ? 2 * HN : L),
                  E3 = LO + (3 * HN < L ? 3 * HN : L);
    if (h == 0) pipe_pencil<N, DIR, KOFF, LO, E1, SC, XL, CM>(D, U, R, cofs, cm, w3, pa, pb, sg);
    if (SPLIT > 1 && h == 1) pipe_pencil<N, DIR, KOFF, E1, E2, SC, XL, CM>(D, U, R, cofs, cm, w3, pa, pb, sg);
    if (SPLIT > 2 && h == 2) pipe_pencil<N, DIR, KOFF, E2, E3, SC, XL, CM>(D, U, R, cofs, cm, w3, pa, pb, sg);
    if (SPLIT > 3 && h == 3) pipe_pencil<N, DIR, KOFF, E3, HI, SC, XL, CM>(D, U, R, cofs, cm, w3, pa, pb, sg);
}

// t-pencils of slab S (compile-time k range).  cofs = tx..tz of the slab in shared memory.
template <int N, int KS, int S, bool CM>
__device__ __forceinline__ void pipe_t_phase(const double (&D)[N * N], const StageArgs &a,
                                             const double *U, double *R, const double *cofs,
                                             const double (&cm)[6], const double *W3,
                                             long long ebase, int tid)
{
    using C = PT<N, KS>;
    constexpr int K0 = C::k0(S), K1 = K0 + C::kb(S);
    constexpr int HE = C::HE;
    if constexpr (KS == 1) {
#pragma unroll 1
        for (int w = tid; w < C::T_ITEMS; w += C::NT) {
            const int h = w / C::T_BLK, r = w - h * C::T_BLK;
            int g, pa, pb;
            if (PIPE_IG_T) { // lanes: i (N) x group (2) x j: both groups read one cofactor address
                const int q = r / N;
                pa = r - q * N;
                g = r >= 2 * C::N2 ? 2 : (q & 1);
                pb = q >> 1;
            } else {
                g = r / C::N2;
                const int p = r - g * C::N2;
                pa = p % N; pb = p / N;
            }
            if (g > 1) continue;
            pipe_split<N, 2, K0, K0, K1, C::TSPLIT, C::SC, C::XL, CM>(
                D, U + (g ? 3 * C::SC + HE : 0), R + (g ? 0 : 3 * C::SC + HE), cofs, cm,
                W3 + K0 * C::N2, pa, pb, g ? -1.0 : 1.0, h);
        }
    } else {
        // the lines run through the other slabs of the element: they come from global memory (L2).
        // Software pipeline over the flat sequence of (item, component) steps: the N loads of
        // step+1 are issued before the FMAs of step.
        constexpr int NO = K1 - K0;
        constexpr int NPASS = (2 * C::N2 + C::NT - 1) / C::NT;
        constexpr int NSTEP = 3 * NPASS;
        double ul[2][N];
        auto item = [&](int ps, int &g, int &nd, bool &ok) {
            const int w = tid + ps * C::NT;
            g = w / C::N2;
            nd = w - g * C::N2; // = pa + N*pb
            ok = w < 2 * C::N2;
            if (!ok) { g = 0; nd = 0; }
        };
        auto issue = [&](int step, int buf) {
            int g, nd; bool ok;
            item(step / 3, g, nd, ok);
            const double *gp = a.u_in + ((g ? 3 : 0) + step % 3) * a.ld + ebase + nd;
#pragma unroll
            for (int m = 0; m < N; m++) ul[buf][m] = ok ? ldg(gp + C::N2 * m) : 0.0;
        };
        issue(0, 0);
        double acc[3][NO];
#pragma unroll
        for (int step = 0; step < NSTEP; step++) {
            const int c = step % 3, buf = step & 1;
            int g, nd; bool ok;
            item(step / 3, g, nd, ok);
            if (step + 1 < NSTEP) issue(step + 1, buf ^ 1);
#pragma unroll
            for (int m = 0; m < N; m++)
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    const double dv = D[(K0 + o) + N * m];
                    if (m == 0) acc[c][o] = dv * ul[buf][m];
                    else acc[c][o] = acc[c][o] + dv * ul[buf][m];
                }
            if (c == 2 && ok) {
                double *Rd = R + (g ? 0 : 3 * C::SC + HE);
                const int pa = nd % N, pb = nd / N;
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    const double d[3] = {acc[0][o], acc[1][o], acc[2][o]};
                    const int nl = nd + C::N2 * o;
                    const double wv = (g ? -1.0 : 1.0) * W3[nd + C::N2 * (K0 + o)];
                    double cc[3];
                    if constexpr (CM) curl_part(d, cm[0], cm[1], cm[2], cc);
                    else curl_part(d, cofs[nl], cofs[C::XL + nl], cofs[2 * C::XL + nl], cc);
                    double *Ro = Rd + PLay<N>::at(pa, pb, o);
                    Ro[0] = Ro[0] + wv * cc[0];
                    Ro[C::SC] = Ro[C::SC] + wv * cc[1];
                    Ro[2 * C::SC] = Ro[2 * C::SC] + wv * cc[2];
                }
            }
        }
    }
}

template <int N, int KS, bool PML, bool CM>
__global__ void __launch_bounds__(PT<N, KS>::NT, PT<N, KS>::MINB)
    pipe_kernel(const __grid_constant__ StageParams<N> prm)
{
    using C = PT<N, KS>;
    constexpr int N2 = C::N2, N3 = C::N3, NF = C::NF;
    constexpr int NT = C::NT, SC = C::SC, KB = C::KB, FPT = C::FPT, XL = C::XL, FL = C::FL;
    constexpr int HE = C::HE;
    const StageArgs &a = prm.a;
    extern __shared__ __align__(16) double smem[];
    double *U = smem + C::OFF_U; // [6][SC] H,E of the slab at stage start (layout Lay<N>)
    double *R = smem + C::OFF_R; // [6][SC] residuals resH,resE
    double *X = smem + C::OFF_X; // [6][XL] rx..sz, then kH,kE
    double *Y = smem + C::OFF_Y; // fields landing zone [6][XL], then face arrays [8][FL]
    double *Z = smem + C::OFF_Z; // [3][XL] tx..tz
    const double *W3 = C::W3S ? smem + C::OFF_W : a.w3; // w3mn of one element
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint64_t *b_fld = bar, *b_face = bar + 1, *b_cof = bar + 2, *b_k = bar + 3, *b_cot = bar + 4;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool w0 = tid < 32;
    const int nitems = a.nel * KS;
    const int G = (int)gridDim.x;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < C::NBAR; b++) mbar_init(bar + b, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (C::W3S)
        for (int i = tid; i < N3; i += NT) smem[C::OFF_W + i] = ldg(a.w3 + i);
    __syncthreads();

    // ---- producers: every lane of warp 0 arrives on the barrier with the bytes of its own copies
    auto slab_base = [&](int e, int s) { return (long long)e * N3 + C::k0(s) * N2; };
    auto issue_fields = [&](int e, int s) {
        const int cnt = N2 * C::kb(s);
        uint32_t bytes = 0;
        const double *src = a.u_in + (long long)lane * a.ld + slab_base(e, s);
        if (lane < 6) bytes = bulk_bytes(src, cnt);
        mbar_arrive_tx(b_fld, bytes);
        if (lane < 6) bulk_load(Y + lane * XL, src, cnt, b_fld);
    };
    auto issue_cof = [&](int e, int s) { // rx..sz -> X
        const int cnt = N2 * C::kb(s);
        uint32_t bytes = 0;
        const double *src = a.met[lane < 6 ? lane : 0] + slab_base(e, s);
        if (lane < 6) bytes = bulk_bytes(src, cnt);
        mbar_arrive_tx(b_cof, bytes);
        if (lane < 6) bulk_load(X + lane * XL, src, cnt, b_cof);
    };
    auto issue_cot = [&](int e, int s) { // tx..tz -> Z
        const int cnt = N2 * C::kb(s);
        uint32_t bytes = 0;
        const double *src = a.met[lane < 3 ? 6 + lane : 6] + slab_base(e, s);
        if (lane < 3) bytes = bulk_bytes(src, cnt);
        mbar_arrive_tx(b_cot, bytes);
        if (lane < 3) bulk_load(Z + lane * XL, src, cnt, b_cot);
    };
    auto issue_k = [&](int e, int s) { // kH,kE -> X
        const int cnt = N2 * C::kb(s);
        uint32_t bytes = 0;
        const double *src = a.kf + (long long)lane * a.ld + slab_base(e, s);
        if (lane < 6) bytes = bulk_bytes(src, cnt);
        mbar_arrive_tx(b_k, bytes);
        if (lane < 6) bulk_load(X + lane * XL, src, cnt, b_k);
    };
    auto face_array = [&](int arr) -> const double * {
        return arr == 0 ? a.unx : arr == 1 ? a.uny : arr == 2 ? a.unz : arr == 3 ? a.area
             : arr == 4 ? a.hY : arr == 5 ? a.Y1 : arr == 6 ? a.hZ : a.Z1;
    };
    // slot (reference face order -y,+x,+y,-x,-z,+z) of strip f4 = 0..3: -x,+x,-y,+y
    auto strip_slot = [](int f4) { return f4 == 0 ? 3 : (f4 == 1 ? 1 : (f4 == 2 ? 0 : 2)); };
    auto issue_face = [&](int e, int s) {
        const long long fbase = (long long)e * NF;
        uint32_t bytes = 0;
        if constexpr (KS == 1) {
            const double *src = face_array(lane & 7) + fbase;
            if (lane < 8) bytes = bulk_bytes(src, NF);
            mbar_arrive_tx(b_face, bytes);
            if (lane < 8) bulk_load(Y + lane * FL, src, NF, b_face);
        } else {
            // 8 arrays x (4 strips + 1 z face): copy id = lane, lane + 32
            const int k0 = C::k0(s), kb = C::kb(s);
            const bool zf = s == 0 || s == KS - 1;
            const double *srcs[2]; double *dsts[2]; int cnts[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int id = lane + 32 * r, arr = id / 5, sg = id - arr * 5;
                cnts[r] = 0; srcs[r] = a.unx; dsts[r] = Y;
                if (id < 40) {
                    const double *fa = face_array(arr);
                    if (sg < 4) {
                        srcs[r] = fa + fbase + strip_slot(sg) * N2 + N * k0;
                        dsts[r] = Y + arr * FL + sg * C::SEGXY;
                        cnts[r] = N * kb;
                    } else if (zf) {
                        srcs[r] = fa + fbase + (s == 0 ? 4 : 5) * N2;
                        dsts[r] = Y + arr * FL + 4 * C::SEGXY;
                        cnts[r] = N2;
                    }
                }
                if (cnts[r] > 0) bytes += bulk_bytes(srcs[r], cnts[r]);
            }
            mbar_arrive_tx(b_face, bytes);
#pragma unroll
            for (int r = 0; r < 2; r++)
                if (cnts[r] > 0) bulk_load(dsts[r], srcs[r], cnts[r], b_face);
        }
    };
    // face point f of this thread in slab s: slot*N2 + point (-1: none), its node in U/R, its
    // index in the staged face arrays (without the parity of the source address).
    // Face slots in the reference's order (cemface, cem_common.F:234-260): -y,+x,+y,-x,-z,+z.
    // The second and later passes start in the middle of the block, so that a partly filled
    // pass does not land on the warps that own the strided x-face gathers of the first pass.
    auto face_point = [&](int f, int s, int &fjs, int &fsn, int &fyi) {
        const int k0 = C::k0(s), kb = C::kb(s);
        const int qf = f == 0 ? tid : f * NT + (tid + NT / 2) % NT;
        int slot = -1, fp0 = 0, ci = 0, cj = 0, kl = 0, yi = 0;
        if (qf < C::FXY) {
            const int f4 = qf / (N * KB), r = qf - f4 * (N * KB);
            const int fa = r % N;
            kl = r / N;
            if (kl < kb) {
                fp0 = fa + N * (k0 + kl);
                if (f4 == 0) { slot = 3; ci = 0; cj = fa; }
                else if (f4 == 1) { slot = 1; ci = N - 1; cj = fa; }
                else if (f4 == 2) { slot = 0; ci = fa; cj = 0; }
                else { slot = 2; ci = fa; cj = N - 1; }
                yi = f4 * C::SEGXY + r;
            }
        } else if (qf < C::F_ITEMS) {
            const int q2 = qf - C::FXY, zf = q2 / N2;
            fp0 = q2 - zf * N2;
            ci = fp0 % N; cj = fp0 / N;
            if (KS == 1) { slot = 4 + zf; kl = zf ? N - 1 : 0; }
            else if (s == 0) { slot = 4; kl = 0; }
            else if (s == KS - 1) { slot = 5; kl = kb - 1; }
            yi = 4 * C::SEGXY + fp0;
        }
        fjs = slot < 0 ? -1 : slot * N2 + fp0;
        fsn = PLay<N>::at(ci, cj, kl);
        if (KS == 1) fyi = slot < 0 ? 0 : fjs; // the element's 6*N2 points are staged contiguously
        else {
            // + parity of the strip's first point (the arrays themselves are 16-byte aligned)
            const int first = slot < 0 ? 0 : (slot < 4 ? slot * N2 + N * k0 : slot * N2);
            fyi = yi + (first & 1); // the caller adds the parity of e*NF
        }
    };

    int q = blockIdx.x;
    int e = q < nitems ? a.elist[q / KS] : 0;
    int en = q + G < nitems ? a.elist[(q + G) / KS] : 0; // element of the next item
    int fvp[FPT];
    if (q < nitems) {
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            int fjs, fsn, fyi;
            face_point(f, q % KS, fjs, fsn, fyi);
            fvp[f] = fjs < 0 ? -2 : ldg(a.vmapP + (long long)e * NF + fjs);
        }
        if (w0) {
            issue_fields(e, q % KS);
            if (!CM) { issue_cof(e, q % KS); issue_cot(e, q % KS); }
        }
    }

    // plane mapping of the staging and epilogue loops: thread = (plane slot ps, node (pi,pj))
    constexpr int NPL = NT / N2;
    static_assert(NPL >= 1, "block smaller than one i-j plane");
    const int ps = tid / N2, pnd = tid - ps * N2;
    const int pi = pnd % N, pj = pnd / N;
    const bool pok = ps < NPL;

    // x-face mirror of an item's new fields: R (parked by its epilogue) -> global, 2*N*kb
    // consecutive entries per component and side.  Runs after the item's last barrier and before
    // the next item's P1 touches R.
    auto xout = [&](int pe, int psl) {
        const int pk0 = C::k0(psl), cnt = N * C::kb(psl);
        for (int v = tid; v < 12 * cnt; v += NT) {
            const int cs = v / cnt, p = v - cs * cnt;   // (component, side), point j + N*kl
            const int c = cs >> 1, side = cs & 1;
            const int j = p % N, kl = p / N;
            a.xtr_out[c * a.ldx + (2ll * pe + side) * N2 + N * pk0 + p] =
                R[c * SC + (c >= 3 ? HE : 0) + PLay<N>::at(side ? N - 1 : 0, j, kl)];
        }
    };
    int xo_e = -1, xo_s = 0; // item whose mirror entries still sit in R

    uint32_t par = 0;
#pragma unroll 1
    for (; q < nitems; q += G, par ^= 1) {
        const int s = q % KS;
        const int k0 = C::k0(s), kb = C::kb(s);
        const long long ebase = (long long)e * N3;
        const long long sbase = ebase + k0 * N2; // first node of the slab
        const int qn = q + G, sn_ = qn % KS;
        const bool more = qn < nitems;
        // element of the item after the next (the producers need the next one's early)
        const int enn = qn + G < nitems ? ldg(a.elist + (qn + G) / KS) : 0;

        int fsn[FPT], fjs[FPT], fyi[FPT];
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            face_point(f, s, fjs[f], fsn[f], fyi[f]);
            if (KS > 1) fyi[f] += (int)(((long long)e * NF) & 1);
        }

        if (PIPE_XOUT && xo_e >= 0) xout(xo_e, xo_s);

        // ---- P0: fields of the slab: landing zone -> U (bank-conflict-free layout) -------------
        mbar_wait(b_fld, par);
        {
            const int off = par_of(a.u_in + sbase);
            constexpr int SPER = (6 * KB + NPL - 1) / NPL;
#pragma unroll 4
            for (int x = 0; x < SPER; x++) {
                const int pl = ps + NPL * x, c = pl / KB, kl = pl - c * KB;
                if (pok && pl < 6 * KB && kl < kb && !(PIPE_SKIP & 16))
                    U[c * SC + (c >= 3 ? HE : 0) + PLay<N>::at(pi, pj, kl)] = Y[c * XL + off + kl * N2 + pnd];
            }
        }
        // neighbour traces of this thread's face points -> L2
#pragma unroll
        for (int f = 0; f < FPT; f++)
            if (PIPE_NBR_PREFETCH && (fvp[f] >= 0 || -(fvp[f] + 3) >= XTR_BIAS)) {
                long long st;
                const double *nb = nbr_trace(a, fvp[f], sbase, st);
#pragma unroll
                for (int c = 0; c < 6; c++) prefetch_l2(nb + c * st);
            }
        __syncthreads();
        if (w0) {
            issue_face(e, s);
            if (PML && PIPE_PML_L2 && (a.elflag[e] & 1)) {
                // PML elements: the epilogue reads 18 more arrays per node through the LSU (sigma,
                // eps, mu, bm, B, D and their RK registers); start them towards L2 now
                const int cnt = N2 * kb;
                const double *src = nullptr;
                if (lane < 3) src = a.sig + (long long)lane * a.npts;
                else if (lane < 6) src = a.pB + (long long)(lane - 3) * a.npts;
                else if (lane < 9) src = a.pD + (long long)(lane - 6) * a.npts;
                else if (lane < 12) src = a.kB + (long long)(lane - 9) * a.npts;
                else if (lane < 15) src = a.kD + (long long)(lane - 12) * a.npts;
                else if (lane == 15) src = a.eps;
                else if (lane == 16) src = a.mu;
                else if (lane == 17) src = a.bmn;
                if (src != nullptr) bulk_prefetch(src + sbase, cnt);
            }
            // the two streams whose shared-memory region frees late (fields: after the flux phase,
            // rx..sz: after the epilogue) start their trip from HBM now, into L2
            if (PIPE_L2_AHEAD && more) {
                const int cnt = N2 * C::kb(sn_);
                const long long nb = slab_base(en, sn_);
                if (lane < 6) bulk_prefetch(a.u_in + (long long)lane * a.ld + nb, cnt);
                else if (lane < 12 && !CM) bulk_prefetch(a.met[lane - 6] + nb, cnt);
            }
        }

        // constant-metric elements: the nine cofactors of the element (first node)
        double cmrs[6] = {0, 0, 0, 0, 0, 0}, cmt[6] = {0, 0, 0, 0, 0, 0};
        if constexpr (CM) {
#pragma unroll
            for (int qq = 0; qq < 6; qq++) cmrs[qq] = ldg(a.met[qq] + ebase);
#pragma unroll
            for (int qq = 0; qq < 3; qq++) cmt[qq] = ldg(a.met[6 + qq] + ebase);
        }

        // ---- P1: r-pencils, thread (g,h,j,k): raw derivatives --------------------------------------
#pragma unroll 1
        for (int w = tid; w < C::RS_ITEMS; w += NT) {
            int h, g, pa, pb;
            C::template rs_item<false>(w, h, g, pa, pb);
            if (g < 2 && pb < kb && !(PIPE_SKIP & 1))
                pipe_split<N, 0, 0, 0, N, C::SPLIT, SC, XL, CM>(prm.D, U + (g ? 3 * SC + HE : 0),
                                                              R + (g ? 0 : 3 * SC + HE), nullptr, cmrs,
                                                              nullptr, pa, pb, g ? -1.0 : 1.0, h);
        }
        __syncthreads();

        // ---- P2: s-pencils, thread (g,h,i,k): r- and s-parts of the weighted curl ------------------
        if (!CM) mbar_wait(b_cof, par);
        {
            const double *cofs = X + par_of(a.met[0] + sbase);
#pragma unroll 1
            for (int w = tid; w < C::RS_ITEMS; w += NT) {
                int h, g, pa, pb;
                C::template rs_item<(PIPE_IG != 0)>(w, h, g, pa, pb);
                if (g < 2 && pb < kb && !(PIPE_SKIP & 2))
                    pipe_split<N, 1, 0, 0, N, C::SPLIT, SC, XL, CM>(
                        prm.D, U + (g ? 3 * SC + HE : 0), R + (g ? 0 : 3 * SC + HE), cofs, cmrs,
                        W3 + k0 * N2, pa, pb, g ? -1.0 : 1.0, h);
            }
        }
        __syncthreads();
        if (w0) issue_k(e, s);

        // masses of this thread's epilogue nodes: in flight during the flux and t phases
        constexpr int PER = (2 * KB + NPL - 1) / NPL; // (g, local k) planes per thread
        double mb[PER];
#pragma unroll
        for (int x = 0; x < PER; x++) {
            const int pl = ps + NPL * x, g = pl / KB ? 1 : 0;
            int kl = pl - g * KB;
            kl = kl < kb ? kl : kb - 1;
            const int nl = (pok ? pnd : 0) + kl * N2;
            mb[x] = ldg((g == 0 ? a.ebm1 : a.hbm1) + sbase + nl);
        }
        // neighbour ids of the next item's face points (consumed at its P0)
        int fvn[FPT];
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            int njs = fjs[f], nsn, nyi;
            if (KS > 1) face_point(f, sn_, njs, nsn, nyi);
            fvn[f] = (njs < 0 || !more) ? -2 : ldg(a.vmapP + (long long)en * NF + njs);
        }

        // neighbour traces (gather; L2 hits after the prefetch of P0)
        double pv[FPT][6];
        auto gather = [&]() {
#pragma unroll
            for (int f = 0; f < ((PIPE_SKIP & 4) ? 0 : FPT); f++) {
                const int vp = fvp[f];
                // volume node, x-face mirror, halo slot, or (unused) the first node of the slab
                long long st;
                const double *nb = nbr_trace(a, vp, sbase, st);
#pragma unroll
                for (int c = 0; c < 6; c++) pv[f][c] = (PIPE_SKIP & 32) ? (double)(vp + c) : ldg(nb + c * st);
            }
        };

        // ---- P3: surface flux: both fluxes of a face point from one neighbour gather ---------------
        auto flux_phase = [&]() {
            mbar_wait(b_face, par);
            double fl[FPT][6];
#pragma unroll
            for (int f = 0; f < ((PIPE_SKIP & 4) ? 0 : FPT); f++) {
                const bool valid = fjs[f] >= 0;
                const long long jf = (long long)e * NF + (valid ? fjs[f] : 0);
                const int vp = fvp[f];
                const int sn = fsn[f];
                const int yi = fyi[f];
                const double unx = Y[yi], uny = Y[FL + yi], unz = Y[2 * FL + yi];
                const double ar = Y[3 * FL + yi];
                const double hY = Y[4 * FL + yi], Y1 = Y[5 * FL + yi];
                const double hZ = Y[6 * FL + yi], Z1 = Y[7 * FL + yi];
                double Hx = U[sn], Hy = U[SC + sn], Hz = U[2 * SC + sn];
                double Ex = U[3 * SC + HE + sn], Ey = U[4 * SC + HE + sn], Ez = U[5 * SC + HE + sn];
                double pHx = pv[f][0], pHy = pv[f][1], pHz = pv[f][2];
                double pEx = pv[f][3], pEy = pv[f][4], pEz = pv[f][5];
                if (a.inc_own != nullptr && valid) { // userinc hook (src/cem_maxwell.F:498)
                    const int qo = a.inc_own[jf], qi = a.inc_nbr[jf];
                    if (qo >= 0) {
                        const double ui = cos(a.inc_phase[qo] - a.inc_wt);
                        Hx += a.inc_amp[qo] * ui; Hy += a.inc_amp[a.inc_n + qo] * ui;
                        Hz += a.inc_amp[2 * a.inc_n + qo] * ui;
                        Ex += a.inc_amp[3 * a.inc_n + qo] * ui;
                        Ey += a.inc_amp[4 * a.inc_n + qo] * ui;
                        Ez += a.inc_amp[5 * a.inc_n + qo] * ui;
                    }
                    if (qi >= 0) {
                        const double ui = cos(a.inc_phase[qi] - a.inc_wt);
                        pHx += a.inc_amp[qi] * ui; pHy += a.inc_amp[a.inc_n + qi] * ui;
                        pHz += a.inc_amp[2 * a.inc_n + qi] * ui;
                        pEx += a.inc_amp[3 * a.inc_n + qi] * ui;
                        pEy += a.inc_amp[4 * a.inc_n + qi] * ui;
                        pEz += a.inc_amp[5 * a.inc_n + qi] * ui;
                    }
                }
                // -n x E, -n x H of the own side (flux3d :946-955)
                double s0 = -uny * Ez + unz * Ey;
                double s1 = -unz * Ex + unx * Ez;
                double s2 = -unx * Ey + uny * Ex;
                double s3 = -uny * Hz + unz * Hy;
                double s4 = -unz * Hx + unx * Hz;
                double s5 = -unx * Hy + uny * Hx;
                int gqn = -1;
                if constexpr (PML) { // userfsrc hook (:958): graphene sheet current, own side
                    if (a.fs_own != nullptr && valid) {
                        const int gqo = a.fs_own[jf];
                        gqn = a.fs_nbr[jf];
                        if (gqo >= 0) {
                            s3 = s3 - a.fs_val[gqo];
                            s4 = s4 - a.fs_val[a.fs_n + gqo];
                            s5 = s5 - a.fs_val[2 * a.fs_n + gqo];
                        }
                    }
                }
                if (vp >= 0 || vp <= -3) {
                    // neighbour's (-n+ x E+) with n+ = -n-  (the gs_op_fields sum of :962)
                    s0 = s0 - (-uny * pEz + unz * pEy);
                    s1 = s1 - (-unz * pEx + unx * pEz);
                    s2 = s2 - (-unx * pEy + uny * pEx);
                    s3 = s3 - (-uny * pHz + unz * pHy);
                    s4 = s4 - (-unz * pHx + unx * pHz);
                    s5 = s5 - (-unx * pHy + uny * pHx);
                    if constexpr (PML) { // the neighbour's face source arrives through the same sum
                        if (gqn >= 0) {
                            s3 = s3 - a.fs_val[gqn];
                            s4 = s4 - a.fs_val[a.fs_n + gqn];
                            s5 = s5 - a.fs_val[2 * a.fs_n + gqn];
                        }
                    }
                } else if (vp == -1) { // 'PEC' / 'PML' outer face: cem_maxwell_flux_pec :1397-1405
                    s0 = 2.0 * s0; s1 = 2.0 * s1; s2 = 2.0 * s2;
                    s3 = 0.0; s4 = 0.0; s5 = 0.0;
                }
                { // flux into resH (:976-986)
                    const double Y02 = -(hY * Y1), C02Y = hY * a.C0;
                    const double fu1 = uny * s5 - unz * s4;
                    const double fu2 = unz * s3 - unx * s5;
                    const double fu3 = unx * s4 - uny * s3;
                    fl[f][0] = ar * (Y02 * s0 - C02Y * fu1);
                    fl[f][1] = ar * (Y02 * s1 - C02Y * fu2);
                    fl[f][2] = ar * (Y02 * s2 - C02Y * fu3);
                }
                { // flux into resE (:987-997)
                    const double Z02 = hZ * Z1, C02Z = hZ * a.C0;
                    const double fw1 = uny * s2 - unz * s1;
                    const double fw2 = unz * s0 - unx * s2;
                    const double fw3 = unx * s1 - uny * s0;
                    fl[f][3] = ar * (Z02 * s3 - C02Z * fw1);
                    fl[f][4] = ar * (Z02 * s4 - C02Z * fw2);
                    fl[f][5] = ar * (Z02 * s5 - C02Z * fw3);
                }
            }
            // lifts into the residual: x-, y-, z-faces in turn (edge/corner nodes get 2/3 of them)
#pragma unroll
            for (int rd = 0; rd < 3; rd++) {
#pragma unroll
                for (int f = 0; f < FPT; f++) {
                    const int slot = fjs[f] < 0 ? -1 : fjs[f] / N2;
                    const int myrd = (slot == 1 || slot == 3) ? 0 : ((slot == 0 || slot == 2) ? 1 : 2);
                    if (slot >= 0 && myrd == rd && !(PIPE_SKIP & 4) &&
                        (!(PIPE_SKIP & 64) || fl[f][0] + fl[f][3] == 1.2345e300)) {
                        const int sn = fsn[f];
#pragma unroll
                        for (int c = 0; c < 6; c++) R[c * SC + (c >= 3 ? HE : 0) + sn] += fl[f][c];
                    }
                }
                __syncthreads();
                // every read of the staged face arrays is done: the landing zone takes the
                // fields of the next item
                if (rd == 0 && w0 && more) issue_fields(en, sn_);
            }
        };

        // ---- P4: t-pencils, thread (g,h,i,j) ---------------------------------------------------------
        auto t_phase = [&]() {
            if (!CM) mbar_wait(b_cot, par);
            const double *cofs = Z + par_of(a.met[6] + sbase);
            if constexpr ((PIPE_SKIP & 8) != 0) {
            } else if constexpr (KS == 1) pipe_t_phase<N, KS, 0, CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
            else {
                if (s == 0) pipe_t_phase<N, KS, 0, CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 1 && s == 1) pipe_t_phase<N, KS, (KS > 1 ? 1 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 2 && s == 2) pipe_t_phase<N, KS, (KS > 2 ? 2 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 3 && s == 3) pipe_t_phase<N, KS, (KS > 3 ? 3 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 4 && s == 4) pipe_t_phase<N, KS, (KS > 4 ? 4 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 5 && s == 5) pipe_t_phase<N, KS, (KS > 5 ? 5 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 6 && s == 6) pipe_t_phase<N, KS, (KS > 6 ? 6 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
                if (KS > 7 && s == 7) pipe_t_phase<N, KS, (KS > 7 ? 7 : 0), CM>(prm.D, a, U, R, cofs, cmt, W3, ebase, tid);
            }
            __syncthreads();
            if (w0 && more && !CM) issue_cot(en, sn_);
        };

        if constexpr (PIPE_FLUX_LAST) {
            // the gathers fly during the t-pencils; lifts are added after the whole volume curl
            // (the reference's own order: add_flux_to_res follows maxwell_wght_curl)
            gather();
            t_phase();
            flux_phase();
        } else {
            gather();
            flux_phase();
            t_phase();
        }

        // ---- P5: epilogue: auxiliary ODEs, inverse mass, low-storage RK update ---------------------
        mbar_wait(b_k, par);
        {
            const double *kx = X + par_of(a.kf + sbase);
#pragma unroll
            for (int x = 0; x < PER; x++) {
                const int pl = ps + NPL * x, g = pl / KB ? 1 : 0;
                const int kl = pl - g * KB;
                if (pok && pl < 2 * KB && kl < kb) {
                    const int nl = pnd + kl * N2;
                    const int sn = PLay<N>::at(pi, pj, kl) + (g == 0 ? HE : 0);
                    const long long gi = sbase + nl;
                    const int cb0 = g == 0 ? 3 : 0; // components being updated
                    double r[3] = {R[cb0 * SC + sn], R[(cb0 + 1) * SC + sn], R[(cb0 + 2) * SC + sn]};
                    const double o[3] = {U[cb0 * SC + sn], U[(cb0 + 1) * SC + sn], U[(cb0 + 2) * SC + sn]};
                    if (PML) { // auxiliary ODEs of this node, in the reference's order:
                        // pml_step (+ PML half of rk_maxwell_ab), then the usersrc ADEs
                        const int ef = a.elflag[e];
                        if (ef & 1) pml_node3(a, gi, g == 0, r, o);
                        if ((ef & 2) && g == 0 && a.ade_mask[gi]) {
#pragma unroll
                            for (int c = 0; c < 3; c++) r[c] = ade_component(a, gi, c, r[c], o[c]);
                        }
                    }
                    if (a.src_prof != nullptr) { // usersrc hook: res(comp) -= profile*(tfac*bm)
                        const int cs = a.src_comp - cb0;
                        if (cs >= 0 && cs < 3) {
                            const double sv2 = ldg(a.src_prof + gi) * (a.src_tfac * ldg(a.bmn + gi));
                            if (cs == 0) r[0] -= sv2;
                            else if (cs == 1) r[1] -= sv2;
                            else r[2] -= sv2;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double t = a.ca * kx[(cb0 + c) * XL + nl] + a.dt * (r[c] * mb[x]);
                        const double un = o[c] + a.cb * t;
                        __stcs(a.kf + (cb0 + c) * a.ld + gi, t);
                        __stcs(a.u_out + (cb0 + c) * a.ld + gi, un);
                        // x-face mirror of the new fields (stage_args.h): parked in this node's
                        // (now dead) residual slot, written out coalesced by xout() below
                        if (!(PIPE_SKIP & 128) && a.xtr_out != nullptr && (pi == 0 || pi == N - 1)) {
                            if (PIPE_XOUT) R[(cb0 + c) * SC + sn] = un;
                            else
                                a.xtr_out[(cb0 + c) * a.ldx + (2ll * e + (pi ? 1 : 0)) * N2 + pj +
                                          N * (k0 + kl)] = un;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (w0 && more && !CM) issue_cof(en, sn_);
        if (PIPE_XOUT && !(PIPE_SKIP & 128) && a.xtr_out != nullptr) { xo_e = e; xo_s = s; }
        e = en;
        en = enn;
#pragma unroll
        for (int f = 0; f < FPT; f++) fvp[f] = fvn[f];
    }
    if (PIPE_XOUT && xo_e >= 0) xout(xo_e, xo_s);
}

template <int N, bool PML, bool CM>
int pipe_launch_inst(const StageParams<N> &prm, cudaStream_t st)
{
    constexpr int KS = pipe_ks_for(N);
    using C = PT<N, KS>;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    static int grid_per_dev[64] = {};
    if (dev < 0 || dev >= 64) return 1;
    if (grid_per_dev[dev] == 0) {
        if (cudaFuncSetAttribute(pipe_kernel<N, KS, PML, CM>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::SMEM) != cudaSuccess)
            return 1;
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pipe_kernel<N, KS, PML, CM>, C::NT,
                                                          C::SMEM) != cudaSuccess || occ < 1)
            return 1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
        grid_per_dev[dev] = occ * sms;
    }
    const int items = KS * prm.a.nel;
    int grid = items < grid_per_dev[dev] ? items : grid_per_dev[dev];
    if (prm.a.grid_cap > 0 && grid > prm.a.grid_cap) grid = prm.a.grid_cap;
    pipe_kernel<N, KS, PML, CM><<<grid, C::NT, C::SMEM, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

template <int N>
int pipe_launch_n(const StageArgs &a, const double *Dhost, bool pml, bool cm, cudaStream_t st)
{
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    if (pml) return cm ? pipe_launch_inst<N, true, true>(prm, st) : pipe_launch_inst<N, true, false>(prm, st);
    return cm ? pipe_launch_inst<N, false, true>(prm, st) : pipe_launch_inst<N, false, false>(prm, st);
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

} // namespace

// One translation unit per order: the Makefile compiles this file once per nx1 with
// -DPIPE_ONLY_N=<nx1> (entry point launch_stage_pipe_n<nx1>); stage_pipe_dispatch.cu selects.
// -DPIPE_SINGLE (developer variants, scripts/build_variants.sh): this unit is the only order and
// also provides launch_stage_pipe itself.
#ifndef PIPE_ONLY_N
#error "compile with -DPIPE_ONLY_N=<nx1>"
#endif
#define PIPE_CAT2(a, b) a##b
#define PIPE_CAT(a, b) PIPE_CAT2(a, b)

// returns 0 ok, -1 the arrays do not meet the alignment the bulk copies need (the caller uses
// launch_stage_slab), >0 CUDA failure.  Dhost = dxm1 (n*n, column-major).
int PIPE_CAT(launch_stage_pipe_n, PIPE_ONLY_N)(const StageArgs &a, const double *Dhost, bool pml,
                                               bool cm, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    // the bulk copies assume 16-byte aligned arrays and an even leading dimension
    bool ok = aligned16(a.u_in) && aligned16(a.kf) && (a.ld % 2 == 0) && aligned16(a.unx) &&
              aligned16(a.uny) && aligned16(a.unz) && aligned16(a.area) && aligned16(a.hY) &&
              aligned16(a.Y1) && aligned16(a.hZ) && aligned16(a.Z1);
    for (int q = 0; q < 9; q++) ok = ok && aligned16(a.met[q]);
    if (!ok) return -1;
    return pipe_launch_n<PIPE_ONLY_N>(a, Dhost, pml, cm, st);
}

#ifdef PIPE_SINGLE
int launch_stage_pipe(const StageArgs &a, const double *Dhost, int nx1, bool pml, bool cm,
                      void *stream)
{
    if (nx1 != PIPE_ONLY_N) return -1;
    return PIPE_CAT(launch_stage_pipe_n, PIPE_ONLY_N)(a, Dhost, pml, cm, stream);
}
#endif

} // namespace nkb
