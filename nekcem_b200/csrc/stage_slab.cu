// Fused Maxwell RK-stage kernel for sm_100a, "element-slab" formulation, 2 <= nx1 <= 24
// (the reference ships mxf1..mxf24, src/nek5_mxm_std.F; orders above 16 are covered for
// completeness with slabs of three planes and are not tuned).
//
// One launch = one RK stage over a list of elements.  One CTA = one k-SLAB of one element
// (slab = all nodes with k0 <= k < k0+kb; KS slabs per element, KS = 1 for nx1 <= 8), and it
// updates ALL SIX components of those nodes, so every array of the stage crosses L2->SM once per
// node: fields, RK registers, the nine cofactors, the two masses, and one neighbour trace per
// face point.  The slabs of an element are adjacent CTAs; what they
// share (the t-lines through the other slabs) is served by L2.
//
//   P0  stage H,E of the slab in shared memory U[6]
//   P1  r-pencils: thread (g,j,k) holds nothing but 3*NO accumulators; the line of 3 source
//       components streams from smem, D(i,m) is a constant-bank operand; raw r-derivatives -> R
//       [local_grad3, src/nek5_grad.F:2-19; mxfK left-to-right sums, src/nek5_mxm_std.F:173-190]
//       g = 0: resE <- curl H,  g = 1: resH <- -curl E      [cem_maxwell :510-602]
//   P2  s-pencils: thread (g,i,k): s-derivatives, then the r- and s-parts of the weighted curl
//       with the cofactors rx..sz (coalesced along i) and w3mn  [maxwell_wght_curl :1428-1497]
//   P3  surface flux: one thread per face point of the slab computes BOTH numerical fluxes from
//       one gather of the neighbour trace (vmapP = the gs_op_fields pair-sum of :962, or the
//       NCCL halo), own trace from smem, PEC mirror, upwind/central; the lifts are added to R in
//       three barrier-separated rounds (x-, y-, z-faces) so that edge nodes are race-free
//       [restrict_to_face :604-652, flux3d :922-1002, flux_pec :1368-1426, add_flux_to_res :725-735]
//   P4  t-pencils: thread (g,i,j) adds the weighted t-part for the slab's k; the line comes from
//       smem (KS = 1) or from global memory/L2 (KS > 1: the other slabs' nodes)
//   P5  streaming epilogue: PML ADEs [pml_step, src/cem_maxwell_pml.F:508-592], Drude/Lorentz
//       currents [cem_maxwell_drude/_lorentz, src/cem_maxwell.F:3095-3211], volume source
//       [usersrc hook :503], inverse mass [invqmass :1878-1886], low-storage RK update
//       [rk4_upd, src/cem_common.F:18-76]; old fields come from smem, k/masses are loaded
//       EPI nodes ahead, results go to the ping-pong buffer.
//
// Arithmetic: same products as the reference; the 6-term curl sum is associated by direction
// ((r-part + s-part)*w + lift) + w*t-part, and nvcc contracts a*b+c into FMA -- both are
// <= 1e-15 relative effects per operation (DESIGN.md "Numerics").
#include "stage_aux.h"

namespace nkb {
namespace {

// ---- compile-time geometry -----------------------------------------------------------------
// Build-time tunables (developer sweeps: scripts/build_variants.sh)
#ifndef SLAB_EPI
#define SLAB_EPI 4
#endif
#ifndef SLAB_PF_MODE
#define SLAB_PF_MODE 1 // 0: prefetch everything at CTA start; 1: staged (just in time); 2: none
#endif
#ifndef SLAB_ST_CS
#define SLAB_ST_CS 1 // 1: streaming (evict-first) stores of the results
#endif
#ifndef SLAB_T_PIPE
#define SLAB_T_PIPE 1 // 1: software-pipelined t-lines from global memory (KS > 1)
#endif
#ifndef SLAB_W3S
#define SLAB_W3S 1 // 1: the slab's share of w3mn in shared memory where that pays; 0: always through L1
#endif
#ifndef SLAB_COF_SLOTS
#define SLAB_COF_SLOTS 1
#endif
__host__ __device__ constexpr int ks_for(int n)
{
#ifdef KS_OVERRIDE_N
    if (n == KS_OVERRIDE_N) return KS_OVERRIDE;
#endif
    // nx1 = 13: three slabs (4, 4, 5 planes) on 352 threads (two i-j planes per pass of the
    // plane-mapped loops; measured, same box: 0.485 with four slabs on 256 threads, 0.53 with four
    // on 352, 0.56 with three on 352, 0.45-0.53 with five on 192).  Thinner slabs lose at nx1 = 14, 15
    // as well (five slabs: 0.54 -> 0.51, 0.565 -> 0.53): the t-lines are re-read once per slab.
    if (n > 16) return (n + 2) / 3; // nx1 = 17..24: three planes per slab (83-166 KB of shared memory)
    return n <= 8 ? 1 : (n <= 12 ? 2 : (n == 13 ? 3 : 4));
}
__host__ __device__ constexpr int rsplit_for(int n) { return (n <= 5 || n > 16) ? 4 : 2; }
__host__ __device__ constexpr int round32(int x) { return ((x + 31) / 32) * 32; }
__host__ __device__ constexpr int nt_for(int items)
{
    if (items <= 320) return round32(items);
    int passes = (items + 255) / 256;
    return round32((items + passes - 1) / passes);
}

template <int N, int KS>
struct Slab {
    static constexpr int N2 = N * N;
    static constexpr int KB = (N + KS - 1) / KS; // thickest slab
    static constexpr int SPLIT = rsplit_for(N);  // threads per r/s pencil
    static constexpr int NO = (N + SPLIT - 1) / SPLIT;
    // work items of a pencil phase: (h, g, pencil) with the output share h slowest and the
    // (g, pencil) block padded to whole warps, so that h -- the only index that selects
    // different code -- is warp-uniform
    static constexpr int RS_BLK = round32(2 * N * KB);
    static constexpr int RS_ITEMS = SPLIT * RS_BLK;
    static constexpr int TSPLIT = KS == 1 ? SPLIT : 1; // threads per t pencil
    static constexpr int T_BLK = round32(2 * N2);
    static constexpr int T_ITEMS = TSPLIT * T_BLK;
#ifdef SLAB_NT
    static constexpr int NT = SLAB_NT;
#else
    // nx1 = 13: two i-j planes (338 nodes) per pass of the plane-mapped staging and epilogue loops
    // instead of one plane on 256 threads (see ks_for; the same move at nx1 = 14, 416 threads at 72
    // registers, loses: 0.539 -> 0.501, and at nx1 = 11, 384 threads at 80 registers: 0.580 -> 0.576).
    // Every order: at least one i-j plane of threads (only nx1 > 16 has fewer pencil items than
    // nodes in a plane).
    static constexpr int NT = N == 13 ? 352 : (nt_for(RS_ITEMS) > round32(N2) ? nt_for(RS_ITEMS) : round32(N2));
#endif
    static constexpr int SC = Lay<N>::SK * KB; // component stride in smem
    static constexpr int FXY = 4 * N * KB;     // face points on the x/y faces of a slab
    static constexpr int FZ = KS == 1 ? 2 * N2 : N2;
    static constexpr int F_ITEMS = FXY + FZ;
    static constexpr int FPT = (F_ITEMS + NT - 1) / NT;
    // U[6][SC], R[6][SC] and, for nx1 = 12, 13, the slab's share of w3mn (linear, KB*N2).  Every
    // node reads its weight four times (s- and t-part of both curl groups) and the streaming loads
    // keep evicting the table from L1: the pipelined kernel gained 2-5 % at nx1 = 9, 10 from the
    // move.  Here, same box, general path: nx1 = 12, 13 +2 %; nx1 = 11, 14, 15 +-0; nx1 = 16 -11 %
    // (two CTAs then need the 228 KB carve-out and the kernel's other loads lose half their L1).
    static constexpr bool W3S = SLAB_W3S && (N == 12 || N == 13);
    static constexpr size_t SMEM = sizeof(double) * (12 * SC + (W3S ? KB * N2 : 0));
    __host__ __device__ static constexpr int k0(int s) { return s * N / KS; }
    __host__ __device__ static constexpr int kb(int s) { return (s + 1) * N / KS - s * N / KS; }
    // cofactor batches: outputs whose cofactors are loaded together
#ifdef SLAB_PB_S
    static constexpr int PB_S = SLAB_PB_S < NO ? SLAB_PB_S : NO;
#else
    static constexpr int PB_S = NO <= 2 ? NO : (NO <= 6 ? (NO + 1) / 2 : (NO + 3) / 4);
#endif
#ifdef SLAB_REG_CAP
    static constexpr int REG_CAP = SLAB_REG_CAP;
#else
    static constexpr int REG_CAP = N == 13 ? 88 : (NO >= 7 ? 128 : 96);
#endif
    static constexpr int MINB_SMEM = (227 * 1024) / ((int)SMEM + 1024);
    static constexpr int MINB_THR = 2048 / NT;
    static constexpr int MINB_REG = 65536 / (NT * REG_CAP);
    static constexpr int MINB0 = MINB_SMEM < MINB_THR ? MINB_SMEM : MINB_THR;
    static constexpr int MINB1 = MINB0 < MINB_REG ? MINB0 : MINB_REG;
    static constexpr int MINB = MINB1 < 1 ? 1 : (MINB1 > 8 ? 8 : MINB1);
    // constant-metric instantiation: no cofactor batches in registers, so 80 registers hold the
    // pencil phases without spilling and a third CTA fits per SM where shared memory allows
    // (n <= 10): measured +10 % at n=8, +3..5 % at n=9,10; the general path loses at n=10
#ifdef SLAB_REG_CAP_CM
    static constexpr int REG_CAP_CM = SLAB_REG_CAP_CM;
#else
    static constexpr int REG_CAP_CM = (MINB_SMEM >= 3 && NO <= 5) ? 80 : REG_CAP;
#endif
    static constexpr int MINB_REG_CM = 65536 / (NT * REG_CAP_CM);
    static constexpr int MINB2 = MINB0 < MINB_REG_CM ? MINB0 : MINB_REG_CM;
    static constexpr int MINB_CM = MINB2 < 1 ? 1 : (MINB2 > 8 ? 8 : MINB2);
    static constexpr int EPI = SLAB_EPI; // nodes in flight per thread in the epilogue
};

// ---- one pencil phase -------------------------------------------------------------------------
// DIR 0/1/2 = r/s/t.  The thread produces outputs O0..O1-1 (positions along the pencil; for
// DIR 2 these are element k indices, and KOFF = k0 of the slab) of pencil (pa,pb):
//   d_c = sum_m D(o,m) u_c(m)                               (mxfK order, left to right)
//   DIR 0: R = d                                             (raw r-derivatives)
//   DIR 1: R = (curl_part(R; rx,ry,rz) + curl_part(d; sx,sy,sz)) * (sg*w3)
//   DIR 2: R = R + (sg*w3) * curl_part(d; tx,ty,tz)
// gsrc != nullptr (DIR 2 only): the line is read from global memory (3 components, stride ld).
template <int N, int DIR, int KOFF>
__device__ __forceinline__ int s_at(int m, int pa, int pb)
{
    return DIR == 0 ? Lay<N>::at(m, pa, pb)
                    : (DIR == 1 ? Lay<N>::at(pa, m, pb) : Lay<N>::at(pa, pb, m - KOFF));
}
template <int N, int DIR>
__device__ __forceinline__ int n_at(int m, int pa, int pb)
{
    return DIR == 0 ? m + N * pa + N * N * pb : (DIR == 1 ? pa + N * m + N * N * pb : pa + N * pb + N * N * m);
}

template <int N, int DIR, int KOFF, int O0, int O1, int PB, int SC, bool GSRC, bool CM>
__device__ __forceinline__ void pencil_phase(const double (&D)[N * N], const StageArgs &a,
                                             const double *U, const double *gsrc, double *R,
                                             const double *w3s, int pa, int pb, long long gbase,
                                             int wbase, double sg, long long mbase)
{
    // cofactor of node nd: a.met[q][mbase + nd] in general (mbase = gbase).  CM = elements
    // with constant metrics (found bitwise at setup): mbase = first node of the element and
    // every node reads that one (cached) value -- the same numbers, 72 B/node less traffic
    constexpr int NO = O1 - O0;
    constexpr int NCOF = DIR == 1 ? 6 : 3;
    if constexpr (NO > 0) {
        constexpr int NB = (NO + PB - 1) / PB;                   // cofactor batches
        constexpr int NS = SLAB_COF_SLOTS < NB ? SLAB_COF_SLOTS : NB; // batches in flight
        double cof[NS][PB][NCOF], wv[NS][PB];
        auto load_cof = [&](int b, int sl) {
#pragma unroll
            for (int x = 0; x < PB; x++) {
                const int o = O0 + b * PB + x < O1 ? O0 + b * PB + x : O1 - 1;
                const int nd = n_at<N, DIR>(o, pa, pb);
                const long long gi = CM ? mbase : mbase + nd;
                if constexpr (DIR == 1) {
#pragma unroll
                    for (int q = 0; q < 6; q++) cof[sl][x][q] = ldg(a.met[q] + gi);
                } else if constexpr (DIR == 2) {
#pragma unroll
                    for (int q = 0; q < 3; q++) cof[sl][x][q] = ldg(a.met[6 + q] + gi);
                }
                if constexpr (DIR != 0) wv[sl][x] = sg * w3s[wbase + nd];
            }
        };
        // cofactors of the first batch(es): in flight during the contraction
        if constexpr (DIR != 0) {
#pragma unroll
            for (int b = 0; b < NS; b++) load_cof(b, b);
        }
        // output-stationary contraction: the line streams by once, each point feeds the 3*NO
        // accumulators of this thread (sum over m left to right, as mxfK)
        double acc[3][NO];
        if constexpr (GSRC) {
            // line from global memory: all N loads of a component are issued before its FMAs
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double ul[N];
#pragma unroll
                for (int m = 0; m < N; m++) ul[m] = ldg(gsrc + c * a.ld + n_at<N, DIR>(m, pa, pb));
#pragma unroll
                for (int m = 0; m < N; m++)
#pragma unroll
                    for (int o = 0; o < NO; o++) {
                        const double dv = D[(O0 + o) + N * m];
                        if (m == 0) acc[c][o] = dv * ul[m];
                        else acc[c][o] = acc[c][o] + dv * ul[m];
                    }
            }
        } else {
#pragma unroll
            for (int m = 0; m < N; m++) {
                const int so = s_at<N, DIR, KOFF>(m, pa, pb);
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
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int sl = b % NS;
#pragma unroll
            for (int x = 0; x < PB; x++) {
                const int oo = b * PB + x;
                if (oo < NO) {
                    const double d[3] = {acc[0][oo], acc[1][oo], acc[2][oo]};
                    double c[3];
                    double *Ro = R + s_at<N, DIR, KOFF>(O0 + oo, pa, pb);
                    if constexpr (DIR == 0) {
                        Ro[0] = d[0]; Ro[SC] = d[1]; Ro[2 * SC] = d[2];
                    } else if constexpr (DIR == 1) {
                        const double dr[3] = {Ro[0], Ro[SC], Ro[2 * SC]};
                        double cr[3];
                        curl_part(dr, cof[sl][x][0], cof[sl][x][1], cof[sl][x][2], cr);
                        curl_part(d, cof[sl][x][3], cof[sl][x][4], cof[sl][x][5], c);
                        Ro[0] = (cr[0] + c[0]) * wv[sl][x];
                        Ro[SC] = (cr[1] + c[1]) * wv[sl][x];
                        Ro[2 * SC] = (cr[2] + c[2]) * wv[sl][x];
                    } else {
                        curl_part(d, cof[sl][x][0], cof[sl][x][1], cof[sl][x][2], c);
                        Ro[0] = Ro[0] + wv[sl][x] * c[0];
                        Ro[SC] = Ro[SC] + wv[sl][x] * c[1];
                        Ro[2 * SC] = Ro[2 * SC] + wv[sl][x] * c[2];
                    }
                }
            }
            // refill the slot just consumed with the batch NS ahead
            if constexpr (DIR != 0) {
                if (b + NS < NB) load_cof(b + NS, sl);
            }
        }
    }
}

// dispatch on the thread's share h of the outputs LO..HI-1 (compile-time ranges)
template <int N, int DIR, int KOFF, int LO, int HI, int SPLIT, int PB, int SC, bool GSRC, bool CM>
__device__ __forceinline__ void pencil_split(const double (&D)[N * N], const StageArgs &a,
                                             const double *U, const double *gsrc, double *R,
                                             const double *w3s, int pa, int pb, long long gbase,
                                             int wbase, double sg, int h, long long mbase)
{
    constexpr int L = HI - LO, HN = (L + SPLIT - 1) / SPLIT;
    constexpr int E1 = LO + (HN < L ? HN : L), E2 = LO + (2 * HN < L ? 2 * HN : L),
                  E3 = LO + (3 * HN < L ? 3 * HN : L);
    if (h == 0) pencil_phase<N, DIR, KOFF, LO, E1, PB, SC, GSRC, CM>(D, a, U, gsrc, R, w3s, pa, pb, gbase, wbase, sg, mbase);
    if (SPLIT > 1 && h == 1)
        pencil_phase<N, DIR, KOFF, E1, E2, PB, SC, GSRC, CM>(D, a, U, gsrc, R, w3s, pa, pb, gbase, wbase, sg, mbase);
    if (SPLIT > 2 && h == 2)
        pencil_phase<N, DIR, KOFF, E2, E3, PB, SC, GSRC, CM>(D, a, U, gsrc, R, w3s, pa, pb, gbase, wbase, sg, mbase);
    if (SPLIT > 3 && h == 3)
        pencil_phase<N, DIR, KOFF, E3, HI, PB, SC, GSRC, CM>(D, a, U, gsrc, R, w3s, pa, pb, gbase, wbase, sg, mbase);
}

// t-pencils of slab S (compile-time k range)
template <int N, int KS, int S, bool CM>
__device__ __forceinline__ void t_phase(const double (&D)[N * N], const StageArgs &a, const double *U,
                                        double *R, const double *w3s, long long ebase, int tid)
{
    using C = Slab<N, KS>;
    constexpr int K0 = C::k0(S), K1 = K0 + C::kb(S);
    // nx1 > 16: no software pipeline (two t-lines of 17..24 values would not fit the registers)
    if constexpr (KS == 1 || !SLAB_T_PIPE || (N > 16)) {
        constexpr int NOT = (K1 - K0 + C::TSPLIT - 1) / C::TSPLIT;
        constexpr int PB_T = NOT <= 8 ? NOT : (NOT + 1) / 2;
#pragma unroll 1
        for (int w = tid; w < C::T_ITEMS; w += C::NT) {
            const int h = w / C::T_BLK, r = w - h * C::T_BLK;
            const int g = r / C::N2, p = r - g * C::N2;
            if (g > 1) continue;
            const int pa = p % N, pb = p / N;
            const double *Us = U + (g ? 3 : 0) * C::SC;
            double *Rd = R + (g ? 0 : 3) * C::SC;
            const double *gs = a.u_in + (g ? 3 : 0) * a.ld + ebase;
            pencil_split<N, 2, K0, K0, K1, C::TSPLIT, PB_T, C::SC, (KS > 1), CM>(
                D, a, Us, gs, Rd, w3s, pa, pb, ebase, -K0 * C::N2, g ? -1.0 : 1.0, h, ebase);
        }
    } else {
        // KS > 1: the lines come from global memory (L2).  Software pipeline over the flat
        // sequence of (item, component) steps: the N loads of step+1 are issued before the FMAs
        // of step, so one line is always in flight behind the arithmetic.
        constexpr int NO = K1 - K0;
        constexpr int NPASS = (C::T_ITEMS + C::NT - 1) / C::NT;
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
        double acc[3][NO], cof[NO][3], wv[NO];
#pragma unroll
        for (int step = 0; step < NSTEP; step++) {
            const int c = step % 3, buf = step & 1;
            int g, nd; bool ok;
            item(step / 3, g, nd, ok);
            if (step + 1 < NSTEP) issue(step + 1, buf ^ 1);
            if (c == 0) { // cofactors and weight of this item's outputs
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    const long long gi = CM ? ebase : ebase + nd + C::N2 * (K0 + o);
#pragma unroll
                    for (int q = 0; q < 3; q++) cof[o][q] = ldg(a.met[6 + q] + gi);
                    wv[o] = (g ? -1.0 : 1.0) * w3s[nd + C::N2 * o];
                }
            }
#pragma unroll
            for (int m = 0; m < N; m++)
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    const double dv = D[(K0 + o) + N * m];
                    if (m == 0) acc[c][o] = dv * ul[buf][m];
                    else acc[c][o] = acc[c][o] + dv * ul[buf][m];
                }
            if (c == 2 && ok) {
                double *Rd = R + (g ? 0 : 3) * C::SC;
                const int pa = nd % N, pb = nd / N;
#pragma unroll
                for (int o = 0; o < NO; o++) {
                    const double d[3] = {acc[0][o], acc[1][o], acc[2][o]};
                    double cc[3];
                    curl_part(d, cof[o][0], cof[o][1], cof[o][2], cc);
                    double *Ro = Rd + Lay<N>::at(pa, pb, o);
                    Ro[0] = Ro[0] + wv[o] * cc[0];
                    Ro[C::SC] = Ro[C::SC] + wv[o] * cc[1];
                    Ro[2 * C::SC] = Ro[2 * C::SC] + wv[o] * cc[2];
                }
            }
        }
    }
}

template <int N, int KS, bool PML, bool CM>
__global__ void __launch_bounds__(Slab<N, KS>::NT, CM ? Slab<N, KS>::MINB_CM : Slab<N, KS>::MINB)
    slab_kernel(const __grid_constant__ StageParams<N> prm)
{
    using C = Slab<N, KS>;
    constexpr int N2 = N * N, N3 = N2 * N, NF = 6 * N2;
    constexpr int NT = C::NT, SC = C::SC, KB = C::KB, FPT = C::FPT;
    const StageArgs &a = prm.a;
    extern __shared__ double smem[];
    double *U = smem;          // [6][SC] H,E of the slab at stage start
    double *R = smem + 6 * SC; // [6][SC] residuals resH,resE
    double *W3sm = smem + 12 * SC; // [kb*N2] w3mn of the slab's nodes

    const int tid = threadIdx.x;
    const int e = a.elist[blockIdx.x / KS];
    const int s = blockIdx.x % KS;
    const int k0 = C::k0(s), kb = C::kb(s);
    const long long ebase = (long long)e * N3;
    const long long sbase = ebase + k0 * N2; // first node of the slab
    const int nslab = N2 * kb;

    // L2 prefetch of the slab's share of the volume arrays [first,last) of the list
    // rx..sz (0-5), tx..tz (6-8), kH,kE (9-14), hbm1, ebm1 (15,16): one warp per array
    auto pf_vol = [&](int first, int last) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int arr = first + warp; arr < last; arr += NT / 32) {
            const double *base;
            if (arr < 9) base = a.met[arr];
            else if (arr < 15) base = a.kf + (arr - 9) * a.ld;
            else base = arr == 15 ? a.hbm1 : a.ebm1;
            prefetch_chunk(base + sbase, nslab * 8, lane);
        }
    };

    // ---- P0: stage the six field components of the slab (loads issued before anything else) ----
    // plane mapping (no per-value index arithmetic): thread = (plane slot ps, node (pi,pj) of an
    // i-j plane); the 6*KB planes (component c, local k) of the slab are handled NPL at a time
    constexpr int NPL = NT / N2;
    static_assert(NPL >= 1, "block smaller than one i-j plane");
    const int ps = tid / N2, pnd = tid - ps * N2;
    const int pi = pnd % N, pj = pnd / N;
    const bool pok = ps < NPL;
    constexpr int SPER = (6 * KB + NPL - 1) / NPL;
    constexpr int SUNR = SPER < 24 ? SPER : 24;
    double sv[SUNR];
#pragma unroll
    for (int x = 0; x < SUNR; x++) {
        const int pl = ps + NPL * x, c = pl / KB, kl = pl - c * KB;
        sv[x] = (pok && pl < 6 * KB && kl < kb) ? ldg(a.u_in + c * a.ld + sbase + kl * N2 + pnd) : 0.0;
    }

    // ---- prologue: put every other HBM request of this slab in flight now -----------------------
    // (a) L2 prefetch of the slab's metric, mass and RK-register arrays: one warp per array
    {
        if (!CM) pf_vol(0, SLAB_PF_MODE == 0 ? 17 : (SLAB_PF_MODE == 1 ? 6 : 0));
        else if (SLAB_PF_MODE == 0) pf_vol(9, 17);
        // face geometry / impedances / vmapP of the slab's part of the four x/y faces and of
        // its z face(s): 9 arrays x (4 strips of N*kb points + whole z faces)
        constexpr int LXY = (N * KB * 8 + 127) / 128 + 1; // lines per strip (any alignment)
        constexpr int LZ = (N2 * 8 + 127) / 128 + 1;
        const long long fbase = (long long)e * NF;
        for (int t = tid; t < 9 * 4 * LXY; t += NT) {
            const int arr = t / (4 * LXY), r = t - arr * (4 * LXY), f = r / LXY, ln = r - f * LXY;
            const long long off = fbase + f * N2 + N * k0;
            const int bytes = (arr == 8 ? 4 : 8) * N * kb;
            const char *b;
            if (arr == 8) b = (const char *)(a.vmapP + off);
            else {
                const double *fa = arr == 0 ? a.unx : arr == 1 ? a.uny : arr == 2 ? a.unz
                                 : arr == 3 ? a.area : arr == 4 ? a.hY : arr == 5 ? a.Y1
                                 : arr == 6 ? a.hZ : a.Z1;
                b = (const char *)(fa + off);
            }
            if (ln * 128 < bytes + 127) prefetch_l2(b + (ln * 128 < bytes ? ln * 128 : bytes - 4));
        }
        if (s == 0 || s == KS - 1) {
            for (int t = tid; t < 9 * 2 * LZ; t += NT) {
                const int arr = t / (2 * LZ), r = t - arr * (2 * LZ), zf = r / LZ, ln = r - zf * LZ;
                if ((zf == 0 && s != 0) || (zf == 1 && s != KS - 1)) continue;
                const long long off = fbase + (4 + zf) * N2;
                const int bytes = (arr == 8 ? 4 : 8) * N2;
                const char *b;
                if (arr == 8) b = (const char *)(a.vmapP + off);
                else {
                    const double *fa = arr == 0 ? a.unx : arr == 1 ? a.uny : arr == 2 ? a.unz
                                     : arr == 3 ? a.area : arr == 4 ? a.hY : arr == 5 ? a.Y1
                                     : arr == 6 ? a.hZ : a.Z1;
                    b = (const char *)(fa + off);
                }
                if (ln * 128 < bytes + 127) prefetch_l2(b + (ln * 128 < bytes ? ln * 128 : bytes - 4));
            }
        }
    }
    if (C::W3S)
        for (int i = tid; i < nslab; i += NT) W3sm[i] = ldg(a.w3 + k0 * N2 + i);
    const double *W3s = C::W3S ? W3sm : a.w3 + k0 * N2;
    // (b) this thread's face points: slot, smem node, round; neighbour ids
    //     face slots in the reference's order (cemface, cem_common.F:234-260): -y,+x,+y,-x,-z,+z
    int fvp[FPT], fsn[FPT], fjs[FPT];
#pragma unroll
    for (int f = 0; f < FPT; f++) {
        const int q = tid + f * NT;
        int slot = -1, fp0 = 0, ci = 0, cj = 0, kl = 0;
        if (q < C::FXY) {
            const int f4 = q / (N * KB), r = q - f4 * (N * KB);
            const int fa = r % N;
            kl = r / N;
            if (kl < kb) {
                fp0 = fa + N * (k0 + kl);
                if (f4 == 0) { slot = 3; ci = 0; cj = fa; }
                else if (f4 == 1) { slot = 1; ci = N - 1; cj = fa; }
                else if (f4 == 2) { slot = 0; ci = fa; cj = 0; }
                else { slot = 2; ci = fa; cj = N - 1; }
            }
        } else if (q < C::F_ITEMS) {
            const int q2 = q - C::FXY, zf = q2 / N2;
            fp0 = q2 - zf * N2;
            ci = fp0 % N; cj = fp0 / N;
            if (KS == 1) { slot = 4 + zf; kl = zf ? N - 1 : 0; }
            else if (s == 0) { slot = 4; kl = 0; }
            else if (s == KS - 1) { slot = 5; kl = kb - 1; }
        }
        fjs[f] = slot < 0 ? -1 : slot * N2 + fp0;
        fsn[f] = Lay<N>::at(ci, cj, kl);
        fvp[f] = slot < 0 ? -2 : ldg(a.vmapP + (long long)e * NF + fjs[f]);
    }
    // staged values -> smem
#pragma unroll 1
    for (int q0 = 0; q0 < SPER; q0 += SUNR) {
        if (q0 > 0) {
#pragma unroll
            for (int x = 0; x < SUNR; x++) {
                const int pl = ps + NPL * (q0 + x), c = pl / KB, kl = pl - c * KB;
                sv[x] = (pok && pl < 6 * KB && kl < kb) ? ldg(a.u_in + c * a.ld + sbase + kl * N2 + pnd) : 0.0;
            }
        }
#pragma unroll
        for (int x = 0; x < SUNR; x++) {
            const int pl = ps + NPL * (q0 + x), c = pl / KB, kl = pl - c * KB;
            if (pok && pl < 6 * KB && kl < kb) U[c * SC + Lay<N>::at(pi, pj, kl)] = sv[x];
        }
    }
    // (c) neighbour traces of those face points -> L2
#pragma unroll
    for (int f = 0; f < FPT; f++)
        if (fvp[f] >= 0 || -(fvp[f] + 3) >= XTR_BIAS) {
            long long st;
            const double *nb = nbr_trace(a, fvp[f], sbase, st);
#pragma unroll
            for (int c = 0; c < 6; c++) prefetch_l2(nb + c * st);
        }
    __syncthreads();

    // ---- P1: r-pencils, thread (g,h,j,k): raw derivatives ----------------------------------------
#pragma unroll 1
    for (int w = tid; w < C::RS_ITEMS; w += NT) {
        const int h = w / C::RS_BLK, r = w - h * C::RS_BLK;
        const int g = r / (N * KB), p = r - g * (N * KB);
        const int pa = p % N, pb = p / N;
        if (g < 2 && pb < kb)
            pencil_split<N, 0, 0, 0, N, C::SPLIT, C::NO, SC, false, CM>(
                prm.D, a, U + (g ? 3 : 0) * SC, nullptr, R + (g ? 0 : 3) * SC, W3s, pa, pb, sbase,
                0, g ? -1.0 : 1.0, h, sbase);
    }
    __syncthreads();
    if (SLAB_PF_MODE == 1 && !CM) pf_vol(6, 9);
    // ---- P2: s-pencils, thread (g,h,i,k): r- and s-parts of the weighted curl ------------------
#pragma unroll 1
    for (int w = tid; w < C::RS_ITEMS; w += NT) {
        const int h = w / C::RS_BLK, r = w - h * C::RS_BLK;
        const int g = r / (N * KB), p = r - g * (N * KB);
        const int pa = p % N, pb = p / N;
        if (g < 2 && pb < kb)
            pencil_split<N, 1, 0, 0, N, C::SPLIT, C::PB_S, SC, false, CM>(
                prm.D, a, U + (g ? 3 : 0) * SC, nullptr, R + (g ? 0 : 3) * SC, W3s, pa, pb, sbase,
                0, g ? -1.0 : 1.0, h, CM ? ebase : sbase);
    }
    __syncthreads();
    if (SLAB_PF_MODE == 1) pf_vol(9, 17);

    // ---- P3: surface flux: both fluxes of a face point from one neighbour gather ---------------
    {
        double fl[FPT][6];
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            const bool valid = fjs[f] >= 0;
            const long long jf = (long long)e * NF + (valid ? fjs[f] : 0);
            const int vp = fvp[f];
            const int sn = fsn[f];
            // all loads first
            const double unx = ldg(a.unx + jf), uny = ldg(a.uny + jf), unz = ldg(a.unz + jf);
            const double ar = ldg(a.area + jf);
            const double hY = ldg(a.hY + jf), Y1 = ldg(a.Y1 + jf);
            const double hZ = ldg(a.hZ + jf), Z1 = ldg(a.Z1 + jf);
            // neighbour trace: volume node, halo slot, or (unused) the first node of the slab
            long long st;
            const double *nb = nbr_trace(a, vp, sbase, st);
            double pv[6];
#pragma unroll
            for (int c = 0; c < 6; c++) pv[c] = ldg(nb + c * st);
            double Hx = U[sn], Hy = U[SC + sn], Hz = U[2 * SC + sn];
            double Ex = U[3 * SC + sn], Ey = U[4 * SC + sn], Ez = U[5 * SC + sn];
            double pHx = pv[0], pHy = pv[1], pHz = pv[2], pEx = pv[3], pEy = pv[4], pEz = pv[5];
            if (a.inc_own != nullptr && valid) { // userinc hook (src/cem_maxwell.F:498)
                const int qo = a.inc_own[jf], qn = a.inc_nbr[jf];
                if (qo >= 0) {
                    const double ui = cos(a.inc_phase[qo] - a.inc_wt);
                    Hx += a.inc_amp[qo] * ui; Hy += a.inc_amp[a.inc_n + qo] * ui;
                    Hz += a.inc_amp[2 * a.inc_n + qo] * ui;
                    Ex += a.inc_amp[3 * a.inc_n + qo] * ui;
                    Ey += a.inc_amp[4 * a.inc_n + qo] * ui;
                    Ez += a.inc_amp[5 * a.inc_n + qo] * ui;
                }
                if (qn >= 0) {
                    const double ui = cos(a.inc_phase[qn] - a.inc_wt);
                    pHx += a.inc_amp[qn] * ui; pHy += a.inc_amp[a.inc_n + qn] * ui;
                    pHz += a.inc_amp[2 * a.inc_n + qn] * ui;
                    pEx += a.inc_amp[3 * a.inc_n + qn] * ui;
                    pEy += a.inc_amp[4 * a.inc_n + qn] * ui;
                    pEz += a.inc_amp[5 * a.inc_n + qn] * ui;
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
                if (slot >= 0 && myrd == rd) {
                    const int sn = fsn[f];
#pragma unroll
                    for (int c = 0; c < 6; c++) R[c * SC + sn] += fl[f][c];
                }
            }
            __syncthreads();
        }
    }

    // ---- P4: t-pencils, thread (g,h,i,j) ---------------------------------------------------------
    if constexpr (KS == 1) t_phase<N, KS, 0, CM>(prm.D, a, U, R, W3s, ebase, tid);
    else {
        if (s == 0) t_phase<N, KS, 0, CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 1 && s == 1) t_phase<N, KS, (KS > 1 ? 1 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 2 && s == 2) t_phase<N, KS, (KS > 2 ? 2 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 3 && s == 3) t_phase<N, KS, (KS > 3 ? 3 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 4 && s == 4) t_phase<N, KS, (KS > 4 ? 4 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 5 && s == 5) t_phase<N, KS, (KS > 5 ? 5 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 6 && s == 6) t_phase<N, KS, (KS > 6 ? 6 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
        if (KS > 7 && s == 7) t_phase<N, KS, (KS > 7 ? 7 : 0), CM>(prm.D, a, U, R, W3s, ebase, tid);
    }
    __syncthreads();

    // ---- P5: streaming epilogue --------------------------------------------------------------------
    {
        constexpr int PER = (2 * KB + NPL - 1) / NPL; // (g, local k) planes per thread
        constexpr int UNR = PER < C::EPI ? PER : C::EPI;
#pragma unroll 1
        for (int q0 = 0; q0 < PER; q0 += UNR) {
            double kk[UNR][3], mb[UNR];
#pragma unroll
            for (int x = 0; x < UNR; x++) {
                const int pl = ps + NPL * (q0 + x), g = pl / KB ? 1 : 0;
                int kl = pl - g * KB;
                kl = kl < kb ? kl : kb - 1;
                const int nl = (pok ? pnd : 0) + kl * N2;
                const long long cold = (g == 0 ? 3 : 0) * a.ld;
#pragma unroll
                for (int c = 0; c < 3; c++) kk[x][c] = a.kf[cold + c * a.ld + sbase + nl];
                mb[x] = ldg((g == 0 ? a.ebm1 : a.hbm1) + sbase + nl);
            }
#pragma unroll
            for (int x = 0; x < UNR; x++) {
                const int pl = ps + NPL * (q0 + x), g = pl / KB ? 1 : 0;
                const int kl = pl - g * KB;
                if (pok && pl < 2 * KB && kl < kb) {
                    const int nl = pnd + kl * N2;
                    const int sn = Lay<N>::at(pi, pj, kl);
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
                        const double t = a.ca * kk[x][c] + a.dt * (r[c] * mb[x]);
                        const double un = o[c] + a.cb * t;
                        if (SLAB_ST_CS) {
                            __stcs(a.kf + (cb0 + c) * a.ld + gi, t);
                            __stcs(a.u_out + (cb0 + c) * a.ld + gi, un);
                        } else {
                            a.kf[(cb0 + c) * a.ld + gi] = t;
                            a.u_out[(cb0 + c) * a.ld + gi] = un;
                        }
                        // x-face mirror of the new fields (stage_args.h)
                        if (a.xtr_out != nullptr && (pi == 0 || pi == N - 1))
                            a.xtr_out[(cb0 + c) * a.ldx + (2ll * e + (pi ? 1 : 0)) * N2 + pj +
                                      N * (k0 + kl)] = un;
                    }
                }
            }
        }
    }
}

template <int N, bool PML, bool CM>
int launch_inst(const StageParams<N> &prm, cudaStream_t st)
{
    constexpr int KS = ks_for(N);
    using C = Slab<N, KS>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(slab_kernel<N, KS, PML, CM>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::SMEM) != cudaSuccess)
            return 1;
        configured[dev] = true;
    }
    slab_kernel<N, KS, PML, CM><<<KS * prm.a.nel, C::NT, C::SMEM, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

template <int N>
int launch_n(const StageArgs &a, const double *Dhost, bool pml, bool cm, cudaStream_t st)
{
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    if (pml) return cm ? launch_inst<N, true, true>(prm, st) : launch_inst<N, true, false>(prm, st);
    return cm ? launch_inst<N, false, true>(prm, st) : launch_inst<N, false, false>(prm, st);
}

} // namespace

// returns 0 ok, -1 unsupported order, >0 CUDA failure.  Dhost = dxm1 (n*n, column-major).
// cm: every element of the list has constant metrics (elflag bit 2).
// Compiled four times (Makefile): nx1 = 2..16 and (-DSLAB_HI) nx1 = 17..24, each as is and with
// -fmad=false -DNKB_STRICT (desc.strict: every product and sum rounded separately, as the
// reference's x86-64 build does).
#ifdef NKB_STRICT
#define SLAB_LAUNCH launch_stage_slab_strict
#define SLAB_LAUNCH_HI launch_stage_slab_hi_strict
#else
#define SLAB_LAUNCH launch_stage_slab
#define SLAB_LAUNCH_HI launch_stage_slab_hi
#endif
int SLAB_LAUNCH_HI(const StageArgs &a, const double *Dhost, int nx1, bool pml, bool cm, void *stream);

#ifdef SLAB_HI
int SLAB_LAUNCH_HI(const StageArgs &a, const double *Dhost, int nx1, bool pml, bool cm, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx1) {
    case 17: return launch_n<17>(a, Dhost, pml, cm, st);
    case 18: return launch_n<18>(a, Dhost, pml, cm, st);
    case 19: return launch_n<19>(a, Dhost, pml, cm, st);
    case 20: return launch_n<20>(a, Dhost, pml, cm, st);
    case 21: return launch_n<21>(a, Dhost, pml, cm, st);
    case 22: return launch_n<22>(a, Dhost, pml, cm, st);
    case 23: return launch_n<23>(a, Dhost, pml, cm, st);
    case 24: return launch_n<24>(a, Dhost, pml, cm, st);
    default: return -1;
    }
}
#else
int SLAB_LAUNCH(const StageArgs &a, const double *Dhost, int nx1, bool pml, bool cm, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
#ifdef SLAB_ONLY_N
    if (nx1 == SLAB_ONLY_N) return launch_n<SLAB_ONLY_N>(a, Dhost, pml, cm, st);
    return -1;
#else
    switch (nx1) {
    case 2: return launch_n<2>(a, Dhost, pml, cm, st);
    case 3: return launch_n<3>(a, Dhost, pml, cm, st);
    case 4: return launch_n<4>(a, Dhost, pml, cm, st);
    case 5: return launch_n<5>(a, Dhost, pml, cm, st);
    case 6: return launch_n<6>(a, Dhost, pml, cm, st);
    case 7: return launch_n<7>(a, Dhost, pml, cm, st);
    case 8: return launch_n<8>(a, Dhost, pml, cm, st);
    case 9: return launch_n<9>(a, Dhost, pml, cm, st);
    case 10: return launch_n<10>(a, Dhost, pml, cm, st);
    case 11: return launch_n<11>(a, Dhost, pml, cm, st);
    case 12: return launch_n<12>(a, Dhost, pml, cm, st);
    case 13: return launch_n<13>(a, Dhost, pml, cm, st);
    case 14: return launch_n<14>(a, Dhost, pml, cm, st);
    case 15: return launch_n<15>(a, Dhost, pml, cm, st);
    case 16: return launch_n<16>(a, Dhost, pml, cm, st);
    default: return SLAB_LAUNCH_HI(a, Dhost, nx1, pml, cm, stream);
    }
#endif
}
#endif

} // namespace nkb
