// Fused Maxwell RK-stage kernel for sm_100a, "plane sweep" formulation for the high orders
// (3D, nx1 = 11..16, plain elements, general per-node metrics).
//
// Why another formulation: from nx1 = 11 on a whole element (6 components + 6 residuals) no longer
// fits one SM's shared memory next to the staging regions of stage_pipe.cu, and k-slabs re-read
// the element through L2 for the t-derivative (stage_slab.cu).  Here the element lives in the
// REGISTER FILE instead:
//
//   * work item = (element, curl group g): g = 0 advances E from curl H, g = 1 advances H from
//     -curl E.  The two items of an element are adjacent work items, i.e. they run at the same time
//     on neighbouring CTAs, so what both read (cofactors, fields, face data) is fetched from HBM
//     once and served to the second reader by L2;
//   * thread (i, j, c) holds the k-line of source component c through node (i, j) in registers
//     (2*nx1 registers): the t-derivative of every plane is a register dot product with a row of D;
//   * the element is then swept plane by plane.  Every per-node operand of a plane (the plane of
//     the three source components, nine cofactors, three RK registers, three old fields, one mass:
//     19 rows of nx1^2 doubles) arrives in a shared-memory ring through 1-D bulk asynchronous
//     copies (cp.async.bulk, mbarrier completion), issued NSTG planes ahead across item
//     boundaries, so the sweep never waits for HBM or L2 and the ring is the only large buffer;
//   * per step k: (A) r- and s-pencils of plane k+1 (one thread per line and component, outputs
//     split in four shares, D(i,m) a constant-bank operand) and the register t-derivative of plane
//     k+1 go to a double-buffered derivative plane; (B) thread (i, j, c') combines the six
//     derivatives and six cofactors of output component c' of plane k, adds the lifts, applies the
//     inverse mass and the low-storage RK update and stores the results; one barrier per plane;
//   * the surface flux of the item's three output components is computed at the start of the item
//     (own and neighbour traces gathered from global memory / L2, face data from global memory)
//     into a shared-memory lift table.
//
// Arithmetic: the products and their association are those of stage_pipe.cu
// (((r-part + s-part)*w + w*t-part) + lifts in x, y, z face order); sums run left to right as in
// the reference's mxm.  Reference semantics: SURVEY.md 8a rows a4-a18; citations at the phases in
// stage_slab.cu / stage_pipe.cu.
#include <cstdint>

#include "stage_common.h"

namespace nkb {
namespace {

#ifndef SWEEP_NSTG_MAX
#define SWEEP_NSTG_MAX 6
#endif
#ifndef SWEEP_L2_AHEAD
#define SWEEP_L2_AHEAD 1 // L2 prefetch of the next item's k-lines and face data during the sweep
#endif

__host__ __device__ constexpr int s_round32(int x) { return ((x + 31) / 32) * 32; }
__host__ __device__ constexpr int s_even(int x) { return (x + 1) / 2 * 2; }
__host__ __device__ constexpr int s_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int s_min(int a, int b) { return a < b ? a : b; }

template <int N>
struct SW {
    static constexpr int N2 = N * N, N3 = N2 * N, NF = 6 * N2;
    static constexpr int NO = 2;                     // outputs per pencil share
    static constexpr int H = (N + NO - 1) / NO;      // output shares of a pencil
    static constexpr int RS_BLK = s_round32(2 * N);  // (direction, line) pencils of a share
    static constexpr int RS_ITEMS = H * RS_BLK;
    static constexpr int NT = s_max(s_round32(N2), RS_ITEMS);
    static constexpr int NARR = 16;                  // rows of a ring stage
    static constexpr int PL = s_even(N2 + 2);        // one row (+ alignment slack of the bulk copy)
    static constexpr int STG = NARR * PL;
    static constexpr int RW = N | 1;                 // row stride of a source / derivative plane (odd)
    static constexpr int DPL = s_even(RW * N);
    static constexpr int LF = s_even(NF);
    static constexpr int FPT = (NF + NT - 1) / NT;   // face points per thread
    static constexpr int FIXED = 12 * DPL + 6 * DPL + 3 * LF + 16;
    static constexpr int CAP = (227 * 1024 - 1024) / 8;
#ifdef SWEEP_MINB
    static constexpr int MINB = SWEEP_MINB;
#else
    // CTAs per SM: two when the ring keeps >= 3 stages and the k-lines fit the registers
    static constexpr int MINB = ((CAP / 2 - FIXED) / STG >= 3 && 65536 / (2 * NT) >= 160) ? 2 : 1;
#endif
    static constexpr int NSTG0 = (CAP / MINB - FIXED) / STG;
    static constexpr int NSTG = s_min(s_min(NSTG0, SWEEP_NSTG_MAX), N);
    static_assert(NSTG >= 2, "ring too small");
    static constexpr int OFF_RING = 0, OFF_DD = NSTG * STG, OFF_PP = OFF_DD + 12 * DPL,
                         OFF_L = OFF_PP + 6 * DPL, OFF_BAR = OFF_L + 3 * LF;
    static constexpr size_t SMEM = sizeof(double) * (OFF_BAR + NSTG + 2);
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
// the first double lands at dst + (src & 15)/8
__device__ __forceinline__ uint32_t bulk_bytes(const double *src, int cnt)
{
    const uint32_t head = (uint32_t)((uintptr_t)src & 15);
    return (head + (uint32_t)cnt * 8u + 15u) & ~15u;
}
__device__ __forceinline__ void bulk_load(double *dst, const double *src, int cnt, uint64_t *bar)
{
    const uintptr_t s = (uintptr_t)src;
    const uint32_t head = (uint32_t)(s & 15);
    const uint32_t bytes = (head + (uint32_t)cnt * 8u + 15u) & ~15u;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(s - head), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch(const double *src, int cnt)
{
    const uintptr_t s = (uintptr_t)src;
    const uint32_t head = (uint32_t)(s & 15);
    const uint32_t bytes = (head + (uint32_t)cnt * 8u + 15u) & ~15u;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s - head), "r"(bytes) : "memory");
}

// One thread, outputs O0..O1-1 of one line of the three source components:
//   out_c(o) = sum_m D(o,m) in_c(m)      (mxfK order, left to right)
template <int N, int O0, int O1, int CS>
__device__ __forceinline__ void sweep_pencil(const double (&D)[N * N], const double *in, int sm,
                                             double *out, int so)
{
    constexpr int NO = O1 - O0;
    if constexpr (NO > 0) {
        double acc[3][NO];
#pragma unroll
        for (int m = 0; m < N; m++) {
            const double u0 = in[m * sm], u1 = in[CS + m * sm], u2 = in[2 * CS + m * sm];
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
            out[(O0 + o) * so] = acc[0][o];
            out[CS + (O0 + o) * so] = acc[1][o];
            out[2 * CS + (O0 + o) * so] = acc[2][o];
        }
    }
}

// dispatch on the (warp-uniform) output share h: outputs [2h, 2h+2) of 0..N-1
template <int N, int CS, int HH = 0>
__device__ __forceinline__ void sweep_share(const double (&D)[N * N], const double *in, int sm,
                                            double *out, int so, int h)
{
    if constexpr (2 * HH < N) {
        if (h == HH) sweep_pencil<N, 2 * HH, s_min(2 * HH + 2, N), CS>(D, in, sm, out, so);
        else sweep_share<N, CS, HH + 1>(D, in, sm, out, so, h);
    }
}

template <int N>
__global__ void __launch_bounds__(SW<N>::NT, SW<N>::MINB)
    sweep_kernel(const __grid_constant__ StageParams<N> prm)
{
    using C = SW<N>;
    constexpr int N2 = C::N2, N3 = C::N3, NF = C::NF, NT = C::NT, PL = C::PL, STG = C::STG;
    constexpr int RW = C::RW, DPL = C::DPL, LF = C::LF, NSTG = C::NSTG, FPT = C::FPT;
    const StageArgs &a = prm.a;
    extern __shared__ __align__(16) double smem[];
    double *ring = smem + C::OFF_RING; // [NSTG][16][PL]: 9 cofactors, 3 RK registers, 3 old fields, mass
    double *DD = smem + C::OFF_DD;     // [2][r,s][3 components][DPL] derivatives of a plane
    double *PP = smem + C::OFF_PP;     // [2][3 components][DPL] source components of a plane
    double *L = smem + C::OFF_L;       // [3][LF] lifts of the item's output components
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool w0 = tid < 32;
    const int G = (int)gridDim.x; // even: the group of a CTA's items is fixed
    const int nitems = 2 * a.nel;
    const int g = blockIdx.x & 1;
    const int srcc = g ? 3 : 0, outc = g ? 0 : 3; // source / output components
    const double sg = g ? -1.0 : 1.0;
    const double *mass = g ? a.hbm1 : a.ebm1;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NSTG; b++) mbar_init(full + b, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // node thread (i, j): all three components of the nodes (i, j, .)
    const bool node = tid < N2;
    const int pnd = node ? tid : 0;
    const int pi = pnd % N, pj = pnd / N;
    const int pp = pi + RW * pj; // its place in a source / derivative plane
    // pencil of this thread: share h of the outputs (warp-uniform), direction d (0: r, 1: s), line
    const int ph = tid / C::RS_BLK, pr = tid - ph * C::RS_BLK;
    const bool pen = tid < C::RS_ITEMS && pr < 2 * N;
    const int pd = pr / N, pline = pr - pd * N;
    const int p_in = pd ? pline : pline * RW, p_sm = pd ? RW : 1; // r: along i, s: along j

    // producer: every lane of warp 0 arrives with the bytes of its own copy (one row each)
    auto issue_plane = [&](int e, int k, int stg) {
        const long long off = (long long)e * N3 + (long long)k * N2;
        const double *src = nullptr;
        if (lane < 9) src = a.met[lane];
        else if (lane < 12) src = a.kf + (long long)(outc + lane - 9) * a.ld;
        else if (lane < 15) src = a.u_in + (long long)(outc + lane - 12) * a.ld;
        else if (lane == 15) src = mass;
        uint32_t bytes = 0;
        if (src != nullptr) bytes = bulk_bytes(src + off, N2);
        mbar_arrive_tx(full + stg, bytes);
        if (src != nullptr) bulk_load(ring + stg * STG + lane * PL, src + off, N2, full + stg);
    };
    // face point f of this thread: slot (reference order -y,+x,+y,-x,-z,+z, cemface
    // cem_common.F:234-260), point within the face, own volume node
    auto face_point = [&](int f, int &fj, int &nd) {
        const int q = tid + f * NT;
        fj = q < NF ? q : -1;
        const int qq = q < NF ? q : 0;
        const int slot = qq / N2, p = qq - slot * N2;
        const int pa = p % N, pb = p / N;
        int ci, cj, ck;
        if (slot == 0) { ci = pa; cj = 0; ck = pb; }
        else if (slot == 2) { ci = pa; cj = N - 1; ck = pb; }
        else if (slot == 1) { ci = N - 1; cj = pa; ck = pb; }
        else if (slot == 3) { ci = 0; cj = pa; ck = pb; }
        else if (slot == 4) { ci = pa; cj = pb; ck = 0; }
        else { ci = pa; cj = pb; ck = N - 1; }
        nd = ci + N * cj + N2 * ck;
    };

    int q = blockIdx.x;
    int e = q < nitems ? ldg(a.elist + (q >> 1)) : 0;
    int en = q + G < nitems ? ldg(a.elist + ((q + G) >> 1)) : 0;
    int fvp[FPT];
#pragma unroll
    for (int f = 0; f < FPT; f++) {
        int fj, nd;
        face_point(f, fj, nd);
        fvp[f] = (fj < 0 || q >= nitems) ? -2 : ldg(a.vmapP + (long long)e * NF + fj);
    }

    // ring bookkeeping: planes are numbered through the items of this CTA; plane t lives in stage
    // t % NSTG and completes phase (t / NSTG) & 1 of that stage's barrier
    int ik = 0;             // next plane to issue (>= N: of the next item)
    int cstg = 0;           // stage of the plane being consumed
    uint32_t parbits = 0;   // per stage: parity of the phase its next consumer waits for
    if (q < nitems && w0) {
        for (int x = 0; x < NSTG; x++) issue_plane(e, x, x);
    }
    if (q < nitems) ik = NSTG; // NSTG <= N

#pragma unroll 1
    for (; q < nitems; q += G) {
        const long long ebase = (long long)e * N3;
        const bool more = q + G < nitems;
        const int enn = q + 2 * G < nitems ? ldg(a.elist + ((q + 2 * G) >> 1)) : 0;

        // ---- k-lines of the three source components through node (i, j) -> registers --------------
        double col[3][N];
        {
            const double *gp = a.u_in + (long long)srcc * a.ld + ebase + pnd;
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int m = 0; m < N; m++) col[c][m] = node ? ldg(gp + c * a.ld + N2 * m) : 0.0;
        }

        // ---- surface flux of the item's output components -> L (flux3d, src/cem_maxwell.F:922-1002)
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            int fj, nd;
            face_point(f, fj, nd);
            if (fj >= 0) {
                const long long jf = (long long)e * NF + fj;
                const int vp = fvp[f];
                long long st;
                const double *nb = nbr_trace(a, vp, ebase, st);
                double pv[6], ov[6];
#pragma unroll
                for (int c = 0; c < 6; c++) pv[c] = ldg(nb + c * st);
#pragma unroll
                for (int c = 0; c < 6; c++) ov[c] = ldg(a.u_in + (long long)c * a.ld + ebase + nd);
                const double unx = ldg(a.unx + jf), uny = ldg(a.uny + jf), unz = ldg(a.unz + jf);
                const double ar = ldg(a.area + jf);
                const double h0 = ldg((g ? a.hY : a.hZ) + jf), i1 = ldg((g ? a.Y1 : a.Z1) + jf);
                double Hx = ov[0], Hy = ov[1], Hz = ov[2], Ex = ov[3], Ey = ov[4], Ez = ov[5];
                double pHx = pv[0], pHy = pv[1], pHz = pv[2], pEx = pv[3], pEy = pv[4], pEz = pv[5];
                if (a.inc_own != nullptr) { // userinc hook (src/cem_maxwell.F:498)
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
                if (vp >= 0 || vp <= -3) {
                    // neighbour's (-n+ x E+) with n+ = -n-  (the gs_op_fields sum of :962)
                    s0 = s0 - (-uny * pEz + unz * pEy);
                    s1 = s1 - (-unz * pEx + unx * pEz);
                    s2 = s2 - (-unx * pEy + uny * pEx);
                    s3 = s3 - (-uny * pHz + unz * pHy);
                    s4 = s4 - (-unz * pHx + unx * pHz);
                    s5 = s5 - (-unx * pHy + uny * pHx);
                } else if (vp == -1) { // 'PEC' / 'PML' outer face: cem_maxwell_flux_pec :1397-1405
                    s0 = 2.0 * s0; s1 = 2.0 * s1; s2 = 2.0 * s2;
                    s3 = 0.0; s4 = 0.0; s5 = 0.0;
                }
                if (g) { // flux into resH (:976-986)
                    const double hY = h0, Y1 = i1;
                    const double Y02 = -(hY * Y1), C02Y = hY * a.C0;
                    const double fu1 = uny * s5 - unz * s4;
                    const double fu2 = unz * s3 - unx * s5;
                    const double fu3 = unx * s4 - uny * s3;
                    L[fj] = ar * (Y02 * s0 - C02Y * fu1);
                    L[LF + fj] = ar * (Y02 * s1 - C02Y * fu2);
                    L[2 * LF + fj] = ar * (Y02 * s2 - C02Y * fu3);
                } else { // flux into resE (:987-997)
                    const double hZ = h0, Z1 = i1;
                    const double Z02 = hZ * Z1, C02Z = hZ * a.C0;
                    const double fw1 = uny * s2 - unz * s1;
                    const double fw2 = unz * s0 - unx * s2;
                    const double fw3 = unx * s1 - uny * s0;
                    L[fj] = ar * (Z02 * s3 - C02Z * fw1);
                    L[LF + fj] = ar * (Z02 * s4 - C02Z * fw2);
                    L[2 * LF + fj] = ar * (Z02 * s5 - C02Z * fw3);
                }
            }
        }
        // neighbour ids of the next item's face points (consumed at its start)
        int fvn[FPT];
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            int fj, nd;
            face_point(f, fj, nd);
            fvn[f] = (fj < 0 || !more) ? -2 : ldg(a.vmapP + (long long)en * NF + fj);
        }

        // source planes 0 and 1 -> PP (from the registers)
        if (node) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                PP[c * DPL + pp] = col[c][0];
                PP[(3 + c) * DPL + pp] = col[c][1];
            }
        }
        __syncthreads();
        // r/s pencils of plane kk: PP[kk & 1] -> DD[kk & 1]
        auto pencils = [&](int kk) {
            if (pen)
                sweep_share<N, DPL>(prm.D, PP + (kk & 1) * 3 * DPL + p_in, p_sm,
                                    DD + (kk & 1) * 6 * DPL + pd * 3 * DPL + p_in, p_sm, ph);
        };
        pencils(0);
        __syncthreads(); // L and the derivatives of plane 0 are complete

#pragma unroll 1
        for (int k = 0; k < N; k++) {
            const int nstg = cstg + 1 == NSTG ? 0 : cstg + 1;
            if (k + 1 < N) pencils(k + 1);

            // ---- plane k: weighted curl, lifts, inverse mass, low-storage RK update ---------------
            if (node) {
                const long long pbase = ebase + (long long)k * N2;
                const long long gi = pbase + pnd;
                // source components of plane k+2 (for its pencils, next step but one)
                double pn[3] = {0.0, 0.0, 0.0};
                if (k + 2 < N) {
#pragma unroll
                    for (int c = 0; c < 3; c++) pn[c] = ldg(a.u_in + (long long)(srcc + c) * a.ld + gi + 2 * N2);
                }
                const double wv = sg * ldg(a.w3 + pnd + k * N2);
                // t-derivatives from the registers: row k of D, left to right
                double dt[3];
                {
                    const double d0 = prm.D[k];
                    dt[0] = d0 * col[0][0]; dt[1] = d0 * col[1][0]; dt[2] = d0 * col[2][0];
#pragma unroll
                    for (int m = 1; m < N; m++) {
                        const double dm = prm.D[k + N * m];
                        dt[0] = dt[0] + dm * col[0][m];
                        dt[1] = dt[1] + dm * col[1][m];
                        dt[2] = dt[2] + dm * col[2][m];
                    }
                }
                const double *dd = DD + (k & 1) * 6 * DPL + pp;
                const double dr[3] = {dd[0], dd[DPL], dd[2 * DPL]};
                const double ds[3] = {dd[3 * DPL], dd[4 * DPL], dd[5 * DPL]};
                mbar_wait(full + cstg, (parbits >> cstg) & 1u);
                const double *S = ring + cstg * STG + (int)(pbase & 1) + pnd;
                double cr[3], cs[3], ct[3], r[3];
                curl_part(dr, S[0], S[PL], S[2 * PL], cr);
                curl_part(ds, S[3 * PL], S[4 * PL], S[5 * PL], cs);
                curl_part(dt, S[6 * PL], S[7 * PL], S[8 * PL], ct);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    r[c] = (cr[c] + cs[c]) * wv;
                    r[c] = r[c] + wv * ct[c];
                }
                // lifts: x-, y-, z-faces in turn
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double *Lc = L + c * LF;
                    if (pi == 0) r[c] += Lc[3 * N2 + pj + N * k];
                    if (pi == N - 1) r[c] += Lc[1 * N2 + pj + N * k];
                    if (pj == 0) r[c] += Lc[0 * N2 + pi + N * k];
                    if (pj == N - 1) r[c] += Lc[2 * N2 + pi + N * k];
                    if (k == 0) r[c] += Lc[4 * N2 + pnd];
                    if (k == N - 1) r[c] += Lc[5 * N2 + pnd];
                }
                if (a.src_prof != nullptr) { // usersrc hook: res(comp) -= profile*(tfac*bm)
                    const int cs2 = a.src_comp - outc;
                    if (cs2 >= 0 && cs2 < 3) {
                        const double sv2 = ldg(a.src_prof + gi) * (a.src_tfac * ldg(a.bmn + gi));
                        if (cs2 == 0) r[0] -= sv2;
                        else if (cs2 == 1) r[1] -= sv2;
                        else r[2] -= sv2;
                    }
                }
                const double mb = S[15 * PL];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double t = a.ca * S[(9 + c) * PL] + a.dt * (r[c] * mb);
                    const double un = S[(12 + c) * PL] + a.cb * t;
                    __stcs(a.kf + (long long)(outc + c) * a.ld + gi, t);
                    __stcs(a.u_out + (long long)(outc + c) * a.ld + gi, un);
                    if (a.xtr_out != nullptr && (pi == 0 || pi == N - 1))
                        a.xtr_out[(long long)(outc + c) * a.ldx + (2ll * e + (pi ? 1 : 0)) * N2 + pj + N * k] = un;
                }
                if (k + 2 < N) {
#pragma unroll
                    for (int c = 0; c < 3; c++) PP[(k & 1) * 3 * DPL + c * DPL + pp] = pn[c];
                }
            }
            __syncthreads();
            // the stage of plane k is free: it takes the plane NSTG ahead (of this or the next item)
            if (w0) {
                if (ik < N) issue_plane(e, ik, cstg);
                else if (more) issue_plane(en, ik - N, cstg);
#if SWEEP_L2_AHEAD
                if (more && k == N / 2) {
                    // next item: k-lines of the source components and the face data towards L2
                    if (lane < 3) bulk_prefetch(a.u_in + (long long)(srcc + lane) * a.ld + (long long)en * N3, N3);
                    else if (lane < 9) {
                        const double *fa = lane == 3 ? a.unx : lane == 4 ? a.uny : lane == 5 ? a.unz
                                         : lane == 6 ? a.area : lane == 7 ? (g ? a.hY : a.hZ)
                                         : (g ? a.Y1 : a.Z1);
                        bulk_prefetch(fa + (long long)en * NF, NF);
                    }
                }
#endif
            }
            ik++;
            parbits ^= 1u << cstg;
            cstg = nstg;
        }
        ik -= N;
        e = en;
        en = enn;
#pragma unroll
        for (int f = 0; f < FPT; f++) fvp[f] = fvn[f];
    }
}

template <int N>
int sweep_launch_n(const StageArgs &a, const double *Dhost, cudaStream_t st)
{
    using C = SW<N>;
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    static int grid_per_dev[64] = {};
    if (dev < 0 || dev >= 64) return 1;
    if (grid_per_dev[dev] == 0) {
        if (cudaFuncSetAttribute(sweep_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::SMEM) != cudaSuccess)
            return 1;
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel<N>, C::NT, C::SMEM) !=
                cudaSuccess || occ < 1)
            return 1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
        grid_per_dev[dev] = (occ * sms) & ~1;
    }
    const int items = 2 * a.nel;
    int grid = items < grid_per_dev[dev] ? items : grid_per_dev[dev];
    if (a.grid_cap > 0 && grid > a.grid_cap) grid = a.grid_cap;
    grid &= ~1; // even: the two groups of an element on adjacent CTAs, a CTA keeps its group
    if (grid < 2) grid = 2;
    sweep_kernel<N><<<grid, C::NT, C::SMEM, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

bool sw_aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

} // namespace

// returns 0 ok, -1 not covered by this kernel (order, auxiliary or constant-metric list,
// alignment: the caller uses the other stage kernels), >0 CUDA failure.
// Dhost = dxm1 (n*n, column-major).
int launch_stage_sweep(const StageArgs &a, const double *Dhost, int nx1, bool aux, bool cm,
                       void *stream)
{
    if (aux || cm || nx1 < 11 || nx1 > 16) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = sw_aligned16(a.u_in) && sw_aligned16(a.kf) && (a.ld % 2 == 0) &&
              sw_aligned16(a.hbm1) && sw_aligned16(a.ebm1);
    for (int q = 0; q < 9; q++) ok = ok && sw_aligned16(a.met[q]);
    if (!ok) return -1;
    switch (nx1) {
    case 11: return sweep_launch_n<11>(a, Dhost, st);
    case 12: return sweep_launch_n<12>(a, Dhost, st);
    case 13: return sweep_launch_n<13>(a, Dhost, st);
    case 14: return sweep_launch_n<14>(a, Dhost, st);
    case 15: return sweep_launch_n<15>(a, Dhost, st);
    case 16: return sweep_launch_n<16>(a, Dhost, st);
    default: return -1;
    }
}

} // namespace nkb
