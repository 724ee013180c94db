// Fused Maxwell RK-stage kernel for sm_100a: "pencil" formulation, 2 <= nx1 <= 16.
//
// One launch = one RK stage over a list of elements.  One CTA = one HALF-TASK (element e,
// group g):  g = 0 updates E from curl(H), g = 1 updates H from -curl(E).  The two half-tasks
// of an element are independent given the stage-start fields (ping-pong buffer u_in), sit in
// adjacent CTAs (shared reads hit L2) and together make ONE pass over HBM per stage
// (SURVEY.md 8a rows a4-a18):
//
//   P0  stage the 3 source components of the element (n^3 nodes each) in shared memory
//   P1  r-pencils: thread (j,k) holds the n-point line of each source component in registers,
//       applies D (kernel-parameter constant bank, no loads) and stores the r-part of the curl
//       [maxwell_wght_curl, src/cem_maxwell.F:1428-1497; local_grad3, src/nek5_grad.F:2-19;
//        mxfK left-to-right sums, src/nek5_mxm_std.F:173-190]
//   P2  s-pencils: thread (i,k), adds the s-part, multiplies by the quadrature weight w3mn
//   P3  surface flux in three rounds (+-x, +-y, +-z faces; no node is touched twice in a
//       round): own trace from smem/global, neighbour trace through vmapP (the gs_op_fields
//       pair-sum of src/cem_maxwell.F:962) or the NCCL halo, PEC mirror, upwind/central flux,
//       times face area, lifted into the smem residual
//       [restrict_to_face :604-652, flux3d :922-1002, flux_pec :1368-1426, add_flux_to_res
//        :725-735]
//   P4  t-pencils: thread (i,j), adds the t-part and finishes every node of its line:
//       PML ADEs [pml_step, src/cem_maxwell_pml.F:508-592], volume source [usersrc hook :503],
//       inverse mass [invqmass :1878-1886] and the low-storage RK update [rk4_upd,
//       src/cem_common.F:18-76], written to the ping-pong field buffer.
//
// Shared-memory traffic is ~33 accesses per node per half-task (vs 21n for a naive
// per-node dot product), which is what keeps n = 16 off the shared-memory roofline.
// Arithmetic: same products as the reference; the 6-term curl sum is associated by
// direction ((r-part + s-part)*w + lift) + w*t-part, and nvcc contracts a*b+c into FMA --
// both are <= 1e-15 relative effects per operation (DESIGN.md "Numerics").
#include <cuda_runtime.h>

#include "stage_args.h"

namespace nkb {

// read-only data (everything except the RK registers and the PML auxiliaries) goes through
// ld.global.nc so that the compiler may hoist the loads above earlier stores
__device__ __forceinline__ double ldg(const double *p) { return __ldg(p); }
__device__ __forceinline__ int ldg(const int *p) { return __ldg(p); }

// ---- compile-time geometry of one half-task --------------------------------------------
__host__ __device__ constexpr int pad_j(int n)
{
    // row padding that makes the r- and s-pencil accesses bank-conflict free (64-bit banks):
    // found by exhaustive search (scripts/smem_banks.py)
    return (n == 6 || n == 14) ? 3 : ((n == 8 || n == 12 || n == 16) ? 1 : 0);
}
__host__ __device__ constexpr int pad_k(int n)
{
    return (n == 3 || n == 4 || n == 7) ? 3 : (n == 10 ? 7 : 0);
}
__host__ __device__ constexpr int split_for(int n) { return n <= 5 ? 4 : 2; }
__host__ __device__ constexpr int threads_for(int n)
{
    return ((n * n * split_for(n) + 31) / 32) * 32;
}
__host__ __device__ constexpr int min_blocks_for(int n)
{
    // occupancy target used for the register cap: limited by smem (227 KB) and 2048 threads
    int sj = n + pad_j(n), sc = (sj * n + pad_k(n)) * n;
    int by_smem = (227 * 1024) / (6 * sc * 8 + 1024);
    int by_thr = 2048 / threads_for(n);
    int b = by_smem < by_thr ? by_smem : by_thr;
    int by_reg = 65536 / (threads_for(n) * (6 * n + 40)); // 3n doubles of pencil + working set
    if (b > by_reg) b = by_reg;
    if (b > 16) b = 16;
    return b < 1 ? 1 : b;
}

template <int N>
struct StageParams {
    StageArgs a;
    double D[N * N]; // dxm1, column-major: D(i,m) at i + N*m  (constant bank operand)
};

// d[c] = sum_m D(OUT,m) * u[c][m], left to right (mxfK order)
template <int N, int OUT>
__device__ __forceinline__ void deriv3(const double (&D)[N * N], const double (&u)[3][N],
                                       double (&d)[3])
{
#pragma unroll
    for (int c = 0; c < 3; c++) d[c] = D[OUT] * u[c][0];
#pragma unroll
    for (int m = 1; m < N; m++) {
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = d[c] + D[OUT + N * m] * u[c][m];
    }
}

// curl contribution of one direction: (u1d,u2d,u3d) derivatives with cofactors (mx,my,mz)
__device__ __forceinline__ void curl_part(const double (&d)[3], double mx, double my, double mz,
                                          double (&c)[3])
{
    c[0] = d[2] * my - d[1] * mz;
    c[1] = d[0] * mz - d[2] * mx;
    c[2] = d[1] * mx - d[0] * my;
}

// ---- P1: r-pencil outputs i in [I0, I1) ----------------------------------------------------
template <int N, int I0, int I1>
__device__ __forceinline__ void r_outputs(const double (&D)[N * N], const double (&u)[3][N],
                                          const StageArgs &a, long long grow, double *Rrow,
                                          int SC)
{
    if constexpr (I0 < I1) {
        double d[3], c[3];
        deriv3<N, I0>(D, u, d);
        curl_part(d, ldg(a.rx + grow + I0), ldg(a.ry + grow + I0), ldg(a.rz + grow + I0), c);
        Rrow[I0] = c[0];
        Rrow[SC + I0] = c[1];
        Rrow[2 * SC + I0] = c[2];
        r_outputs<N, I0 + 1, I1>(D, u, a, grow, Rrow, SC);
    }
}

// ---- P2: s-pencil outputs j in [J0, J1) ------------------------------------------------------
template <int N, int J0, int J1, int SJ>
__device__ __forceinline__ void s_outputs(const double (&D)[N * N], const double (&u)[3][N],
                                          const StageArgs &a, long long gcol, int ncol, double sg,
                                          double *Rcol, int SC)
{
    if constexpr (J0 < J1) {
        double d[3], c[3];
        deriv3<N, J0>(D, u, d);
        curl_part(d, ldg(a.sx + gcol + N * J0), ldg(a.sy + gcol + N * J0), ldg(a.sz + gcol + N * J0), c);
        const double wv = sg * ldg(a.w3 + ncol + N * J0);
        Rcol[SJ * J0] = (Rcol[SJ * J0] + c[0]) * wv;
        Rcol[SC + SJ * J0] = (Rcol[SC + SJ * J0] + c[1]) * wv;
        Rcol[2 * SC + SJ * J0] = (Rcol[2 * SC + SJ * J0] + c[2]) * wv;
        s_outputs<N, J0 + 1, J1, SJ>(D, u, a, gcol, ncol, sg, Rcol, SC);
    }
}

// ---- P4: t-pencil outputs k in [K0, K1): finishes the node ------------------------------------
template <int N, int K0, int K1, int SK, bool PML>
__device__ __forceinline__ void t_outputs(const double (&D)[N * N], const double (&u)[3][N],
                                          const StageArgs &a, long long gnode, int node, int g,
                                          double sg, const double *Rt, int SC)
{
    if constexpr (K0 < K1) {
        constexpr int N2 = N * N;
        const long long gi = gnode + (long long)N2 * K0;
        double d[3], c[3], r[3];
        deriv3<N, K0>(D, u, d);
        curl_part(d, ldg(a.tx + gi), ldg(a.ty + gi), ldg(a.tz + gi), c);
        const double wv = sg * ldg(a.w3 + node + N2 * K0);
        r[0] = Rt[SK * K0] + wv * c[0];
        r[1] = Rt[SC + SK * K0] + wv * c[1];
        r[2] = Rt[2 * SC + SK * K0] + wv * c[2];
        const long long cold = (g == 0 ? 3 : 0) * a.ld; // components being updated
        const double o0 = ldg(a.u_in + cold + gi), o1 = ldg(a.u_in + cold + a.ld + gi),
                     o2 = ldg(a.u_in + cold + 2 * a.ld + gi);
        if (PML) { // pml_step, src/cem_maxwell_pml.F:540-585, then the PML half of rk_maxwell_ab
            const double bm1 = ldg(a.bmn + gi);
            const double bm1inv = 1.0 / bm1;
            const double sigx = a.sig[gi], sigy = a.sig[a.npts + gi], sigz = a.sig[2 * a.npts + gi];
            const double permitt = a.eps[gi];
            const double sxp = sigx / permitt, syp = sigy / permitt, szp = sigz / permitt;
            double *pF = g == 0 ? a.pD : a.pB;
            double *kF = g == 0 ? a.kD : a.kB;
            const double b0 = pF[gi], b1 = pF[a.npts + gi], b2 = pF[2 * a.npts + gi];
            const double rb0 = r[0] * bm1inv - syp * b0;
            const double rb1 = r[1] * bm1inv - szp * b1;
            const double rb2 = r[2] * bm1inv - sxp * b2;
            double p0, p1, p2;
            if (g == 0) {
                p0 = -syp * b0 + sxp * b0 - sigz * o0;
                p1 = -szp * b1 + syp * b1 - sigx * o1;
                p2 = -sxp * b2 + szp * b2 - sigy * o2;
            } else {
                const double permeab = a.mu[gi];
                p0 = -syp * b0 + sxp * b0 - szp * permeab * o0;
                p1 = -szp * b1 + syp * b1 - sxp * permeab * o1;
                p2 = -sxp * b2 + szp * b2 - syp * permeab * o2;
            }
            r[0] = r[0] + p0 * bm1; r[1] = r[1] + p1 * bm1; r[2] = r[2] + p2 * bm1;
            double kk;
            kk = a.ca * kF[gi] + a.dt * rb0; kF[gi] = kk; pF[gi] = b0 + a.cb * kk;
            kk = a.ca * kF[a.npts + gi] + a.dt * rb1; kF[a.npts + gi] = kk;
            pF[a.npts + gi] = b1 + a.cb * kk;
            kk = a.ca * kF[2 * a.npts + gi] + a.dt * rb2; kF[2 * a.npts + gi] = kk;
            pF[2 * a.npts + gi] = b2 + a.cb * kk;
        }
        if (a.src_prof != nullptr) { // usersrc hook: res(comp) -= profile*(tfac*bm)
            const int cs = a.src_comp - (g == 0 ? 3 : 0);
            if (cs >= 0 && cs < 3) {
                const double sv = ldg(a.src_prof + gi) * (a.src_tfac * ldg(a.bmn + gi));
                if (cs == 0) r[0] -= sv;
                else if (cs == 1) r[1] -= sv;
                else r[2] -= sv;
            }
        }
        const double mb = ldg((g == 0 ? a.ebm1 : a.hbm1) + gi);
        r[0] *= mb; r[1] *= mb; r[2] *= mb;
        double kk;
        kk = a.ca * a.kf[cold + gi] + a.dt * r[0]; a.kf[cold + gi] = kk;
        a.u_out[cold + gi] = o0 + a.cb * kk;
        kk = a.ca * a.kf[cold + a.ld + gi] + a.dt * r[1]; a.kf[cold + a.ld + gi] = kk;
        a.u_out[cold + a.ld + gi] = o1 + a.cb * kk;
        kk = a.ca * a.kf[cold + 2 * a.ld + gi] + a.dt * r[2]; a.kf[cold + 2 * a.ld + gi] = kk;
        a.u_out[cold + 2 * a.ld + gi] = o2 + a.cb * kk;
        t_outputs<N, K0 + 1, K1, SK, PML>(D, u, a, gnode, node, g, sg, Rt, SC);
    }
}

template <int N, bool PML>
__global__ void __launch_bounds__(threads_for(N), min_blocks_for(N))
    stage_kernel(const __grid_constant__ StageParams<N> prm)
{
    constexpr int N2 = N * N, N3 = N2 * N, NF = 6 * N2;
    constexpr int SPLIT = split_for(N), NT = threads_for(N);
    constexpr int SJ = N + pad_j(N), SK = SJ * N + pad_k(N), SC = SK * N;
    constexpr int HN = (N + SPLIT - 1) / SPLIT; // outputs per thread of a pencil
    const StageArgs &a = prm.a;
    extern __shared__ double smem[];
    double *U = smem;          // [3][SC] source components at stage start
    double *R = smem + 3 * SC; // [3][SC] residual of the updated components

    const int tid = threadIdx.x;
    const int e = a.elist[blockIdx.x >> 1];
    const int g = blockIdx.x & 1; // 0: E <- curl H ; 1: H <- -curl E
    const long long ebase = (long long)e * N3;
    const double *__restrict__ src = a.u_in + (g == 0 ? 0 : 3) * a.ld;
    const double *__restrict__ oth = a.u_in + (g == 0 ? 3 : 0) * a.ld;
    const double sg = g == 0 ? 1.0 : -1.0;

    // ---- P0: stage the source components ---------------------------------------------------
#pragma unroll 8
    for (int q = tid; q < 3 * N3; q += NT) {
        const int c = q / N3, r = q - c * N3;
        const int i = r % N, j = (r / N) % N, k = r / N2;
        U[c * SC + i + SJ * j + SK * k] = ldg(src + c * a.ld + ebase + r);
    }
    __syncthreads();

    const int p = tid % N2, h = tid / N2;
    const int pa = p % N, pb = p / N;

    // ---- P1: r-pencils, thread (j,k) = (pa,pb) ------------------------------------------------
    if (h < SPLIT) {
        double u[3][N];
        const double *Urow = U + SJ * pa + SK * pb;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int m = 0; m < N; m++) u[c][m] = Urow[c * SC + m];
        const long long grow = ebase + N * pa + N2 * pb;
        double *Rrow = R + SJ * pa + SK * pb;
        if (h == 0) r_outputs<N, 0, (HN < N ? HN : N)>(prm.D, u, a, grow, Rrow, SC);
        if (SPLIT > 1 && h == 1)
            r_outputs<N, HN, (2 * HN < N ? 2 * HN : N)>(prm.D, u, a, grow, Rrow, SC);
        if (SPLIT > 2 && h == 2)
            r_outputs<N, 2 * HN, (3 * HN < N ? 3 * HN : N)>(prm.D, u, a, grow, Rrow, SC);
        if (SPLIT > 3 && h == 3) r_outputs<N, 3 * HN, N>(prm.D, u, a, grow, Rrow, SC);
    }
    __syncthreads();

    // ---- P2: s-pencils, thread (i,k) = (pa,pb) ------------------------------------------------
    if (h < SPLIT) {
        double u[3][N];
        const double *Ucol = U + pa + SK * pb;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int m = 0; m < N; m++) u[c][m] = Ucol[c * SC + SJ * m];
        const long long gcol = ebase + pa + N2 * pb;
        const int ncol = pa + N2 * pb;
        double *Rcol = R + pa + SK * pb;
        if (h == 0) s_outputs<N, 0, (HN < N ? HN : N), SJ>(prm.D, u, a, gcol, ncol, sg, Rcol, SC);
        if (SPLIT > 1 && h == 1)
            s_outputs<N, HN, (2 * HN < N ? 2 * HN : N), SJ>(prm.D, u, a, gcol, ncol, sg, Rcol, SC);
        if (SPLIT > 2 && h == 2)
            s_outputs<N, 2 * HN, (3 * HN < N ? 3 * HN : N), SJ>(prm.D, u, a, gcol, ncol, sg, Rcol,
                                                              SC);
        if (SPLIT > 3 && h == 3) s_outputs<N, 3 * HN, N, SJ>(prm.D, u, a, gcol, ncol, sg, Rcol, SC);
    }
    __syncthreads();

    // ---- P3: surface flux, three rounds of two opposite faces ---------------------------------
    // slot order of the reference (cemface, cem_common.F:234-260): -y,+x,+y,-x,-z,+z
#pragma unroll 1
    for (int rd = 0; rd < 3; rd++) {
        for (int q = tid; q < 2 * N2; q += NT) {
            const int hi = q / N2, fp0 = q - hi * N2;
            const int fa = fp0 % N, fb = fp0 / N;
            int s, node, sn;
            if (rd == 0) { // +-x
                s = hi ? 1 : 3;
                node = (hi ? N - 1 : 0) + N * fa + N2 * fb;
                sn = (hi ? N - 1 : 0) + SJ * fa + SK * fb;
            } else if (rd == 1) { // +-y
                s = hi ? 2 : 0;
                node = fa + N * (hi ? N - 1 : 0) + N2 * fb;
                sn = fa + SJ * (hi ? N - 1 : 0) + SK * fb;
            } else { // +-z
                s = hi ? 5 : 4;
                node = fa + N * fb + N2 * (hi ? N - 1 : 0);
                sn = fa + SJ * fb + SK * (hi ? N - 1 : 0);
            }
            const long long jf = (long long)e * NF + s * N2 + fp0;
            const double unx = ldg(a.unx + jf), uny = ldg(a.uny + jf), unz = ldg(a.unz + jf);
            const int vp = ldg(a.vmapP + jf);
            const double S0 = U[sn], S1 = U[SC + sn], S2 = U[2 * SC + sn];
            const long long gn = ebase + node;
            const double O0 = ldg(oth + gn), O1 = ldg(oth + a.ld + gn), O2 = ldg(oth + 2 * a.ld + gn);
            // own (H,E)
            const double Hx = g == 0 ? S0 : O0, Hy = g == 0 ? S1 : O1, Hz = g == 0 ? S2 : O2;
            const double Ex = g == 0 ? O0 : S0, Ey = g == 0 ? O1 : S1, Ez = g == 0 ? O2 : S2;
            // -n x E, -n x H of the own side (flux3d :946-955)
            double s0 = -uny * Ez + unz * Ey;
            double s1 = -unz * Ex + unx * Ez;
            double s2 = -unx * Ey + uny * Ex;
            double s3 = -uny * Hz + unz * Hy;
            double s4 = -unz * Hx + unx * Hz;
            double s5 = -unx * Hy + uny * Hx;
            if (vp >= 0 || vp <= -3) {
                double pHx, pHy, pHz, pEx, pEy, pEz;
                if (vp >= 0) {
                    pHx = ldg(a.u_in + vp);
                    pHy = ldg(a.u_in + a.ld + vp);
                    pHz = ldg(a.u_in + 2 * a.ld + vp);
                    pEx = ldg(a.u_in + 3 * a.ld + vp);
                    pEy = ldg(a.u_in + 4 * a.ld + vp);
                    pEz = ldg(a.u_in + 5 * a.ld + vp);
                } else {
                    const double *hp = a.halo + 6ll * (long long)(-(vp + 3));
                    pHx = hp[0]; pHy = hp[1]; pHz = hp[2];
                    pEx = hp[3]; pEy = hp[4]; pEz = hp[5];
                }
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
            const double ar = ldg(a.area + jf);
            double f0, f1, f2;
            if (g == 1) { // flux into resH (:976-986)
                const double hY = ldg(a.hY + jf), Y1 = ldg(a.Y1 + jf);
                const double Y02 = -(hY * Y1), C02Y = hY * a.C0;
                const double fu1 = uny * s5 - unz * s4;
                const double fu2 = unz * s3 - unx * s5;
                const double fu3 = unx * s4 - uny * s3;
                f0 = ar * (Y02 * s0 - C02Y * fu1);
                f1 = ar * (Y02 * s1 - C02Y * fu2);
                f2 = ar * (Y02 * s2 - C02Y * fu3);
            } else { // flux into resE (:987-997)
                const double hZ = ldg(a.hZ + jf), Z1 = ldg(a.Z1 + jf);
                const double Z02 = hZ * Z1, C02Z = hZ * a.C0;
                const double fw1 = uny * s2 - unz * s1;
                const double fw2 = unz * s0 - unx * s2;
                const double fw3 = unx * s1 - uny * s0;
                f0 = ar * (Z02 * s3 - C02Z * fw1);
                f1 = ar * (Z02 * s4 - C02Z * fw2);
                f2 = ar * (Z02 * s5 - C02Z * fw3);
            }
            R[sn] += f0;
            R[SC + sn] += f1;
            R[2 * SC + sn] += f2;
        }
        __syncthreads();
    }

    // ---- P4: t-pencils, thread (i,j) = (pa,pb): finish the nodes ------------------------------
    if (h < SPLIT) {
        double u[3][N];
        const double *Ut = U + pa + SJ * pb;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int m = 0; m < N; m++) u[c][m] = Ut[c * SC + SK * m];
        const int node = pa + N * pb;
        const long long gnode = ebase + node;
        const double *Rt = R + pa + SJ * pb;
        if (h == 0)
            t_outputs<N, 0, (HN < N ? HN : N), SK, PML>(prm.D, u, a, gnode, node, g, sg, Rt, SC);
        if (SPLIT > 1 && h == 1)
            t_outputs<N, HN, (2 * HN < N ? 2 * HN : N), SK, PML>(prm.D, u, a, gnode, node, g, sg,
                                                               Rt, SC);
        if (SPLIT > 2 && h == 2)
            t_outputs<N, 2 * HN, (3 * HN < N ? 3 * HN : N), SK, PML>(prm.D, u, a, gnode, node, g,
                                                                   sg, Rt, SC);
        if (SPLIT > 3 && h == 3)
            t_outputs<N, 3 * HN, N, SK, PML>(prm.D, u, a, gnode, node, g, sg, Rt, SC);
    }
}

template <int N>
static int launch_n(const StageArgs &a, const double *Dhost, bool pml, cudaStream_t st)
{
    constexpr int SJ = N + pad_j(N), SK = SJ * N + pad_k(N), SC = SK * N;
    constexpr size_t smem = sizeof(double) * 6 * SC;
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(stage_kernel<N, false>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(stage_kernel<N, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return 1;
        configured = true;
    }
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    if (pml)
        stage_kernel<N, true><<<2 * a.nel, threads_for(N), smem, st>>>(prm);
    else
        stage_kernel<N, false><<<2 * a.nel, threads_for(N), smem, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// returns 0 ok, -1 unsupported order, >0 CUDA failure.  Dhost = dxm1 (n*n, column-major).
int launch_stage(const StageArgs &a, const double *Dhost, int nx1, bool pml, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx1) {
    case 2: return launch_n<2>(a, Dhost, pml, st);
    case 3: return launch_n<3>(a, Dhost, pml, st);
    case 4: return launch_n<4>(a, Dhost, pml, st);
    case 5: return launch_n<5>(a, Dhost, pml, st);
    case 6: return launch_n<6>(a, Dhost, pml, st);
    case 7: return launch_n<7>(a, Dhost, pml, st);
    case 8: return launch_n<8>(a, Dhost, pml, st);
    case 9: return launch_n<9>(a, Dhost, pml, st);
    case 10: return launch_n<10>(a, Dhost, pml, st);
    case 11: return launch_n<11>(a, Dhost, pml, st);
    case 12: return launch_n<12>(a, Dhost, pml, st);
    case 13: return launch_n<13>(a, Dhost, pml, st);
    case 14: return launch_n<14>(a, Dhost, pml, st);
    case 15: return launch_n<15>(a, Dhost, pml, st);
    case 16: return launch_n<16>(a, Dhost, pml, st);
    default: return -1;
    }
}

} // namespace nkb
