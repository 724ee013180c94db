// Fused Maxwell RK-stage kernels for sm_100a (generic-order version, 2 <= nx1 <= 14).
//
// One launch = one RK stage over a list of elements.  One CTA = one element.  Per element
// the kernel does, in ONE pass over HBM:
//   phase 1  stage the element's 6 x n^3 field nodes (H,E at stage start) in shared memory
//   phase 2  surface flux: own trace from smem, neighbour trace gathered through vmapP
//            (or the NCCL halo), PEC mirror, upwind/central flux, times face area -> smem
//            [cem_maxwell_restrict_to_face + flux3d + flux_pec, src/cem_maxwell.F:604-652,
//             922-1002, 1368-1426; the gs_op_fields pair-sum of :962 is the vmapP gather]
//   phase 3  per node: tensor-product derivatives D_r,D_s,D_t of all six components,
//            weighted curl with the metric cofactors [maxwell_wght_curl :1428-1497], lift
//            [add_flux_to_res :725-735], PML ADEs [pml_step, src/cem_maxwell_pml.F:508-592],
//            volume source [usersrc hook :503], inverse mass [invqmass :1878-1886] and the
//            low-storage RK update [rk4_upd src/cem_common.F:18-76], written to the
//            ping-pong field buffer.
//
// Arithmetic follows the reference's grouping and left-to-right summation order (mxfK,
// src/nek5_mxm_std.F:173-190); nvcc contracts a*b+c into FMA, which the reference's CPU
// build does not -- a <=1e-15 relative effect per operation (DESIGN.md "Numerics").
#include <cuda_runtime.h>

#include "stage_args.h"

namespace nkb {

__host__ __device__ constexpr int kt_for(int n)
{
    // <= 1024 threads (64 regs) up to n = 10, <= 512 threads (128 regs) above: the fully
    // unrolled 18-accumulator derivative loop spills at 64 registers for n >= 11
    int kt = (n <= 10 ? 1024 : 512) / (n * n);
    if (kt > n) kt = n;
    if (kt < 1) kt = 1;
    return kt;
}

template <int N, bool PML>
__global__ void __launch_bounds__(N * N * kt_for(N))
    stage_kernel(const StageArgs a)
{
    constexpr int N2 = N * N, N3 = N2 * N, KT = kt_for(N), NT = N2 * KT, NF = 6 * N2;
    extern __shared__ double smem[];
    double *U = smem;        // [6][N3]   H,E at stage start
    double *F = U + 6 * N3;  // [6][NF]   area * numerical flux, component-major
    double *Ds = F + 6 * NF; // [N*N]     D(i,m) at i + N*m

    const int tid = threadIdx.x;
    const int e = a.elist[blockIdx.x];
    const long long ebase = (long long)e * N3;

    // ---- phase 1: stage fields -----------------------------------------------------
    for (int q = tid; q < 6 * N3; q += NT) {
        const int c = q / N3, r = q - c * N3;
        U[q] = a.u_in[c * a.ld + ebase + r];
    }
    for (int q = tid; q < N * N; q += NT) Ds[q] = a.D[q];
    __syncthreads();

    // ---- phase 2: surface flux -----------------------------------------------------
    for (int fp = tid; fp < NF; fp += NT) {
        const int s = fp / N2, p = fp - s * N2;
        const int pa = p % N, pb = p / N;
        int node;
        switch (s) { // preprocessor face order: -y,+x,+y,-x,-z,+z (cemface, cem_common.F:234-260)
        case 0: node = pa + N2 * pb; break;
        case 1: node = (N - 1) + N * pa + N2 * pb; break;
        case 2: node = pa + N * (N - 1) + N2 * pb; break;
        case 3: node = N * pa + N2 * pb; break;
        case 4: node = pa + N * pb; break;
        default: node = pa + N * pb + N2 * (N - 1); break;
        }
        const long long jf = (long long)e * NF + fp;
        const double unx = a.unx[jf], uny = a.uny[jf], unz = a.unz[jf];
        const double Hx = U[node], Hy = U[N3 + node], Hz = U[2 * N3 + node];
        const double Ex = U[3 * N3 + node], Ey = U[4 * N3 + node], Ez = U[5 * N3 + node];
        // -n x E, -n x H of the own side (flux3d :946-955)
        double s0 = -uny * Ez + unz * Ey;
        double s1 = -unz * Ex + unx * Ez;
        double s2 = -unx * Ey + uny * Ex;
        double s3 = -uny * Hz + unz * Hy;
        double s4 = -unz * Hx + unx * Hz;
        double s5 = -unx * Hy + uny * Hx;
        const int vp = a.vmapP[jf];
        if (vp >= 0 || vp <= -3) {
            double pHx, pHy, pHz, pEx, pEy, pEz;
            if (vp >= 0) {
                pHx = a.u_in[vp];
                pHy = a.u_in[a.ld + vp];
                pHz = a.u_in[2 * a.ld + vp];
                pEx = a.u_in[3 * a.ld + vp];
                pEy = a.u_in[4 * a.ld + vp];
                pEz = a.u_in[5 * a.ld + vp];
            } else {
                const double *h = a.halo + 6ll * (long long)(-(vp + 3));
                pHx = h[0]; pHy = h[1]; pHz = h[2];
                pEx = h[3]; pEy = h[4]; pEz = h[5];
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
        const double hY = a.hY[jf], Y1 = a.Y1[jf], hZ = a.hZ[jf], Z1 = a.Z1[jf];
        const double Y02 = -(hY * Y1);
        const double Z02 = hZ * Z1;
        const double C02Y = hY * a.C0;
        const double C02Z = hZ * a.C0;
        const double fu1 = uny * s5 - unz * s4;
        const double fu2 = unz * s3 - unx * s5;
        const double fu3 = unx * s4 - uny * s3;
        const double fw1 = uny * s2 - unz * s1;
        const double fw2 = unz * s0 - unx * s2;
        const double fw3 = unx * s1 - uny * s0;
        const double ar = a.area[jf];
        F[0 * NF + fp] = ar * (Y02 * s0 - C02Y * fu1);
        F[1 * NF + fp] = ar * (Y02 * s1 - C02Y * fu2);
        F[2 * NF + fp] = ar * (Y02 * s2 - C02Y * fu3);
        F[3 * NF + fp] = ar * (Z02 * s3 - C02Z * fw1);
        F[4 * NF + fp] = ar * (Z02 * s4 - C02Z * fw2);
        F[5 * NF + fp] = ar * (Z02 * s5 - C02Z * fw3);
    }
    __syncthreads();

    // ---- phase 3: volume + lift + ADEs + inverse mass + RK -----------------------------
    const int i = tid % N, j = (tid / N) % N, kz = tid / N2;
    for (int k = kz; k < N; k += KT) {
        const int node = i + N * j + N2 * k;
        const long long g = ebase + node;
        double ur[6], us[6], ut[6];
        {
            const double di = Ds[i], dj = Ds[j], dk = Ds[k];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                ur[c] = di * U[c * N3 + N * j + N2 * k];
                us[c] = U[c * N3 + i + N2 * k] * dj;
                ut[c] = U[c * N3 + i + N * j] * dk;
            }
        }
#pragma unroll
        for (int m = 1; m < N; m++) {
            const double di = Ds[i + N * m], dj = Ds[j + N * m], dk = Ds[k + N * m];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                ur[c] = ur[c] + di * U[c * N3 + m + N * j + N2 * k];
                us[c] = us[c] + U[c * N3 + i + N * m + N2 * k] * dj;
                ut[c] = ut[c] + U[c * N3 + i + N * j + N2 * m] * dk;
            }
        }
        const double w = a.w3[node];
        const double rx = a.rx[g], ry = a.ry[g], rz = a.rz[g];
        const double sx = a.sx[g], sy = a.sy[g], sz = a.sz[g];
        const double tx = a.tx[g], ty = a.ty[g], tz = a.tz[g];
        double rH0, rH1, rH2, rE0, rE1, rE2;
        { // resEN = wcurl(HN)
            const double u1rw = ur[0] * w, u1sw = us[0] * w, u1tw = ut[0] * w;
            const double u2rw = ur[1] * w, u2sw = us[1] * w, u2tw = ut[1] * w;
            const double u3rw = ur[2] * w, u3sw = us[2] * w, u3tw = ut[2] * w;
            rE0 = u3rw * ry + u3sw * sy + u3tw * ty - u2rw * rz - u2sw * sz - u2tw * tz;
            rE1 = u1rw * rz + u1sw * sz + u1tw * tz - u3rw * rx - u3sw * sx - u3tw * tx;
            rE2 = u2rw * rx + u2sw * sx + u2tw * tx - u1rw * ry - u1sw * sy - u1tw * ty;
        }
        { // resHN = -wcurl(EN)
            const double u1rw = ur[3] * w, u1sw = us[3] * w, u1tw = ut[3] * w;
            const double u2rw = ur[4] * w, u2sw = us[4] * w, u2tw = ut[4] * w;
            const double u3rw = ur[5] * w, u3sw = us[5] * w, u3tw = ut[5] * w;
            rH0 = -(u3rw * ry + u3sw * sy + u3tw * ty - u2rw * rz - u2sw * sz - u2tw * tz);
            rH1 = -(u1rw * rz + u1sw * sz + u1tw * tz - u3rw * rx - u3sw * sx - u3tw * tx);
            rH2 = -(u2rw * rx + u2sw * sx + u2tw * tx - u1rw * ry - u1sw * sy - u1tw * ty);
        }
        // lift, ascending face slot like the reference's sequential j loop (:726-735)
#define NKB_LIFT(fp_)                                                                          \
    do {                                                                                       \
        const int fq = (fp_);                                                                  \
        rH0 += F[0 * NF + fq]; rH1 += F[1 * NF + fq]; rH2 += F[2 * NF + fq];                   \
        rE0 += F[3 * NF + fq]; rE1 += F[4 * NF + fq]; rE2 += F[5 * NF + fq];                   \
    } while (0)
        if (j == 0) NKB_LIFT(0 * N2 + i + N * k);
        if (i == N - 1) NKB_LIFT(1 * N2 + j + N * k);
        if (j == N - 1) NKB_LIFT(2 * N2 + i + N * k);
        if (i == 0) NKB_LIFT(3 * N2 + j + N * k);
        if (k == 0) NKB_LIFT(4 * N2 + i + N * j);
        if (k == N - 1) NKB_LIFT(5 * N2 + i + N * j);
#undef NKB_LIFT

        const double h0 = U[node], h1 = U[N3 + node], h2 = U[2 * N3 + node];
        const double e0 = U[3 * N3 + node], e1 = U[4 * N3 + node], e2 = U[5 * N3 + node];

        if (PML) { // pml_step, src/cem_maxwell_pml.F:540-585, then the PML half of rk_maxwell_ab
            const double bm1 = a.bmn[g];
            const double bm1inv = 1.0 / bm1;
            const double sigx = a.sig[g], sigy = a.sig[a.npts + g], sigz = a.sig[2 * a.npts + g];
            const double permitt = a.eps[g];
            const double sxp = sigx / permitt, syp = sigy / permitt, szp = sigz / permitt;
            const double permeab = a.mu[g];
            const double b0 = a.pB[g], b1 = a.pB[a.npts + g], b2 = a.pB[2 * a.npts + g];
            const double d0 = a.pD[g], d1 = a.pD[a.npts + g], d2 = a.pD[2 * a.npts + g];
            const double rb0 = rH0 * bm1inv - syp * b0;
            const double rb1 = rH1 * bm1inv - szp * b1;
            const double rb2 = rH2 * bm1inv - sxp * b2;
            const double rd0 = rE0 * bm1inv - syp * d0;
            const double rd1 = rE1 * bm1inv - szp * d1;
            const double rd2 = rE2 * bm1inv - sxp * d2;
            const double ph0 = -syp * b0 + sxp * b0 - szp * permeab * h0;
            const double ph1 = -szp * b1 + syp * b1 - sxp * permeab * h1;
            const double ph2 = -sxp * b2 + szp * b2 - syp * permeab * h2;
            const double pe0 = -syp * d0 + sxp * d0 - sigz * e0;
            const double pe1 = -szp * d1 + syp * d1 - sigx * e1;
            const double pe2 = -sxp * d2 + szp * d2 - sigy * e2;
            rH0 = rH0 + ph0 * bm1; rH1 = rH1 + ph1 * bm1; rH2 = rH2 + ph2 * bm1;
            rE0 = rE0 + pe0 * bm1; rE1 = rE1 + pe1 * bm1; rE2 = rE2 + pe2 * bm1;
            double kk;
            kk = a.ca * a.kB[g] + a.dt * rb0; a.kB[g] = kk; a.pB[g] = b0 + a.cb * kk;
            kk = a.ca * a.kB[a.npts + g] + a.dt * rb1; a.kB[a.npts + g] = kk;
            a.pB[a.npts + g] = b1 + a.cb * kk;
            kk = a.ca * a.kB[2 * a.npts + g] + a.dt * rb2; a.kB[2 * a.npts + g] = kk;
            a.pB[2 * a.npts + g] = b2 + a.cb * kk;
            kk = a.ca * a.kD[g] + a.dt * rd0; a.kD[g] = kk; a.pD[g] = d0 + a.cb * kk;
            kk = a.ca * a.kD[a.npts + g] + a.dt * rd1; a.kD[a.npts + g] = kk;
            a.pD[a.npts + g] = d1 + a.cb * kk;
            kk = a.ca * a.kD[2 * a.npts + g] + a.dt * rd2; a.kD[2 * a.npts + g] = kk;
            a.pD[2 * a.npts + g] = d2 + a.cb * kk;
        }
        if (a.src_prof != nullptr) { // usersrc hook: res(c) -= profile*(tfac*bm)
            const double sv = a.src_prof[g] * (a.src_tfac * a.bmn[g]);
            switch (a.src_comp) {
            case 0: rH0 -= sv; break;
            case 1: rH1 -= sv; break;
            case 2: rH2 -= sv; break;
            case 3: rE0 -= sv; break;
            case 4: rE1 -= sv; break;
            default: rE2 -= sv; break;
            }
        }
        const double hb = a.hbm1[g], eb = a.ebm1[g];
        rH0 *= hb; rH1 *= hb; rH2 *= hb;
        rE0 *= eb; rE1 *= eb; rE2 *= eb;
        double kk;
        kk = a.ca * a.kf[g] + a.dt * rH0; a.kf[g] = kk; a.u_out[g] = h0 + a.cb * kk;
        kk = a.ca * a.kf[a.ld + g] + a.dt * rH1; a.kf[a.ld + g] = kk;
        a.u_out[a.ld + g] = h1 + a.cb * kk;
        kk = a.ca * a.kf[2 * a.ld + g] + a.dt * rH2; a.kf[2 * a.ld + g] = kk;
        a.u_out[2 * a.ld + g] = h2 + a.cb * kk;
        kk = a.ca * a.kf[3 * a.ld + g] + a.dt * rE0; a.kf[3 * a.ld + g] = kk;
        a.u_out[3 * a.ld + g] = e0 + a.cb * kk;
        kk = a.ca * a.kf[4 * a.ld + g] + a.dt * rE1; a.kf[4 * a.ld + g] = kk;
        a.u_out[4 * a.ld + g] = e1 + a.cb * kk;
        kk = a.ca * a.kf[5 * a.ld + g] + a.dt * rE2; a.kf[5 * a.ld + g] = kk;
        a.u_out[5 * a.ld + g] = e2 + a.cb * kk;
    }
}

template <int N>
static int launch_n(const StageArgs &a, bool pml, cudaStream_t st)
{
    constexpr int N2 = N * N, N3 = N2 * N, NT = N2 * kt_for(N);
    constexpr size_t smem = sizeof(double) * (6 * N3 + 36 * N2 + N * N);
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
    if (pml)
        stage_kernel<N, true><<<a.nel, NT, smem, st>>>(a);
    else
        stage_kernel<N, false><<<a.nel, NT, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// returns 0 ok, -1 unsupported order, >0 CUDA failure
int launch_stage(const StageArgs &a, int nx1, bool pml, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx1) {
    case 2: return launch_n<2>(a, pml, st);
    case 3: return launch_n<3>(a, pml, st);
    case 4: return launch_n<4>(a, pml, st);
    case 5: return launch_n<5>(a, pml, st);
    case 6: return launch_n<6>(a, pml, st);
    case 7: return launch_n<7>(a, pml, st);
    case 8: return launch_n<8>(a, pml, st);
    case 9: return launch_n<9>(a, pml, st);
    case 10: return launch_n<10>(a, pml, st);
    case 11: return launch_n<11>(a, pml, st);
    case 12: return launch_n<12>(a, pml, st);
    case 13: return launch_n<13>(a, pml, st);
    case 14: return launch_n<14>(a, pml, st);
    default: return -1;
    }
}

} // namespace nkb
