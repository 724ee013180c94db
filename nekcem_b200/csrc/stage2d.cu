// Fused 2D (TE / TM) Maxwell RK-stage kernel for sm_100a, 2 <= nx1 <= 24.
//
// The 2D modes evolve three components: TM (imode 2) Hx, Hy, Ez; TE (imode 1) Ex, Ey, Hz
// [cem_maxwell, src/cem_maxwell.F:584-596].  One thread owns one node of one element for the whole
// stage, EB elements share a CTA:
//   P0  stage the three active components and dxm1 in shared memory
//   P1  face threads: numerical flux of every face point of the CTA's elements
//       [restrict_to_face :604-652, flux2d :811-920 incl. the PEC/PML fix :1368-1426; the
//        gs_op_fields pair-sum of :891/:912 is the vmapP gather or the NCCL halo] -> smem
//   P2  node threads: (ur,us) = local_grad2 (src/nek5_grad.F:21-34) as n-term dot products in the
//       mxm order, weighted curl exactly as maxwell_wght_curl :1498-1534 (each derivative times
//       w3mn first), sign, then the lifts of the node's faces in ascending face order
//       [add_flux_to_res :725-752], pml_step, Drude/Lorentz, volume source, inverse mass, RK
//       update -- every array crosses HBM once per stage.
// Face slots are in preprocessor order -y,+x,+y,-x like the reference's cemface
// (src/cem_common.F:234-260); a face point p of slot s sits at node (p,0), (n-1,p), (p,n-1), (0,p).
#include "stage_aux.h"

namespace nkb {
namespace {

template <int N>
struct Cfg2 {
    static constexpr int N2 = N * N;
    static constexpr int EB = 256 / N2 > 0 ? (256 / N2 > 8 ? 8 : 256 / N2) : 1; // elements per CTA
    static constexpr int NT = ((EB * N2 + 31) / 32) * 32;
    static constexpr int NF = 4 * N;
#ifndef S2D_MINB
#define S2D_MINB 4
#endif
    // resident CTAs per SM asked of the compiler for the plain instantiation: the kernel is bound
    // by load latency (long_scoreboard 6.8 stall cycles per issue at three CTAs of 80 registers);
    // four CTAs at 64 registers, a few spills included: 0.70 -> 0.79 of the roofline at N=7 (512^2
    // elements, TE); five or six CTAs spill too much (0.66 / 0.65)
    static constexpr int MINB = NT <= 256 ? S2D_MINB : 1;
};

template <int N, bool AUX>
__global__ void __launch_bounds__(Cfg2<N>::NT, AUX ? 1 : Cfg2<N>::MINB)
    stage2d_kernel(const __grid_constant__ StageParams<N> prm)
{
    using C = Cfg2<N>;
    constexpr int N2 = C::N2, EB = C::EB, NT = C::NT, NF = C::NF;
    const StageArgs &a = prm.a;
    __shared__ double Ds[N * N];
    __shared__ double U[3][EB * N2];
    __shared__ double F[3][EB * NF];

    const int tid = threadIdx.x;
    const bool tm = a.imode == 2;
    // active components in the (H0,H1,H2,E0,E1,E2) numbering; the third is the "z" one
    const int cA = tm ? 0 : 3, cB = tm ? 1 : 4, cC = tm ? 5 : 2;
    const int el = tid / N2, nd = tid - el * N2;
    const int slot_e = blockIdx.x * EB + el;
    const bool live = el < EB && slot_e < a.nel;
    const int e = live ? a.elist[slot_e] : 0;
    const long long gi = (long long)e * N2 + nd;

    for (int q = tid; q < N * N; q += NT) Ds[q] = prm.D[q];
    double oA = 0.0, oB = 0.0, oC = 0.0;
    if (live) {
        oA = ldg(a.u_in + cA * a.ld + gi);
        oB = ldg(a.u_in + cB * a.ld + gi);
        oC = ldg(a.u_in + cC * a.ld + gi);
        U[0][tid] = oA; U[1][tid] = oB; U[2][tid] = oC;
    }
    // loads of the epilogue, issued early
    double kA = 0.0, kB = 0.0, kC = 0.0, mH = 0.0, mE = 0.0, met[4] = {0, 0, 0, 0}, w3 = 0.0;
    if (live) {
        kA = a.kf[cA * a.ld + gi]; kB = a.kf[cB * a.ld + gi]; kC = a.kf[cC * a.ld + gi];
        mH = ldg(a.hbm1 + gi); mE = ldg(a.ebm1 + gi);
        met[0] = ldg(a.met[0] + gi); met[1] = ldg(a.met[1] + gi); // rx, ry
        met[2] = ldg(a.met[3] + gi); met[3] = ldg(a.met[4] + gi); // sx, sy
        w3 = ldg(a.w3 + nd);
    }
    __syncthreads();

    // ---- P1: fluxes of the face points (flux2d) ---------------------------------------------------
    for (int t = tid; t < EB * NF; t += NT) {
        const int fe = t / NF, fp = t - fe * NF;
        const int se = blockIdx.x * EB + fe;
        if (se >= a.nel) continue;
        const int ee = a.elist[se];
        const int s = fp / N, p = fp - s * N;
        const int node = s == 0 ? p : (s == 1 ? (N - 1) + N * p : (s == 2 ? p + N * (N - 1) : N * p));
        const long long jf = (long long)ee * NF + fp;
        const int vp = ldg(a.vmapP + jf);
        const double unx = ldg(a.unx + jf), uny = ldg(a.uny + jf), ar = ldg(a.area + jf);
        const double hY = ldg(a.hY + jf), Y1 = ldg(a.Y1 + jf), hZ = ldg(a.hZ + jf), Z1 = ldg(a.Z1 + jf);
        double oa = U[0][fe * N2 + node], ob = U[1][fe * N2 + node], oc = U[2][fe * N2 + node];
        double pa = 0.0, pb = 0.0, pc = 0.0;
        if (vp >= 0) {
            pa = ldg(a.u_in + cA * a.ld + vp); pb = ldg(a.u_in + cB * a.ld + vp);
            pc = ldg(a.u_in + cC * a.ld + vp);
        } else if (vp <= -3) {
            const double *hp = a.halo + 6ll * (long long)(-(vp + 3));
            pa = ldg(hp + cA); pb = ldg(hp + cB); pc = ldg(hp + cC);
        }
        if (a.inc_own != nullptr) { // userinc hook (src/cem_maxwell.F:498)
            const int qo = a.inc_own[jf], qn = a.inc_nbr[jf];
            if (qo >= 0) {
                const double ui = cos(a.inc_phase[qo] - a.inc_wt);
                oa += a.inc_amp[cA * a.inc_n + qo] * ui; ob += a.inc_amp[cB * a.inc_n + qo] * ui;
                oc += a.inc_amp[cC * a.inc_n + qo] * ui;
            }
            if (qn >= 0) {
                const double ui = cos(a.inc_phase[qn] - a.inc_wt);
                pa += a.inc_amp[cA * a.inc_n + qn] * ui; pb += a.inc_amp[cB * a.inc_n + qn] * ui;
                pc += a.inc_amp[cC * a.inc_n + qn] * ui;
            }
        }
        // own side (:825-828 TM, :880-883 TE): f0 = -ny*z, f1 = nx*z, f2 = -nx*B + ny*A
        double f0 = -uny * oc, f1 = unx * oc, f2 = -unx * ob + uny * oa;
        // userfsrc hook (:850-851 TM, :887-888 TE): graphene sheet current subtracted from the
        // -(n x H) slots -- f2 in TM (component 3), f0,f1 in TE (components 1,2)
        int gqn = -1;
        if (AUX) {
            if (a.fs_own != nullptr) {
                const int gqo = a.fs_own[jf];
                gqn = a.fs_nbr[jf];
                if (gqo >= 0) {
                    if (tm) f2 = f2 - a.fs_val[2 * a.fs_n + gqo];
                    else { f0 = f0 - a.fs_val[gqo]; f1 = f1 - a.fs_val[a.fs_n + gqo]; }
                }
            }
        }
        if (vp >= 0 || vp <= -3) {
            // neighbour's contribution with n+ = -n- (the gs_op_fields sum)
            f0 = f0 + uny * pc; f1 = f1 - unx * pc; f2 = f2 - (-unx * pb + uny * pa);
            if (AUX) {
                if (gqn >= 0) {
                    if (tm) f2 = f2 - a.fs_val[2 * a.fs_n + gqn];
                    else { f0 = f0 - a.fs_val[gqn]; f1 = f1 - a.fs_val[a.fs_n + gqn]; }
                }
            }
        } else if (vp == -1) { // PEC / PML outer face (cem_maxwell_flux_pec :1407-1421)
            if (tm) { f0 = 2.0 * f0; f1 = 2.0 * f1; f2 = 0.0; }
            else { f0 = 0.0; f1 = 0.0; f2 = 2.0 * f2; }
        }
        const double g1 = uny * f2, g2 = -unx * f2, g3 = unx * f1 - uny * f0;
        double r0, r1, r2;
        if (tm) { // :899-905
            r0 = hY * (-Y1 * f0 - a.C0 * g1);
            r1 = hY * (-Y1 * f1 - a.C0 * g2);
            r2 = hZ * (Z1 * f2 - a.C0 * g3);
        } else { // :909-915
            r0 = hZ * (Z1 * f0 - a.C0 * g1);
            r1 = hZ * (Z1 * f1 - a.C0 * g2);
            r2 = hY * (-Y1 * f2 - a.C0 * g3);
        }
        F[0][t] = ar * r0; F[1][t] = ar * r1; F[2][t] = ar * r2;
    }
    __syncthreads();

    // ---- P2: node threads ---------------------------------------------------------------------------
    if (!live) return;
    const int i = nd % N, j = nd / N;
    double ur[3], us[3];
    {
        const double *Ue = &U[0][el * N2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double *Uc = Ue + c * (EB * N2);
            double sr = Ds[i] * Uc[N * j], ss = Ds[j] * Uc[i];
#pragma unroll
            for (int m = 1; m < N; m++) {
                sr = sr + Ds[i + N * m] * Uc[m + N * j];
                ss = ss + Ds[j + N * m] * Uc[i + N * m];
            }
            ur[c] = sr; us[c] = ss;
        }
    }
    const double rx = met[0], ry = met[1], sx = met[2], sy = met[3];
    const double u1rw = ur[0] * w3, u1sw = us[0] * w3, u2rw = ur[1] * w3, u2sw = us[1] * w3;
    const double u3rw = ur[2] * w3, u3sw = us[2] * w3;
    double rA = (u3rw * ry + u3sw * sy);
    double rB = -(u3rw * rx + u3sw * sx);
    double rC = (u2rw * rx + u2sw * sx - u1rw * ry - u1sw * sy);
    // chsign of the H residuals (cem_maxwell :587-588, :595)
    if (tm) { rA = -rA; rB = -rB; }
    else rC = -rC;
    // lifts in ascending face order
    {
        const int fb = el * NF;
        if (j == 0) { rA = rA + F[0][fb + i]; rB = rB + F[1][fb + i]; rC = rC + F[2][fb + i]; }
        if (i == N - 1) { rA = rA + F[0][fb + N + j]; rB = rB + F[1][fb + N + j]; rC = rC + F[2][fb + N + j]; }
        if (j == N - 1) { rA = rA + F[0][fb + 2 * N + i]; rB = rB + F[1][fb + 2 * N + i]; rC = rC + F[2][fb + 2 * N + i]; }
        if (i == 0) { rA = rA + F[0][fb + 3 * N + j]; rB = rB + F[1][fb + 3 * N + j]; rC = rC + F[2][fb + 3 * N + j]; }
    }
    if (AUX) {
        const int ef = a.elflag[e];
        if (ef & 1) {
            rA = pml_component(a, gi, 0, !tm, rA, oA);
            rB = pml_component(a, gi, 1, !tm, rB, oB);
            rC = pml_component(a, gi, 2, tm, rC, oC);
        }
        if ((ef & 2) && a.ade_mask[gi]) {
            if (tm) rC = ade_component(a, gi, 2, rC, oC);
            else {
                rA = ade_component(a, gi, 0, rA, oA);
                rB = ade_component(a, gi, 1, rB, oB);
            }
        }
    }
    if (a.src_prof != nullptr) { // usersrc hook: res(comp) -= profile*(tfac*bm)
        const double sv = ldg(a.src_prof + gi) * (a.src_tfac * ldg(a.bmn + gi));
        if (a.src_comp == cA) rA -= sv;
        else if (a.src_comp == cB) rB -= sv;
        else if (a.src_comp == cC) rC -= sv;
    }
    // invqmass (:1888-1901) and rk4_upd
    const double mAB = tm ? mH : mE, mC = tm ? mE : mH;
    double t;
    t = a.ca * kA + a.dt * (rA * mAB); a.kf[cA * a.ld + gi] = t; a.u_out[cA * a.ld + gi] = oA + a.cb * t;
    t = a.ca * kB + a.dt * (rB * mAB); a.kf[cB * a.ld + gi] = t; a.u_out[cB * a.ld + gi] = oB + a.cb * t;
    t = a.ca * kC + a.dt * (rC * mC); a.kf[cC * a.ld + gi] = t; a.u_out[cC * a.ld + gi] = oC + a.cb * t;
}

template <int N>
int launch_n(const StageArgs &a, const double *Dhost, bool aux, cudaStream_t st)
{
    using C = Cfg2<N>;
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    const int grid = (a.nel + C::EB - 1) / C::EB;
    if (aux) stage2d_kernel<N, true><<<grid, C::NT, 0, st>>>(prm);
    else stage2d_kernel<N, false><<<grid, C::NT, 0, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

} // namespace

// returns 0 ok, -1 unsupported order, >0 CUDA failure.  Dhost = dxm1 (n*n, column-major).
// compiled twice (Makefile): as is, and with -fmad=false -DNKB_STRICT (desc.strict)
#ifdef NKB_STRICT
int launch_stage2d_strict(const StageArgs &a, const double *Dhost, int nx1, bool aux, void *stream)
#else
int launch_stage2d(const StageArgs &a, const double *Dhost, int nx1, bool aux, void *stream)
#endif
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx1) {
    case 2: return launch_n<2>(a, Dhost, aux, st);
    case 3: return launch_n<3>(a, Dhost, aux, st);
    case 4: return launch_n<4>(a, Dhost, aux, st);
    case 5: return launch_n<5>(a, Dhost, aux, st);
    case 6: return launch_n<6>(a, Dhost, aux, st);
    case 7: return launch_n<7>(a, Dhost, aux, st);
    case 8: return launch_n<8>(a, Dhost, aux, st);
    case 9: return launch_n<9>(a, Dhost, aux, st);
    case 10: return launch_n<10>(a, Dhost, aux, st);
    case 11: return launch_n<11>(a, Dhost, aux, st);
    case 12: return launch_n<12>(a, Dhost, aux, st);
    case 13: return launch_n<13>(a, Dhost, aux, st);
    case 14: return launch_n<14>(a, Dhost, aux, st);
    case 15: return launch_n<15>(a, Dhost, aux, st);
    case 16: return launch_n<16>(a, Dhost, aux, st);
    case 17: return launch_n<17>(a, Dhost, aux, st);
    case 18: return launch_n<18>(a, Dhost, aux, st);
    case 19: return launch_n<19>(a, Dhost, aux, st);
    case 20: return launch_n<20>(a, Dhost, aux, st);
    case 21: return launch_n<21>(a, Dhost, aux, st);
    case 22: return launch_n<22>(a, Dhost, aux, st);
    case 23: return launch_n<23>(a, Dhost, aux, st);
    case 24: return launch_n<24>(a, Dhost, aux, st);
    default: return -1;
    }
}

} // namespace nkb
