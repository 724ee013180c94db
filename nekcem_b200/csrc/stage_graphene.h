// Graphene sheets: the surface-current auxiliary differential equations of one face point
// (internal; not part of the C ABI).  Host-compilable on purpose: tests/test_graphene_point.py
// builds this header with g++ and checks it against the oracle without a GPU.
//
// Reference: cem_3d_graphene_current / cem_te_graphene_current / cem_tm_graphene_current
// (src/cem_maxwell.F:2827-2931, 2933-3022, 3024-3093), called once per RK stage from the .usr's
// userfsrc (tests/3dgraphene/3dgraphene.usr:236-266, tests/2dgraphene/2dgraphene.usr:249-290),
// i.e. inside cem_maxwell_flux3d/2d after the own-side -(n x H), -(n x E) are formed and before
// the gs_op_fields sum (src/cem_maxwell.F:958-962).
#pragma once

#if defined(__CUDACC__)
#define NKB_HD __host__ __device__ __forceinline__
#else
#define NKB_HD inline
#endif

namespace nkb {

// State of one face point: fj[c + 3*m], kj[c + 3*m] = fjn(j,c+1,m+1), kfjn(j,c+1,m+1);
// m = 0 is the algebraic total current (the value userfsrc subtracts from -(n x H)), m = 1 the
// Drude term, m = 2,3 and m = 4,5 the two critical-point pairs.  par[0..11] = params(j,1:12) =
// (a_d, b_d, b_cp1, a_211, a_221, b_11, b_21, b_cp2, a_212, a_222, b_12, b_22).
// H, E: the own-side face values fHN(j,:), fEN(j,:) (after userinc); n: unit normal;
// Yfac = 0.5/Y_0(j); yc = yconduc(j).  imode 3: all components; 1 (TE): components 0,1 from
// (Hz; Ex,Ey); 2 (TM): component 2 from (Hx,Hy; Ez).  ca, cb, dt: rk4_upd (src/cem_common.F:18-76)
// with ca = rk4a(rkstep), cb = rk4b(rkstep).
// IMODE is a template parameter so that the component range is a compile-time constant: the
// loops unroll and fj, kj, par stay in registers on the device (no local memory).
template <int IMODE>
NKB_HD void graphene_point_t(const double H[3], const double E[3], const double n[3], double Yfac,
                             double yc, const double par[12], double fj[18], double kj[18],
                             double ca, double cb, double dt)
{
    double nH[3] = {0.0, 0.0, 0.0}, nEn[3] = {0.0, 0.0, 0.0};
    constexpr int imode = IMODE;
    constexpr int c0 = IMODE == 2 ? 2 : 0, c1 = IMODE == 1 ? 2 : 3;
    if (imode == 3) { // :2862-2871
        nH[0] = -n[1] * H[2] + n[2] * H[1];
        nH[1] = n[0] * H[2] - n[2] * H[0];
        nH[2] = -n[0] * H[1] + n[1] * H[0];
        const double ndotE = n[0] * E[0] + n[1] * E[1] + n[2] * E[2];
        nEn[0] = E[0] - n[0] * ndotE;
        nEn[1] = E[1] - n[1] * ndotE;
        nEn[2] = E[2] - n[2] * ndotE;
    } else if (imode == 1) { // TE :2969-2976
        nH[0] = -n[1] * H[2];
        nH[1] = n[0] * H[2];
        nEn[0] = (n[1] * n[1]) * E[0] - (n[0] * n[1]) * E[1];
        nEn[1] = (n[0] * n[0]) * E[1] - (n[0] * n[1]) * E[0];
    } else { // TM :3061-3065: n x (E x n) = E
        nH[2] = -n[0] * H[1] + n[1] * H[0];
        nEn[2] = E[2];
    }
    const double a_d = par[0], b_d = par[1], b_cp1 = par[2], a_211 = par[3], a_221 = par[4],
                 b_11 = par[5], b_21 = par[6], b_cp2 = par[7], a_212 = par[8], a_222 = par[9],
                 b_12 = par[10], b_22 = par[11];
    const double cpfac = b_cp1 + b_cp2;
    const double jnfac = 1.0 - cpfac * Yfac;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = c0; c < c1; c++) {
        const double tmp = Yfac * (nH[c] + yc * nEn[c]);
        const double j1 = fj[c + 3], j2 = fj[c + 6], j3 = fj[c + 9], j4 = fj[c + 12],
                     j5 = fj[c + 15];
        const double j0 = (j1 + j2 + j4 - cpfac * tmp) / jnfac; // total current
        fj[c] = j0;
        const double f = tmp - Yfac * j0; // forcing term
        const double r1 = -a_d * j1 + b_d * f;                // Drude
        const double r2 = j3 + b_11 * f;                      // first critical point
        const double r3 = -a_211 * j2 - a_221 * j3 + b_21 * f;
        const double r4 = j5 + b_12 * f;                      // second critical point
        const double r5 = -a_212 * j4 - a_222 * j5 + b_22 * f;
        double t;
        t = ca * kj[c + 3] + dt * r1;  kj[c + 3] = t;  fj[c + 3] = j1 + cb * t;
        t = ca * kj[c + 6] + dt * r2;  kj[c + 6] = t;  fj[c + 6] = j2 + cb * t;
        t = ca * kj[c + 9] + dt * r3;  kj[c + 9] = t;  fj[c + 9] = j3 + cb * t;
        t = ca * kj[c + 12] + dt * r4; kj[c + 12] = t; fj[c + 12] = j4 + cb * t;
        t = ca * kj[c + 15] + dt * r5; kj[c + 15] = t; fj[c + 15] = j5 + cb * t;
    }
}

NKB_HD void graphene_point(int imode, const double H[3], const double E[3], const double n[3],
                           double Yfac, double yc, const double par[12], double fj[18],
                           double kj[18], double ca, double cb, double dt)
{
    if (imode == 3) graphene_point_t<3>(H, E, n, Yfac, yc, par, fj, kj, ca, cb, dt);
    else if (imode == 1) graphene_point_t<1>(H, E, n, Yfac, yc, par, fj, kj, ca, cb, dt);
    else graphene_point_t<2>(H, E, n, Yfac, yc, par, fj, kj, ca, cb, dt);
}

} // namespace nkb
