// Per-node auxiliary ODE updates shared by the 3D and 2D stage kernels: PML split fields and the
// Drude / Lorentz polarisation currents, fused into the RK epilogue of the node that owns them.
#pragma once
#include "stage_common.h"

namespace nkb {

// pml_step (src/cem_maxwell_pml.F:508-592) for field component c (0..2) of H (isE = false) or
// E (isE = true) at global node gi, fused with the PML half of rk_maxwell_ab
// (src/cem_maxwell.F:1935-1960).  r is the residual of that component before invqmass, o the
// field value at stage start.  Returns the corrected residual.
// The sigma pattern is cyclic: component 0 uses (sy, sx, sz), 1 (sz, sy, sx), 2 (sx, sz, sy).
__device__ __forceinline__ double pml_component(const StageArgs &a, long long gi, int c, bool isE,
                                                double r, double o)
{
    const double bm1 = ldg(a.bmn + gi);
    const double bm1inv = 1.0 / bm1;
    const double sg[3] = {a.sig[gi], a.sig[a.npts + gi], a.sig[2 * a.npts + gi]};
    const double permitt = a.eps[gi];
    const double sA = sg[(c + 1) % 3], sB = sg[c], sC = sg[(c + 2) % 3];
    const double sAp = sA / permitt, sBp = sB / permitt, sCp = sC / permitt;
    double *pF = (isE ? a.pD : a.pB) + (long long)c * a.npts;
    double *kF = (isE ? a.kD : a.kB) + (long long)c * a.npts;
    const double b = pF[gi];
    const double rb = r * bm1inv - sAp * b;
    double p;
    if (isE) p = -sAp * b + sBp * b - sC * o;
    else p = -sAp * b + sBp * b - sCp * a.mu[gi] * o;
    const double t = a.ca * kF[gi] + a.dt * rb;
    kF[gi] = t;
    pF[gi] = b + a.cb * t;
    return r + p * bm1;
}

// The same for the three components of H (isE = false) or E (isE = true) at one node: every
// operand of the node is loaded before the arithmetic starts (fourteen independent loads in
// flight instead of three dependent rounds), shared operands once.  Operation for operation
// identical to three pml_component calls.
__device__ __forceinline__ void pml_node3(const StageArgs &a, long long gi, bool isE, double (&r)[3],
                                          const double (&o)[3])
{
    const double bm1 = ldg(a.bmn + gi);
    const double sg[3] = {ldg(a.sig + gi), ldg(a.sig + a.npts + gi), ldg(a.sig + 2 * a.npts + gi)};
    const double permitt = ldg(a.eps + gi);
    const double mu = isE ? 0.0 : ldg(a.mu + gi);
    double *pF = isE ? a.pD : a.pB, *kF = isE ? a.kD : a.kB;
    double b[3], k[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        b[c] = pF[(long long)c * a.npts + gi];
        k[c] = kF[(long long)c * a.npts + gi];
    }
    const double bm1inv = 1.0 / bm1;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double sA = sg[(c + 1) % 3], sB = sg[c], sC = sg[(c + 2) % 3];
        const double sAp = sA / permitt, sBp = sB / permitt, sCp = sC / permitt;
        const double rb = r[c] * bm1inv - sAp * b[c];
        double p;
        if (isE) p = -sAp * b[c] + sBp * b[c] - sC * o[c];
        else p = -sAp * b[c] + sBp * b[c] - sCp * mu * o[c];
        const double t = a.ca * k[c] + a.dt * rb;
        kF[(long long)c * a.npts + gi] = t;
        pF[(long long)c * a.npts + gi] = b[c] + a.cb * t;
        r[c] = r[c] + p * bm1;
    }
}

// cem_maxwell_drude / cem_maxwell_lorentz (src/cem_maxwell.F:3095-3211) for E component c at a
// node of the user's list: resE -= J*bm, the current's own ODE, and its rk4_upd.  Returns the
// corrected residual.  e_old = E(c) at stage start.
__device__ __forceinline__ double ade_component(const StageArgs &a, long long gi, int c, double r,
                                                double e_old)
{
    const long long np = a.npts;
    const double bm = ldg(a.bmn + gi);
    if (a.ade_kind == 1) {
        const double pa = ldg(a.ade_par + gi), pb = ldg(a.ade_par + np + gi);
        const double j = a.ade_j[c * np + gi];
        r = r - j * bm;
        const double rj = -pa * j + pb * e_old;
        const double t = a.ca * a.ade_k[c * np + gi] + a.dt * rj;
        a.ade_k[c * np + gi] = t;
        a.ade_j[c * np + gi] = j + a.cb * t;
    } else {
        const double pa = ldg(a.ade_par + gi), pb = ldg(a.ade_par + np + gi),
                     pc = ldg(a.ade_par + 2 * np + gi);
        const double j0 = a.ade_j[c * np + gi], j1 = a.ade_j[(3 + c) * np + gi];
        r = r - j0 * bm;
        const double r0 = -pa * j0 - pb * j1 + pc * e_old;
        const double r1 = j0;
        double t = a.ca * a.ade_k[c * np + gi] + a.dt * r0;
        a.ade_k[c * np + gi] = t;
        a.ade_j[c * np + gi] = j0 + a.cb * t;
        t = a.ca * a.ade_k[(3 + c) * np + gi] + a.dt * r1;
        a.ade_k[(3 + c) * np + gi] = t;
        a.ade_j[(3 + c) * np + gi] = j1 + a.cb * t;
    }
    return r;
}

} // namespace nkb
