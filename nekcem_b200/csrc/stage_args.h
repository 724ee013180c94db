// Arguments of one fused RK-stage launch (internal; not part of the C ABI).
#pragma once
#include <cstdint>

namespace nkb {

constexpr int XTR_BIAS = 1 << 30;

struct StageArgs {
    // fields: ping-pong H,E (6 components, leading dimension ld) and the RK register k
    const double *u_in;
    double *u_out;
    double *kf;
    long long ld;
    // volume geometry (npts each) -- src/GEOM:30-45
    const double *met[9]; // rxmn,rymn,rzmn,sxmn,symn,szmn,txmn,tymn,tzmn
    const double *hbm1, *ebm1, *bmn;
    const double *D;  // dxm1, column-major n x n
    const double *w3; // w3mn(nxyz)
    // face geometry/material (nxzfl each).  hY = 0.5/Y_0, hZ = 0.5/Z_0 (computed once at setup:
    // the reference evaluates 0.5/Y0 first in every product, src/cem_maxwell.F:976-979)
    const double *unx, *uny, *unz, *area, *hY, *Y1, *hZ, *Z1;
    const int *vmapP;   // >=0: local volume node of the neighbour trace; -1: PEC mirror;
                        // -2: unpaired non-PEC face; <=-3: s = -(v+3): halo slot s, or, for
                        // s >= XTR_BIAS (3D), entry s - XTR_BIAS of the x-face mirror
    const double *halo; // [nhalo][6] traces received from peer ranks
    // 3D: compact mirror of the fields on the -x/+x faces of every element, (6, ldx), entry
    // (2e + side)*n^2 + j + n*k.  A neighbour trace across an x face is a stride-n gather in the
    // volume array (one 32-byte sector and one L1 tag per value); from the mirror it is a
    // contiguous read.  Written by the epilogue next to the fields (ping-pong like them).
    const double *xtr_in;
    double *xtr_out;
    long long ldx;
    const int *elist;   // element ids handled by this launch
    int nel;
    int grid_cap;       // persistent kernels: at most this many CTAs (0: fill the device)
    double ca, cb, dt, C0;
    // PML auxiliary fields (PML launches only) -- src/PML
    const double *sig, *eps, *mu;
    double *pB, *pD, *kB, *kD;
    long long npts;
    // incident field (userinc hook): slot of the own / the neighbour's face point in the
    // incident list (-1: none), amplitudes [6][inc_n], phases [inc_n], inc_wt = omega*rktime
    const int *inc_own, *inc_nbr;
    const double *inc_amp, *inc_phase;
    int inc_n;
    double inc_wt;
    // Drude / Lorentz auxiliary differential equations (cem_maxwell_drude / _lorentz,
    // src/cem_maxwell.F:3095-3211), AUX launches only.  ade_kind 0 none, 1 Drude, 2 Lorentz;
    // ade_j, ade_k: (npts,3) or (npts,3,2); ade_par: (npts,2) or (npts,3); ade_mask: 1 on the
    // nodes of the user's dindex list
    const unsigned char *elflag; // per element: bit 0 = PML element, bit 1 = has ADE nodes
    int ade_kind;
    double *ade_j, *ade_k;
    const double *ade_par;
    const unsigned char *ade_mask;
    // 2D: imode 2 = TM (Hx,Hy,Ez), 1 = TE (Ex,Ey,Hz)
    int imode;
    // separable volume source (usersrc hook)
    const double *src_prof;
    int src_comp;
    double src_tfac;
    // face sources of graphene sheets (userfsrc hook, AUX launches only): slot of the own / the
    // neighbour's face point in the graphene list (-1: none); fs_val[c*fs_n + q] = fjn(j,c,1) of
    // this stage, written by graphene_kernel before the stage kernel starts
    const int *fs_own, *fs_nbr;
    const double *fs_val;
    int fs_n;
};

} // namespace nkb
