// Arguments of graphene_kernel (internal; not part of the C ABI).
#pragma once

namespace nkb {

struct GrapheneArgs {
    const double *u;
    long long ld;
    int ng, imode;
    const int *fp, *node;
    const double *unx, *uny, *unz, *hY, *yc, *par;
    double *fj, *kj;
    const int *inc_own;
    const double *inc_amp, *inc_phase;
    int inc_n;
    double inc_wt, ca, cb, dt;
};

} // namespace nkb
