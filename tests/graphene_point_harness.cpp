// CPU harness for nekcem_b200/csrc/stage_graphene.h (TEST INFRASTRUCTURE): the per-face-point
// graphene update is a __host__ __device__ function, so the arithmetic the GPU kernel runs can
// be compiled with g++ and compared with the oracle without a GPU (tests/test_graphene_point.py).
#include "../nekcem_b200/csrc/stage_graphene.h"

extern "C" void graphene_points(int imode, int n, const double *H, const double *E,
                                const double *nrm, const double *Yfac, const double *yc,
                                const double *par, double *fj, double *kj, double ca, double cb,
                                double dt)
{
    // arrays are [m][n] like the library's compact device layout
    for (int q = 0; q < n; q++) {
        double h[3], e[3], nn[3], p[12], f[18], k[18];
        for (int c = 0; c < 3; c++) {
            h[c] = H[c * n + q];
            e[c] = E[c * n + q];
            nn[c] = nrm[c * n + q];
        }
        for (int m = 0; m < 12; m++) p[m] = par[m * n + q];
        for (int m = 0; m < 18; m++) {
            f[m] = fj[m * n + q];
            k[m] = kj[m * n + q];
        }
        nkb::graphene_point(imode, h, e, nn, Yfac[q], yc[q], p, f, k, ca, cb, dt);
        for (int m = 0; m < 18; m++) {
            fj[m * n + q] = f[m];
            kj[m * n + q] = k[m];
        }
    }
}
