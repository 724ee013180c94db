// Graphene sheets: one thread per face point of the user's graphindex list advances the surface
// current ADEs of that point by one RK stage (stage_graphene.h) from the stage-start fields,
// exactly where the reference's userfsrc runs (own-side face values after userinc, before the
// face sum).  Slot m = 0 of fj then holds fjn(j,:,1), which the stage kernels subtract from
// -(n x H) on both sides of the face.
//
// Compiled twice (Makefile): as is, and with -fmad=false -DNKB_STRICT (desc.strict: no FMA
// contraction, the arithmetic of the reference's x86-64 build operation for operation).
#include <cuda_runtime.h>

#include "graphene_args.h"
#include "stage_graphene.h"

namespace nkb {
namespace {

__global__ void graphene_kernel(GrapheneArgs g)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.ng) return;
    const int j = g.fp[q];
    const long long nd = g.node[q];
    double H[3], E[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        H[c] = g.u[c * g.ld + nd];
        E[c] = g.u[(3 + c) * g.ld + nd];
    }
    if (g.inc_own != nullptr) { // userinc precedes the flux (src/cem_maxwell.F:498)
        const int qi = g.inc_own[j];
        if (qi >= 0) {
            const double ui = cos(g.inc_phase[qi] - g.inc_wt);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                H[c] += g.inc_amp[c * g.inc_n + qi] * ui;
                E[c] += g.inc_amp[(3 + c) * g.inc_n + qi] * ui;
            }
        }
    }
    const double n[3] = {g.unx[j], g.uny[j], g.unz ? g.unz[j] : 0.0};
    double par[12], fj[18], kj[18];
#pragma unroll
    for (int m = 0; m < 12; m++) par[m] = g.par[(long long)m * g.ng + q];
#pragma unroll
    for (int m = 0; m < 18; m++) {
        fj[m] = g.fj[(long long)m * g.ng + q];
        kj[m] = g.kj[(long long)m * g.ng + q];
    }
    nkb::graphene_point(g.imode, H, E, n, g.hY[j], g.yc[q], par, fj, kj, g.ca, g.cb, g.dt);
#pragma unroll
    for (int m = 0; m < 18; m++) {
        g.fj[(long long)m * g.ng + q] = fj[m];
        g.kj[(long long)m * g.ng + q] = kj[m];
    }
}

} // namespace

#ifdef NKB_STRICT
int launch_graphene_strict(const GrapheneArgs &g, void *stream)
#else
int launch_graphene(const GrapheneArgs &g, void *stream)
#endif
{
    graphene_kernel<<<(unsigned)((g.ng + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

} // namespace nkb
