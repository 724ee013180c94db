// Order dispatch of the pipelined stage kernel: one translation unit per nx1 (stage_pipe.cu
// compiled with -DPIPE_ONLY_N=<nx1>, see the Makefile's PIPE_ORDERS).
#include "stage_args.h"

namespace nkb {

#define PIPE_DECL(n) \
    int launch_stage_pipe_n##n(const StageArgs &a, const double *Dhost, bool pml, bool cm, void *stream);
#define PIPE_CASE(n) \
    case n: return launch_stage_pipe_n##n(a, Dhost, pml, cm, stream);

PIPE_ORDER_LIST(PIPE_DECL)

// returns 0 ok, -1 order (or alignment) not covered by this kernel (the caller uses
// launch_stage_slab), >0 CUDA failure
int launch_stage_pipe(const StageArgs &a, const double *Dhost, int nx1, bool pml, bool cm,
                      void *stream)
{
    // nx1 = 5, 7: general lists only (measured against the slab kernel, same box: nx1 = 7 general
    // 0.58 -> 0.615, constant-metric 0.715 -> 0.67; nx1 = 5: 0.45 -> 0.53 / +-0; nx1 = 6 loses both)
    if (nx1 < 8 && cm) return -1;
    switch (nx1) {
        PIPE_ORDER_LIST(PIPE_CASE)
    default: return -1;
    }
}

} // namespace nkb
