/* Harness around the translated reference routines (TEST INFRASTRUCTURE ONLY; see
 * oracle/build_ref.py).  It supplies what the reference's hot path calls but that lies outside
 * the path: the wall clock, the communication timer, the OpenACC presence query, the fatal
 * exit, and the three user callbacks of a case's .usr file (userinc / usersrc / userfsrc),
 * which default to the empty bodies of tests/3dboxper/3dboxper.usr and can be pointed at
 * functions supplied by the tests.  The face exchange gs_op_fields_ is NOT stubbed: it is the
 * reference's own src/jl/gs.c compiled next to the translated Fortran. */
#include <stdio.h>
#include <stdlib.h>
#define EXPORT __attribute__((visibility("default")))

typedef void (*user_cb)(double *t, double *a1, double *a2, double *a3, double *a4, double *a5,
                        double *a6);
static user_cb cb_inc, cb_src, cb_fsrc;

EXPORT void ref_set_user(int which, user_cb fn)
{
    if (which == 0) cb_inc = fn;
    else if (which == 1) cb_src = fn;
    else cb_fsrc = fn;
}

EXPORT void userinc_(double *t, double *a1, double *a2, double *a3, double *a4, double *a5, double *a6)
{
    if (cb_inc) cb_inc(t, a1, a2, a3, a4, a5, a6);
}
EXPORT void usersrc_(double *t, double *a1, double *a2, double *a3, double *a4, double *a5, double *a6)
{
    if (cb_src) cb_src(t, a1, a2, a3, a4, a5, a6);
}
EXPORT void userfsrc_(double *t, double *a1, double *a2, double *a3, double *a4, double *a5, double *a6)
{
    if (cb_fsrc) cb_fsrc(t, a1, a2, a3, a4, a5, a6);
}

/* src/nek5_comm_mpi.F:501-513 returns mpi_wtime(); timing is irrelevant to the results */
EXPORT double dclock_(void) { return 0.0; }
/* src/cem_common.F:331-369 only accumulates comm_t */
EXPORT void measure_comm_(double *t0) { (void)t0; }
/* src/nek5_acc_dummy.F: .false. without OpenACC */
EXPORT int acc_nek_present_(double *a, int *n) { (void)a; (void)n; return 0; }
EXPORT void exitt_(int *rc) { fprintf(stderr, "reference called exitt(%d)\n", rc ? *rc : 0); exit(1); }
EXPORT void q_filter_(double *w) { (void)w; fprintf(stderr, "q_filter: drive filterq from the test (see build_ref.py)\n"); exit(1); }
/* dealiased curl (src/cem_maxwell.F:1541-1729): ifdealias is off in every case of the path */
EXPORT void maxwell_wght_dcurl_(void) { fprintf(stderr, "maxwell_wght_dcurl is outside the path\n"); exit(1); }
/* global integer sum over ranks (src/nek5_mat1.F): one process here */
EXPORT int iglsum_(int *a, int *n) { (void)n; return *a; }
/* global reduction over ranks (src/nek5_comm_mpi.F:379-420): the identity on one process */
EXPORT void gop_(double *x, double *w, const char *op, int *n) { (void)x; (void)w; (void)op; (void)n; }
/* MPI broadcast of the NCCL id (src/nek5_comm_mpi.F bcast): never reached on one process */
EXPORT void bcast_(void *buf, int *len) { (void)buf; (void)len; }
/* curved-side generators (src/nek5_genxyz.F): no case of the path has curved sides */
EXPORT void sphsrf_(void) { fprintf(stderr, "sphsrf is outside the path\n"); exit(1); }
EXPORT void gensrf_(void) { fprintf(stderr, "gensrf is outside the path\n"); exit(1); }
