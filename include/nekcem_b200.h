/*
 * nekcem_b200.h -- C ABI of libnekcem_b200.so
 *
 * A B200-native (sm_100a) implementation of NekCEM's time-domain Maxwell SEDG
 * right-hand side fused with its 5-stage low-storage RK update.  This header is
 * the drop-in boundary: every entry point names the reference interface it
 * replaces (paths relative to the NekCEM tree).  Plain pointers and sizes only;
 * no C++/torch types.  Host language on the reference side is Fortran; the
 * `ISO_C_BINDING` module that binds these symbols is in fortran/nekcem_b200_mod.F90
 * and INTEGRATION.md shows the call sites.
 *
 * Conventions (identical to the reference's COMMON blocks):
 *   - real == double; integer == 32-bit; gs ids are 64-bit (INTEGER*8).
 *   - arrays are column-major, element-major: node (i,j,k,e) at
 *     i + nx1*(j + nx1*(k + nx1*e)); vector fields are (npts,3).
 *   - face arrays are (nx1*nz1, 2*ldim, nelt) with the face slot in PREPROCESSOR
 *     order 1..6 = (-y,+x,+y,-x,-z,+z), exactly like `area/unx` (src/nek5_coef.F:1187-1235)
 *     and `cemface` (src/cem_common.F:214-283).
 *   - index arrays (cempec, pmlptr, dindex) hold 1-BASED indices as in Fortran.
 *   - every function returns 0 on success; on failure it returns nonzero and
 *     nekcem_b200_last_error() describes it.  The trailing-underscore Fortran twins
 *     (src: nekcem_b200/csrc/fortran_abi.cu) print the message and exit(1), which is the
 *     reference's own error behaviour (`exitt`, src/nek5_comm_mpi.F:650-692; jl `fail`,
 *     src/jl/fail.c:11-17).
 *   - one host thread per context (the reference runs one thread per MPI rank).
 *   - There is NO CPU fallback: a context created with a device needs a B200; compute
 *     entry points fail loudly otherwise.
 */
#ifndef NEKCEM_B200_H
#define NEKCEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEKCEM_B200_ABI_VERSION 1

/* ---- array ids for nekcem_b200_set_array / get_array ------------------------------
 * Each id names the reference COMMON array it mirrors; this is the list that
 * `acc_copy_all_in` uploads (src/cem_drive.F:397-480). */
enum nekcem_b200_array {
    NKB_DXM1 = 0,      /* dxm1(nx1,nx1)        src/DXYZ:4-7                           */
    NKB_W3MN,          /* w3mn(nxyz)           src/GEOM (flat w3m1), cem_maxwell.F:153 */
    NKB_RXMN, NKB_RYMN, NKB_RZMN, /* rxmn..  (npts) unnormalised cofactors, src/GEOM:30-45 */
    NKB_SXMN, NKB_SYMN, NKB_SZMN,
    NKB_TXMN, NKB_TYMN, NKB_TZMN,
    NKB_BMN,           /* bmn(npts) = jac*w3   src/GEOM:33                            */
    NKB_HBM1, NKB_EBM1,/* 1/(mu*bm), 1/(eps*bm) src/EMWAVE:20-21, cem_maxwell.F:183-186 */
    NKB_UNXM, NKB_UNYM, NKB_UNZM, NKB_AREAM, /* (nxzfl) src/GEOM:52-56                   */
    NKB_Y_0, NKB_Y_1, NKB_Z_0, NKB_Z_1,      /* (nxzfl) src/EMWAVE:40-47                 */
    NKB_HN, NKB_EN,    /* (npts,3)             src/EMWAVE:5-6                          */
    NKB_KHN, NKB_KEN,  /* (npts,3) RK registers src/EMWAVE:9-10                        */
    NKB_PERMITTIVITY, NKB_PERMEABILITY,      /* (npts) src/EMWAVE:31-32 (PML only)      */
    NKB_PMLSIGMA,      /* (npts,3)             src/PML:11                              */
    NKB_PMLBN, NKB_PMLDN, NKB_KPMLBN, NKB_KPMLDN, /* (npts,3) src/PML:14-28              */
    NKB_XMN, NKB_YMN, NKB_ZMN, /* xmn,ymn,zmn(npts) node coordinates, src/GEOM:30-33; only needed
                                  by nekcem_b200_error_sums_mode                             */
    NKB_YCONDUC,       /* yconduc(nxzfl) own-side conductance on the faces, src/EMWAVE,
                          cem_maxwell.F:285; only needed by graphene sheets                  */
    NKB_ARRAY_COUNT
};

/* ---- problem description (the scalars the hot path reads from COMMON) -------------- */
typedef struct nekcem_b200_desc {
    int32_t abi_version; /* NEKCEM_B200_ABI_VERSION */
    int32_t ldim;        /* 3, or 2 for the TE/TM modes (cem_maxwell_flux2d path)     */
    int32_t nx1;         /* points per direction, N+1 (SIZE: lx1): 2..24, the range of the
                            reference's mxm (mxf1..mxf24, src/nek5_mxm_std.F)          */
    int32_t nelt;        /* local element count (SIZE/DIMN)                          */
    int32_t imode;       /* 3 = 3D, 2 = TM, 1 = TE (src/INPUT:18-47, cem_param.F)    */
    int32_t ifupwind;    /* param(19)=0 -> 1 (C0=1); central flux -> 0 (C0=0)        */
    int32_t ifpec;       /* any 'PEC' face (setlog, src/nek5_bdry.F:68-92)           */
    int32_t ifpml;       /* any 'PML' face                                           */
    int32_t device;      /* CUDA device ordinal; -1 = host-only planning context:
                            face pairing, exchange plan and registrations work, uploaded
                            arrays are kept on the host for inspection, every compute call
                            fails.  NEKCEM_B200_HOST_ONLY in the environment forces this
                            mode (test aid for driving the Fortran shim without a GPU); it
                            is not a CPU fallback -- nothing is ever computed on the host  */
    int32_t strict;      /* 1: the stage kernels and the graphene kernel compiled without FMA
                            contraction (-fmad=false): every product and sum is rounded
                            separately, as in the reference's x86-64 build.  3D contexts
                            then run the slab formulation at every order (slower)        */
    int32_t rank;        /* MPI rank (nid) and size (np): one rank <-> one GPU       */
    int32_t nranks;
} nekcem_b200_desc;

const char *nekcem_b200_last_error(void);

/* Lifetime.  Replaces the device-residency hooks `acc_copy_all_in` /
 * `acc_copy_all_out` (src/cem_drive.F:161-165, 397-566). */
int nekcem_b200_create(const nekcem_b200_desc *desc, int *handle);
int nekcem_b200_destroy(int handle);

/* Upload one COMMON array (count = number of elements of that array). */
int nekcem_b200_set_array(int handle, int which, const double *host, int64_t count);
/* Download (the `!$ACC UPDATE HOST(hn,en,...)` points: tests/3dboxper/3dboxper.usr:199,
 * src/io.F:193-195). */
int nekcem_b200_get_array(int handle, int which, double *host, int64_t count);

/* Fortran leading dimensions.  The reference dimensions its arrays by SIZE with lelt >= nelt
 * (lelt = lelg/lpmin + 3, tests/3dboxper/SIZE:15), so a vector field is hn(lpts1,3) with
 * lpts1 = lx1*ly1*lz1*lelt >= npts (src/EMWAVE:5-16, src/PML), a .usr's ADE arrays are jn(lpts,3),
 * params(lpts,2) and its graphene arrays fjn(lxzfl,3,6), params(lxzfl,12).  set_array_ld /
 * get_array_ld transfer the three components of an (ld,3) array (ids NKB_HN .. NKB_KEN,
 * NKB_PMLSIGMA .. NKB_KPMLDN; one-component arrays: same as set_array on their first entries);
 * set_leading_dims declares lpts and lxzfl for the arrays that set_drude / set_lorentz / get_ade
 * and set_graphene / get_graphene receive (default: npts, nxzfl -- compact arrays). */
int nekcem_b200_set_leading_dims(int handle, int64_t lpts, int64_t lxzfl);
int nekcem_b200_set_array_ld(int handle, int which, const double *host, int64_t ld);
int nekcem_b200_get_array_ld(int handle, int which, double *host, int64_t ld);

/* Face connectivity.  glo_num are the face-point global ids the reference hands to
 * `gs_setup(gsh_face,glo_num,ntot,nekcomm,np)` (src/nek5_connect11.F:2217-2223,
 * src/jl/gs.c:1898-1907): two face points are paired iff they carry the same non-zero id.
 * cempec/ncempec is the 1-based face-point list of `cem_maxwell_pec_init`
 * (src/cem_maxwell.F:1338-1366; 'PEC' and 'PML' faces). */
int nekcem_b200_set_faces(int handle, const int64_t *glo_num, int64_t nxzfl,
                          const int32_t *cempec, int32_t ncempec);

/* PML element list `pmlptr(1:maxpml)` (1-based), src/cem_maxwell_pml.F:461-468. */
int nekcem_b200_set_pml(int handle, const int32_t *pmlptr, int32_t maxpml);

/* Multi-rank face exchange (replaces gs_op_fields on gsh_face between ranks,
 * src/cem_maxwell.F:962, src/jl/gs.c:471-512, with grouped ncclSend/ncclRecv).
 * Rank 0 obtains an id, the host code broadcasts its 128 bytes (MPI_Bcast in the
 * Fortran shim), every rank calls comm_init. */
int nekcem_b200_comm_unique_id(char id[128]);
int nekcem_b200_comm_init(int handle, const char id[128]);

/* Host-side planning of the inter-rank exchange; exposed so that the host logic can be
 * driven (and tested) with any transport.  face_singletons: ids that are unpaired on this
 * rank.  face_remote: given every rank's singleton ids (concatenated, counts[r] each),
 * builds per-peer send/recv lists.  nekcem_b200_setup() calls these itself over NCCL when
 * nranks > 1 and a communicator exists. */
int nekcem_b200_face_singletons(int handle, int64_t *ids, int64_t capacity, int64_t *count);
int nekcem_b200_face_remote(int handle, const int64_t *counts, const int64_t *all_ids);
/* Plan queries (for tests / diagnostics). */
int nekcem_b200_plan_npeers(int handle, int32_t *npeers, int64_t *nhalo);
int nekcem_b200_plan_peer(int handle, int32_t ipeer, int32_t *peer_rank, int64_t *count,
                          int64_t *send_facepts /* capacity count, 0-based face points */);
int nekcem_b200_plan_vmap(int handle, int32_t *vmapP, int64_t nxzfl);
int nekcem_b200_plan_elements(int handle, int32_t *n_interior, int32_t *n_boundary);

/* Finish setup: builds vmapP / element lists, uploads them.  Replaces the setup half of
 * cem_maxwell_init that depends on connectivity (src/cem_maxwell.F:165-186). */
int nekcem_b200_setup(int handle);

/* Incident-field hook (8f rank 1; covers `userinc` of tests/3ddielectric/3ddielectric.usr:6-50,
 * called at src/cem_maxwell.F:498 after restrict_to_face): on the ninc face points
 * facepts[] (1-based indices into the (nxzf,nfaces,nelt) face arrays, the reference's
 * `incindex`) the own trace gets, in every stage,
 *     f(comp)(j) += amp[comp*ninc + q] * cos(phase[q] - omega*rktime),   comp 0..2 = H, 3..5 = E
 * before the flux is formed; the neighbour sees the modified trace exactly as through the
 * reference's gs_op_fields sum.  Call before nekcem_b200_setup; ninc = 0 removes it. */
int nekcem_b200_set_incident(int handle, int32_t ninc, const int32_t *facepts, const double *amp,
                             const double *phase, double omega);

/* Volume source hook (8f rank 1; covers tests/3dboxpml/3dboxpml.usr:30-88):
 * res(comp) -= profile(i) * (amp*sin(omega*rktime+phase) * bmn(i)) inside every stage,
 * at the reference's `usersrc` position (src/cem_maxwell.F:503).  comp: 0..2 = H, 3..5 = E. */
int nekcem_b200_set_volume_source(int handle, int comp, const double *profile, double amp,
                                  double omega, double phase);

/* Drude / Lorentz auxiliary differential equations.  Replace the per-stage calls
 * `cem_maxwell_drude(jn,kjn,resjn,params,dindex,n)` (src/cem_maxwell.F:3095-3147) and
 * `cem_maxwell_lorentz(jn,kjn,resjn,params,lindex,n)` (:3149-3211) that the reference's .usr
 * files make from `usersrc` (tests/drude/drude.usr:153-170, tests/lorentz/lorentz.usr): the
 * polarisation current is kept on the device and advanced inside the fused stage kernel at the
 * reference's position (after pml_step, before the inverse mass):
 *     Drude:   resE -= J*bm;  dJ/dt = -a*J + b*E;                    params(npts,2) = (a,b)
 *     Lorentz: resE -= J*bm;  dJ/dt = -a*J - b*P + c*E;  dP/dt = J;   params(npts,3) = (a,b,c)
 * jn,kjn: (npts,3) Drude, (npts,3,2) Lorentz (the user's COMMON arrays; NULL = zeros); index:
 * the user's 1-based node list; n = 0 removes the ADE.  Call before nekcem_b200_setup.
 * get_ade downloads jn and/or kjn (the `!$ACC UPDATE HOST` seam). */
int nekcem_b200_set_drude(int handle, const double *jn, const double *kjn, const double *params,
                          const int32_t *dindex, int32_t n);
int nekcem_b200_set_lorentz(int handle, const double *jn, const double *kjn, const double *params,
                            const int32_t *lindex, int32_t n);
int nekcem_b200_get_ade(int handle, double *jn, double *kjn);

/* Graphene sheets (SURVEY.md 8f rank 1 `userfsrc`, rank 4 surface-current ADEs).  Replace the
 * per-stage calls `cem_3d_graphene_current / cem_te_graphene_current / cem_tm_graphene_current
 * (fjn,kfjn,resfjn,params,gindex,n)` (src/cem_maxwell.F:2827-2931, 2933-3022, 3024-3093; which
 * one follows from desc.imode) that the reference's .usr files make from `userfsrc`
 * (tests/3dgraphene/3dgraphene.usr:236-266, tests/2dgraphene/2dgraphene.usr:249-290), together
 * with that userfsrc's `srcfh(c)(j) -= fjn(j,c,1)`: the sheet currents live on the device, are
 * advanced once per RK stage from the stage-start face values (after userinc), and the total
 * current is subtracted from -(n x H) on both sides of the face before the flux is formed.
 * fjn,kfjn: (nxzfl,3,6) (NULL = zeros); params: (nxzfl,12); yconduc: (nxzfl) COMMON /EMWAVE/
 * yconduc, or NULL to use the uploaded NKB_YCONDUC; gindex: the user's 1-based face-point list,
 * n entries; n = 0 removes the sheets.  Call before nekcem_b200_setup.  A sheet may lie on an
 * inter-rank face: its (tangential) current then travels folded into the packed H trace,
 * H' = H - n x f, so that the peer's n x H' reproduces the face sum.  get_graphene writes the listed face points of fjn / kfjn
 * (the `!$ACC UPDATE HOST` seam); other entries are left untouched. */
int nekcem_b200_set_graphene(int handle, const double *fjn, const double *kfjn,
                             const double *params, const double *yconduc, const int32_t *gindex,
                             int32_t n);
int nekcem_b200_get_graphene(int handle, double *fjn, double *kfjn);

/* RK tables of COMMON /RKCOEF/ (src/RK5): rk4a(5), rk4b(5), rk4c(6).  A new context holds the
 * LSRK(5,4) values of rk_storage's ifrk45 branch (src/cem_common.F:86-104); the shim uploads the
 * host's own arrays after rk_storage so that the device uses bit-identical coefficients, and so
 * that param(17) = 22 behaves as it does in the reference: rk_storage fills only rk4a(1:2),
 * rk4b(1:2) (:106-110), cem_maxwell_op_rk still runs five stages (src/cem_maxwell.F:336-340),
 * of which the last three then leave the fields unchanged (b = 0) and rktime = time (c = 0). */
int nekcem_b200_set_rk_coefficients(int handle, const double a[5], const double b[5],
                                    const double c[6]);
int nekcem_b200_get_rk_coefficients(int handle, double a[5], double b[5], double c[6]);

/* Optional modal filter (`if (iffilter) call q_filter(0.01)` at the end of
 * cem_maxwell_op_rk, src/cem_maxwell.F:342; param(18) = 1).  intv: the nx1 x nx1 matrix the
 * reference's own build_new_filter(intv,zgm1,nx1,ncut,wght,nid) returns (src/nek5_filter.F:171-249;
 * q_filter uses ncut = 2, wght = 0.01), column-major; NULL switches the filter off.  Once set,
 * every time step of nekcem_b200_step ends with filterq (src/nek5_filter.F:92-144) on ex,ey,ez,
 * hx,hy,hz on the device.  apply_filter runs it once (for callers that drive single stages). */
int nekcem_b200_set_filter(int handle, const double *intv);
int nekcem_b200_apply_filter(int handle);

/* The hot path.  Replaces `cem_maxwell_op_rk` (src/cem_maxwell.F:327-345): nsteps time
 * steps of 5 x {rk_c; cem_maxwell_op; rk_maxwell_ab} (+ q_filter when a filter is set); advances the context's time by
 * nsteps*dt like time_advancing_pde (src/cem_drive.F:618-654). */
int nekcem_b200_set_time(int handle, double time, double dt);
int nekcem_b200_get_time(int handle, double *time);
int nekcem_b200_step(int handle, int nsteps);
/* Transport of the inter-rank face exchange chosen at setup: 0 = no inter-rank faces, 1 = grouped
 * ncclSend/ncclRecv on a side stream, 2 = stores into the peers' halo buffers over NVLink (CUDA
 * IPC), fused with the pack kernel. */
int nekcem_b200_transport(int handle, int32_t *kind);

/* Number of CUDA devices visible to the process (rank -> device mapping of the shims; replaces
 * the reference's `devid = rank % 2`, src/cem_mxm_gpu.cu:430-437). */
int nekcem_b200_device_count(void);

/* The same time step for callers that keep the fields on the HOST and exchange them every step
 * (the `!$ACC UPDATE DEVICE(hn,en)` ... `!$ACC UPDATE HOST(hn,en)` seams around cem_maxwell_op_rk,
 * tests/drude/drude.usr:94, tests/3dboxper/3dboxper.usr:199) on a STREAM of inputs: each call
 * uploads one input state (hn_in, en_in: 3*npts doubles each, pinned host memory for the copies to
 * be asynchronous), advances the state the PREVIOUS call uploaded by one time step, and returns
 * the result the previous call computed (hn_out, en_out) -- so the host-to-device copy of input
 * k+1, the five fused stages of input k and the device-to-host copy of result k-1 run
 * concurrently (PCIe is full duplex; the copies use staging buffers no kernel touches, roles
 * change by pointer swaps).  NULL inputs drain the pipeline; NULL outputs discard a result.
 * A context used this way must not mix in nekcem_b200_step calls without draining first. */
int nekcem_b200_step_streamed(int handle, const double *hn_in, const double *en_in, double *hn_out,
                              double *en_out);
/* One RK stage (rkstep = 1..5) for stage-level parity tests. */
int nekcem_b200_stage(int handle, int rkstep);
int nekcem_b200_synchronize(int handle);

/* The right-hand side alone.  Replaces `call cem_maxwell_op` (src/cem_maxwell.F:484-508) for the
 * callers other than the RK loop -- the exponential and eigenvalue drivers apply it as an operator
 * (amult, src/cem_maxwell.F:2310-2365: fields in, reshn/resen out; cem_maxwell_op_exp :2158-2242,
 * cem_maxwell_op_eig :1970-1984).  One fused stage with (rk4a, rk4b, dt) = (0, 0, 1) at
 * rktime: afterwards NKB_KHN / NKB_KEN hold reshn / resen (after invqmass), the PML and ADE
 * registers hold their residuals, the fields are unchanged. */
int nekcem_b200_apply_rhs(int handle, double rktime);

/* Transport-independent stage (option "external_exchange" = 1; no communicator needed): the
 * caller performs the inter-rank face exchange that replaces gs_op_fields between ranks
 * (src/cem_maxwell.F:962) itself -- MPI from the Fortran host, or a device copy between two
 * contexts of one process.  stage_pack advances the graphene sheet currents and packs the
 * stage-start traces of the inter-rank faces into the send buffer (synchronised on return);
 * halo_buffers returns the DEVICE pointers of peer ipeer's send / receive slices (6 doubles per
 * shared face point, ordered by global face id on both sides) and their length in doubles;
 * stage_compute then runs the fused stage on all elements.  halo_exchange_local copies, for two
 * contexts of this process on one device, src's slice for dst's rank into dst's halo -- the
 * single-GPU stand-in for the NCCL exchange used by the tests. */
int nekcem_b200_stage_pack(int handle, int rkstep);
int nekcem_b200_stage_compute(int handle, int rkstep);
int nekcem_b200_halo_buffers(int handle, int32_t ipeer, double **send, double **recv,
                             int64_t *count);
int nekcem_b200_halo_exchange_local(int dst_handle, int src_handle);

/* cem_error (src/cem_common.F:1335-1355) on device: l2[c] = sqrt(sum(err*bm*err)/vol),
 * linf[c] = max|err| for the six components against exact (npts,3)+(npts,3) host arrays.
 * Sums are LOCAL to the rank (the caller reduces like glsc3/glamax). */
int nekcem_b200_error_sums(int handle, const double *exact_hn, const double *exact_en,
                           double sumsq[6], double linf[6]);

/* The same norms against an analytic solution evaluated ON THE DEVICE (device-side `usersol`,
 * SURVEY.md 8f rank 1): the separable standing modes of the shipped periodic / PEC box tests
 * (usersol of tests/3dboxper/3dboxper.usr:45-86, 3dboxpec.usr:64-115, 2dboxper.usr, 2dboxpec.usr),
 *   exact_c(x,y,z) = amp[c] * f(kind[3c+0], k[0]*x + ph[0]) * f(kind[3c+1], k[1]*y + ph[1])
 *                           * f(kind[3c+2], k[2]*z + ph[2]),   f(0,.) = 1, f(1,.) = sin, f(2,.) = cos,
 * c = 0..2 H, 3..5 E; the caller folds the time factor (tmph, tmpe of the .usr) into amp.
 * Needs NKB_XMN, NKB_YMN (and NKB_ZMN in 3D) uploaded; no host array crosses PCIe. */
int nekcem_b200_error_sums_mode(int handle, const int32_t kind[18], const double k[3],
                                const double ph[3], const double amp[6], double sumsq[6],
                                double linf[6]);

/* The same for the plane-wave solutions of the layered-media tests (usersol of
 * tests/3ddielectric/3ddielectric.usr:83-151, 2ddielectric.usr, drude.usr, lorentz.usr,
 * 3dgraphene.usr:148-222, 2dgraphene.usr): two regions of elements (region[e] = 0 / 1, e.g.
 * upper / lower half space), in each a wave travelling along y with a complex wavenumber and
 * complex per-component amplitudes, damped inside PML elements (inpml[e] != 0) by the graded
 * profile the .usr files use:
 *   exact_c = Re( amp[r][c] * exp( i*(k[r]*y - omega*time) - pml_eta[r]*pmlfac ) ),
 *   pmlfac  = pml_smax[r]*pml_d[r]/(pml_order+1) * (pml_sign[r]*(y - pml_y0[r])/pml_d[r])^(pml_order+1)
 * c = 0..2 H, 3..5 E.  The caller folds reflection / transmission coefficients, impedances and
 * the direction of travel into amp and k.  Needs NKB_YMN uploaded. */
typedef struct nekcem_b200_planewave {
    double omega;
    double k_re[2], k_im[2];
    double amp_re[2][6], amp_im[2][6];
    double pml_eta[2], pml_smax[2], pml_d[2], pml_y0[2], pml_sign[2];
    double pml_order;
} nekcem_b200_planewave;
int nekcem_b200_error_sums_planewave(int handle, const nekcem_b200_planewave *wave,
                                     const unsigned char *region, const unsigned char *inpml,
                                     double time, double sumsq[6], double linf[6]);

/* Output hand-off (the `!$ACC UPDATE HOST(HN,EN)` seam of cem_out,
 * src/io.F:193-195): the payload of one VTK "VECTORS" block exactly as the reference's writer
 * assembles it on the host -- vtk_nonswap_field interleaves the three components per node
 * (src/io_dumpvtk.F:858-878; call site src/io.F:207-208), writefield4 casts each value to float,
 * writefield4_double keeps double, both byte-swap to big-endian (src/io_co.c:443-456, 511-524).
 * which: 0 = EN ("VECTORS" block of vtkout1), 1 = HN (vtkout2); as_double: 0 = float32 (param(87)
 * != 0), 1 = float64.  out receives 3*npts values of 4 or 8 bytes, ready for
 * MPI_File_write_at_all; interleave, cast and swap run on the device, so a float dump moves
 * 12 B/node over PCIe instead of 48 and the host touches no node. */
int nekcem_b200_vtk_payload(int handle, int which, int as_double, void *out);

/* Restart hand-off, read side.  Replaces the field part of `restart_swap` (src/io.F:637-781; called
 * from cem_maxwell_init_fields when ifrestart, src/cem_maxwell.F:200-202): payload = the bytes
 * readfield4 / readfield4_double deliver for one "VECTORS" section of a restart file written by
 * cem_restart_out / cem_out (src/io.F:411-491) -- three values per node, big-endian, float32 (as_double
 * = 0, param(87) != 0) or float64 --, in this rank's element order (i.e. after the reference's
 * swap_real_backward, which stays with the file I/O on the host).  Byte swap, cast and
 * de-interleave (save2vectors, src/io.F:790-810) run on the device: which = 0 fills EN, 1 fills HN.
 * The exact inverse of nekcem_b200_vtk_payload.  The reference's restart files hold EN and HN only;
 * the rest of the state (RK registers, PML fields, ADE and sheet currents) travels through
 * nekcem_b200_get_array / set_array / get_ade / get_graphene, see tests
 * test_checkpoint_resume_full_state_bitwise. */
int nekcem_b200_restart_ingest(int handle, int which, int as_double, const void *payload);

/* Device-time of the last nekcem_b200_step call in milliseconds (CUDA events on the
 * compute stream) and the number of kernels it launched. */
int nekcem_b200_last_step_ms(int handle, float *ms, int64_t *launches);

/* Options.  "external_exchange": see nekcem_b200_stage_pack.  Performance tunables (no effect on
 * results): "pipeline" (default 1): 1 = the persistent bulk-copy stage kernel (stage_pipe.cu) for
 * the orders it covers, 0 = the slab kernel (stage_slab.cu) for every order; "pipeline_ctas"
 * (default 0 = fill the device): upper bound on the persistent kernel's grid; "p2p" (default 1, read
 * at setup, the same on every rank): inter-GPU face exchange by stores into the peers' halo
 * buffers over NVLink (CUDA IPC) instead of ncclSend/ncclRecv; "xtrace" (default 1, read at
 * setup): neighbour traces across x faces from a compact mirror of the fields.
 * "const_metrics" (default 1): exploit exact, bitwise redundancy found in the geometry at setup
 * -- elements whose nine cofactors rxmn..tzmn (src/GEOM:30-45) hold one value each over the
 * whole element read them once per element instead of once per node, and bitwise identical
 * hbm1/ebm1 (src/cem_maxwell.F:183-186 with eps = mu) share one array.  The numbers entering
 * the arithmetic are the same, so results do not change by a single bit; 0 streams every
 * array per node exactly as the reference's loops do. */
int nekcem_b200_set_option(int handle, const char *name, int value);
/* What the setup scan found: elements with constant cofactors, and whether hbm1 == ebm1. */
int nekcem_b200_geometry_info(int handle, int64_t *n_const_metric_elements, int32_t *masses_shared);

/* Algorithmic HBM bytes per stage for this context: 280 B/node +
 * 116 B/face point (+ PML add-on for PML elements). */
int nekcem_b200_algorithmic_bytes(int handle, double *bytes_per_stage);

#ifdef __cplusplus
}
#endif
#endif /* NEKCEM_B200_H */
