// libnekcem_b200: context, host-side planning (face pairing, element lists, inter-rank
// exchange plan) and the C ABI declared in include/nekcem_b200.h.
//
// Host logic restated from the reference (no code copied):
//   face pairing      = semantics of gs_setup/gs_op on gsh_face (src/jl/gs.c:1898-1907,
//                       src/nek5_connect11.F:2179-2233): equal non-zero ids are one point
//   face -> volume    = cem_set_fc_ptr (src/cem_common.F:214-283)
//   boundary handling = cem_maxwell_pec_init / flux_pec (src/cem_maxwell.F:1338-1426)
//   step driver       = cem_maxwell_op_rk + rk_c + rk_storage (src/cem_maxwell.F:327-345,
//                       src/cem_common.F:2-16, 78-114)
#include <cuda_runtime.h>
#include <nccl.h>  // types and enums only: the entry points are bound at run time (nccl_api below)
#include <dlfcn.h>
#include <type_traits>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nekcem_b200.h"
#include "stage_args.h"
#include "graphene_args.h"
#include "stage_graphene.h"

namespace nkb {
int launch_stage_slab(const StageArgs &a, const double *Dhost, int nx1, bool aux, bool cm, void *stream);
int launch_stage_pipe(const StageArgs &a, const double *Dhost, int nx1, bool aux, bool cm, void *stream);
int launch_stage_sweep(const StageArgs &a, const double *Dhost, int nx1, bool aux, bool cm, void *stream);
int launch_stage2d(const StageArgs &a, const double *Dhost, int nx1, bool aux, void *stream);
// the same kernels compiled with -fmad=false (desc.strict; 3D contexts then use the slab
// formulation at every order)
int launch_stage_slab_strict(const StageArgs &a, const double *Dhost, int nx1, bool aux, bool cm, void *stream);
int launch_stage2d_strict(const StageArgs &a, const double *Dhost, int nx1, bool aux, void *stream);
int launch_graphene(const GrapheneArgs &g, void *stream);
int launch_graphene_strict(const GrapheneArgs &g, void *stream);
}

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__,          \
                        __LINE__, cudaGetErrorString(e_));                                     \
    } while (0)
#define NCCL_OK(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess)                                                                 \
            return fail("NCCL error at %s:%d: %s", __FILE__, __LINE__, (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
    } while (0)

// one destination of pack_push_kernel (device table)
struct PushPeer {
    double *dst[2];            // the peer's halo slice for this rank, even / odd stages
    unsigned long long *flag;  // the peer's flag for this rank
    long long src_off, count;  // this rank's send slice: face points [src_off, src_off + count)
};

struct Peer {
    int rank = -1;
    std::vector<int64_t> send_fp; // local face points, sorted by shared id
    int64_t off = 0;              // offset (in face points) into sendbuf / halo
};

// NCCL is bound at run time instead of at link time.  A process may already hold another copy
// of libnccl.so.2 (PyTorch bundles its own, newer than the system's): two copies under one soname
// cannot coexist, and whichever loads first would break the other.  So the library takes the
// copy that is already loaded if there is one (RTLD_NOLOAD), else loads libnccl.so.2 itself --
// only when a communicator is actually requested (single-GPU runs never touch NCCL).
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.ok) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail("cannot load libnccl.so.2: %s", dlerror());
    bool all = true;
    auto bind = [&](auto &fp, const char *name) {
        fp = reinterpret_cast<std::remove_reference_t<decltype(fp)>>(dlsym(h, name));
        if (!fp) all = false;
    };
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
    bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommDestroy, "ncclCommDestroy");
    bind(g_nccl.AllGather, "ncclAllGather");
    bind(g_nccl.Send, "ncclSend");
    bind(g_nccl.Recv, "ncclRecv");
    bind(g_nccl.GroupStart, "ncclGroupStart");
    bind(g_nccl.GroupEnd, "ncclGroupEnd");
    bind(g_nccl.GetErrorString, "ncclGetErrorString");
    if (!all) return fail("libnccl.so.2 lacks a required entry point");
    g_nccl.ok = true;
    return 0;
}

struct Ctx {
    nekcem_b200_desc d{};
    int n = 0, nxyz = 0, nxzf = 0, nfaces = 0;
    int64_t npts = 0, nxzfl = 0, ld = 0;
    bool host_only = false;
    std::vector<double> host_copy[NKB_ARRAY_COUNT]; // host-only contexts: what was uploaded
    bool setup_done = false;
    // device arrays mirroring COMMON blocks (nullptr when not set)
    double *dev[NKB_ARRAY_COUNT] = {};
    bool have[NKB_ARRAY_COUNT] = {};
    double *u[2] = {nullptr, nullptr}; // ping-pong (6,ld): H then E
    int cur = 0;
    double *kf = nullptr; // (6,ld)
    double *hY = nullptr, *hZ = nullptr;
    std::vector<double> D_host; // dxm1: passed to the stage kernel by value (constant bank)
    int *vmapP_d = nullptr;
    int *elist_d = nullptr; // concatenated lists
    // element lists, index = 4*boundary + 2*constant-metrics + aux (aux = PML and/or ADE elements)
    int list_off[8] = {}, list_n[8] = {};
    std::vector<unsigned char> elflag_h; // host copy of elflag_d after the geometry scan
    // host planning data
    std::vector<int64_t> glo;
    std::vector<int32_t> cempec;
    std::vector<int32_t> vmapP;
    std::vector<int32_t> pml_el; // 0-based
    std::vector<std::pair<int64_t, int64_t>> singles; // (id, face point)
    bool local_matched = false, remote_planned = false;
    std::vector<Peer> peers;
    int64_t nhalo = 0;
    int n_interior = 0, n_boundary = 0;
    // comm
    ncclComm_t comm = nullptr;
    bool has_comm = false;
    cudaStream_t s_compute = nullptr, s_comm = nullptr;
    cudaEvent_t ev_stage = nullptr, ev_halo = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    double *sendbuf = nullptr, *halo = nullptr;
    int *send_node = nullptr;
    // time stepping
    double time = 0.0, dt = 0.0;
    double rk4a[5], rk4b[5], rk4c[6];
    // incident field (userinc hook)
    std::vector<int32_t> inc_fp;   // 0-based face points
    std::vector<int64_t> pairfp;   // local face point paired with each face point (-1: none)
    int *inc_own_d = nullptr, *inc_nbr_d = nullptr, *inc_send_d = nullptr;
    double *inc_amp_d = nullptr, *inc_phase_d = nullptr;
    std::vector<double> inc_amp, inc_phase;
    double inc_omega = 0.0;
    // volume source
    double *src_prof = nullptr;
    int src_comp = 0;
    double src_amp = 0, src_omega = 0, src_phase = 0;
    // diagnostics
    float last_ms = 0.f;
    int64_t last_launches = 0;
    // Drude / Lorentz ADE state (user COMMON arrays, mirrored on the device)
    int ade_kind = 0; // 0 none, 1 Drude, 2 Lorentz
    double *ade_j = nullptr, *ade_k = nullptr, *ade_par = nullptr;
    unsigned char *ade_mask = nullptr;
    unsigned char *elflag_d = nullptr; // per element: bit 0 PML, bit 1 ADE, bit 2 constant metrics
    std::vector<char> ade_el; // per element: contains ADE nodes
    std::vector<double> ade_host_j, ade_host_k; // host-only contexts: what set_drude/lorentz received
    // graphene sheets (userfsrc hook): compact per-face-point state, index q = position in the
    // user's graphindex list.  Host staging until setup, then device-resident.
    std::vector<int32_t> g_fp;              // 0-based face points
    std::vector<double> g_fj_h, g_kj_h, g_par_h, g_yc_h; // [18][ng], [18][ng], [12][ng], [ng]
    bool g_yc_given = false;
    bool g_dev_current = false; // the device copy of the sheet state is newer than the staging
    int *g_send_d = nullptr, *send_fp_d = nullptr; // per halo entry: sheet slot (-1: none), face point
    cudaEvent_t ev_sheet = nullptr;
    // leading dimensions of the caller's multi-component Fortran arrays: (lpts,k) for the ADE
    // arrays of a .usr, (lxzfl,3,6) / (lxzfl,12) for its graphene arrays (SIZE: lelt >= nelt)
    int64_t ld_pts = 0, ld_fac = 0; // 0: npts / nxzfl
    // optional modal filter at the end of every time step (q_filter, param(18) = 1)
    double *filter_d = nullptr;
    // transport-independent stepping (nekcem_b200_stage_pack / stage_compute): the caller moves
    // sendbuf -> the peers' halo itself; no communicator needed
    bool opt_external_exchange = false;
    int *g_fp_d = nullptr, *g_node_d = nullptr, *fs_own_d = nullptr, *fs_nbr_d = nullptr;
    double *g_fj = nullptr, *g_kj = nullptr, *g_par = nullptr, *g_yc = nullptr;
    // redundancy found in the geometry at setup (exact, bitwise): elements whose nine cofactors
    // do not vary over the element read them once per element; identical hbm1/ebm1 share one array
    bool opt_const_metrics = true;
    // 1 (default): the persistent bulk-copy kernel (stage_pipe.cu) for the orders it covers;
    // 0: the slab kernel (stage_slab.cu) everywhere
    bool opt_pipeline = true;
    bool opt_sweep = false; // plane-sweep kernel for nx1 = 11..16 (stage_sweep.cu)
    // 3D: mirror of the fields on the x faces (StageArgs::xtr_in); 0 = gather from the volume
    bool opt_xtrace = true;
    double *xtr[2] = {nullptr, nullptr};
    int64_t ldx = 0;
    bool xtr_valid = false; // xtr[cur] holds the traces of u[cur]
    int opt_pipeline_ctas = 0;
    // Inter-GPU face exchange over peer memory (default when every peer's buffers can be opened
    // through CUDA IPC): pack_push_kernel stores the packed traces straight into the peer's halo
    // buffer over NVLink and raises the peer's flag; no NCCL call in the time loop.
    bool opt_p2p = true, p2p_ok = false;
    double *halo2 = nullptr;                 // [2][6*nhalo]: halo of even / odd stages
    unsigned long long *flags_d = nullptr;   // [nranks]: last stage whose traces rank r delivered
    unsigned int *done_d = nullptr;          // [npeers] block counters of pack_push_kernel
    struct PushPeer *push_d = nullptr;       // [npeers]
    int *peer_rank_d = nullptr;              // [npeers]
    std::vector<void *> ipc_opened;
    unsigned long long stage_no = 0;
    // nekcem_b200_step_streamed: staging buffers (next input / previous result) and their streams
    double *st_in = nullptr, *st_out = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    bool st_have_input = false, st_have_result = false; // > 0: cap the persistent grid (tests: many items per CTA on small meshes)
    bool masses_same = false;
    bool geom_scanned = false;
    int64_t n_const_metric_el = 0;
    double *red_d = nullptr; // reduction scratch
    int red_blocks = 0;
};

std::vector<std::unique_ptr<Ctx>> g_ctx;

Ctx *get(int h)
{
    if (h < 0 || h >= (int)g_ctx.size() || !g_ctx[h]) {
        fail("invalid nekcem_b200 handle %d", h);
        return nullptr;
    }
    return g_ctx[h].get();
}

int64_t array_count(const Ctx *c, int which)
{
    switch (which) {
    case NKB_DXM1: return (int64_t)c->n * c->n;
    case NKB_W3MN: return c->nxyz;
    case NKB_RXMN: case NKB_RYMN: case NKB_RZMN: case NKB_SXMN: case NKB_SYMN: case NKB_SZMN:
    case NKB_TXMN: case NKB_TYMN: case NKB_TZMN: case NKB_BMN: case NKB_HBM1: case NKB_EBM1:
    case NKB_PERMITTIVITY: case NKB_PERMEABILITY: case NKB_XMN: case NKB_YMN: case NKB_ZMN:
        return c->npts;
    case NKB_UNXM: case NKB_UNYM: case NKB_UNZM: case NKB_AREAM:
    case NKB_Y_0: case NKB_Y_1: case NKB_Z_0: case NKB_Z_1: case NKB_YCONDUC:
        return c->nxzfl;
    case NKB_HN: case NKB_EN: case NKB_KHN: case NKB_KEN:
    case NKB_PMLSIGMA: case NKB_PMLBN: case NKB_PMLDN: case NKB_KPMLBN: case NKB_KPMLDN:
        return 3 * c->npts;
    default: return -1;
    }
}

// face point (slot s, point p) -> node inside the element, = cemface (cem_common.F:234-260)
inline int face_node(int n, int s, int p)
{
    const int n2 = n * n, pa = p % n, pb = p / n;
    switch (s) {
    case 0: return pa + n2 * pb;
    case 1: return (n - 1) + n * pa + n2 * pb;
    case 2: return pa + n * (n - 1) + n2 * pb;
    case 3: return n * pa + n2 * pb;
    case 4: return pa + n * pb;
    default: return pa + n * pb + n2 * (n - 1);
    }
}

// rk_storage, ifrk45 branch (src/cem_common.F:86-104)
void rk_storage(Ctx *c)
{
    c->rk4a[0] = 0.0;
    c->rk4a[1] = -567301805773.0 / 1357537059087.0;
    c->rk4a[2] = -2404267990393.0 / 2016746695238.0;
    c->rk4a[3] = -3550918686646.0 / 2091501179385.0;
    c->rk4a[4] = -1275806237668.0 / 842570457699.0;
    c->rk4b[0] = 1432997174477.0 / 9575080441755.0;
    c->rk4b[1] = 5161836677717.0 / 13612068292357.0;
    c->rk4b[2] = 1720146321549.0 / 2090206949498.0;
    c->rk4b[3] = 3134564353537.0 / 4481467310338.0;
    c->rk4b[4] = 2277821191437.0 / 14882151754819.0;
    c->rk4c[0] = 0.0;
    c->rk4c[1] = 1432997174477.0 / 9575080441755.0;
    c->rk4c[2] = 2526269341429.0 / 6820363962896.0;
    c->rk4c[3] = 2006345519317.0 / 3224310063776.0;
    c->rk4c[4] = 2802321613138.0 / 2924317926251.0;
    c->rk4c[5] = 1.0;
}

// ---------------------------------------------------------------------------------------
// small device kernels
// ---------------------------------------------------------------------------------------
// x-face mirror of the fields (StageArgs::xtr_in): entry t = (2e + side)*n^2 + (j + n*k)
__global__ void xtrace_fill_kernel(const double *u, long long ld, double *xtr, long long ldx, int n,
                                   long long nent)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nent) return;
    const int n2 = n * n;
    const long long es = t / n2;
    const int jk = (int)(t - es * n2);
    const long long node = (es >> 1) * (long long)n2 * n + ((es & 1) ? n - 1 : 0) + (long long)n * jk;
    for (int c = 0; c < 6; c++) xtr[c * ldx + t] = u[c * ld + node];
}

__global__ void half_inverse_kernel(const double *x, double *y, long long n)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = 0.5 / x[i];
}

// Setup scan: element e has "constant metrics" when each of its cofactor arrays holds one
// bit pattern over the whole element (affine elements whose metrics were generated without
// round-off noise).  The stage kernel then reads 9 values per element instead of 9 per node --
// the same numbers, so the results are unchanged bit for bit.
__global__ void metric_const_kernel(const double *m0, const double *m1, const double *m2,
                                    const double *m3, const double *m4, const double *m5,
                                    const double *m6, const double *m7, const double *m8,
                                    int nxyz, unsigned char *elflag, unsigned long long *count)
{
    const double *met[9] = {m0, m1, m2, m3, m4, m5, m6, m7, m8};
    const long long base = (long long)blockIdx.x * nxyz;
    int same = 1;
    for (int q = 0; q < 9; q++) {
        const long long ref = __double_as_longlong(met[q][base]);
        for (int i = threadIdx.x; i < nxyz; i += blockDim.x)
            same &= (__double_as_longlong(met[q][base + i]) == ref);
    }
    same = __syncthreads_and(same);
    if (threadIdx.x == 0) {
        unsigned char f = elflag[blockIdx.x] & (unsigned char)~4;
        if (same) { f |= 4; atomicAdd(count, 1ull); }
        elflag[blockIdx.x] = f;
    }
}

__global__ void arrays_differ_kernel(const double *a, const double *b, long long n, int *differ)
{
    int d = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        d |= (__double_as_longlong(a[i]) != __double_as_longlong(b[i]));
    if (__syncthreads_or(d) && threadIdx.x == 0) *differ = 1;
}

// pack the traces of the send face points: sendbuf[q][c] = u[c][send_node[q]]
// (+ the incident field of tagged send face points, so that the peer sees the same trace the
// reference's gs_op_fields sum would give it)
// packed trace value t = 6*q + c of the send list (what the peer's face sum needs from this side)
__device__ __forceinline__ double packed_trace(long long t, const double *u, long long ld,
                                               const int *send_node, const int *inc_send,
                                               const double *inc_amp, const double *inc_phase,
                                               int inc_n, double inc_wt, const int *g_send,
                                               const int *send_fp, const double *fs_val, int fs_n,
                                               const double *unx, const double *uny, const double *unz)
{
    const long long q = t / 6;
    const int c = (int)(t - q * 6);
    double v = u[c * ld + send_node[q]];
    if (inc_send != nullptr) {
        const int qi = inc_send[q];
        if (qi >= 0) v += inc_amp[c * inc_n + qi] * cos(inc_phase[qi] - inc_wt);
    }
    if (g_send != nullptr && c < 3) {
        // graphene sheet on an inter-rank face: the peer must see -(n+ x H+) - f+ in its face sum
        // (userfsrc precedes gs_op_fields, src/cem_maxwell.F:958-962).  It forms the sum from the
        // received trace as n x H (n = -n+), so the tangential sheet current f+ travels folded
        // into the H trace: H' = H+ - n+ x f+, because n x (n x f) = -f for tangential f.
        const int gq = g_send[q];
        if (gq >= 0) {
            const int jf = send_fp[q];
            const double nx = unx[jf], ny = uny[jf], nz = unz ? unz[jf] : 0.0;
            const double fx = fs_val[gq], fy = fs_val[fs_n + gq], fz = fs_val[2 * fs_n + gq];
            const double d = c == 0 ? ny * fz - nz * fy : (c == 1 ? nz * fx - nx * fz : nx * fy - ny * fx);
            v -= d;
        }
    }
    return v;
}

__global__ void pack_kernel(const double *u, long long ld, const int *send_node, double *sendbuf,
                            long long nsend, const int *inc_send, const double *inc_amp,
                            const double *inc_phase, int inc_n, double inc_wt,
                            const int *g_send, const int *send_fp, const double *fs_val, int fs_n,
                            const double *unx, const double *uny, const double *unz)
{
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nsend * 6) return;
    sendbuf[t] = packed_trace(t, u, ld, send_node, inc_send, inc_amp, inc_phase, inc_n, inc_wt,
                              g_send, send_fp, fs_val, fs_n, unx, uny, unz);
}

// The exchange that replaces gs_op_fields between ranks (src/cem_maxwell.F:962), fused with the
// pack: blockIdx.y = peer; every value goes straight into the peer GPU's halo buffer (a store over
// NVLink into memory opened through CUDA IPC), and the last block of a peer's slice publishes the
// stage number in the peer's flag after a system-scope fence.  The peer's boundary-element launch
// is ordered behind wait_flags_kernel, which spins on those flags.
__global__ void pack_push_kernel(const double *u, long long ld, const int *send_node,
                                 const PushPeer *tab, int parity, unsigned long long stage_no,
                                 int me, unsigned int *done, const int *inc_send,
                                 const double *inc_amp, const double *inc_phase, int inc_n,
                                 double inc_wt, const int *g_send, const int *send_fp,
                                 const double *fs_val, int fs_n, const double *unx,
                                 const double *uny, const double *unz)
{
    const PushPeer pp = tab[blockIdx.y];
    const long long tot = 6 * pp.count;
    const unsigned int nb = (unsigned int)((tot + blockDim.x - 1) / blockDim.x);
    if (blockIdx.x >= nb) return;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < tot)
        pp.dst[parity][t] = packed_trace(6 * pp.src_off + t, u, ld, send_node, inc_send, inc_amp,
                                         inc_phase, inc_n, inc_wt, g_send, send_fp, fs_val, fs_n,
                                         unx, uny, unz);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done + blockIdx.y, 1u);
        if (prev == nb - 1) {
            done[blockIdx.y] = 0;
            __threadfence_system();
            *(volatile unsigned long long *)(pp.flag + me) = stage_no;
        }
    }
}

__global__ void wait_flags_kernel(const unsigned long long *flags, const int *peer_rank, int npeers,
                                  unsigned long long stage_no)
{
    if ((int)threadIdx.x < npeers) {
        const volatile unsigned long long *f = flags + peer_rank[threadIdx.x];
        while (*f < stage_no) { }
    }
    __threadfence_system();
}

// (graphene_kernel: graphene.cu)

// cem_error partial sums (src/cem_common.F:1335-1355): per block, per component
__global__ void error_kernel(const double *u, long long ld, const double *exact, long long npts,
                             const double *bm, double *part /* [blocks][12] */)
{
    __shared__ double ssum[256], smax[256];
    for (int c = 0; c < 6; c++) {
        double sum = 0.0, mx = 0.0;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npts;
             i += (long long)gridDim.x * blockDim.x) {
            double err = exact[c * npts + i] - u[c * ld + i];
            sum += err * bm[i] * err;
            mx = fmax(mx, fabs(err));
        }
        ssum[threadIdx.x] = sum;
        smax[threadIdx.x] = mx;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                ssum[threadIdx.x] += ssum[threadIdx.x + s];
                smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[blockIdx.x * 12 + c] = ssum[0];
            part[blockIdx.x * 12 + 6 + c] = smax[0];
        }
        __syncthreads();
    }
}

// cem_error against a separable standing mode evaluated on the fly (device-side usersol)
struct ModeSol {
    int kind[18];
    double k[3], ph[3], amp[6];
};
__device__ __forceinline__ double mode_f(int kind, double a)
{
    return kind == 0 ? 1.0 : (kind == 1 ? sin(a) : cos(a));
}
__global__ void error_mode_kernel(const double *u, long long ld, ModeSol m, const double *x,
                                  const double *y, const double *z, long long npts,
                                  const double *bm, double *part /* [blocks][12] */)
{
    __shared__ double ssum[256], smax[256];
    for (int c = 0; c < 6; c++) {
        double sum = 0.0, mx = 0.0;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npts;
             i += (long long)gridDim.x * blockDim.x) {
            const double zz = z ? z[i] : 0.0;
            const double ex = m.amp[c] * mode_f(m.kind[3 * c], m.k[0] * x[i] + m.ph[0]) *
                              mode_f(m.kind[3 * c + 1], m.k[1] * y[i] + m.ph[1]) *
                              mode_f(m.kind[3 * c + 2], m.k[2] * zz + m.ph[2]);
            double err = ex - u[c * ld + i];
            sum += err * bm[i] * err;
            mx = fmax(mx, fabs(err));
        }
        ssum[threadIdx.x] = sum;
        smax[threadIdx.x] = mx;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                ssum[threadIdx.x] += ssum[threadIdx.x + s];
                smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[blockIdx.x * 12 + c] = ssum[0];
            part[blockIdx.x * 12 + 6 + c] = smax[0];
        }
        __syncthreads();
    }
}

// q_filter / filterq (src/nek5_filter.F:2-144): v <- (F (x) F (x) F) v for one component of one
// element per CTA; the three contractions in the order of the reference's mxm calls (r, then s,
// then t; contracted index ascending), ping-ponging between two shared-memory copies of the
// element.  F: the n x n matrix of build_new_filter, column-major.  2D: two contractions.
__global__ void filter_kernel(double *u, long long ld, const double *F, int n, int nz, int nxyz)
{
    extern __shared__ double fsm[];
    double *A = fsm, *B = fsm + nxyz, *Fs = fsm + 2 * nxyz;
    const int e = blockIdx.x, c = blockIdx.y;
    double *v = u + (long long)c * ld + (long long)e * nxyz;
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) Fs[q] = F[q];
    for (int q = threadIdx.x; q < nxyz; q += blockDim.x) A[q] = v[q];
    __syncthreads();
    const int n2 = n * n;
    // r: B(i,j,k) = sum_m F(i,m) A(m,j,k)
    for (int q = threadIdx.x; q < nxyz; q += blockDim.x) {
        const int i = q % n, jk = q / n;
        double sum = Fs[i] * A[n * jk];
        for (int m = 1; m < n; m++) sum = sum + Fs[i + n * m] * A[m + n * jk];
        B[q] = sum;
    }
    __syncthreads();
    // s: A(i,j,k) = sum_m B(i,m,k) F(j,m)
    for (int q = threadIdx.x; q < nxyz; q += blockDim.x) {
        const int i = q % n, j = (q / n) % n, k = q / n2;
        const double *b = B + n2 * k;
        double sum = b[i] * Fs[j];
        for (int m = 1; m < n; m++) sum = sum + b[i + n * m] * Fs[j + n * m];
        A[q] = sum;
    }
    __syncthreads();
    if (nz > 1) {
        // t: v(i,j,k) = sum_m A(i,j,m) F(k,m)
        for (int q = threadIdx.x; q < nxyz; q += blockDim.x) {
            const int ij = q % n2, k = q / n2;
            double sum = A[ij] * Fs[k];
            for (int m = 1; m < n; m++) sum = sum + A[ij + n2 * m] * Fs[k + n * m];
            v[q] = sum;
        }
    } else {
        for (int q = threadIdx.x; q < nxyz; q += blockDim.x) v[q] = A[q];
    }
}

// Output hand-off (cem_out): the payload of one VTK "VECTORS" block exactly as the reference's
// writer assembles it on the host -- vtk_nonswap_field interleaves the three components per node
// (src/io_dumpvtk.F:858-878), writefield4 / writefield4_double cast each value to float (or keep
// double) and byte-swap it to big-endian (src/io_co.c:443-456, 511-524, src/io_util.c:99-146).
// One thread per (node, component); 4- or 8-byte stores are coalesced along the output.
template <typename OUT>
__global__ void vtk_payload_kernel(const double *u, long long ld, long long npts, OUT *out)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * npts) return;
    const long long i = t / 3;
    const int c = (int)(t - 3 * i);
    const double v = u[c * ld + i];
    if constexpr (sizeof(OUT) == 4) {
        const unsigned int b = __float_as_uint(__double2float_rn(v));
        out[t] = __byte_perm(b, 0, 0x0123);
    } else {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        const unsigned int lo = (unsigned int)b, hi = (unsigned int)(b >> 32);
        out[t] = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    }
}

// The inverse of vtk_payload_kernel: big-endian float32 / float64 triples of a restart file's field
// section -> the three components of EN or HN (readfield4[_double] + save2vectors,
// src/io.F:764-775, 790-810).  A float32 restart recovers 7-8 digits, as the reference notes (:664).
template <typename IN>
__global__ void restart_ingest_kernel(const IN *in, double *u, long long ld, long long npts)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * npts) return;
    const long long i = t / 3;
    const int c = (int)(t - 3 * i);
    double v;
    if constexpr (sizeof(IN) == 4) {
        v = (double)__uint_as_float(__byte_perm(in[t], 0, 0x0123));
    } else {
        const unsigned long long b = in[t];
        const unsigned int lo = (unsigned int)b, hi = (unsigned int)(b >> 32);
        v = __longlong_as_double((long long)(((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) |
                                             __byte_perm(hi, 0, 0x0123)));
    }
    u[c * ld + i] = v;
}

// cem_error against a plane wave in two half spaces with a graded PML decay (device-side usersol
// of the layered-media tests): exact_c = Re( amp[r][c] * exp(i (k_r y - omega t) - eta_r pmlfac) )
__global__ void error_planewave_kernel(const double *u, long long ld, nekcem_b200_planewave w,
                                       double wt, const unsigned char *region,
                                       const unsigned char *inpml, int nxyz, const double *y,
                                       long long npts, const double *bm,
                                       double *part /* [blocks][12] */)
{
    __shared__ double ssum[256], smax[256];
    double sum[6] = {0, 0, 0, 0, 0, 0}, mx[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npts;
         i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / nxyz;
        const int r = region[e] ? 1 : 0;
        const double yy = y[i];
        double pmlfac = 0.0;
        if (inpml[e]) {
            const double d = w.pml_d[r];
            pmlfac = (w.pml_smax[r] * d / (w.pml_order + 1.0)) *
                     pow(w.pml_sign[r] * (yy - w.pml_y0[r]) / d, w.pml_order + 1.0);
        }
        const double th = w.k_re[r] * yy - wt;
        const double mag = exp(-w.k_im[r] * yy - w.pml_eta[r] * pmlfac);
        const double cr = mag * cos(th), ci = mag * sin(th);
        const double b = bm[i];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const double ex = w.amp_re[r][c] * cr - w.amp_im[r][c] * ci;
            const double err = ex - u[c * ld + i];
            sum[c] += err * b * err;
            mx[c] = fmax(mx[c], fabs(err));
        }
    }
    for (int c = 0; c < 6; c++) {
        ssum[threadIdx.x] = sum[c];
        smax[threadIdx.x] = mx[c];
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                ssum[threadIdx.x] += ssum[threadIdx.x + s];
                smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[blockIdx.x * 12 + c] = ssum[0];
            part[blockIdx.x * 12 + 6 + c] = smax[0];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// host planning
// ---------------------------------------------------------------------------------------
int match_local(Ctx *c)
{
    if (c->glo.empty()) return fail("nekcem_b200_set_faces has not been called");
    const int64_t nf = c->nxzfl;
    const int n = c->n, nfp = c->nxzf * c->nfaces;
    c->vmapP.assign(nf, -2);
    c->pairfp.assign(nf, -1);
    std::vector<std::pair<int64_t, int64_t>> v;
    v.reserve(nf);
    for (int64_t j = 0; j < nf; j++)
        if (c->glo[j] != 0) v.emplace_back(c->glo[j], j);
    std::sort(v.begin(), v.end());
    c->singles.clear();
    for (size_t a = 0; a < v.size();) {
        size_t b = a;
        while (b < v.size() && v[b].first == v[a].first) b++;
        if (b - a == 2) {
            const int64_t j0 = v[a].second, j1 = v[a + 1].second;
            const int64_t e0 = j0 / nfp, e1 = j1 / nfp;
            const int f0 = (int)(j0 - e0 * nfp), f1 = (int)(j1 - e1 * nfp);
            const int64_t n0 = e0 * c->nxyz + face_node(n, f0 / c->nxzf, f0 % c->nxzf);
            const int64_t n1 = e1 * c->nxyz + face_node(n, f1 / c->nxzf, f1 % c->nxzf);
            c->vmapP[j0] = (int32_t)n1;
            c->vmapP[j1] = (int32_t)n0;
            c->pairfp[j0] = j1;
            c->pairfp[j1] = j0;
        } else if (b - a == 1) {
            c->singles.emplace_back(v[a].first, v[a].second);
        } else {
            return fail("face id %lld is shared by %zu local face points (expected <= 2)",
                        (long long)v[a].first, b - a);
        }
        a = b;
    }
    // physical boundaries: PEC-like faces get the mirror code (-1); refined later for ids
    // that turn out to be shared with another rank
    for (int32_t q : c->cempec) {
        if (q < 1 || q > nf) return fail("cempec entry %d out of range 1..%lld", q, (long long)nf);
        if (c->vmapP[q - 1] == -2) c->vmapP[q - 1] = -1;
    }
    c->local_matched = true;
    return 0;
}

int plan_remote(Ctx *c, const int64_t *counts, const int64_t *all_ids)
{
    if (!c->local_matched) return fail("face_remote called before local matching");
    const int R = c->d.nranks, me = c->d.rank;
    std::map<int64_t, int64_t> mine; // id -> face point
    for (auto &s : c->singles) mine[s.first] = s.second;
    c->peers.clear();
    int64_t off = 0, pos = 0;
    for (int r = 0; r < R; r++) {
        const int64_t cnt = counts[r];
        if (r != me) {
            std::vector<std::pair<int64_t, int64_t>> shared;
            for (int64_t q = 0; q < cnt; q++) {
                auto it = mine.find(all_ids[pos + q]);
                if (it != mine.end()) shared.emplace_back(it->first, it->second);
            }
            if (!shared.empty()) {
                std::sort(shared.begin(), shared.end());
                Peer p;
                p.rank = r;
                p.off = off;
                for (auto &s : shared) p.send_fp.push_back(s.second);
                off += (int64_t)shared.size();
                c->peers.push_back(std::move(p));
            }
        }
        pos += cnt;
    }
    c->nhalo = off;
    for (auto &p : c->peers)
        for (size_t q = 0; q < p.send_fp.size(); q++) {
            const int64_t slot = p.off + (int64_t)q;
            if (slot + 3 > 2147483647LL) return fail("halo too large");
            c->vmapP[p.send_fp[q]] = (int32_t)(-(slot + 3));
        }
    c->remote_planned = true;
    return 0;
}

int build_lists(Ctx *c, std::vector<int32_t> &lists)
{
    const int nfp = c->nxzf * c->nfaces;
    std::vector<char> is_b(c->d.nelt, 0), is_pml(c->d.nelt, 0);
    for (auto &p : c->peers)
        for (int64_t fp : p.send_fp) is_b[fp / nfp] = 1;
    for (int32_t e : c->pml_el) is_pml[e] = 1;
    if (c->ade_kind)
        for (int e = 0; e < c->d.nelt; e++)
            if (c->ade_el[e]) is_pml[e] = 1;
    // elements holding a graphene face point, or paired with one, take the AUX instantiation
    // (the only one that carries the face-source code)
    for (int32_t fp : c->g_fp) {
        is_pml[fp / nfp] = 1;
        if (c->pairfp[fp] >= 0) is_pml[c->pairfp[fp] / nfp] = 1;
    }
    const bool have_flags = (int)c->elflag_h.size() == c->d.nelt;
    std::vector<int32_t> L[8];
    for (int e = 0; e < c->d.nelt; e++) {
        const int cm = have_flags && (c->elflag_h[e] & 4) ? 2 : 0;
        L[(is_b[e] ? 4 : 0) + cm + (is_pml[e] ? 1 : 0)].push_back(e);
    }
    lists.clear();
    for (int q = 0; q < 8; q++) {
        c->list_off[q] = (int)lists.size();
        c->list_n[q] = (int)L[q].size();
        lists.insert(lists.end(), L[q].begin(), L[q].end());
    }
    c->n_interior = c->list_n[0] + c->list_n[1] + c->list_n[2] + c->list_n[3];
    c->n_boundary = c->list_n[4] + c->list_n[5] + c->list_n[6] + c->list_n[7];
    return 0;
}

int exchange_singletons_nccl(Ctx *c)
{
    const int R = c->d.nranks;
    // counts
    int64_t mycnt = (int64_t)c->singles.size();
    int64_t *d_cnt = nullptr;
    CUDA_OK(cudaMalloc(&d_cnt, sizeof(int64_t) * (R + 1)));
    CUDA_OK(cudaMemcpy(d_cnt + R, &mycnt, sizeof(int64_t), cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllGather(d_cnt + R, d_cnt, 1, ncclInt64, c->comm, c->s_comm));
    CUDA_OK(cudaStreamSynchronize(c->s_comm));
    std::vector<int64_t> counts(R);
    CUDA_OK(cudaMemcpy(counts.data(), d_cnt, sizeof(int64_t) * R, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(d_cnt));
    int64_t mx = 0, tot = 0;
    for (auto v : counts) {
        mx = std::max(mx, v);
        tot += v;
    }
    if (mx == 0) {
        std::vector<int64_t> none;
        return plan_remote(c, counts.data(), none.data());
    }
    int64_t *d_ids = nullptr;
    CUDA_OK(cudaMalloc(&d_ids, sizeof(int64_t) * mx * (R + 1)));
    std::vector<int64_t> ids(mx, 0);
    for (size_t q = 0; q < c->singles.size(); q++) ids[q] = c->singles[q].first;
    CUDA_OK(cudaMemcpy(d_ids + mx * R, ids.data(), sizeof(int64_t) * mx, cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllGather(d_ids + mx * R, d_ids, mx, ncclInt64, c->comm, c->s_comm));
    CUDA_OK(cudaStreamSynchronize(c->s_comm));
    std::vector<int64_t> padded(mx * R), all;
    CUDA_OK(cudaMemcpy(padded.data(), d_ids, sizeof(int64_t) * mx * R, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(d_ids));
    all.reserve(tot);
    for (int r = 0; r < R; r++)
        all.insert(all.end(), padded.begin() + r * mx, padded.begin() + r * mx + counts[r]);
    return plan_remote(c, counts.data(), all.data());
}

// Peer-memory transport of the face exchange: every rank publishes CUDA IPC handles of its halo
// and flag buffers plus the offsets at which it expects each peer's traces (one NCCL all-gather at
// setup); afterwards the time loop needs no NCCL call.  All ranks agree on the outcome: if any
// rank cannot open a peer's buffers, everybody keeps the NCCL send/recv path.
void release_p2p(Ctx *c)
{
    for (void *q : c->ipc_opened) cudaIpcCloseMemHandle(q);
    c->ipc_opened.clear();
    cudaFree(c->halo2); cudaFree(c->flags_d); cudaFree(c->done_d); cudaFree(c->push_d);
    cudaFree(c->peer_rank_d);
    c->halo2 = nullptr; c->flags_d = nullptr; c->done_d = nullptr; c->push_d = nullptr;
    c->peer_rank_d = nullptr;
    c->p2p_ok = false;
}

int setup_p2p(Ctx *c)
{
    release_p2p(c);
    if (!c->has_comm || c->d.nranks < 2 || c->opt_external_exchange) return 0;
    const int R = c->d.nranks, me = c->d.rank;
    const size_t nh = (size_t)std::max<int64_t>(c->nhalo, 1);
    int64_t ok = c->opt_p2p ? 1 : 0;
    CUDA_OK(cudaMalloc(&c->halo2, sizeof(double) * 2 * 6 * nh));
    CUDA_OK(cudaMemset(c->halo2, 0, sizeof(double) * 2 * 6 * nh));
    CUDA_OK(cudaMalloc(&c->flags_d, sizeof(unsigned long long) * R));
    CUDA_OK(cudaMemset(c->flags_d, 0, sizeof(unsigned long long) * R));
    cudaIpcMemHandle_t hh{}, hf{};
    if (cudaIpcGetMemHandle(&hh, c->halo2) != cudaSuccess ||
        cudaIpcGetMemHandle(&hf, c->flags_d) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    // blob: ok, nhalo, halo handle, flag handle, offsets of every rank's traces in my halo (-1: none)
    const size_t B = 16 + 2 * sizeof(cudaIpcMemHandle_t) + 8 * (size_t)R;
    std::vector<char> mine(B, 0), all(B * R);
    int64_t nhalo64 = c->nhalo;
    memcpy(mine.data(), &ok, 8);
    memcpy(mine.data() + 8, &nhalo64, 8);
    memcpy(mine.data() + 16, &hh, sizeof(hh));
    memcpy(mine.data() + 16 + sizeof(hh), &hf, sizeof(hf));
    std::vector<int64_t> offs(R, -1);
    for (auto &p : c->peers) offs[p.rank] = p.off;
    memcpy(mine.data() + 16 + 2 * sizeof(hh), offs.data(), 8 * (size_t)R);
    char *d_all = nullptr;
    CUDA_OK(cudaMalloc(&d_all, B * (R + 1)));
    CUDA_OK(cudaMemcpy(d_all + B * R, mine.data(), B, cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllGather(d_all + B * R, d_all, B, ncclInt8, c->comm, c->s_comm));
    CUDA_OK(cudaStreamSynchronize(c->s_comm));
    CUDA_OK(cudaMemcpy(all.data(), d_all, B * R, cudaMemcpyDeviceToHost));
    std::vector<PushPeer> tab;
    std::vector<int> pranks;
    for (auto &p : c->peers) {
        const char *rb = all.data() + B * (size_t)p.rank;
        int64_t rok = 0, rnh = 0, roff = -1;
        memcpy(&rok, rb, 8);
        memcpy(&rnh, rb + 8, 8);
        memcpy(&roff, rb + 16 + 2 * sizeof(hh) + 8 * (size_t)me, 8);
        if (!ok || !rok || roff < 0) { ok = 0; break; }
        cudaIpcMemHandle_t rhh, rhf;
        memcpy(&rhh, rb + 16, sizeof(rhh));
        memcpy(&rhf, rb + 16 + sizeof(rhh), sizeof(rhf));
        void *ph = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&ph, rhh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); ok = 0; break;
        }
        c->ipc_opened.push_back(ph);
        if (cudaIpcOpenMemHandle(&pf, rhf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); ok = 0; break;
        }
        c->ipc_opened.push_back(pf);
        PushPeer pp;
        pp.dst[0] = (double *)ph + 6 * roff;
        pp.dst[1] = (double *)ph + 6 * (size_t)std::max<int64_t>(rnh, 1) + 6 * roff;
        pp.flag = (unsigned long long *)pf;
        pp.src_off = p.off;
        pp.count = (long long)p.send_fp.size();
        tab.push_back(pp);
        pranks.push_back(p.rank);
    }
    // everybody must take the same path: all-gather the outcome
    CUDA_OK(cudaMemcpy(d_all + 8 * R, &ok, 8, cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllGather(d_all + 8 * R, d_all, 1, ncclInt64, c->comm, c->s_comm));
    CUDA_OK(cudaStreamSynchronize(c->s_comm));
    std::vector<int64_t> oks(R);
    CUDA_OK(cudaMemcpy(oks.data(), d_all, 8 * (size_t)R, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(d_all));
    for (auto v : oks) if (!v) ok = 0;
    if (!ok) {
        release_p2p(c);
        return 0;
    }
    if (!tab.empty()) {
        CUDA_OK(cudaMalloc(&c->push_d, sizeof(PushPeer) * tab.size()));
        CUDA_OK(cudaMemcpy(c->push_d, tab.data(), sizeof(PushPeer) * tab.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->peer_rank_d, sizeof(int) * pranks.size()));
        CUDA_OK(cudaMemcpy(c->peer_rank_d, pranks.data(), sizeof(int) * pranks.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->done_d, sizeof(unsigned int) * tab.size()));
        CUDA_OK(cudaMemset(c->done_d, 0, sizeof(unsigned int) * tab.size()));
    }
    c->stage_no = 0;
    c->p2p_ok = true;
    return 0;
}

int require(Ctx *c, std::initializer_list<int> ids)
{
    static const char *names[NKB_ARRAY_COUNT] = {
        "dxm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn", "txmn", "tymn", "tzmn",
        "bmn", "hbm1", "ebm1", "unxm", "unym", "unzm", "aream", "Y_0", "Y_1", "Z_0", "Z_1", "hn",
        "en", "khn", "ken", "permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn",
        "kpmlbn", "kpmldn", "xmn", "ymn", "zmn"};
    for (int id : ids)
        if (!c->have[id]) return fail("array '%s' has not been uploaded", names[id]);
    return 0;
}

int ensure_dev(Ctx *c, int which)
{
    if (c->dev[which]) return 0;
    const int64_t cnt = array_count(c, which);
    CUDA_OK(cudaMalloc(&c->dev[which], sizeof(double) * cnt));
    CUDA_OK(cudaMemsetAsync(c->dev[which], 0, sizeof(double) * cnt, c->s_compute));
    return 0;
}

// (re)scan the geometry for exact redundancy; called from setup and lazily after a geometry
// array was replaced
int scan_geometry(Ctx *c)
{
    c->geom_scanned = true;
    c->masses_same = false;
    c->n_const_metric_el = 0;
    if (!c->elflag_d) return 0;
    unsigned long long *cnt = nullptr;
    int *differ = nullptr;
    CUDA_OK(cudaMalloc(&cnt, sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc(&differ, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->s_compute));
    CUDA_OK(cudaMemsetAsync(differ, 0, sizeof(int), c->s_compute));
    if (c->d.ldim == 3 && c->opt_const_metrics) {
        const double *m[9];
        for (int q = 0; q < 9; q++) m[q] = c->dev[NKB_RXMN + q];
        metric_const_kernel<<<c->d.nelt, 128, 0, c->s_compute>>>(
            m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], c->nxyz, c->elflag_d, cnt);
    } else if (c->d.ldim == 3) {
        // option switched off: clear bit 2
        std::vector<unsigned char> ef(c->d.nelt);
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
        CUDA_OK(cudaMemcpy(ef.data(), c->elflag_d, ef.size(), cudaMemcpyDeviceToHost));
        for (auto &f : ef) f &= (unsigned char)~4;
        CUDA_OK(cudaMemcpy(c->elflag_d, ef.data(), ef.size(), cudaMemcpyHostToDevice));
    }
    if (c->opt_const_metrics)
        arrays_differ_kernel<<<1184, 256, 0, c->s_compute>>>(c->dev[NKB_HBM1], c->dev[NKB_EBM1],
                                                            c->npts, differ);
    unsigned long long hc = 0;
    int hd = 1;
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    CUDA_OK(cudaMemcpy(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(&hd, differ, sizeof(hd), cudaMemcpyDeviceToHost));
    cudaFree(cnt); cudaFree(differ);
    c->n_const_metric_el = (int64_t)hc;
    c->masses_same = c->opt_const_metrics && hd == 0;
    // the element lists depend on the flags: rebuild and upload them
    c->elflag_h.resize(c->d.nelt);
    CUDA_OK(cudaMemcpy(c->elflag_h.data(), c->elflag_d, c->elflag_h.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> lists;
    build_lists(c, lists);
    cudaFree(c->elist_d);
    c->elist_d = nullptr;
    CUDA_OK(cudaMalloc(&c->elist_d, sizeof(int) * std::max<size_t>(lists.size(), 1)));
    CUDA_OK(cudaMemcpy(c->elist_d, lists.data(), sizeof(int) * lists.size(), cudaMemcpyHostToDevice));
    return 0;
}

// phase 0: the whole stage (NCCL exchange between ranks); phase 1: sheet currents + pack of the
// send buffer only; phase 2: the element launches only (halo filled by the caller)
int run_stage(Ctx *c, int rkstep /*1..5*/, int phase = 0)
{
    if (!c->geom_scanned && scan_geometry(c)) return 1;
    nkb::StageArgs a{};
    a.u_in = c->u[c->cur];
    a.u_out = c->u[c->cur ^ 1];
    a.kf = c->kf;
    a.ld = c->ld;
    for (int q = 0; q < 9; q++) a.met[q] = c->dev[NKB_RXMN + q];
    a.hbm1 = c->dev[NKB_HBM1]; a.bmn = c->dev[NKB_BMN];
    // bitwise identical inverse masses (eps = mu): both reads hit the same lines
    a.ebm1 = c->masses_same ? c->dev[NKB_HBM1] : c->dev[NKB_EBM1];
    a.D = c->dev[NKB_DXM1];
    a.w3 = c->dev[NKB_W3MN];
    a.unx = c->dev[NKB_UNXM]; a.uny = c->dev[NKB_UNYM]; a.unz = c->dev[NKB_UNZM];
    a.area = c->dev[NKB_AREAM];
    a.hY = c->hY; a.Y1 = c->dev[NKB_Y_1]; a.hZ = c->hZ; a.Z1 = c->dev[NKB_Z_1];
    a.vmapP = c->vmapP_d;
    a.halo = c->halo;
    if (c->xtr[0]) {
        if (!c->xtr_valid) {
            const long long nent = 2ll * c->n * c->n * c->d.nelt;
            xtrace_fill_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, c->s_compute>>>(
                c->u[c->cur], c->ld, c->xtr[c->cur], c->ldx, c->n, nent);
            CUDA_OK(cudaGetLastError());
            c->xtr_valid = true;
        }
        a.xtr_in = c->xtr[c->cur];
        a.xtr_out = c->xtr[c->cur ^ 1];
        a.ldx = c->ldx;
    }
    a.grid_cap = c->opt_pipeline_ctas;
    a.ca = c->rk4a[rkstep - 1];
    a.cb = c->rk4b[rkstep - 1];
    a.dt = c->dt;
    a.C0 = c->d.ifupwind ? 1.0 : 0.0;
    a.sig = c->dev[NKB_PMLSIGMA]; a.eps = c->dev[NKB_PERMITTIVITY]; a.mu = c->dev[NKB_PERMEABILITY];
    a.pB = c->dev[NKB_PMLBN]; a.pD = c->dev[NKB_PMLDN];
    a.kB = c->dev[NKB_KPMLBN]; a.kD = c->dev[NKB_KPMLDN];
    a.npts = c->npts;
    a.elflag = c->elflag_d;
    a.ade_kind = c->ade_kind;
    a.ade_j = c->ade_j; a.ade_k = c->ade_k; a.ade_par = c->ade_par; a.ade_mask = c->ade_mask;
    a.imode = c->d.imode;
    a.src_prof = c->src_prof;
    a.src_comp = c->src_comp;
    // rk_c (src/cem_common.F:12): rktime = time + dt*rk4c(i)
    const double rktime = c->time + c->dt * c->rk4c[rkstep - 1];
    a.src_tfac = c->src_prof ? c->src_amp * sin(c->src_omega * rktime + c->src_phase) : 0.0;
    a.inc_own = c->inc_own_d; a.inc_nbr = c->inc_nbr_d;
    a.inc_amp = c->inc_amp_d; a.inc_phase = c->inc_phase_d;
    a.inc_n = (int)c->inc_fp.size();
    a.inc_wt = c->inc_omega * rktime;
    a.fs_own = c->fs_own_d; a.fs_nbr = c->fs_nbr_d;
    a.fs_val = c->g_fj; // slot m = 0: fjn(:,1:3,1)
    a.fs_n = (int)c->g_fp.size();
    if (!c->g_fp.empty() && phase != 2) {
        // userfsrc -> cem_*_graphene_current: needs only stage-start data, so it runs first on
        // the compute stream; every stage launch of this stage is ordered after it
        nkb::GrapheneArgs g{};
        g.u = a.u_in; g.ld = c->ld;
        g.ng = a.fs_n; g.imode = c->d.imode;
        g.fp = c->g_fp_d; g.node = c->g_node_d;
        g.unx = a.unx; g.uny = a.uny; g.unz = c->d.ldim == 3 ? a.unz : nullptr;
        g.hY = c->hY; g.yc = c->g_yc; g.par = c->g_par;
        g.fj = c->g_fj; g.kj = c->g_kj;
        g.inc_own = c->inc_own_d; g.inc_amp = c->inc_amp_d; g.inc_phase = c->inc_phase_d;
        g.inc_n = a.inc_n; g.inc_wt = a.inc_wt;
        g.ca = a.ca; g.cb = a.cb; g.dt = a.dt;
        if ((c->d.strict ? nkb::launch_graphene_strict(g, c->s_compute)
                         : nkb::launch_graphene(g, c->s_compute)) != 0)
            return fail("graphene kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        c->last_launches++;
        if (c->g_send_d) CUDA_OK(cudaEventRecord(c->ev_sheet, c->s_compute));
    }

    auto launch_list = [&](int q) -> int {
        if (c->list_n[q] == 0) return 0;
        nkb::StageArgs b = a;
        b.elist = c->elist_d + c->list_off[q];
        b.nel = c->list_n[q];
        int rc = -1;
        if (c->d.ldim == 3) {
            if (c->d.strict)
                rc = nkb::launch_stage_slab_strict(b, c->D_host.data(), c->n, (q & 1) != 0,
                                                   (q & 2) != 0, c->s_compute);
            else if (c->opt_pipeline && c->opt_sweep)
                rc = nkb::launch_stage_sweep(b, c->D_host.data(), c->n, (q & 1) != 0, (q & 2) != 0,
                                             c->s_compute);
            if (rc < 0 && c->opt_pipeline && !c->d.strict)
                rc = nkb::launch_stage_pipe(b, c->D_host.data(), c->n, (q & 1) != 0, (q & 2) != 0,
                                            c->s_compute);
            if (rc < 0 && !c->d.strict)
                rc = nkb::launch_stage_slab(b, c->D_host.data(), c->n, (q & 1) != 0, (q & 2) != 0,
                                            c->s_compute);
        } else {
            rc = c->d.strict
                     ? nkb::launch_stage2d_strict(b, c->D_host.data(), c->n, (q & 1) != 0, c->s_compute)
                     : nkb::launch_stage2d(b, c->D_host.data(), c->n, (q & 1) != 0, c->s_compute);
        }
        if (rc < 0) return fail("nx1=%d is not supported by the stage kernels (2..24)", c->n);
        if (rc > 0) return fail("stage kernel launch failed: %s",
                                cudaGetErrorString(cudaGetLastError()));
        c->last_launches++;
        return 0;
    };

    const bool exchange = !c->peers.empty();
    if (exchange && phase == 0 && c->opt_external_exchange)
        return fail("external_exchange is set: drive the stages with nekcem_b200_stage_pack / "
                    "nekcem_b200_stage_compute");
    const bool p2p = exchange && phase == 0 && c->p2p_ok;
    if (p2p) {
        // peer-memory transport: traces go straight into the peers' halo buffers
        c->stage_no++;
        const int parity = (int)(c->stage_no & 1);
        a.halo = c->halo2 + (size_t)parity * 6 * c->nhalo;
        CUDA_OK(cudaStreamWaitEvent(c->s_comm, c->ev_stage, 0));
        if (c->g_send_d) CUDA_OK(cudaStreamWaitEvent(c->s_comm, c->ev_sheet, 0));
        long long mx = 0;
        for (auto &p : c->peers) mx = std::max<long long>(mx, 6 * (long long)p.send_fp.size());
        pack_push_kernel<<<dim3((unsigned)((mx + 255) / 256), (unsigned)c->peers.size()), 256, 0,
                           c->s_comm>>>(
            a.u_in, c->ld, c->send_node, c->push_d, parity, c->stage_no, c->d.rank, c->done_d,
            c->inc_send_d, c->inc_amp_d, c->inc_phase_d, (int)c->inc_fp.size(), a.inc_wt,
            c->g_send_d, c->send_fp_d, c->g_fj, a.fs_n, a.unx, a.uny,
            c->d.ldim == 3 ? a.unz : nullptr);
        CUDA_OK(cudaGetLastError());
        c->last_launches++;
    }
    if (exchange && phase != 2 && !p2p) {
        // side stream: pack stage-start traces, grouped send/recv (replaces gs_op_fields
        // between ranks, src/cem_maxwell.F:962); overlaps the interior-element launches
        CUDA_OK(cudaStreamWaitEvent(c->s_comm, c->ev_stage, 0));
        if (c->g_send_d) CUDA_OK(cudaStreamWaitEvent(c->s_comm, c->ev_sheet, 0));
        const long long tot = c->nhalo * 6;
        pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->s_comm>>>(
            a.u_in, c->ld, c->send_node, c->sendbuf, c->nhalo, c->inc_send_d, c->inc_amp_d,
            c->inc_phase_d, (int)c->inc_fp.size(), a.inc_wt, c->g_send_d, c->send_fp_d, c->g_fj,
            a.fs_n, a.unx, a.uny, c->d.ldim == 3 ? a.unz : nullptr);
        c->last_launches++;
    }
    if (phase == 1) {
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
        CUDA_OK(cudaStreamSynchronize(c->s_comm));
        return 0;
    }
    if (exchange && phase == 0 && !p2p) {
        NCCL_OK(g_nccl.GroupStart());
        for (auto &p : c->peers) {
            const size_t cnt = p.send_fp.size() * 6;
            NCCL_OK(g_nccl.Send(c->sendbuf + 6 * p.off, cnt, ncclDouble, p.rank, c->comm, c->s_comm));
            NCCL_OK(g_nccl.Recv(c->halo + 6 * p.off, cnt, ncclDouble, p.rank, c->comm, c->s_comm));
        }
        NCCL_OK(g_nccl.GroupEnd());
        CUDA_OK(cudaEventRecord(c->ev_halo, c->s_comm));
    }
    for (int q = 0; q < 4; q++)
        if (launch_list(q)) return 1;
    if (p2p) {
        // boundary elements start once every peer has published this stage's traces
        wait_flags_kernel<<<1, 32, 0, c->s_compute>>>(c->flags_d, c->peer_rank_d,
                                                      (int)c->peers.size(), c->stage_no);
        CUDA_OK(cudaGetLastError());
        c->last_launches++;
    } else if (exchange && phase == 0) {
        CUDA_OK(cudaStreamWaitEvent(c->s_compute, c->ev_halo, 0));
    }
    for (int q = 4; q < 8; q++)
        if (launch_list(q)) return 1;
    CUDA_OK(cudaEventRecord(c->ev_stage, c->s_compute));
    c->cur ^= 1;
    return 0;
}

// Graphene state: host staging -> device (called from nekcem_b200_setup, after the face plan)
int upload_graphene(Ctx *c)
{
    if (c->g_dev_current && c->g_fj && !c->g_fp.empty()) {
        // setup re-run after time steps (e.g. geometry replaced): keep the advanced state
        const size_t ng0 = c->g_fp.size();
        CUDA_OK(cudaMemcpy(c->g_fj_h.data(), c->g_fj, sizeof(double) * 18 * ng0, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(c->g_kj_h.data(), c->g_kj, sizeof(double) * 18 * ng0, cudaMemcpyDeviceToHost));
    }
    c->g_dev_current = false;
    cudaFree(c->g_fp_d); cudaFree(c->g_node_d); cudaFree(c->fs_own_d); cudaFree(c->fs_nbr_d);
    cudaFree(c->g_fj); cudaFree(c->g_kj); cudaFree(c->g_par); cudaFree(c->g_yc);
    cudaFree(c->g_send_d); cudaFree(c->send_fp_d);
    c->g_fp_d = c->g_node_d = c->fs_own_d = c->fs_nbr_d = nullptr;
    c->g_fj = c->g_kj = c->g_par = c->g_yc = nullptr;
    c->g_send_d = c->send_fp_d = nullptr;
    const size_t ng = c->g_fp.size();
    if (ng == 0) return 0;
    const int nfp = c->nxzf * c->nfaces;
    std::vector<int32_t> own(c->nxzfl, -1), nbr(c->nxzfl, -1), node(ng);
    for (size_t q = 0; q < ng; q++) {
        const int64_t fp = c->g_fp[q], e = fp / nfp;
        const int f = (int)(fp - e * nfp);
        if (own[fp] >= 0) return fail("graphene index lists face point %lld twice", (long long)fp + 1);
        own[fp] = (int32_t)q;
        node[q] = (int32_t)(e * c->nxyz + face_node(c->n, f / c->nxzf, f % c->nxzf));
    }
    for (int64_t j = 0; j < c->nxzfl; j++)
        if (c->pairfp[j] >= 0) nbr[j] = own[c->pairfp[j]];
    if (!c->g_yc_given) {
        // yconduc from the reference's COMMON /EMWAVE/ (src/cem_maxwell.F:285), uploaded as
        // NKB_YCONDUC by the host code
        if (!c->have[NKB_YCONDUC])
            return fail("graphene sheets need yconduc: pass it to nekcem_b200_set_graphene or "
                        "upload NKB_YCONDUC");
        std::vector<double> yc(c->nxzfl);
        CUDA_OK(cudaMemcpy(yc.data(), c->dev[NKB_YCONDUC], sizeof(double) * c->nxzfl,
                           cudaMemcpyDeviceToHost));
        c->g_yc_h.resize(ng);
        for (size_t q = 0; q < ng; q++) c->g_yc_h[q] = yc[c->g_fp[q]];
    }
    CUDA_OK(cudaMalloc(&c->g_fp_d, sizeof(int) * ng));
    CUDA_OK(cudaMalloc(&c->g_node_d, sizeof(int) * ng));
    CUDA_OK(cudaMalloc(&c->fs_own_d, sizeof(int) * c->nxzfl));
    CUDA_OK(cudaMalloc(&c->fs_nbr_d, sizeof(int) * c->nxzfl));
    CUDA_OK(cudaMalloc(&c->g_fj, sizeof(double) * 18 * ng));
    CUDA_OK(cudaMalloc(&c->g_kj, sizeof(double) * 18 * ng));
    CUDA_OK(cudaMalloc(&c->g_par, sizeof(double) * 12 * ng));
    CUDA_OK(cudaMalloc(&c->g_yc, sizeof(double) * ng));
    CUDA_OK(cudaMemcpy(c->g_fp_d, c->g_fp.data(), sizeof(int) * ng, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->g_node_d, node.data(), sizeof(int) * ng, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->fs_own_d, own.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->fs_nbr_d, nbr.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->g_fj, c->g_fj_h.data(), sizeof(double) * 18 * ng, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->g_kj, c->g_kj_h.data(), sizeof(double) * 18 * ng, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->g_par, c->g_par_h.data(), sizeof(double) * 12 * ng, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->g_yc, c->g_yc_h.data(), sizeof(double) * ng, cudaMemcpyHostToDevice));
    if (c->nhalo > 0) {
        // sheet points on inter-rank faces: their current is folded into the packed H trace
        std::vector<int32_t> gs(c->nhalo, -1), sfp(c->nhalo, 0);
        bool any = false;
        for (auto &p : c->peers)
            for (size_t q = 0; q < p.send_fp.size(); q++) {
                gs[p.off + q] = own[p.send_fp[q]];
                sfp[p.off + q] = (int32_t)p.send_fp[q];
                any = any || gs[p.off + q] >= 0;
            }
        if (any) {
            CUDA_OK(cudaMalloc(&c->g_send_d, sizeof(int) * c->nhalo));
            CUDA_OK(cudaMalloc(&c->send_fp_d, sizeof(int) * c->nhalo));
            CUDA_OK(cudaMemcpy(c->g_send_d, gs.data(), sizeof(int) * c->nhalo, cudaMemcpyHostToDevice));
            CUDA_OK(cudaMemcpy(c->send_fp_d, sfp.data(), sizeof(int) * c->nhalo, cudaMemcpyHostToDevice));
        }
    }
    c->g_dev_current = true;
    return 0;
}

// `if (iffilter) call q_filter(0.01)` at the end of cem_maxwell_op_rk (src/cem_maxwell.F:342):
// all six components of the current fields, in place
int apply_filter(Ctx *c)
{
    if (!c->filter_d) return 0;
    const int n = c->n, nz = c->d.ldim == 3 ? n : 1;
    const size_t smem = sizeof(double) * (2 * (size_t)c->nxyz + (size_t)n * n);
    static bool configured[64] = {}; // per device
    const int dev = c->d.device;
    if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
    if (!configured[dev]) {
        // two copies of the largest element (nx1 = 24) and the filter matrix: 225.8 KB
        CUDA_OK(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(double) * (2 * 13824 + 576))));
        configured[dev] = true;
    }
    const int nt = c->nxyz >= 256 ? 256 : ((c->nxyz + 31) / 32) * 32;
    filter_kernel<<<dim3((unsigned)c->d.nelt, 6), nt, smem, c->s_compute>>>(
        c->u[c->cur], c->ld, c->filter_d, n, nz, c->nxyz);
    CUDA_OK(cudaGetLastError());
    c->last_launches++;
    c->xtr_valid = false;
    if (c->d.ldim == 2) {
        // the 2D stage kernels never write the three inactive components, so both ping-pong
        // buffers must hold the same (now filtered) values of them
        const bool tm = c->d.imode == 2;
        const int inact[3] = {tm ? 2 : 0, tm ? 3 : 1, tm ? 4 : 5};
        for (int q = 0; q < 3; q++)
            CUDA_OK(cudaMemcpyAsync(c->u[c->cur ^ 1] + inact[q] * c->ld,
                                    c->u[c->cur] + inact[q] * c->ld, sizeof(double) * c->npts,
                                    cudaMemcpyDeviceToDevice, c->s_compute));
    }
    return 0;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {
#pragma GCC visibility push(default)

const char *nekcem_b200_last_error(void) { return g_err.c_str(); }

int nekcem_b200_create(const nekcem_b200_desc *desc, int *handle)
{
    if (!desc || !handle) return fail("null argument");
    if (desc->abi_version != NEKCEM_B200_ABI_VERSION)
        return fail("ABI version mismatch: caller %d, library %d", desc->abi_version,
                    NEKCEM_B200_ABI_VERSION);
    if (!((desc->ldim == 3 && desc->imode == 3) ||
          (desc->ldim == 2 && (desc->imode == 1 || desc->imode == 2))))
        return fail("need ldim=3 with imode=3, or ldim=2 with imode=1 (TE) / 2 (TM); got ldim=%d "
                    "imode=%d", desc->ldim, desc->imode);
    // the reference's own range: mxf1..mxf24 (src/nek5_mxm_std.F)
    if (desc->nx1 < 2 || desc->nx1 > 24)
        return fail("nx1=%d outside the supported range 2..24", desc->nx1);
    if (desc->nelt < 1) return fail("nelt must be >= 1");
    if (desc->nranks < 1 || desc->rank < 0 || desc->rank >= desc->nranks)
        return fail("bad rank/nranks %d/%d", desc->rank, desc->nranks);
    auto c = std::make_unique<Ctx>();
    c->d = *desc;
    c->n = desc->nx1;
    c->nxyz = desc->ldim == 3 ? c->n * c->n * c->n : c->n * c->n;
    c->nxzf = desc->ldim == 3 ? c->n * c->n : c->n;
    c->nfaces = 2 * desc->ldim;
    c->npts = (int64_t)c->nxyz * desc->nelt;
    c->nxzfl = (int64_t)c->nxzf * c->nfaces * desc->nelt;
    if (c->npts > 2147483647LL - 64) return fail("npts exceeds 32-bit node indexing");
    c->ld = ((c->npts + 31) / 32) * 32;
    // device < 0: host-only planning context (face pairing, exchange plan, registration checks; it
    // keeps host copies of uploaded arrays so that the upload path can be inspected, and refuses
    // every compute call).  NEKCEM_B200_HOST_ONLY in the environment forces it whatever the
    // caller asked for -- a test aid for driving the Fortran shim on a machine without a GPU.
    c->host_only = desc->device < 0 || getenv("NEKCEM_B200_HOST_ONLY") != nullptr;
    rk_storage(c.get());
    if (!c->host_only) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return fail("no CUDA device available (%s); libnekcem_b200 has no CPU fallback",
                        cudaGetErrorString(e));
        if (desc->device >= ndev) return fail("device %d out of range (%d devices)", desc->device, ndev);
        CUDA_OK(cudaSetDevice(desc->device));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, desc->device));
        if (prop.major != 10)
            return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                        desc->device, prop.major, prop.minor);
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
        {   // the exchange kernels are tiny and on the critical path of the boundary elements:
            // their stream outranks the persistent stage kernels
            int lo = 0, hi = 0;
            CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_OK(cudaStreamCreateWithPriority(&c->s_comm, cudaStreamNonBlocking, hi));
        }
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_sheet, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreate(&c->ev_t0));
        CUDA_OK(cudaEventCreate(&c->ev_t1));
        for (int q = 0; q < 2; q++) {
            CUDA_OK(cudaMalloc(&c->u[q], sizeof(double) * 6 * c->ld));
            CUDA_OK(cudaMemset(c->u[q], 0, sizeof(double) * 6 * c->ld));
        }
        CUDA_OK(cudaMalloc(&c->kf, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaMemset(c->kf, 0, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaEventRecord(c->ev_stage, c->s_compute));
    }
    g_ctx.push_back(std::move(c));
    *handle = (int)g_ctx.size() - 1;
    return 0;
}

int nekcem_b200_destroy(int handle)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->host_only) {
        cudaSetDevice(c->d.device);
        cudaDeviceSynchronize();
        if (c->has_comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
        for (auto &p : c->dev) cudaFree(p);
        cudaFree(c->u[0]); cudaFree(c->u[1]); cudaFree(c->kf);
        cudaFree(c->xtr[0]); cudaFree(c->xtr[1]);
        release_p2p(c);
        cudaFree(c->st_in); cudaFree(c->st_out);
        if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
        if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
        cudaFree(c->hY); cudaFree(c->hZ); cudaFree(c->vmapP_d); cudaFree(c->elist_d);
        cudaFree(c->sendbuf); cudaFree(c->halo); cudaFree(c->send_node);
        cudaFree(c->src_prof); cudaFree(c->red_d);
        cudaFree(c->ade_j); cudaFree(c->ade_k); cudaFree(c->ade_par); cudaFree(c->ade_mask);
        cudaFree(c->inc_own_d); cudaFree(c->inc_nbr_d); cudaFree(c->inc_send_d);
        cudaFree(c->inc_amp_d); cudaFree(c->inc_phase_d);
        cudaFree(c->g_fp_d); cudaFree(c->g_node_d); cudaFree(c->fs_own_d); cudaFree(c->fs_nbr_d);
        cudaFree(c->g_fj); cudaFree(c->g_kj); cudaFree(c->g_par); cudaFree(c->g_yc);
        cudaFree(c->g_send_d); cudaFree(c->send_fp_d);
        cudaFree(c->filter_d);
        cudaEventDestroy(c->ev_sheet);
        cudaEventDestroy(c->ev_stage); cudaEventDestroy(c->ev_halo);
        cudaEventDestroy(c->ev_t0); cudaEventDestroy(c->ev_t1);
        cudaStreamDestroy(c->s_compute); cudaStreamDestroy(c->s_comm);
    }
    g_ctx[handle].reset();
    return 0;
}

int nekcem_b200_set_array(int handle, int which, const double *host, int64_t count)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (which < 0 || which >= NKB_ARRAY_COUNT) return fail("bad array id %d", which);
    if (!host) return fail("null host pointer");
    const int64_t want = array_count(c, which);
    if (count != want)
        return fail("array id %d: count %lld, expected %lld", which, (long long)count, (long long)want);
    if (c->host_only) {
        c->host_copy[which].assign(host, host + count);
        c->have[which] = true;
        if (which == NKB_DXM1) c->D_host.assign(host, host + count);
        return 0;
    }
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    if (which == NKB_HN || which == NKB_EN || which == NKB_KHN || which == NKB_KEN) {
        double *base = (which == NKB_HN || which == NKB_EN) ? c->u[c->cur] : c->kf;
        const int c0 = (which == NKB_HN || which == NKB_KHN) ? 0 : 3;
        c->xtr_valid = false;
        // on the compute stream (the kernels' streams are non-blocking: a copy on the legacy
        // stream would not be ordered against them), completed before returning
        CUDA_OK(cudaMemcpy2DAsync(base + c0 * c->ld, sizeof(double) * c->ld, host,
                                  sizeof(double) * c->npts, sizeof(double) * c->npts, 3,
                                  cudaMemcpyHostToDevice, c->s_compute));
        // 2D modes never write the inactive components: keep both ping-pong buffers alike
        if (c->d.ldim == 2 && (which == NKB_HN || which == NKB_EN))
            CUDA_OK(cudaMemcpy2DAsync(c->u[c->cur ^ 1] + c0 * c->ld, sizeof(double) * c->ld, host,
                                      sizeof(double) * c->npts, sizeof(double) * c->npts, 3,
                                      cudaMemcpyHostToDevice, c->s_compute));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
    } else {
        if (ensure_dev(c, which)) return 1;
        CUDA_OK(cudaMemcpyAsync(c->dev[which], host, sizeof(double) * count, cudaMemcpyHostToDevice,
                                c->s_compute));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
        if (which == NKB_DXM1) c->D_host.assign(host, host + count);
        if ((which >= NKB_RXMN && which <= NKB_TZMN) || which == NKB_HBM1 || which == NKB_EBM1)
            c->geom_scanned = false;
        if (which == NKB_Y_0 || which == NKB_Z_0) {
            double *&h = (which == NKB_Y_0) ? c->hY : c->hZ;
            if (!h) CUDA_OK(cudaMalloc(&h, sizeof(double) * count));
            half_inverse_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->s_compute>>>(
                c->dev[which], h, count);
            CUDA_OK(cudaStreamSynchronize(c->s_compute));
        }
    }
    c->have[which] = true;
    return 0;
}

int nekcem_b200_get_array(int handle, int which, double *host, int64_t count)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (which < 0 || which >= NKB_ARRAY_COUNT) return fail("bad array id %d", which);
    if (!host) return fail("null host pointer");
    const int64_t want = array_count(c, which);
    if (count != want)
        return fail("array id %d: count %lld, expected %lld", which, (long long)count, (long long)want);
    if (c->host_only) {
        if ((int64_t)c->host_copy[which].size() != count) return fail("array id %d was never set", which);
        memcpy(host, c->host_copy[which].data(), sizeof(double) * count);
        return 0;
    }
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    if (which == NKB_HN || which == NKB_EN || which == NKB_KHN || which == NKB_KEN) {
        const double *base = (which == NKB_HN || which == NKB_EN) ? c->u[c->cur] : c->kf;
        const int c0 = (which == NKB_HN || which == NKB_KHN) ? 0 : 3;
        CUDA_OK(cudaMemcpy2D(host, sizeof(double) * c->npts, base + c0 * c->ld,
                             sizeof(double) * c->ld, sizeof(double) * c->npts, 3,
                             cudaMemcpyDeviceToHost));
    } else {
        if (!c->dev[which]) return fail("array id %d was never set", which);
        CUDA_OK(cudaMemcpy(host, c->dev[which], sizeof(double) * count, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int nekcem_b200_set_faces(int handle, const int64_t *glo_num, int64_t nxzfl, const int32_t *cempec,
                          int32_t ncempec)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (nxzfl != c->nxzfl) return fail("nxzfl %lld, expected %lld", (long long)nxzfl, (long long)c->nxzfl);
    if (!glo_num) return fail("null glo_num");
    if (ncempec < 0 || (ncempec > 0 && !cempec)) return fail("bad cempec");
    c->glo.assign(glo_num, glo_num + nxzfl);
    c->cempec.assign(cempec, cempec + ncempec);
    c->local_matched = c->remote_planned = c->setup_done = false;
    return match_local(c);
}

int nekcem_b200_set_pml(int handle, const int32_t *pmlptr, int32_t maxpml)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (maxpml < 0 || (maxpml > 0 && !pmlptr)) return fail("bad pmlptr");
    c->pml_el.clear();
    for (int q = 0; q < maxpml; q++) {
        if (pmlptr[q] < 1 || pmlptr[q] > c->d.nelt) return fail("pmlptr(%d)=%d out of range", q + 1, pmlptr[q]);
        c->pml_el.push_back(pmlptr[q] - 1);
    }
    c->setup_done = false;
    return 0;
}

int nekcem_b200_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId u;
    if (nccl_load()) return 1;
    NCCL_OK(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return 0;
}

int nekcem_b200_comm_init(int handle, const char id[128])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return fail("host-only planning context: no communicator");
    if (c->d.nranks == 1) return 0;
    CUDA_OK(cudaSetDevice(c->d.device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    if (nccl_load()) return 1;
    NCCL_OK(g_nccl.CommInitRank(&c->comm, c->d.nranks, u, c->d.rank));
    c->has_comm = true;
    return 0;
}

int nekcem_b200_face_singletons(int handle, int64_t *ids, int64_t capacity, int64_t *count)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->local_matched) return fail("set_faces first");
    *count = (int64_t)c->singles.size();
    if (ids) {
        if (capacity < *count) return fail("capacity %lld < %lld", (long long)capacity, (long long)*count);
        for (size_t q = 0; q < c->singles.size(); q++) ids[q] = c->singles[q].first;
    }
    return 0;
}

int nekcem_b200_face_remote(int handle, const int64_t *counts, const int64_t *all_ids)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!counts) return fail("null counts");
    return plan_remote(c, counts, all_ids);
}

int nekcem_b200_plan_npeers(int handle, int32_t *npeers, int64_t *nhalo)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    *npeers = (int32_t)c->peers.size();
    *nhalo = c->nhalo;
    return 0;
}

int nekcem_b200_plan_peer(int handle, int32_t ipeer, int32_t *peer_rank, int64_t *count,
                          int64_t *send_facepts)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (ipeer < 0 || ipeer >= (int)c->peers.size()) return fail("bad peer index");
    const Peer &p = c->peers[ipeer];
    *peer_rank = p.rank;
    *count = (int64_t)p.send_fp.size();
    if (send_facepts) std::copy(p.send_fp.begin(), p.send_fp.end(), send_facepts);
    return 0;
}

int nekcem_b200_plan_vmap(int handle, int32_t *vmapP, int64_t nxzfl)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (nxzfl != c->nxzfl || !c->local_matched) return fail("plan_vmap: not ready");
    std::copy(c->vmapP.begin(), c->vmapP.end(), vmapP);
    return 0;
}

int nekcem_b200_plan_elements(int handle, int32_t *n_interior, int32_t *n_boundary)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    std::vector<int32_t> lists;
    build_lists(c, lists);
    *n_interior = c->n_interior;
    *n_boundary = c->n_boundary;
    return 0;
}

int nekcem_b200_setup(int handle)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->local_matched) return fail("nekcem_b200_set_faces has not been called");
    if (c->host_only) return fail("host-only planning context cannot be set up for compute");
    CUDA_OK(cudaSetDevice(c->d.device));
    if (require(c, {NKB_DXM1, NKB_W3MN, NKB_RXMN, NKB_RYMN, NKB_SXMN, NKB_SYMN, NKB_BMN, NKB_HBM1,
                    NKB_EBM1, NKB_UNXM, NKB_UNYM, NKB_AREAM, NKB_Y_0, NKB_Y_1, NKB_Z_0, NKB_Z_1}))
        return 1;
    if (c->d.ldim == 3 &&
        require(c, {NKB_RZMN, NKB_SZMN, NKB_TXMN, NKB_TYMN, NKB_TZMN, NKB_UNZM}))
        return 1;
    if (!c->pml_el.empty()) {
        if (require(c, {NKB_PERMITTIVITY, NKB_PERMEABILITY, NKB_PMLSIGMA})) return 1;
        for (int id : {NKB_PMLBN, NKB_PMLDN, NKB_KPMLBN, NKB_KPMLDN})
            if (ensure_dev(c, id)) return 1;
    }
    if (c->d.nranks > 1 && !c->remote_planned) {
        if (!c->has_comm) return fail("nranks>1: call nekcem_b200_comm_init (or face_remote) before setup");
        if (exchange_singletons_nccl(c)) return 1;
    }
    std::vector<int32_t> lists;
    build_lists(c, lists);
    cudaFree(c->vmapP_d); cudaFree(c->sendbuf); cudaFree(c->halo);
    cudaFree(c->send_node);
    c->vmapP_d = nullptr; c->sendbuf = c->halo = nullptr; c->send_node = nullptr;
    CUDA_OK(cudaMalloc(&c->vmapP_d, sizeof(int) * c->nxzfl));
    cudaFree(c->xtr[0]); cudaFree(c->xtr[1]);
    c->xtr[0] = c->xtr[1] = nullptr;
    c->xtr_valid = false;
    const long long n_xtr = 2ll * c->n * c->n * c->d.nelt;
    if (c->d.ldim == 3 && c->opt_xtrace && n_xtr < nkb::XTR_BIAS - 4) {
        // neighbour traces across an x face come from the compact mirror: own face slot +x/-x
        // (1, 3) and the neighbour node on an x face of its own element (i = 0 or n-1)
        std::vector<int32_t> pm(c->vmapP);
        const int n = c->n, n2 = n * n, nfp = 6 * n2;
        for (int64_t fp = 0; fp < c->nxzfl; fp++) {
            const int slot = (int)(fp % nfp) / n2;
            const int32_t v = c->vmapP[fp];
            if (v < 0 || (slot != 1 && slot != 3)) continue;
            const int node = v % c->nxyz, i = node % n;
            if (i != 0 && i != n - 1) continue;
            const int64_t t = ((int64_t)(v / c->nxyz) * 2 + (i ? 1 : 0)) * n2 + node / n;
            pm[fp] = (int32_t)(-(3 + (int64_t)nkb::XTR_BIAS + t));
        }
        CUDA_OK(cudaMemcpy(c->vmapP_d, pm.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
        c->ldx = (n_xtr + 31) / 32 * 32;
        for (int q = 0; q < 2; q++) {
            CUDA_OK(cudaMalloc(&c->xtr[q], sizeof(double) * 6 * c->ldx));
            CUDA_OK(cudaMemset(c->xtr[q], 0, sizeof(double) * 6 * c->ldx));
        }
    } else {
        CUDA_OK(cudaMemcpy(c->vmapP_d, c->vmapP.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
    }
    {
        std::vector<unsigned char> ef(c->d.nelt, 0);
        for (int32_t e : c->pml_el) ef[e] |= 1;
        if (c->ade_kind)
            for (int e = 0; e < c->d.nelt; e++)
                if (c->ade_el[e]) ef[e] |= 2;
        cudaFree(c->elflag_d);
        c->elflag_d = nullptr;
        CUDA_OK(cudaMalloc(&c->elflag_d, ef.size()));
        CUDA_OK(cudaMemcpy(c->elflag_d, ef.data(), ef.size(), cudaMemcpyHostToDevice));
        if (scan_geometry(c)) return 1;
    }
    if (c->nhalo > 0) {
        if (!c->has_comm && !c->opt_external_exchange)
            return fail("inter-rank faces exist but no communicator was initialised");
        const int nfp = c->nxzf * c->nfaces;
        std::vector<int> send_node(c->nhalo);
        for (auto &p : c->peers)
            for (size_t q = 0; q < p.send_fp.size(); q++) {
                const int64_t fp = p.send_fp[q], e = fp / nfp;
                const int f = (int)(fp - e * nfp);
                send_node[p.off + q] = (int)(e * c->nxyz + face_node(c->n, f / c->nxzf, f % c->nxzf));
            }
        CUDA_OK(cudaMalloc(&c->send_node, sizeof(int) * c->nhalo));
        CUDA_OK(cudaMemcpy(c->send_node, send_node.data(), sizeof(int) * c->nhalo, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->sendbuf, sizeof(double) * 6 * c->nhalo));
        CUDA_OK(cudaMalloc(&c->halo, sizeof(double) * 6 * c->nhalo));
        CUDA_OK(cudaMemset(c->halo, 0, sizeof(double) * 6 * c->nhalo));
    }
    if (setup_p2p(c)) return 1;
    cudaFree(c->inc_own_d); cudaFree(c->inc_nbr_d); cudaFree(c->inc_send_d);
    cudaFree(c->inc_amp_d); cudaFree(c->inc_phase_d);
    c->inc_own_d = c->inc_nbr_d = c->inc_send_d = nullptr;
    c->inc_amp_d = c->inc_phase_d = nullptr;
    if (!c->inc_fp.empty()) {
        const size_t ni = c->inc_fp.size();
        std::vector<int32_t> own(c->nxzfl, -1), nbr(c->nxzfl, -1);
        for (size_t q = 0; q < ni; q++) own[c->inc_fp[q]] = (int32_t)q;
        for (int64_t j = 0; j < c->nxzfl; j++)
            if (c->pairfp[j] >= 0) nbr[j] = own[c->pairfp[j]];
        CUDA_OK(cudaMalloc(&c->inc_own_d, sizeof(int) * c->nxzfl));
        CUDA_OK(cudaMalloc(&c->inc_nbr_d, sizeof(int) * c->nxzfl));
        CUDA_OK(cudaMemcpy(c->inc_own_d, own.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(c->inc_nbr_d, nbr.data(), sizeof(int) * c->nxzfl, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->inc_amp_d, sizeof(double) * 6 * ni));
        CUDA_OK(cudaMalloc(&c->inc_phase_d, sizeof(double) * ni));
        CUDA_OK(cudaMemcpy(c->inc_amp_d, c->inc_amp.data(), sizeof(double) * 6 * ni, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(c->inc_phase_d, c->inc_phase.data(), sizeof(double) * ni, cudaMemcpyHostToDevice));
        if (c->nhalo > 0) {
            std::vector<int32_t> snd(c->nhalo, -1);
            for (auto &p : c->peers)
                for (size_t q = 0; q < p.send_fp.size(); q++) snd[p.off + q] = own[p.send_fp[q]];
            CUDA_OK(cudaMalloc(&c->inc_send_d, sizeof(int) * c->nhalo));
            CUDA_OK(cudaMemcpy(c->inc_send_d, snd.data(), sizeof(int) * c->nhalo, cudaMemcpyHostToDevice));
        }
    }
    if (upload_graphene(c)) return 1;
    if (!c->red_d) {
        c->red_blocks = 1024;
        CUDA_OK(cudaMalloc(&c->red_d, sizeof(double) * 12 * c->red_blocks));
    }
    CUDA_OK(cudaDeviceSynchronize());
    c->setup_done = true;
    return 0;
}

int nekcem_b200_set_incident(int handle, int32_t ninc, const int32_t *facepts, const double *amp,
                             const double *phase, double omega)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (ninc < 0 || (ninc > 0 && (!facepts || !amp || !phase))) return fail("bad incident arguments");
    c->inc_fp.clear();
    for (int q = 0; q < ninc; q++) {
        if (facepts[q] < 1 || facepts[q] > c->nxzfl)
            return fail("incident face point %d out of range 1..%lld", facepts[q], (long long)c->nxzfl);
        c->inc_fp.push_back(facepts[q] - 1);
    }
    c->inc_amp.assign(amp, amp + 6 * (size_t)ninc);
    c->inc_phase.assign(phase, phase + ninc);
    c->inc_omega = omega;
    c->setup_done = false;
    return 0;
}

int nekcem_b200_set_volume_source(int handle, int comp, const double *profile, double amp,
                                  double omega, double phase)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return fail("host-only planning context");
    if (comp < 0 || comp > 5) return fail("source component %d out of range 0..5", comp);
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    if (!profile) {
        cudaFree(c->src_prof);
        c->src_prof = nullptr;
        return 0;
    }
    if (!c->src_prof) CUDA_OK(cudaMalloc(&c->src_prof, sizeof(double) * c->npts));
    CUDA_OK(cudaMemcpy(c->src_prof, profile, sizeof(double) * c->npts, cudaMemcpyHostToDevice));
    c->src_comp = comp;
    c->src_amp = amp;
    c->src_omega = omega;
    c->src_phase = phase;
    return 0;
}

static int set_ade(int handle, int kind, const double *jn, const double *kjn, const double *params,
                   const int32_t *index, int32_t n)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (n < 0 || (n > 0 && (!index || !params))) return fail("bad ADE arguments");
    if (!c->host_only) {
        CUDA_OK(cudaSetDevice(c->d.device));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
    }
    cudaFree(c->ade_j); cudaFree(c->ade_k); cudaFree(c->ade_par); cudaFree(c->ade_mask);
    c->ade_j = c->ade_k = c->ade_par = nullptr;
    c->ade_mask = nullptr;
    c->ade_kind = 0;
    c->setup_done = false;
    if (n == 0) return 0;
    const int nj = kind == 1 ? 3 : 6, np = kind == 1 ? 2 : 3;
    std::vector<unsigned char> mask(c->npts, 0);
    c->ade_el.assign(c->d.nelt, 0);
    for (int q = 0; q < n; q++) {
        if (index[q] < 1 || index[q] > c->npts)
            return fail("ADE index(%d)=%d out of range 1..%lld", q + 1, index[q], (long long)c->npts);
        mask[index[q] - 1] = 1;
        c->ade_el[(index[q] - 1) / c->nxyz] = 1;
    }
    const size_t bj = sizeof(double) * nj * c->npts, bp = sizeof(double) * np * c->npts;
    const int64_t lp = c->ld_pts ? c->ld_pts : c->npts; // the user's (lpts,k) arrays
    // compact (npts,k) view of a user array: the array itself when lpts == npts, else packed here
    std::vector<double> packed;
    auto compact = [&](const double *src, int ncomp) -> const double * {
        if (lp == c->npts) return src;
        packed.resize((size_t)ncomp * c->npts);
        for (int q = 0; q < ncomp; q++)
            memcpy(packed.data() + (size_t)q * c->npts, src + (size_t)q * lp, sizeof(double) * c->npts);
        return packed.data();
    };
    if (c->host_only) {
        // planning context: keep what would be uploaded (inspected through get_ade)
        c->ade_host_j.assign((size_t)nj * c->npts, 0.0);
        c->ade_host_k.assign((size_t)nj * c->npts, 0.0);
        if (jn) { const double *p = compact(jn, nj); c->ade_host_j.assign(p, p + (size_t)nj * c->npts); }
        if (kjn) { const double *p = compact(kjn, nj); c->ade_host_k.assign(p, p + (size_t)nj * c->npts); }
        c->ade_kind = kind;
        return 0;
    }
    CUDA_OK(cudaMalloc(&c->ade_j, bj));
    CUDA_OK(cudaMalloc(&c->ade_k, bj));
    CUDA_OK(cudaMalloc(&c->ade_par, bp));
    CUDA_OK(cudaMalloc(&c->ade_mask, c->npts));
    if (jn) CUDA_OK(cudaMemcpy(c->ade_j, compact(jn, nj), bj, cudaMemcpyHostToDevice));
    else CUDA_OK(cudaMemset(c->ade_j, 0, bj));
    if (kjn) CUDA_OK(cudaMemcpy(c->ade_k, compact(kjn, nj), bj, cudaMemcpyHostToDevice));
    else CUDA_OK(cudaMemset(c->ade_k, 0, bj));
    CUDA_OK(cudaMemcpy(c->ade_par, compact(params, np), bp, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->ade_mask, mask.data(), c->npts, cudaMemcpyHostToDevice));
    c->ade_kind = kind;
    return 0;
}

int nekcem_b200_set_drude(int handle, const double *jn, const double *kjn, const double *params,
                          const int32_t *dindex, int32_t n)
{
    return set_ade(handle, 1, jn, kjn, params, dindex, n);
}

int nekcem_b200_set_lorentz(int handle, const double *jn, const double *kjn, const double *params,
                            const int32_t *lindex, int32_t n)
{
    return set_ade(handle, 2, jn, kjn, params, lindex, n);
}

int nekcem_b200_get_ade(int handle, double *jn, double *kjn)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->ade_kind) return fail("no Drude/Lorentz state has been set");
    const int nj = c->ade_kind == 1 ? 3 : 6;
    const int64_t lp = c->ld_pts ? c->ld_pts : c->npts;
    const size_t cnt = (size_t)nj * c->npts;
    if (!c->host_only) {
        CUDA_OK(cudaSetDevice(c->d.device));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
    }
    // compact (npts,k) state -> the user's (lpts,k) array
    std::vector<double> tmp;
    auto fetch = [&](double *dst, const double *dev, const std::vector<double> &host) -> int {
        if (lp == c->npts && !c->host_only) {
            CUDA_OK(cudaMemcpy(dst, dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
            return 0;
        }
        const double *src = host.data();
        if (!c->host_only) {
            tmp.resize(cnt);
            CUDA_OK(cudaMemcpy(tmp.data(), dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
            src = tmp.data();
        }
        for (int q = 0; q < nj; q++)
            memcpy(dst + (size_t)q * lp, src + (size_t)q * c->npts, sizeof(double) * c->npts);
        return 0;
    };
    if (jn && fetch(jn, c->ade_j, c->ade_host_j)) return 1;
    if (kjn && fetch(kjn, c->ade_k, c->ade_host_k)) return 1;
    return 0;
}

int nekcem_b200_set_graphene(int handle, const double *fjn, const double *kfjn,
                             const double *params, const double *yconduc, const int32_t *gindex,
                             int32_t n)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (n < 0 || (n > 0 && (!gindex || !params))) return fail("bad graphene arguments");
    if (!c->host_only) {
        CUDA_OK(cudaSetDevice(c->d.device));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
    }
    c->g_fp.clear();
    c->setup_done = false;
    c->g_dev_current = false;
    const size_t ng = (size_t)n;
    const int64_t nfp = c->nxzfl;                       // valid face points
    const int64_t nf = c->ld_fac ? c->ld_fac : nfp;     // leading dimension of the user's arrays
    for (int q = 0; q < n; q++) {
        if (gindex[q] < 1 || gindex[q] > nfp)
            return fail("graphene index(%d)=%d out of range 1..%lld", q + 1, gindex[q], (long long)nfp);
        c->g_fp.push_back(gindex[q] - 1);
    }
    // (nxzfl,3,6) / (nxzfl,12) user arrays -> compact [m][q]
    c->g_fj_h.assign(18 * ng, 0.0);
    c->g_kj_h.assign(18 * ng, 0.0);
    c->g_par_h.assign(12 * ng, 0.0);
    c->g_yc_h.assign(ng, 0.0);
    c->g_yc_given = yconduc != nullptr;
    for (size_t q = 0; q < ng; q++) {
        const int64_t j = c->g_fp[q];
        for (int m = 0; m < 18; m++) {
            if (fjn) c->g_fj_h[m * ng + q] = fjn[j + nf * m];
            if (kfjn) c->g_kj_h[m * ng + q] = kfjn[j + nf * m];
        }
        for (int m = 0; m < 12; m++) c->g_par_h[m * ng + q] = params[j + nf * m];
        if (yconduc) c->g_yc_h[q] = yconduc[j];
    }
    return 0;
}

int nekcem_b200_get_graphene(int handle, double *fjn, double *kfjn)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->g_fp.empty()) return fail("no graphene state has been set");
    const size_t ng = c->g_fp.size();
    const int64_t nf = c->ld_fac ? c->ld_fac : c->nxzfl;
    if (c->g_fj && c->g_dev_current) { // device state is current once setup has run
        CUDA_OK(cudaSetDevice(c->d.device));
        CUDA_OK(cudaStreamSynchronize(c->s_compute));
        CUDA_OK(cudaMemcpy(c->g_fj_h.data(), c->g_fj, sizeof(double) * 18 * ng, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(c->g_kj_h.data(), c->g_kj, sizeof(double) * 18 * ng, cudaMemcpyDeviceToHost));
    }
    for (size_t q = 0; q < ng; q++) {
        const int64_t j = c->g_fp[q];
        for (int m = 0; m < 18; m++) {
            if (fjn) fjn[j + nf * m] = c->g_fj_h[m * ng + q];
            if (kfjn) kfjn[j + nf * m] = c->g_kj_h[m * ng + q];
        }
    }
    return 0;
}

int nekcem_b200_set_leading_dims(int handle, int64_t lpts, int64_t lxzfl)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (lpts < c->npts || lxzfl < c->nxzfl)
        return fail("leading dimensions (%lld, %lld) smaller than npts, nxzfl (%lld, %lld)",
                    (long long)lpts, (long long)lxzfl, (long long)c->npts, (long long)c->nxzfl);
    c->ld_pts = lpts;
    c->ld_fac = lxzfl;
    return 0;
}

int nekcem_b200_set_array_ld(int handle, int which, const double *host, int64_t ld)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (which < 0 || which >= NKB_ARRAY_COUNT) return fail("bad array id %d", which);
    if (!host) return fail("null host pointer");
    const int64_t cnt = array_count(c, which);
    if (cnt != 3 * c->npts || ld == c->npts) return nekcem_b200_set_array(handle, which, host, cnt);
    if (ld < c->npts) return fail("leading dimension %lld smaller than npts %lld", (long long)ld, (long long)c->npts);
    std::vector<double> tmp(3 * (size_t)c->npts);
    for (int q = 0; q < 3; q++)
        memcpy(tmp.data() + q * c->npts, host + q * ld, sizeof(double) * c->npts);
    return nekcem_b200_set_array(handle, which, tmp.data(), cnt);
}

int nekcem_b200_get_array_ld(int handle, int which, double *host, int64_t ld)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (which < 0 || which >= NKB_ARRAY_COUNT) return fail("bad array id %d", which);
    if (!host) return fail("null host pointer");
    const int64_t cnt = array_count(c, which);
    if (cnt != 3 * c->npts || ld == c->npts) return nekcem_b200_get_array(handle, which, host, cnt);
    if (ld < c->npts) return fail("leading dimension %lld smaller than npts %lld", (long long)ld, (long long)c->npts);
    std::vector<double> tmp(3 * (size_t)c->npts);
    if (nekcem_b200_get_array(handle, which, tmp.data(), cnt)) return 1;
    for (int q = 0; q < 3; q++)
        memcpy(host + q * ld, tmp.data() + q * c->npts, sizeof(double) * c->npts);
    return 0;
}

int nekcem_b200_set_rk_coefficients(int handle, const double a[5], const double b[5],
                                    const double cc[6])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!a || !b || !cc) return fail("null argument");
    for (int q = 0; q < 5; q++) { c->rk4a[q] = a[q]; c->rk4b[q] = b[q]; }
    for (int q = 0; q < 6; q++) c->rk4c[q] = cc[q];
    return 0;
}

int nekcem_b200_get_rk_coefficients(int handle, double a[5], double b[5], double cc[6])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!a || !b || !cc) return fail("null argument");
    for (int q = 0; q < 5; q++) { a[q] = c->rk4a[q]; b[q] = c->rk4b[q]; }
    for (int q = 0; q < 6; q++) cc[q] = c->rk4c[q];
    return 0;
}

int nekcem_b200_set_filter(int handle, const double *intv)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return fail("host-only planning context");
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    if (!intv) {
        cudaFree(c->filter_d);
        c->filter_d = nullptr;
        return 0;
    }
    const size_t bytes = sizeof(double) * (size_t)c->n * c->n;
    if (!c->filter_d) CUDA_OK(cudaMalloc(&c->filter_d, bytes));
    CUDA_OK(cudaMemcpy(c->filter_d, intv, bytes, cudaMemcpyHostToDevice));
    return 0;
}

int nekcem_b200_apply_filter(int handle)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (!c->filter_d) return fail("no filter matrix has been set");
    CUDA_OK(cudaSetDevice(c->d.device));
    return apply_filter(c);
}

int nekcem_b200_set_option(int handle, const char *name, int value)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!name) return fail("null option name");
    if (strcmp(name, "pipeline") == 0) {
        c->opt_pipeline = value != 0;
        return 0;
    }
    if (strcmp(name, "sweep") == 0) {
        c->opt_sweep = value != 0;
        return 0;
    }
    if (strcmp(name, "p2p") == 0) {
        // takes effect at the next nekcem_b200_setup (every rank must set the same value)
        c->opt_p2p = value != 0;
        return 0;
    }
    if (strcmp(name, "xtrace") == 0) {
        // takes effect at the next nekcem_b200_setup
        c->opt_xtrace = value != 0;
        return 0;
    }
    if (strcmp(name, "pipeline_ctas") == 0) {
        if (value < 0) return fail("pipeline_ctas must be >= 0");
        c->opt_pipeline_ctas = value;
        return 0;
    }
    if (strcmp(name, "const_metrics") == 0) {
        // 1 (default): exploit exact redundancy of the geometry (constant cofactors per element,
        // identical hbm1/ebm1); 0: stream every array per node as the reference does
        c->opt_const_metrics = value != 0;
        c->geom_scanned = false;
        return 0;
    }
    if (strcmp(name, "external_exchange") == 0) {
        // 1: the caller performs the inter-rank halo exchange itself between
        // nekcem_b200_stage_pack and nekcem_b200_stage_compute (any transport); no communicator
        c->opt_external_exchange = value != 0;
        return 0;
    }
    return fail("unknown option '%s'", name);
}

int nekcem_b200_geometry_info(int handle, int64_t *n_const_metric_elements, int32_t *masses_shared)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->geom_scanned && c->elflag_d) {
        CUDA_OK(cudaSetDevice(c->d.device));
        if (scan_geometry(c)) return 1;
    }
    if (n_const_metric_elements) *n_const_metric_elements = c->n_const_metric_el;
    if (masses_shared) *masses_shared = c->masses_same ? 1 : 0;
    return 0;
}

int nekcem_b200_set_time(int handle, double time, double dt)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    c->time = time;
    c->dt = dt;
    return 0;
}

int nekcem_b200_get_time(int handle, double *time)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    *time = c->time;
    return 0;
}

int nekcem_b200_stage(int handle, int rkstep)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (rkstep < 1 || rkstep > 5) return fail("rkstep %d out of range 1..5", rkstep);
    CUDA_OK(cudaSetDevice(c->d.device));
    return run_stage(c, rkstep);
}

int nekcem_b200_apply_rhs(int handle, double rktime)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    CUDA_OK(cudaSetDevice(c->d.device));
    // one fused stage with (a, b, dt) = (0, 0, 1): k <- res (after invqmass), fields unchanged
    const double t0 = c->time, dt0 = c->dt, a0 = c->rk4a[0], b0 = c->rk4b[0], c0 = c->rk4c[0];
    c->time = rktime; c->dt = 1.0;
    c->rk4a[0] = 0.0; c->rk4b[0] = 0.0; c->rk4c[0] = 0.0;
    const int rc = run_stage(c, 1);
    c->time = t0; c->dt = dt0;
    c->rk4a[0] = a0; c->rk4b[0] = b0; c->rk4c[0] = c0;
    return rc;
}

int nekcem_b200_stage_pack(int handle, int rkstep)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (rkstep < 1 || rkstep > 5) return fail("rkstep %d out of range 1..5", rkstep);
    CUDA_OK(cudaSetDevice(c->d.device));
    return run_stage(c, rkstep, 1);
}

int nekcem_b200_stage_compute(int handle, int rkstep)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (rkstep < 1 || rkstep > 5) return fail("rkstep %d out of range 1..5", rkstep);
    CUDA_OK(cudaSetDevice(c->d.device));
    return run_stage(c, rkstep, 2);
}

int nekcem_b200_halo_buffers(int handle, int32_t ipeer, double **send, double **recv,
                             int64_t *count)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (ipeer < 0 || ipeer >= (int)c->peers.size()) return fail("peer index %d out of range", ipeer);
    const Peer &p = c->peers[ipeer];
    if (send) *send = c->sendbuf + 6 * p.off;
    if (recv) *recv = c->halo + 6 * p.off;
    if (count) *count = 6 * (int64_t)p.send_fp.size();
    return 0;
}

int nekcem_b200_halo_exchange_local(int dst_handle, int src_handle)
{
    Ctx *d = get(dst_handle), *sx = get(src_handle);
    if (!d || !sx) return 1;
    if (!d->setup_done || !sx->setup_done) return fail("nekcem_b200_setup has not completed");
    if (d->d.device != sx->d.device) return fail("halo_exchange_local: contexts on different devices");
    const Peer *ps = nullptr, *pd = nullptr;
    for (auto &p : sx->peers) if (p.rank == d->d.rank) ps = &p;
    for (auto &p : d->peers) if (p.rank == sx->d.rank) pd = &p;
    if (!ps && !pd) return 0; // the two ranks share no face
    if (!ps || !pd || ps->send_fp.size() != pd->send_fp.size())
        return fail("halo_exchange_local: ranks %d and %d disagree about their shared faces",
                    sx->d.rank, d->d.rank);
    CUDA_OK(cudaSetDevice(d->d.device));
    CUDA_OK(cudaMemcpy(d->halo + 6 * pd->off, sx->sendbuf + 6 * ps->off,
                       sizeof(double) * 6 * ps->send_fp.size(), cudaMemcpyDeviceToDevice));
    return 0;
}

int nekcem_b200_step(int handle, int nsteps)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (c->dt == 0.0) return fail("dt is zero: call nekcem_b200_set_time");
    CUDA_OK(cudaSetDevice(c->d.device));
    c->last_launches = 0;
    CUDA_OK(cudaEventRecord(c->ev_t0, c->s_compute));
    for (int s = 0; s < nsteps; s++) {
        for (int rk = 1; rk <= 5; rk++)
            if (run_stage(c, rk)) return 1;
        if (apply_filter(c)) return 1;
        c->time = c->time + c->dt;
    }
    CUDA_OK(cudaEventRecord(c->ev_t1, c->s_compute));
    return 0;
}

// One time step per call on a STREAM of host inputs: the fields advanced by this call are the
// ones the previous call uploaded, the result handed back is the one the previous call computed.
// Copies run on their own streams into / out of staging buffers that no kernel touches, so PCIe
// in both directions overlaps the five stage launches; the buffers change roles by pointer swaps.
int nekcem_b200_step_streamed(int handle, const double *hn_in, const double *en_in, double *hn_out,
                              double *en_out)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (c->dt == 0.0) return fail("dt is zero: call nekcem_b200_set_time");
    if ((hn_in == nullptr) != (en_in == nullptr) || (hn_out == nullptr) != (en_out == nullptr))
        return fail("step_streamed: hn/en pointers must be given in pairs");
    CUDA_OK(cudaSetDevice(c->d.device));
    if (!c->st_in) {
        CUDA_OK(cudaMalloc(&c->st_in, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaMalloc(&c->st_out, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaMemset(c->st_in, 0, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaMemset(c->st_out, 0, sizeof(double) * 6 * c->ld));
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        CUDA_OK(cudaDeviceSynchronize());
    }
    const size_t row = sizeof(double) * c->npts, pitch = sizeof(double) * c->ld;
    // (1) next input: host -> st_in
    if (hn_in) {
        CUDA_OK(cudaMemcpy2DAsync(c->st_in, pitch, hn_in, row, row, 3, cudaMemcpyHostToDevice, c->s_h2d));
        CUDA_OK(cudaMemcpy2DAsync(c->st_in + 3 * c->ld, pitch, en_in, row, row, 3,
                                  cudaMemcpyHostToDevice, c->s_h2d));
    }
    // (2) previous result: st_out -> host
    const bool give = c->st_have_result && hn_out;
    if (give) {
        CUDA_OK(cudaMemcpy2DAsync(hn_out, row, c->st_out, pitch, row, 3, cudaMemcpyDeviceToHost, c->s_d2h));
        CUDA_OK(cudaMemcpy2DAsync(en_out, row, c->st_out + 3 * c->ld, pitch, row, 3,
                                  cudaMemcpyDeviceToHost, c->s_d2h));
    }
    // (3) one time step on the input the previous call delivered
    const bool compute = c->st_have_input;
    c->last_launches = 0;
    CUDA_OK(cudaEventRecord(c->ev_t0, c->s_compute));
    if (compute) {
        for (int rk = 1; rk <= 5; rk++)
            if (run_stage(c, rk)) return 1;
        if (apply_filter(c)) return 1;
        c->time = c->time + c->dt;
    }
    CUDA_OK(cudaEventRecord(c->ev_t1, c->s_compute));
    CUDA_OK(cudaStreamSynchronize(c->s_h2d));
    CUDA_OK(cudaStreamSynchronize(c->s_d2h));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    if (c->has_comm) CUDA_OK(cudaStreamSynchronize(c->s_comm));
    // (4) rotate: result -> st_out, next input -> current fields
    if (compute) {
        std::swap(c->st_out, c->u[c->cur]);
        c->st_have_result = true;
    } else if (give) {
        c->st_have_result = false;
    }
    if (hn_in) {
        std::swap(c->st_in, c->u[c->cur]);
        c->xtr_valid = false;
        if (c->d.ldim == 2) // the 2D kernels never write the inactive components
            CUDA_OK(cudaMemcpy(c->u[c->cur ^ 1], c->u[c->cur], sizeof(double) * 6 * c->ld,
                               cudaMemcpyDeviceToDevice));
    }
    c->st_have_input = hn_in != nullptr;
    return 0;
}

int nekcem_b200_transport(int handle, int32_t *kind)
{
    Ctx *c = get(handle);
    if (!c || !kind) return 1;
    *kind = c->peers.empty() ? 0 : (c->p2p_ok ? 2 : 1);
    return 0;
}

int nekcem_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int nekcem_b200_synchronize(int handle)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return 0;
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaStreamSynchronize(c->s_comm));
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    return 0;
}

int nekcem_b200_last_step_ms(int handle, float *ms, int64_t *launches)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    CUDA_OK(cudaSetDevice(c->d.device));
    CUDA_OK(cudaEventSynchronize(c->ev_t1));
    CUDA_OK(cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    if (ms) *ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    return 0;
}

int nekcem_b200_error_sums(int handle, const double *exact_hn, const double *exact_en,
                           double sumsq[6], double linf[6])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    CUDA_OK(cudaSetDevice(c->d.device));
    double *ex = nullptr;
    CUDA_OK(cudaMalloc(&ex, sizeof(double) * 6 * c->npts));
    CUDA_OK(cudaMemcpy(ex, exact_hn, sizeof(double) * 3 * c->npts, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ex + 3 * c->npts, exact_en, sizeof(double) * 3 * c->npts, cudaMemcpyHostToDevice));
    const int nb = (int)std::min<int64_t>(c->red_blocks, (c->npts + 255) / 256);
    error_kernel<<<nb, 256, 0, c->s_compute>>>(c->u[c->cur], c->ld, ex, c->npts, c->dev[NKB_BMN], c->red_d);
    std::vector<double> part(12 * nb);
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    CUDA_OK(cudaMemcpy(part.data(), c->red_d, sizeof(double) * 12 * nb, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(ex));
    for (int q = 0; q < 6; q++) {
        double s = 0.0, m = 0.0;
        for (int b = 0; b < nb; b++) {
            s += part[b * 12 + q];
            m = std::max(m, part[b * 12 + 6 + q]);
        }
        sumsq[q] = s;
        linf[q] = m;
    }
    return 0;
}

int nekcem_b200_error_sums_mode(int handle, const int32_t kind[18], const double k[3],
                                const double ph[3], const double amp[6], double sumsq[6],
                                double linf[6])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (!kind || !k || !ph || !amp || !sumsq || !linf) return fail("null argument");
    if (require(c, {NKB_XMN, NKB_YMN})) return 1;
    if (c->d.ldim == 3 && require(c, {NKB_ZMN})) return 1;
    CUDA_OK(cudaSetDevice(c->d.device));
    ModeSol m;
    for (int q = 0; q < 18; q++) {
        if (kind[q] < 0 || kind[q] > 2) return fail("mode kind must be 0 (one), 1 (sin) or 2 (cos)");
        m.kind[q] = kind[q];
    }
    for (int q = 0; q < 3; q++) { m.k[q] = k[q]; m.ph[q] = ph[q]; }
    for (int q = 0; q < 6; q++) m.amp[q] = amp[q];
    const int nb = (int)std::min<int64_t>(c->red_blocks, (c->npts + 255) / 256);
    error_mode_kernel<<<nb, 256, 0, c->s_compute>>>(c->u[c->cur], c->ld, m, c->dev[NKB_XMN],
                                                    c->dev[NKB_YMN],
                                                    c->d.ldim == 3 ? c->dev[NKB_ZMN] : nullptr,
                                                    c->npts, c->dev[NKB_BMN], c->red_d);
    std::vector<double> part(12 * nb);
    CUDA_OK(cudaStreamSynchronize(c->s_compute));
    CUDA_OK(cudaMemcpy(part.data(), c->red_d, sizeof(double) * 12 * nb, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 6; q++) {
        double s = 0.0, mxv = 0.0;
        for (int b = 0; b < nb; b++) {
            s += part[b * 12 + q];
            mxv = std::max(mxv, part[b * 12 + 6 + q]);
        }
        sumsq[q] = s;
        linf[q] = mxv;
    }
    return 0;
}

int nekcem_b200_error_sums_planewave(int handle, const nekcem_b200_planewave *wave,
                                     const unsigned char *region, const unsigned char *inpml,
                                     double time, double sumsq[6], double linf[6])
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (!c->setup_done) return fail("nekcem_b200_setup has not completed");
    if (!wave || !region || !inpml || !sumsq || !linf) return fail("null argument");
    if (require(c, {NKB_YMN})) return 1;
    for (int r = 0; r < 2; r++)
        if (!(wave->pml_d[r] != 0.0)) return fail("planewave: pml_d[%d] must be non-zero", r);
    CUDA_OK(cudaSetDevice(c->d.device));
    unsigned char *flags = nullptr;
    CUDA_OK(cudaMalloc(&flags, 2 * (size_t)c->d.nelt));
    CUDA_OK(cudaMemcpy(flags, region, c->d.nelt, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(flags + c->d.nelt, inpml, c->d.nelt, cudaMemcpyHostToDevice));
    const int nb = (int)std::min<int64_t>(c->red_blocks, (c->npts + 255) / 256);
    error_planewave_kernel<<<nb, 256, 0, c->s_compute>>>(
        c->u[c->cur], c->ld, *wave, wave->omega * time, flags, flags + c->d.nelt, c->nxyz,
        c->dev[NKB_YMN], c->npts, c->dev[NKB_BMN], c->red_d);
    std::vector<double> part(12 * nb);
    cudaError_t e1 = cudaStreamSynchronize(c->s_compute);
    cudaFree(flags);
    CUDA_OK(e1);
    CUDA_OK(cudaMemcpy(part.data(), c->red_d, sizeof(double) * 12 * nb, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 6; q++) {
        double s = 0.0, mxv = 0.0;
        for (int b = 0; b < nb; b++) {
            s += part[b * 12 + q];
            mxv = std::max(mxv, part[b * 12 + 6 + q]);
        }
        sumsq[q] = s;
        linf[q] = mxv;
    }
    return 0;
}

int nekcem_b200_vtk_payload(int handle, int which, int as_double, void *out)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return fail("host-only planning context: no device arrays");
    if (which != 0 && which != 1) return fail("vtk_payload: which must be 0 (EN) or 1 (HN)");
    if (!out) return fail("null output pointer");
    CUDA_OK(cudaSetDevice(c->d.device));
    const size_t esz = as_double ? 8 : 4, bytes = esz * 3 * (size_t)c->npts;
    void *buf = nullptr;
    CUDA_OK(cudaMalloc(&buf, bytes));
    const double *base = c->u[c->cur] + (which == 0 ? 3 : 0) * c->ld;
    const long long tot = 3 * (long long)c->npts;
    const unsigned grid = (unsigned)((tot + 255) / 256);
    if (as_double)
        vtk_payload_kernel<unsigned long long><<<grid, 256, 0, c->s_compute>>>(
            base, c->ld, c->npts, (unsigned long long *)buf);
    else
        vtk_payload_kernel<unsigned int><<<grid, 256, 0, c->s_compute>>>(base, c->ld, c->npts,
                                                                         (unsigned int *)buf);
    cudaError_t e1 = cudaStreamSynchronize(c->s_compute);
    if (e1 == cudaSuccess) e1 = cudaMemcpy(out, buf, bytes, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    CUDA_OK(e1);
    return 0;
}

int nekcem_b200_restart_ingest(int handle, int which, int as_double, const void *payload)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    if (c->host_only) return fail("host-only planning context: no device arrays");
    if (which != 0 && which != 1) return fail("restart_ingest: which must be 0 (EN) or 1 (HN)");
    if (!payload) return fail("null payload pointer");
    CUDA_OK(cudaSetDevice(c->d.device));
    const size_t esz = as_double ? 8 : 4, bytes = esz * 3 * (size_t)c->npts;
    void *buf = nullptr;
    CUDA_OK(cudaMalloc(&buf, bytes));
    cudaError_t e1 = cudaMemcpyAsync(buf, payload, bytes, cudaMemcpyHostToDevice, c->s_compute);
    const long long tot = 3 * (long long)c->npts;
    const unsigned grid = (unsigned)((tot + 255) / 256);
    const int c0 = which == 0 ? 3 : 0;
    for (int b = 0; b < (c->d.ldim == 2 ? 2 : 1) && e1 == cudaSuccess; b++) {
        // 2D modes never write the inactive components: both ping-pong buffers take the state
        double *base = c->u[c->cur ^ b] + c0 * c->ld;
        if (as_double)
            restart_ingest_kernel<unsigned long long><<<grid, 256, 0, c->s_compute>>>(
                (const unsigned long long *)buf, base, c->ld, c->npts);
        else
            restart_ingest_kernel<unsigned int><<<grid, 256, 0, c->s_compute>>>(
                (const unsigned int *)buf, base, c->ld, c->npts);
        e1 = cudaGetLastError();
    }
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->s_compute);
    cudaFree(buf);
    CUDA_OK(e1);
    c->xtr_valid = false;
    c->have[which == 0 ? NKB_EN : NKB_HN] = true;
    return 0;
}

int nekcem_b200_algorithmic_bytes(int handle, double *bytes_per_stage)
{
    Ctx *c = get(handle);
    if (!c) return 1;
    // volume 280 B/node, 116 B/face point; PML element node +240 B
    double b;
    if (c->d.ldim == 3) {
        b = 280.0 * (double)c->npts + 116.0 * (double)c->nxzfl;
        b += 240.0 * (double)c->pml_el.size() * (double)c->nxyz;
    } else {
        // 2D: 3 active components (read, write, k read+write) 96 B, rx,ry,sx,sy 32 B, masses 16 B;
        // per face point 2 normals + area + 4 impedances + 3 neighbour values + vmapP = 84 B
        b = 144.0 * (double)c->npts + 84.0 * (double)c->nxzfl;
        b += 120.0 * (double)c->pml_el.size() * (double)c->nxyz;
    }
    if (c->ade_kind) {
        // Drude node: J,kJ read+write (3 comps) + 2 params + bm = 120 B; Lorentz twice the currents
        double nade = 0;
        for (char f : c->ade_el) nade += f ? 1.0 : 0.0;
        b += (c->ade_kind == 1 ? 120.0 : 224.0) * nade * (double)c->nxyz;
    }
    *bytes_per_stage = b;
    return 0;
}

#pragma GCC visibility pop
} // extern "C"
