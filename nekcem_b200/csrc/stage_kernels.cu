// Fused Maxwell RK-stage kernel for sm_100a: "pencil" formulation, 2 <= nx1 <= 16.
//
// One launch = one RK stage over a list of elements.  One CTA = one HALF-TASK (element e,
// group g):  g = 0 updates E from curl(H), g = 1 updates H from -curl(E).  The two half-tasks
// of an element are independent given the stage-start fields (ping-pong buffer u_in), sit in
// adjacent CTAs (shared reads hit L2) and together make ONE pass over HBM per stage
// (SURVEY.md 8a rows a4-a18):
//
//   P0  stage the 3 source components of the element (n^3 nodes each) in shared memory
//   P1  r-pencils: thread (j,k) holds the n-point line of each source component in registers,
//       applies D (kernel parameter -> uniform registers, no memory loads) and stores the raw
//       r-derivatives [local_grad3, src/nek5_grad.F:2-19; mxfK left-to-right sums,
//       src/nek5_mxm_std.F:173-190]
//   P2  s-pencils: thread (i,k): s-derivatives, then the r- and s-parts of the weighted curl with
//       the cofactors rx..sz (coalesced along i) and the quadrature weight w3mn
//       [maxwell_wght_curl, src/cem_maxwell.F:1428-1497]
//   P3  surface flux in three rounds (+-x, +-y, +-z faces; no node is touched twice in a
//       round): own trace from smem/global, neighbour trace through vmapP (the gs_op_fields
//       pair-sum of src/cem_maxwell.F:962) or the NCCL halo, PEC mirror, upwind/central flux,
//       times face area, lifted into the smem residual
//       [restrict_to_face :604-652, flux3d :922-1002, flux_pec :1368-1426, add_flux_to_res
//        :725-735]
//   P4  t-pencils: thread (i,j), adds the weighted t-part to the smem residual
//   P5  streaming epilogue, one node per thread and pass, coalesced, every load of a pass
//       issued before the first use: PML ADEs [pml_step, src/cem_maxwell_pml.F:508-592],
//       volume source [usersrc hook :503], inverse mass [invqmass :1878-1886] and the
//       low-storage RK update [rk4_upd, src/cem_common.F:18-76], written to the ping-pong
//       field buffer.
//
// Memory-level parallelism is explicit: the prologue prefetches every array of the half-task
// into L2 (one warp per array), each pencil phase loads the cofactors of all its outputs before
// the first FMA, and the epilogue loads EPI_UNROLL nodes ahead.
// Shared-memory traffic is ~40 accesses per node per half-task (vs 21n for a naive
// per-node dot product), which is what keeps n = 16 off the shared-memory roofline.
// Arithmetic: same products as the reference; the 6-term curl sum is associated by
// direction ((r-part + s-part)*w + lift) + w*t-part, and nvcc contracts a*b+c into FMA --
// both are <= 1e-15 relative effects per operation (DESIGN.md "Numerics").
#include "stage_common.h"

namespace nkb {

// threads per pencil: each computes ceil(n/split) of the pencil's n outputs
__host__ __device__ constexpr int split_for(int n) { return n <= 5 ? 4 : 2; }
__host__ __device__ constexpr int threads_for(int n)
{
    return ((n * n * split_for(n) + 31) / 32) * 32;
}
__host__ __device__ constexpr int outputs_for(int n) { return (n + split_for(n) - 1) / split_for(n); }
// outputs whose cofactors are loaded together (s-phase needs 7 doubles per output, t-phase 4)
__host__ __device__ constexpr int batch_s(int n)
{
    int no = outputs_for(n);
    return no <= 2 ? no : (no + 1) / 2;
}
__host__ __device__ constexpr int batch_t(int n)
{
    int no = outputs_for(n);
    return no <= 8 ? no : (no + 1) / 2;
}
__host__ __device__ constexpr int regs_for(int n)
{
    // 3*NO accumulators + cofactor batch + working set
    int r = 6 * outputs_for(n) + 14 * batch_s(n) + 44;
    return r > 255 ? 255 : r;
}
__host__ __device__ constexpr int min_blocks_for(int n)
{
    // occupancy target used for the register cap: limited by smem (227 KB) and 2048 threads
    int sj = (n == 8 || n == 16) ? n : n + pad_j(n);
    int sc = ((n == 8 || n == 16) ? n * n : sj * n + pad_k(n)) * n;
    int by_smem = (227 * 1024) / (6 * sc * 8 + 1024);
    int by_thr = 2048 / threads_for(n);
    int b = by_smem < by_thr ? by_smem : by_thr;
    int by_reg = 65536 / (threads_for(n) * regs_for(n));
    if (b > by_reg) b = by_reg;
    if (b > 16) b = 16;
    return b < 1 ? 1 : b;
}
__host__ __device__ constexpr int epi_unroll_for(int n) { return regs_for(n) >= 160 ? 8 : 4; }

// smem offset / element-node index of point m of the pencil (pa,pb) in direction DIR
template <int N, int DIR>
__device__ __forceinline__ int pen_at(int m, int pa, int pb)
{
    return DIR == 0 ? Lay<N>::at(m, pa, pb) : (DIR == 1 ? Lay<N>::at(pa, m, pb) : Lay<N>::at(pa, pb, m));
}
template <int N, int DIR>
__device__ __forceinline__ int pen_node(int m, int pa, int pb)
{
    return DIR == 0 ? m + N * pa + N * N * pb : (DIR == 1 ? pa + N * m + N * N * pb : pa + N * pb + N * N * m);
}

// cofactors (and signed weight) of outputs O0+b0 .. O0+b0+PB-1 of a pencil
template <int N, int DIR, int O0, int O1, int PB, int NCOF, int B0>
__device__ __forceinline__ void load_cof(const StageArgs &a, double (&cof)[PB][NCOF],
                                         double (&wv)[PB], int pa, int pb, long long ebase,
                                         double sg)
{
#pragma unroll
    for (int x = 0; x < PB; x++) {
        const int o = O0 + B0 + x < O1 ? O0 + B0 + x : O1 - 1;
        const int nd = pen_node<N, DIR>(o, pa, pb);
        const long long gi = ebase + nd;
        if constexpr (DIR == 1) {
#pragma unroll
            for (int q = 0; q < 6; q++) cof[x][q] = ldg(a.met[q] + gi);
        } else {
#pragma unroll
            for (int q = 0; q < 3; q++) cof[x][q] = ldg(a.met[6 + q] + gi);
        }
        wv[x] = sg * ldg(a.w3 + nd);
    }
}

// combine the derivatives of outputs O0+B0 .. with their cofactors into the smem residual
template <int N, int DIR, int O0, int O1, int PB, int NCOF, int B0>
__device__ __forceinline__ void finish_batch(const double (&acc)[3][O1 - O0],
                                             const double (&cof)[PB][NCOF],
                                             const double (&wv)[PB], double *R, int pa, int pb)
{
    constexpr int NO = O1 - O0, SC = Lay<N>::SC;
#pragma unroll
    for (int x = 0; x < PB; x++) {
        if constexpr (true) {
            const int oo = B0 + x;
            if (oo < NO) {
                const double d[3] = {acc[0][oo], acc[1][oo], acc[2][oo]};
                double c[3];
                double *Ro = R + pen_at<N, DIR>(O0 + oo, pa, pb);
                if constexpr (DIR == 0) {
                    Ro[0] = d[0]; Ro[SC] = d[1]; Ro[2 * SC] = d[2];
                } else if constexpr (DIR == 1) {
                    const double dr[3] = {Ro[0], Ro[SC], Ro[2 * SC]};
                    double cr[3];
                    curl_part(dr, cof[x][0], cof[x][1], cof[x][2], cr);
                    curl_part(d, cof[x][3], cof[x][4], cof[x][5], c);
                    Ro[0] = (cr[0] + c[0]) * wv[x];
                    Ro[SC] = (cr[1] + c[1]) * wv[x];
                    Ro[2 * SC] = (cr[2] + c[2]) * wv[x];
                } else {
                    curl_part(d, cof[x][0], cof[x][1], cof[x][2], c);
                    Ro[0] = Ro[0] + wv[x] * c[0];
                    Ro[SC] = Ro[SC] + wv[x] * c[1];
                    Ro[2 * SC] = Ro[2 * SC] + wv[x] * c[2];
                }
            }
        }
    }
}

// One pencil phase.  DIR 0/1/2 = r/s/t.  The thread streams the line of N points of pencil
// (pa,pb) from shared memory and produces outputs O0..O1-1 of it:
//   d_c = sum_m D(o,m) u_c(m)                               (mxfK order, left to right)
//   DIR 0: R = d                                             (raw r-derivatives)
//   DIR 1: R = (curl_part(R; rx,ry,rz) + curl_part(d; sx,sy,sz)) * (sg*w3)
//   DIR 2: R = R + (sg*w3) * curl_part(d; tx,ty,tz)
// with curl_part(d; mx,my,mz) = (d3*my - d2*mz, d1*mz - d3*mx, d2*mx - d1*my).
template <int N, int DIR, int O0, int O1, int PB>
__device__ __forceinline__ void pencil_phase(const double (&D)[N * N], const StageArgs &a,
                                             const double *U, double *R, int pa, int pb,
                                             long long ebase, double sg)
{
    constexpr int NO = O1 - O0, SC = Lay<N>::SC;
    constexpr int NCOF = DIR == 1 ? 6 : 3;
    if constexpr (NO > 0) {
        // cofactors and weight of the first batch of outputs: in flight during the contraction
        double cof[PB][NCOF], wv[PB];
        if constexpr (DIR != 0) load_cof<N, DIR, O0, O1, PB, NCOF, 0>(a, cof, wv, pa, pb, ebase, sg);
        // output-stationary contraction: the line is streamed from smem once, each point feeds
        // the 3*NO accumulators of this thread (sum over m left to right, as mxfK)
        double acc[3][NO];
#pragma unroll
        for (int m = 0; m < N; m++) {
            const int so = pen_at<N, DIR>(m, pa, pb);
            const double u0 = U[so], u1 = U[SC + so], u2 = U[2 * SC + so];
#pragma unroll
            for (int o = 0; o < NO; o++) {
                const double dv = D[(O0 + o) + N * m];
                if (m == 0) {
                    acc[0][o] = dv * u0; acc[1][o] = dv * u1; acc[2][o] = dv * u2;
                } else {
                    acc[0][o] = acc[0][o] + dv * u0;
                    acc[1][o] = acc[1][o] + dv * u1;
                    acc[2][o] = acc[2][o] + dv * u2;
                }
            }
        }
        finish_batch<N, DIR, O0, O1, PB, NCOF, 0>(acc, cof, wv, R, pa, pb);
        if constexpr (PB < NO) {
            static_assert(2 * PB >= NO, "at most two cofactor batches");
            if constexpr (DIR != 0)
                load_cof<N, DIR, O0, O1, PB, NCOF, PB>(a, cof, wv, pa, pb, ebase, sg);
            finish_batch<N, DIR, O0, O1, PB, NCOF, PB>(acc, cof, wv, R, pa, pb);
        }
    }
}

// dispatch on the thread's output range (h = which 1/SPLIT of the outputs)
template <int N, int DIR, int SPLIT, int PB>
__device__ __forceinline__ void pencil_split(const double (&D)[N * N], const StageArgs &a,
                                             const double *U, double *R, int pa, int pb,
                                             long long ebase, double sg, int h)
{
    constexpr int HN = (N + SPLIT - 1) / SPLIT;
    constexpr int E1 = HN < N ? HN : N, E2 = 2 * HN < N ? 2 * HN : N, E3 = 3 * HN < N ? 3 * HN : N;
    if (h == 0) pencil_phase<N, DIR, 0, E1, PB>(D, a, U, R, pa, pb, ebase, sg);
    if (SPLIT > 1 && h == 1) pencil_phase<N, DIR, E1, E2, PB>(D, a, U, R, pa, pb, ebase, sg);
    if (SPLIT > 2 && h == 2) pencil_phase<N, DIR, E2, E3, PB>(D, a, U, R, pa, pb, ebase, sg);
    if (SPLIT > 3 && h == 3) pencil_phase<N, DIR, E3, N, PB>(D, a, U, R, pa, pb, ebase, sg);
}

template <int N, bool PML>
__global__ void __launch_bounds__(threads_for(N), min_blocks_for(N))
    stage_kernel(const __grid_constant__ StageParams<N> prm)
{
    constexpr int N2 = N * N, N3 = N2 * N, NF = 6 * N2;
    constexpr int SPLIT = split_for(N), NT = threads_for(N), SC = Lay<N>::SC;
    constexpr int FPT = (2 * N2 + NT - 1) / NT; // face points per thread and flux round
    const StageArgs &a = prm.a;
    extern __shared__ double smem[];
    double *U = smem;          // [3][SC] source components at stage start
    double *R = smem + 3 * SC; // [3][SC] residual of the updated components

    const int tid = threadIdx.x;
    const int e = a.elist[blockIdx.x >> 1];
    const int g = blockIdx.x & 1; // 0: E <- curl H ; 1: H <- -curl E
    const long long ebase = (long long)e * N3;
    const long long cold = (g == 0 ? 3 : 0) * a.ld; // components being updated
    const double *__restrict__ src = a.u_in + (g == 0 ? 0 : 3) * a.ld;
    const double *__restrict__ oth = a.u_in + cold;
    const double sg = g == 0 ? 1.0 : -1.0;

    // ---- P0: stage the source components (loads issued before anything else) -----------------
    constexpr int SPER = (3 * N3 + NT - 1) / NT;
    constexpr int SUNR = SPER < 24 ? SPER : 24;
    double sv[SUNR];
#pragma unroll
    for (int x = 0; x < SUNR; x++) {
        const int q = tid + x * NT;
        const int c = q / N3, r = q - c * N3;
        sv[x] = q < 3 * N3 ? ldg(src + c * a.ld + ebase + r) : 0.0;
    }

    // ---- prologue: put every other HBM request of this half-task in flight now -------------
    // (a) L2 prefetch of the element's metric, mass, RK-register, old-field and face arrays:
    //     one warp per array, no registers or shared memory held while the data travels
    {
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int NW = NT / 32;
        for (int arr = warp; arr < 23; arr += NW) {
            const void *base;
            int bytes = N3 * 8;
            if (arr < 9) base = a.met[arr] + ebase;
            else if (arr < 12) base = a.kf + cold + (arr - 9) * a.ld + ebase;
            else if (arr < 15) base = a.u_in + cold + (arr - 12) * a.ld + ebase;
            else if (arr == 15) base = (g == 0 ? a.ebm1 : a.hbm1) + ebase;
            else {
                const long long fbase = (long long)e * NF;
                bytes = NF * 8;
                if (arr == 16) base = a.unx + fbase;
                else if (arr == 17) base = a.uny + fbase;
                else if (arr == 18) base = a.unz + fbase;
                else if (arr == 19) base = a.area + fbase;
                else if (arr == 20) base = (g == 0 ? a.hZ : a.hY) + fbase;
                else if (arr == 21) base = (g == 0 ? a.Z1 : a.Y1) + fbase;
                else { base = a.vmapP + fbase; bytes = NF * 4; }
            }
            prefetch_chunk(base, bytes, lane);
        }
        // (b) the source components of the half-task that will start pf_dist CTAs from now
        if (a.pf_dist > 0) {
            const int b2 = blockIdx.x + a.pf_dist;
            if (b2 < 2 * a.nel && warp < 3) {
                const int e2 = a.elist[b2 >> 1];
                prefetch_chunk(a.u_in + ((b2 & 1) ? 3 : 0) * a.ld + warp * a.ld + (long long)e2 * N3,
                               N3 * 8, lane);
            }
        }
    }
    // (c) neighbour ids of this thread's face points in each of the three flux rounds
    int vpn[3][FPT];
#pragma unroll
    for (int f = 0; f < FPT; f++) {
        const int q = tid + f * NT;
        vpn[0][f] = vpn[1][f] = vpn[2][f] = -2;
        if (q < 2 * N2) {
            const int hi = q / N2, fp0 = q - hi * N2;
            const long long fb = (long long)e * NF + fp0;
            vpn[0][f] = ldg(a.vmapP + fb + (hi ? 1 : 3) * N2);
            vpn[1][f] = ldg(a.vmapP + fb + (hi ? 2 : 0) * N2);
            vpn[2][f] = ldg(a.vmapP + fb + (hi ? 5 : 4) * N2);
        }
    }
    // staged values -> smem
#pragma unroll 1
    for (int q0 = 0; q0 < SPER; q0 += SUNR) {
        if (q0 > 0) {
#pragma unroll
            for (int x = 0; x < SUNR; x++) {
                const int q = tid + (q0 + x) * NT;
                const int c = q / N3, r = q - c * N3;
                sv[x] = q < 3 * N3 ? ldg(src + c * a.ld + ebase + r) : 0.0;
            }
        }
#pragma unroll
        for (int x = 0; x < SUNR; x++) {
            const int q = tid + (q0 + x) * NT;
            const int c = q / N3, r = q - c * N3;
            const int i = r % N, j = (r / N) % N, k = r / N2;
            if (q < 3 * N3) U[c * SC + Lay<N>::at(i, j, k)] = sv[x];
        }
    }
    // (d) neighbour traces of those face points -> L2
#pragma unroll
    for (int rd = 0; rd < 3; rd++)
#pragma unroll
        for (int f = 0; f < FPT; f++)
            if (vpn[rd][f] >= 0) {
#pragma unroll
                for (int c = 0; c < 6; c++) prefetch_l2(a.u_in + c * a.ld + vpn[rd][f]);
            }
    __syncthreads();

    const int p = tid % N2, h = tid / N2;
    const int pa = p % N, pb = p / N;

    // ---- P1: r-pencils, thread (j,k) = (pa,pb): raw derivatives --------------------------------
    if (h < SPLIT) pencil_split<N, 0, SPLIT, outputs_for(N)>(prm.D, a, U, R, pa, pb, ebase, sg, h);
    __syncthreads();
    // ---- P2: s-pencils, thread (i,k) = (pa,pb): r- and s-parts of the weighted curl -----------
    if (h < SPLIT) pencil_split<N, 1, SPLIT, batch_s(N)>(prm.D, a, U, R, pa, pb, ebase, sg, h);
    __syncthreads();

    // ---- P3: surface flux, three rounds of two opposite faces ---------------------------------
    // slot order of the reference (cemface, cem_common.F:234-260): -y,+x,+y,-x,-z,+z
#pragma unroll 1
    for (int rd = 0; rd < 3; rd++) {
        double fn[FPT][3], far[FPT], fi0[FPT], fi1[FPT], fo[FPT][3], fnb[FPT][6];
        int fsn[FPT], fjs[FPT];
        // all loads of the round first ...
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            const int q = tid + f * NT;
            const int qq = q < 2 * N2 ? q : 0;
            const int hi = qq / N2, fp0 = qq - hi * N2;
            const int fa = fp0 % N, fb = fp0 / N;
            const int ex = hi ? N - 1 : 0;
            int s, ci, cj, ck;
            if (rd == 0) { s = hi ? 1 : 3; ci = ex; cj = fa; ck = fb; }      // +-x
            else if (rd == 1) { s = hi ? 2 : 0; ci = fa; cj = ex; ck = fb; } // +-y
            else { s = hi ? 5 : 4; ci = fa; cj = fb; ck = ex; }              // +-z
            fsn[f] = Lay<N>::at(ci, cj, ck);
            const long long gn = ebase + ci + N * cj + N2 * ck;
            const long long jf = (long long)e * NF + s * N2 + fp0;
            fjs[f] = s * N2 + fp0;
            fn[f][0] = ldg(a.unx + jf); fn[f][1] = ldg(a.uny + jf); fn[f][2] = ldg(a.unz + jf);
            far[f] = ldg(a.area + jf);
            fi0[f] = ldg((g == 0 ? a.hZ : a.hY) + jf);
            fi1[f] = ldg((g == 0 ? a.Z1 : a.Y1) + jf);
#pragma unroll
            for (int c = 0; c < 3; c++) fo[f][c] = ldg(oth + c * a.ld + gn);
            const int vp = rd == 0 ? vpn[0][f] : (rd == 1 ? vpn[1][f] : vpn[2][f]);
            // neighbour trace: volume node, halo slot, or (unused) the own node
            const double *nb = vp >= 0 ? a.u_in + vp
                                       : (vp <= -3 ? a.halo + 6ll * (long long)(-(vp + 3))
                                                   : a.u_in + gn);
            const long long st = vp <= -3 ? 1 : a.ld;
#pragma unroll
            for (int c = 0; c < 6; c++) fnb[f][c] = ldg(nb + c * st);
        }
        // ... then the flux arithmetic and the lift into the smem residual
#pragma unroll
        for (int f = 0; f < FPT; f++) {
            const int q = tid + f * NT;
            if (q < 2 * N2) {
                const int sn = fsn[f];
                const int vp = rd == 0 ? vpn[0][f] : (rd == 1 ? vpn[1][f] : vpn[2][f]);
                const double unx = fn[f][0], uny = fn[f][1], unz = fn[f][2];
                const double S0 = U[sn], S1 = U[SC + sn], S2 = U[2 * SC + sn];
                const double O0 = fo[f][0], O1 = fo[f][1], O2 = fo[f][2];
                // own (H,E)
                double Hx = g == 0 ? S0 : O0, Hy = g == 0 ? S1 : O1, Hz = g == 0 ? S2 : O2;
                double Ex = g == 0 ? O0 : S0, Ey = g == 0 ? O1 : S1, Ez = g == 0 ? O2 : S2;
                double pHx = fnb[f][0], pHy = fnb[f][1], pHz = fnb[f][2];
                double pEx = fnb[f][3], pEy = fnb[f][4], pEz = fnb[f][5];
                if (a.inc_own != nullptr) { // userinc hook (src/cem_maxwell.F:498)
                    const long long jf = (long long)e * NF + fjs[f];
                    const int qo = a.inc_own[jf], qn = a.inc_nbr[jf];
                    if (qo >= 0) {
                        const double ui = cos(a.inc_phase[qo] - a.inc_wt);
                        Hx += a.inc_amp[qo] * ui; Hy += a.inc_amp[a.inc_n + qo] * ui;
                        Hz += a.inc_amp[2 * a.inc_n + qo] * ui;
                        Ex += a.inc_amp[3 * a.inc_n + qo] * ui;
                        Ey += a.inc_amp[4 * a.inc_n + qo] * ui;
                        Ez += a.inc_amp[5 * a.inc_n + qo] * ui;
                    }
                    if (qn >= 0) {
                        const double ui = cos(a.inc_phase[qn] - a.inc_wt);
                        pHx += a.inc_amp[qn] * ui; pHy += a.inc_amp[a.inc_n + qn] * ui;
                        pHz += a.inc_amp[2 * a.inc_n + qn] * ui;
                        pEx += a.inc_amp[3 * a.inc_n + qn] * ui;
                        pEy += a.inc_amp[4 * a.inc_n + qn] * ui;
                        pEz += a.inc_amp[5 * a.inc_n + qn] * ui;
                    }
                }
                // -n x E, -n x H of the own side (flux3d :946-955)
                double s0 = -uny * Ez + unz * Ey;
                double s1 = -unz * Ex + unx * Ez;
                double s2 = -unx * Ey + uny * Ex;
                double s3 = -uny * Hz + unz * Hy;
                double s4 = -unz * Hx + unx * Hz;
                double s5 = -unx * Hy + uny * Hx;
                if (vp >= 0 || vp <= -3) {
                    // neighbour's (-n+ x E+) with n+ = -n-  (the gs_op_fields sum of :962)
                    s0 = s0 - (-uny * pEz + unz * pEy);
                    s1 = s1 - (-unz * pEx + unx * pEz);
                    s2 = s2 - (-unx * pEy + uny * pEx);
                    s3 = s3 - (-uny * pHz + unz * pHy);
                    s4 = s4 - (-unz * pHx + unx * pHz);
                    s5 = s5 - (-unx * pHy + uny * pHx);
                } else if (vp == -1) { // 'PEC' / 'PML' outer face: cem_maxwell_flux_pec :1397-1405
                    s0 = 2.0 * s0; s1 = 2.0 * s1; s2 = 2.0 * s2;
                    s3 = 0.0; s4 = 0.0; s5 = 0.0;
                }
                const double ar = far[f];
                double f0, f1, f2;
                if (g == 1) { // flux into resH (:976-986): fi0 = 0.5/Y_0, fi1 = Y_1
                    const double Y02 = -(fi0[f] * fi1[f]), C02Y = fi0[f] * a.C0;
                    const double fu1 = uny * s5 - unz * s4;
                    const double fu2 = unz * s3 - unx * s5;
                    const double fu3 = unx * s4 - uny * s3;
                    f0 = ar * (Y02 * s0 - C02Y * fu1);
                    f1 = ar * (Y02 * s1 - C02Y * fu2);
                    f2 = ar * (Y02 * s2 - C02Y * fu3);
                } else { // flux into resE (:987-997): fi0 = 0.5/Z_0, fi1 = Z_1
                    const double Z02 = fi0[f] * fi1[f], C02Z = fi0[f] * a.C0;
                    const double fw1 = uny * s2 - unz * s1;
                    const double fw2 = unz * s0 - unx * s2;
                    const double fw3 = unx * s1 - uny * s0;
                    f0 = ar * (Z02 * s3 - C02Z * fw1);
                    f1 = ar * (Z02 * s4 - C02Z * fw2);
                    f2 = ar * (Z02 * s5 - C02Z * fw3);
                }
                R[sn] += f0;
                R[SC + sn] += f1;
                R[2 * SC + sn] += f2;
            }
        }
        __syncthreads();
    }

    // ---- P4: t-pencils, thread (i,j) = (pa,pb) ------------------------------------------------
    if (h < SPLIT) pencil_split<N, 2, SPLIT, batch_t(N)>(prm.D, a, U, R, pa, pb, ebase, sg, h);
    __syncthreads();

    // ---- P5: streaming epilogue ------------------------------------------------------------------
    {
        constexpr int PER = (N3 + NT - 1) / NT;
        constexpr int UNR = PER < epi_unroll_for(N) ? PER : epi_unroll_for(N);
        double *__restrict__ kfp = a.kf + cold + ebase;
        double *__restrict__ uop = a.u_out + cold + ebase;
        const double *__restrict__ mbp = (g == 0 ? a.ebm1 : a.hbm1) + ebase;
#pragma unroll 1
        for (int q0 = 0; q0 < PER; q0 += UNR) {
            double o[UNR][3], kk[UNR][3], mb[UNR];
#pragma unroll
            for (int x = 0; x < UNR; x++) {
                const int node = tid + (q0 + x) * NT;
                const int nd = node < N3 ? node : N3 - 1;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    o[x][c] = ldg(oth + c * a.ld + ebase + nd);
                    kk[x][c] = kfp[c * a.ld + nd];
                }
                mb[x] = ldg(mbp + nd);
            }
#pragma unroll
            for (int x = 0; x < UNR; x++) {
                const int node = tid + (q0 + x) * NT;
                if (node < N3) {
                    const int i = node % N, j = (node / N) % N, k = node / N2;
                    const int sn = Lay<N>::at(i, j, k);
                    const long long gi = ebase + node;
                    double r[3] = {R[sn], R[SC + sn], R[2 * SC + sn]};
                    if (PML) { // pml_step (src/cem_maxwell_pml.F:540-585) + PML half of rk_maxwell_ab
                        const double bm1 = ldg(a.bmn + gi);
                        const double bm1inv = 1.0 / bm1;
                        const double sigx = a.sig[gi], sigy = a.sig[a.npts + gi],
                                     sigz = a.sig[2 * a.npts + gi];
                        const double permitt = a.eps[gi];
                        const double sxp = sigx / permitt, syp = sigy / permitt, szp = sigz / permitt;
                        double *pF = g == 0 ? a.pD : a.pB;
                        double *kF = g == 0 ? a.kD : a.kB;
                        const double b0 = pF[gi], b1 = pF[a.npts + gi], b2 = pF[2 * a.npts + gi];
                        const double rb0 = r[0] * bm1inv - syp * b0;
                        const double rb1 = r[1] * bm1inv - szp * b1;
                        const double rb2 = r[2] * bm1inv - sxp * b2;
                        double p0, p1, p2;
                        if (g == 0) {
                            p0 = -syp * b0 + sxp * b0 - sigz * o[x][0];
                            p1 = -szp * b1 + syp * b1 - sigx * o[x][1];
                            p2 = -sxp * b2 + szp * b2 - sigy * o[x][2];
                        } else {
                            const double permeab = a.mu[gi];
                            p0 = -syp * b0 + sxp * b0 - szp * permeab * o[x][0];
                            p1 = -szp * b1 + syp * b1 - sxp * permeab * o[x][1];
                            p2 = -sxp * b2 + szp * b2 - syp * permeab * o[x][2];
                        }
                        r[0] = r[0] + p0 * bm1; r[1] = r[1] + p1 * bm1; r[2] = r[2] + p2 * bm1;
                        double t;
                        t = a.ca * kF[gi] + a.dt * rb0; kF[gi] = t; pF[gi] = b0 + a.cb * t;
                        t = a.ca * kF[a.npts + gi] + a.dt * rb1; kF[a.npts + gi] = t;
                        pF[a.npts + gi] = b1 + a.cb * t;
                        t = a.ca * kF[2 * a.npts + gi] + a.dt * rb2; kF[2 * a.npts + gi] = t;
                        pF[2 * a.npts + gi] = b2 + a.cb * t;
                    }
                    if (a.src_prof != nullptr) { // usersrc hook: res(comp) -= profile*(tfac*bm)
                        const int cs = a.src_comp - (g == 0 ? 3 : 0);
                        if (cs >= 0 && cs < 3) {
                            const double sv2 = ldg(a.src_prof + gi) * (a.src_tfac * ldg(a.bmn + gi));
                            if (cs == 0) r[0] -= sv2;
                            else if (cs == 1) r[1] -= sv2;
                            else r[2] -= sv2;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double t = a.ca * kk[x][c] + a.dt * (r[c] * mb[x]);
                        kfp[c * a.ld + node] = t;
                        uop[c * a.ld + node] = o[x][c] + a.cb * t;
                    }
                }
            }
        }
    }
}

template <int N>
static int launch_n(const StageArgs &a, const double *Dhost, bool pml, cudaStream_t st)
{
    constexpr size_t smem = sizeof(double) * 6 * Lay<N>::SC;
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(stage_kernel<N, false>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(stage_kernel<N, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return 1;
        configured = true;
    }
    if (a.nel <= 0) return 0;
    StageParams<N> prm;
    prm.a = a;
    for (int q = 0; q < N * N; q++) prm.D[q] = Dhost[q];
    if (pml)
        stage_kernel<N, true><<<2 * a.nel, threads_for(N), smem, st>>>(prm);
    else
        stage_kernel<N, false><<<2 * a.nel, threads_for(N), smem, st>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// returns 0 ok, -1 unsupported order, >0 CUDA failure.  Dhost = dxm1 (n*n, column-major).
int launch_stage(const StageArgs &a, const double *Dhost, int nx1, bool pml, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx1) {
    case 2: return launch_n<2>(a, Dhost, pml, st);
    case 3: return launch_n<3>(a, Dhost, pml, st);
    case 4: return launch_n<4>(a, Dhost, pml, st);
    case 5: return launch_n<5>(a, Dhost, pml, st);
    case 6: return launch_n<6>(a, Dhost, pml, st);
    case 7: return launch_n<7>(a, Dhost, pml, st);
    case 8: return launch_n<8>(a, Dhost, pml, st);
    case 9: return launch_n<9>(a, Dhost, pml, st);
    case 10: return launch_n<10>(a, Dhost, pml, st);
    case 11: return launch_n<11>(a, Dhost, pml, st);
    case 12: return launch_n<12>(a, Dhost, pml, st);
    case 13: return launch_n<13>(a, Dhost, pml, st);
    case 14: return launch_n<14>(a, Dhost, pml, st);
    case 15: return launch_n<15>(a, Dhost, pml, st);
    case 16: return launch_n<16>(a, Dhost, pml, st);
    default: return -1;
    }
}

} // namespace nkb
