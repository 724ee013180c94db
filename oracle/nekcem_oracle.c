/*
 * nekcem_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the NekCEM time-domain Maxwell SEDG right-hand side
 * and its 5-stage low-storage RK step, in the reference's loop order and
 * arithmetic grouping.  Every routine cites the reference file:line it follows
 * (paths relative to the NekCEM source tree).
 *
 * PARITY STATUS: PINNED against the reference's own code run here: oracle/_ref
 * (the reference's Fortran routines of the path translated statement by statement
 * by oracle/f2c_lite.py -- no Fortran compiler exists in this image -- plus its own
 * src/jl gather-scatter library; recipe oracle/build_ref.py).  This restatement
 * must reproduce it bit for bit (tests/test_reference_pin.py); both are compiled
 * with -ffp-contract=off.  Unpinned remainder: the code generation of a real
 * Fortran compiler (<= 1e-15 relative per operation).  The reference's known-answer
 * tests (analytic solutions and L2/Linf tolerances of tests/<case>/<case>.usr) are
 * checked besides (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (nekcem_b200/) never does.
 *
 * Conventions: all arrays are laid out exactly as the Fortran COMMON blocks
 * (column-major, element-major), but indices held in integer arrays are
 * 0-based.  real == double (the reference builds with -fdefault-real-8,
 * bin/configurenek:117-123).  Loops are OpenMP-parallel over independent
 * points/elements only; no reduction order depends on the thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* speclib: GLL points / weights / derivative matrix                   */
/* ------------------------------------------------------------------ */

/* src/nek5_speclib.F:373-394 GAMMAF */
static double nek_gammaf(double x)
{
    const double pi = 4.0 * atan(1.0);
    double g = 1.0;
    if (x == -0.5) g = -2.0 * sqrt(pi);
    if (x == 0.5) g = sqrt(pi);
    if (x == 1.0) g = 1.0;
    if (x == 2.0) g = 1.0;
    if (x == 1.5) g = sqrt(pi) / 2.;
    if (x == 2.5) g = 1.5 * sqrt(pi) / 2.;
    if (x == 3.5) g = 0.5 * (2.5 * (1.5 * sqrt(pi)));
    if (x == 3.) g = 2.;
    if (x == 4.) g = 6.;
    if (x == 5.) g = 24.;
    if (x == 6.) g = 120.;
    return g;
}

/* src/nek5_speclib.F:396-419 PNORMJ */
static double pnormj(int n, double alpha, double beta)
{
    double dn = (double)n;
    double cnst = alpha + beta + 1.0;
    double prod;
    if (n <= 1) {
        prod = nek_gammaf(dn + alpha) * nek_gammaf(dn + beta);
        prod = prod / (nek_gammaf(dn) * nek_gammaf(dn + alpha + beta));
        return prod * pow(2.0, cnst) / (2.0 * dn + cnst);
    }
    prod = nek_gammaf(alpha + 1.0) * nek_gammaf(beta + 1.0);
    prod = prod / (2.0 * (1.0 + cnst) * nek_gammaf(cnst + 1.0));
    prod = prod * (1.0 + alpha) * (2.0 + alpha);
    prod = prod * (1.0 + beta) * (2.0 + beta);
    for (int i = 3; i <= n; i++) {
        double dindx = (double)i;
        double frac = (dindx + alpha) * (dindx + beta) / (dindx * (dindx + alpha + beta));
        prod = prod * frac;
    }
    return prod * pow(2.0, cnst) / (2.0 * dn + cnst);
}

/* src/nek5_speclib.F:482-521 JACOBF */
static void jacobf(double *poly, double *pder, double *polym1, double *pderm1,
                   double *polym2, double *pderm2, int n, double alp, double bet, double x)
{
    double apb = alp + bet;
    double polyl, pderl, psave = 0, pdsave = 0;
    *poly = 1.;
    *pder = 0.;
    if (n == 0) return;
    polyl = *poly;
    pderl = *pder;
    *poly = (alp - bet + (apb + 2.) * x) / 2.;
    *pder = (apb + 2.) / 2.;
    if (n == 1) return;
    for (int k = 2; k <= n; k++) {
        double dk = (double)k;
        double a1 = 2. * dk * (dk + apb) * (2. * dk + apb - 2.);
        double a2 = (2. * dk + apb - 1.) * (alp * alp - bet * bet);
        double b3 = (2. * dk + apb - 2.);
        double a3 = b3 * (b3 + 1.) * (b3 + 2.);
        double a4 = 2. * (dk + alp - 1.) * (dk + bet - 1.) * (2. * dk + apb);
        double polyn = ((a2 + a3 * x) * (*poly) - a4 * polyl) / a1;
        double pdern = ((a2 + a3 * x) * (*pder) - a4 * pderl + a3 * (*poly)) / a1;
        psave = polyl;
        pdsave = pderl;
        polyl = *poly;
        *poly = polyn;
        pderl = *pder;
        *pder = pdern;
    }
    *polym1 = polyl;
    *pderm1 = pderl;
    *polym2 = psave;
    *pderm2 = pdsave;
}

/* src/nek5_speclib.F:421-480 JACG: zeros of Jacobi polynomial by deflated Newton */
static void jacg(double *xjac, int np, double alpha, double beta)
{
    const int kstop = 10;
    const double eps = 1.0e-12;
    int n = np - 1;
    double one = 1.;
    double dth = 4. * atan(one) / (2. * ((double)n) + 2.);
    double xlast = 0, x = 0;
    double p, pd, pm1, pdm1, pm2, pdm2;
    for (int j = 1; j <= np; j++) {
        if (j == 1) {
            x = cos((2. * (((double)j) - 1.) + 1.) * dth);
        } else {
            double x1 = cos((2. * (((double)j) - 1.) + 1.) * dth);
            double x2 = xlast;
            x = (x1 + x2) / 2.;
        }
        for (int k = 1; k <= kstop; k++) {
            jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np, alpha, beta, x);
            double recsum = 0.;
            int jm = j - 1;
            for (int i = 1; i <= jm; i++) recsum = recsum + 1. / (x - xjac[np - i + 1 - 1]);
            double delx = -p / (pd - recsum * p);
            x = x + delx;
            if (fabs(delx) < eps) break;
        }
        xjac[np - j + 1 - 1] = x;
        xlast = x;
    }
    for (int i = 1; i <= np; i++) {
        double xmin = 2.;
        int jmin = i;
        for (int j = i; j <= np; j++) {
            if (xjac[j - 1] < xmin) {
                xmin = xjac[j - 1];
                jmin = j;
            }
        }
        if (jmin != i) {
            double swap = xjac[i - 1];
            xjac[i - 1] = xjac[jmin - 1];
            xjac[jmin - 1] = swap;
        }
    }
}

/* src/nek5_speclib.F:156-206 ZWGJD */
static void zwgjd(double *z, double *w, int np, double alpha, double beta)
{
    int n = np - 1;
    double one = 1., two = 2.;
    double apb = alpha + beta;
    if (np == 1) {
        z[0] = (beta - alpha) / (apb + two);
        w[0] = nek_gammaf(alpha + one) * nek_gammaf(beta + one) / nek_gammaf(apb + two) * pow(two, apb + one);
        return;
    }
    jacg(z, np, alpha, beta);
    int np1 = n + 1, np2 = n + 2;
    double dnp1 = (double)np1, dnp2 = (double)np2;
    double fac1 = dnp1 + alpha + beta + one;
    double fac2 = fac1 + dnp1;
    double fac3 = fac2 + one;
    double fnorm = pnormj(np1, alpha, beta);
    double rcoef = (fnorm * fac2 * fac3) / (two * fac1 * dnp2);
    for (int i = 0; i < np; i++) {
        double p, pd, pm1, pdm1, pm2, pdm2;
        jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np2, alpha, beta, z[i]);
        w[i] = -rcoef / (p * pdm1);
    }
}

/* src/nek5_speclib.F:285-327 ENDW1 */
static double endw1(int n, double alpha, double beta)
{
    double zero = 0., one = 1., two = 2., three = 3., four = 4.;
    double apb = alpha + beta;
    double f1, f2, f3 = 0, fint1, fint2;
    if (n == 0) return zero;
    f1 = nek_gammaf(alpha + two) * nek_gammaf(beta + one) / nek_gammaf(apb + three);
    f1 = f1 * (apb + two) * pow(two, apb + two) / two;
    if (n == 1) return f1;
    fint1 = nek_gammaf(alpha + two) * nek_gammaf(beta + one) / nek_gammaf(apb + three);
    fint1 = fint1 * pow(two, apb + two);
    fint2 = nek_gammaf(alpha + two) * nek_gammaf(beta + two) / nek_gammaf(apb + four);
    fint2 = fint2 * pow(two, apb + three);
    f2 = (-two * (beta + two) * fint1 + (apb + four) * fint2) * (apb + three) / four;
    if (n == 2) return f2;
    for (int i = 3; i <= n; i++) {
        double di = (double)(i - 1);
        double abn = alpha + beta + di;
        double abnn = abn + di;
        double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
        double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
        double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

/* src/nek5_speclib.F:329-371 ENDW2 */
static double endw2(int n, double alpha, double beta)
{
    double zero = 0., one = 1., two = 2., three = 3., four = 4.;
    double apb = alpha + beta;
    double f1, f2, f3 = 0, fint1, fint2;
    if (n == 0) return zero;
    f1 = nek_gammaf(alpha + one) * nek_gammaf(beta + two) / nek_gammaf(apb + three);
    f1 = f1 * (apb + two) * pow(two, apb + two) / two;
    if (n == 1) return f1;
    fint1 = nek_gammaf(alpha + one) * nek_gammaf(beta + two) / nek_gammaf(apb + three);
    fint1 = fint1 * pow(two, apb + two);
    fint2 = nek_gammaf(alpha + two) * nek_gammaf(beta + two) / nek_gammaf(apb + four);
    fint2 = fint2 * pow(two, apb + three);
    f2 = (two * (alpha + two) * fint1 - (apb + four) * fint2) * (apb + three) / four;
    if (n == 2) return f2;
    for (int i = 3; i <= n; i++) {
        double di = (double)(i - 1);
        double abn = alpha + beta + di;
        double abnn = abn + di;
        double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
        double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
        double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

/* src/nek5_speclib.F:107-122 ZWGLL -> ZWGLJ -> ZWGLJD (:240-283), alpha=beta=0 */
ORA_API void ora_zwgll(double *z, double *w, int np)
{
    double alpha = 0., beta = 0.;
    int n = np - 1, nm1 = n - 1;
    double one = 1., two = 2.;
    double p, pd, pm1, pdm1, pm2, pdm2;
    if (np <= 1) {
        fprintf(stderr, "ZWGLJD: Minimum number of Gauss-Lobatto points is 2\n");
        exit(1);
    }
    if (nm1 > 0) {
        double alpg = alpha + one, betg = beta + one;
        zwgjd(z + 1, w + 1, nm1, alpg, betg);
    }
    z[0] = -one;
    z[np - 1] = one;
    for (int i = 1; i < np - 1; i++) w[i] = w[i] / (one - z[i] * z[i]);
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[0]);
    w[0] = endw1(n, alpha, beta) / (two * pd);
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[np - 1]);
    w[np - 1] = endw2(n, alpha, beta) / (two * pd);
}

/* src/nek5_speclib.F:882-912 PNLEG */
static double pnleg(double z, int n)
{
    if (fabs(z) < 1.0e-25) z = 0.0;
    double p1 = 1., p2, p3;
    if (n == 0) return p1;
    p2 = z;
    p3 = p2;
    for (int k = 1; k <= n - 1; k++) {
        double fk = (double)k;
        p3 = ((2. * fk + 1.) * z * p2 - fk * p1) / (fk + 1.);
        p1 = p2;
        p2 = p3;
    }
    return p3;
}

/* src/nek5_speclib.F:807-840 DGLL.  d, dt are column-major nz x nz: d(i,j)=d[i+nz*j] */
ORA_API void ora_dgll(double *d, double *dt, const double *z, int nz)
{
    int n = nz - 1;
    if (nz == 1) {
        d[0] = 0.;
        return;
    }
    double fn = (double)n;
    double d0 = fn * (fn + 1.) / 4.;
    for (int i = 0; i < nz; i++)
        for (int j = 0; j < nz; j++) {
            double v = 0.;
            if (i != j) v = pnleg(z[i], n) / (pnleg(z[j], n) * (z[i] - z[j]));
            if (i == j && i == 0) v = -d0;
            if (i == j && i == nz - 1) v = d0;
            d[i + nz * j] = v;
            dt[j + nz * i] = v;
        }
}

/* ------------------------------------------------------------------ */
/* mxm and local gradients                                             */
/* ------------------------------------------------------------------ */

/* src/nek5_mxm_wrapper.F:1-79 -> mxmf2 -> mxfK (src/nek5_mxm_std.F:1-66, e.g. mxf8
 * :173-190): C(n1,n3) = A(n1,n2)*B(n2,n3), column-major, inner sum strictly left to
 * right starting from the first product (no zero-init add). */
ORA_API void ora_mxm(const double *a, int n1, const double *b, int n2, double *c, int n3)
{
    for (int j = 0; j < n3; j++)
        for (int i = 0; i < n1; i++) {
            double s = a[i] * b[n2 * j];
            for (int k = 1; k < n2; k++) s = s + a[i + n1 * k] * b[k + n2 * j];
            c[i + n1 * j] = s;
        }
}

/* src/nek5_grad.F:2-19 local_grad3 */
static void local_grad3(double *ur, double *us, double *ut, const double *u, int N,
                        const double *D, const double *Dt)
{
    int m1 = N + 1, m2 = m1 * m1;
    ora_mxm(D, m1, u, m1, ur, m2);
    for (int k = 0; k <= N; k++) ora_mxm(u + m2 * k, m1, Dt, m1, us + m2 * k, m1);
    ora_mxm(u, m2, Dt, m1, ut, m1);
}

/* src/nek5_grad.F:21-34 local_grad2 */
static void local_grad2(double *ur, double *us, const double *u, int N, const double *D,
                        const double *Dt)
{
    int m1 = N + 1;
    ora_mxm(D, m1, u, m1, ur, m1);
    ora_mxm(u, m1, Dt, m1, us, m1);
}

/* ------------------------------------------------------------------ */
/* Geometry (setup-time; consumed by the hot path)                     */
/* ------------------------------------------------------------------ */

/* src/nek5_genxyz.F:562-680 GENXYZ, straight-sided elements only.
 * xc,yc,zc: (2^ldim, nelt) corner coordinates in PREPROCESSOR corner order. */
ORA_API void ora_genxyz(int ldim, int nx1, int nelt, const double *zgm1, const double *xc,
                        const double *yc, const double *zc, double *xm1, double *ym1,
                        double *zm1)
{
    int ny1 = nx1, nz1 = (ldim == 3) ? nx1 : 1;
    int nxyz = nx1 * ny1 * nz1;
    int ncrn = 1 << ldim;
    static const int indx[8] = {1, 2, 4, 3, 5, 6, 8, 7};
    double *h = (double *)malloc(sizeof(double) * nx1 * 3 * 2);
#define H(i, d, s) h[(i) + nx1 * ((d) + 3 * (s))]
    for (int ix = 0; ix < nx1; ix++) {
        H(ix, 0, 0) = (1.0 - zgm1[ix]) * 0.5;
        H(ix, 0, 1) = (1.0 + zgm1[ix]) * 0.5;
        H(ix, 1, 0) = (1.0 - zgm1[ix]) * 0.5;
        H(ix, 1, 1) = (1.0 + zgm1[ix]) * 0.5;
        if (ldim == 3) {
            H(ix, 2, 0) = (1.0 - zgm1[ix]) * 0.5;
            H(ix, 2, 1) = (1.0 + zgm1[ix]) * 0.5;
        } else {
            H(ix, 2, 0) = 1.0;
            H(ix, 2, 1) = 1.0;
        }
    }
    for (int e = 0; e < nelt; e++) {
        double xcb[8], ycb[8], zcb[8];
        double *x = xm1 + (size_t)nxyz * e, *y = ym1 + (size_t)nxyz * e, *z = zm1 + (size_t)nxyz * e;
        for (int i = 0; i < nxyz; i++) x[i] = y[i] = z[i] = 0.0;
        for (int ix = 0; ix < ncrn; ix++) {
            int i = indx[ix] - 1;
            xcb[ix] = xc[i + ncrn * e];
            ycb[ix] = yc[i + ncrn * e];
            zcb[ix] = (ldim == 3) ? zc[i + ncrn * e] : 0.0;
        }
        int iztmax = ldim - 1;
        for (int izt = 0; izt < iztmax; izt++)
            for (int iyt = 0; iyt < 2; iyt++)
                for (int ixt = 0; ixt < 2; ixt++) {
                    int c = ixt + 2 * iyt + 4 * izt;
                    for (int iz = 0; iz < nz1; iz++)
                        for (int iy = 0; iy < ny1; iy++) {
                            double hh = H(iy, 1, iyt) * H(iz, 2, izt);
                            for (int ix = 0; ix < nx1; ix++) {
                                double hhh = H(ix, 0, ixt) * hh;
                                int l = ix + nx1 * (iy + ny1 * iz);
                                x[l] = x[l] + hhh * xcb[c];
                                y[l] = y[l] + hhh * ycb[c];
                                z[l] = z[l] + hhh * zcb[c];
                            }
                        }
                }
    }
#undef H
    free(h);
}

/* Geometric factors from nodal coordinates.
 *   XYZRST   src/nek5_coef.F:877-925   (derivatives of x,y,z wrt r,s,t by mxm)
 *   GLMAPM1  src/nek5_coef.F:555-636   (Jacobian and UNNORMALISED cofactors rx=J*dr/dx)
 *   GEODAT1  src/nek5_coef.F:756-778   (bm1 = jac*w3m1)
 *   AREA3    src/nek5_coef.F:1150-1237 / AREA2 :1020-1095 (face area, unit normals;
 *            face slots in preprocessor order 1..6 = -y,+x,+y,-x,-z,+z)
 * Outputs are the flat arrays of cem_maxwell_init (src/cem_maxwell.F:135-162). */
ORA_API void ora_geom(int ldim, int nx1, int nelt, const double *dxm1, const double *dxtm1,
                      const double *wxm1, const double *xm1, const double *ym1,
                      const double *zm1, double *rxm1, double *rym1, double *rzm1,
                      double *sxm1, double *sym1, double *szm1, double *txm1, double *tym1,
                      double *tzm1, double *jacm1, double *bm1, double *w3m1, double *area,
                      double *unx, double *uny, double *unz)
{
    int ny1 = nx1, nz1 = (ldim == 3) ? nx1 : 1;
    int nxy1 = nx1 * ny1, nyz1 = ny1 * nz1, nxyz = nx1 * ny1 * nz1;
    int nfaces = 2 * ldim, nxzf = nx1 * nz1;
    /* GENWZ: src/nek5_coef.F:255 (3D) / :25-120 (2D) */
    for (int iz = 0; iz < nz1; iz++)
        for (int iy = 0; iy < ny1; iy++)
            for (int ix = 0; ix < nx1; ix++)
                w3m1[ix + nx1 * (iy + ny1 * iz)] =
                    (ldim == 3) ? wxm1[ix] * wxm1[iy] * wxm1[iz] : wxm1[ix] * wxm1[iy];

#pragma omp parallel
    {
        double *xr = (double *)malloc(sizeof(double) * nxyz * 12);
        double *yr = xr + nxyz, *zr = yr + nxyz, *xs = zr + nxyz, *ys = xs + nxyz,
               *zs = ys + nxyz, *xt = zs + nxyz, *yt = xt + nxyz, *zt = yt + nxyz,
               *A = zt + nxyz, *B = A + nxyz, *C = B + nxyz;
#pragma omp for
        for (int e = 0; e < nelt; e++) {
            size_t o = (size_t)nxyz * e;
            const double *x = xm1 + o, *y = ym1 + o, *z = zm1 + o;
            ora_mxm(dxm1, nx1, x, nx1, xr, nyz1);
            ora_mxm(dxm1, nx1, y, nx1, yr, nyz1);
            ora_mxm(dxm1, nx1, z, nx1, zr, nyz1);
            for (int iz = 0; iz < nz1; iz++) {
                ora_mxm(x + nxy1 * iz, nx1, dxtm1, ny1, xs + nxy1 * iz, ny1);
                ora_mxm(y + nxy1 * iz, nx1, dxtm1, ny1, ys + nxy1 * iz, ny1);
                ora_mxm(z + nxy1 * iz, nx1, dxtm1, ny1, zs + nxy1 * iz, ny1);
            }
            if (ldim == 3) {
                ora_mxm(x, nxy1, dxtm1, nz1, xt, nz1);
                ora_mxm(y, nxy1, dxtm1, nz1, yt, nz1);
                ora_mxm(z, nxy1, dxtm1, nz1, zt, nz1);
            } else {
                for (int i = 0; i < nxy1; i++) {
                    xt[i] = 0.;
                    yt[i] = 0.;
                    zt[i] = 1.;
                }
            }
            double *rx = rxm1 + o, *ry = rym1 + o, *rz = rzm1 + o, *sx = sxm1 + o,
                   *sy = sym1 + o, *sz = szm1 + o, *tx = txm1 + o, *ty = tym1 + o,
                   *tz = tzm1 + o, *jac = jacm1 + o;
            for (int i = 0; i < nxyz; i++) {
                if (ldim == 2) {
                    double j = 0.;
                    j = j + xr[i] * ys[i];
                    j = j - xs[i] * yr[i];
                    jac[i] = j;
                    rx[i] = ys[i];
                    ry[i] = -xs[i];
                    sx[i] = -yr[i];
                    sy[i] = xr[i];
                    rz[i] = 0.;
                    sz[i] = 0.;
                    tz[i] = 1.;
                    tx[i] = 0.;
                    ty[i] = 0.;
                } else {
                    double j = 0.;
                    j = j + xr[i] * ys[i] * zt[i];
                    j = j + xt[i] * yr[i] * zs[i];
                    j = j + xs[i] * yt[i] * zr[i];
                    j = j - xr[i] * yt[i] * zs[i];
                    j = j - xs[i] * yr[i] * zt[i];
                    j = j - xt[i] * ys[i] * zr[i];
                    jac[i] = j;
                    rx[i] = ys[i] * zt[i] - yt[i] * zs[i];
                    ry[i] = xt[i] * zs[i] - xs[i] * zt[i];
                    rz[i] = xs[i] * yt[i] - xt[i] * ys[i];
                    sx[i] = yt[i] * zr[i] - yr[i] * zt[i];
                    sy[i] = xr[i] * zt[i] - xt[i] * zr[i];
                    sz[i] = xt[i] * yr[i] - xr[i] * yt[i];
                    tx[i] = yr[i] * zs[i] - ys[i] * zr[i];
                    ty[i] = xs[i] * zr[i] - xr[i] * zs[i];
                    tz[i] = xr[i] * ys[i] - xs[i] * yr[i];
                }
                bm1[o + i] = jac[i] * w3m1[i];
            }
            double *ar = area + (size_t)nxzf * nfaces * e, *nx_ = unx + (size_t)nxzf * nfaces * e,
                   *ny_ = uny + (size_t)nxzf * nfaces * e, *nz_ = unz + (size_t)nxzf * nfaces * e;
#define FA(arr, a, b, f) arr[(a) + nx1 * (b) + nxzf * ((f)-1)]
#define V3(arr, i, j, k) arr[(i) + nx1 * ((j) + ny1 * (k))]
            if (ldim == 3) {
                /* "R": faces 2 (+x) and 4 (-x) */
                for (int i = 0; i < nxyz; i++) {
                    A[i] = ys[i] * zt[i] - zs[i] * yt[i];
                    B[i] = zs[i] * xt[i] - xs[i] * zt[i];
                    C[i] = xs[i] * yt[i] - ys[i] * xt[i];
                }
                for (int iz = 0; iz < nz1; iz++)
                    for (int iy = 0; iy < ny1; iy++) {
                        double weight = wxm1[iy] * wxm1[iz];
                        int l2 = (nx1 - 1) + nx1 * (iy + ny1 * iz), l4 = nx1 * (iy + ny1 * iz);
                        double d2 = A[l2] * A[l2] + B[l2] * B[l2] + C[l2] * C[l2];
                        double d4 = A[l4] * A[l4] + B[l4] * B[l4] + C[l4] * C[l4];
                        FA(ar, iy, iz, 2) = sqrt(d2) * weight;
                        FA(ar, iy, iz, 4) = sqrt(d4) * weight;
                        FA(nx_, iy, iz, 4) = -A[l4];
                        FA(nx_, iy, iz, 2) = A[l2];
                        FA(ny_, iy, iz, 4) = -B[l4];
                        FA(ny_, iy, iz, 2) = B[l2];
                        FA(nz_, iy, iz, 4) = -C[l4];
                        FA(nz_, iy, iz, 2) = C[l2];
                    }
                /* "S": faces 1 (-y) and 3 (+y) */
                for (int i = 0; i < nxyz; i++) {
                    A[i] = yr[i] * zt[i] - zr[i] * yt[i];
                    B[i] = zr[i] * xt[i] - xr[i] * zt[i];
                    C[i] = xr[i] * yt[i] - yr[i] * xt[i];
                }
                for (int iz = 0; iz < nz1; iz++)
                    for (int ix = 0; ix < nx1; ix++) {
                        double weight = wxm1[ix] * wxm1[iz];
                        int l1 = ix + nx1 * (0 + ny1 * iz), l3 = ix + nx1 * ((ny1 - 1) + ny1 * iz);
                        double d1 = A[l1] * A[l1] + B[l1] * B[l1] + C[l1] * C[l1];
                        double d3 = A[l3] * A[l3] + B[l3] * B[l3] + C[l3] * C[l3];
                        FA(ar, ix, iz, 1) = sqrt(d1) * weight;
                        FA(ar, ix, iz, 3) = sqrt(d3) * weight;
                        FA(nx_, ix, iz, 1) = A[l1];
                        FA(nx_, ix, iz, 3) = -A[l3];
                        FA(ny_, ix, iz, 1) = B[l1];
                        FA(ny_, ix, iz, 3) = -B[l3];
                        FA(nz_, ix, iz, 1) = C[l1];
                        FA(nz_, ix, iz, 3) = -C[l3];
                    }
                /* "T": faces 5 (-z) and 6 (+z) */
                for (int i = 0; i < nxyz; i++) {
                    A[i] = yr[i] * zs[i] - zr[i] * ys[i];
                    B[i] = zr[i] * xs[i] - xr[i] * zs[i];
                    C[i] = xr[i] * ys[i] - yr[i] * xs[i];
                }
                for (int ix = 0; ix < nx1; ix++)
                    for (int iy = 0; iy < ny1; iy++) {
                        double weight = wxm1[ix] * wxm1[iy];
                        int l5 = ix + nx1 * iy, l6 = ix + nx1 * (iy + ny1 * (nz1 - 1));
                        double d5 = A[l5] * A[l5] + B[l5] * B[l5] + C[l5] * C[l5];
                        double d6 = A[l6] * A[l6] + B[l6] * B[l6] + C[l6] * C[l6];
                        FA(ar, ix, iy, 5) = sqrt(d5) * weight;
                        FA(ar, ix, iy, 6) = sqrt(d6) * weight;
                        FA(nx_, ix, iy, 5) = -A[l5];
                        FA(nx_, ix, iy, 6) = A[l6];
                        FA(ny_, ix, iy, 5) = -B[l5];
                        FA(ny_, ix, iy, 6) = B[l6];
                        FA(nz_, ix, iy, 5) = -C[l5];
                        FA(nz_, ix, iy, 6) = C[l6];
                    }
                /* UNITVEC src/nek5_mat1.F:2006-2017 */
                for (int i = 0; i < nxzf * nfaces; i++) {
                    double len = sqrt(nx_[i] * nx_[i] + ny_[i] * ny_[i] + nz_[i] * nz_[i]);
                    if (len != 0.0) {
                        nx_[i] = nx_[i] / len;
                        ny_[i] = ny_[i] / len;
                        nz_[i] = nz_[i] / len;
                    }
                }
            } else {
                /* AREA2 (non-axisymmetric): WGTR = wxm1 */
                for (int iy = 0; iy < ny1; iy++) {
                    double xs2 = V3(xs, nx1 - 1, iy, 0), ys2 = V3(ys, nx1 - 1, iy, 0);
                    double xs4 = V3(xs, 0, iy, 0), ys4 = V3(ys, 0, iy, 0);
                    double ss2 = sqrt(xs2 * xs2 + ys2 * ys2), ss4 = sqrt(xs4 * xs4 + ys4 * ys4);
                    double t1x2 = xs2 / ss2, t1y2 = ys2 / ss2, t1x4 = -xs4 / ss4, t1y4 = -ys4 / ss4;
                    FA(nx_, iy, 0, 2) = t1y2;
                    FA(ny_, iy, 0, 2) = -t1x2;
                    FA(nx_, iy, 0, 4) = t1y4;
                    FA(ny_, iy, 0, 4) = -t1x4;
                    FA(nz_, iy, 0, 2) = 0.;
                    FA(nz_, iy, 0, 4) = 0.;
                    FA(ar, iy, 0, 2) = ss2 * wxm1[iy];
                    FA(ar, iy, 0, 4) = ss4 * wxm1[iy];
                }
                for (int ix = 0; ix < nx1; ix++) {
                    double xr1 = V3(xr, ix, 0, 0), yr1 = V3(yr, ix, 0, 0);
                    double xr3 = V3(xr, ix, ny1 - 1, 0), yr3 = V3(yr, ix, ny1 - 1, 0);
                    double rr1 = sqrt(xr1 * xr1 + yr1 * yr1), rr3 = sqrt(xr3 * xr3 + yr3 * yr3);
                    double t1x1 = xr1 / rr1, t1y1 = yr1 / rr1, t1x3 = -xr3 / rr3, t1y3 = -yr3 / rr3;
                    FA(nx_, ix, 0, 1) = t1y1;
                    FA(ny_, ix, 0, 1) = -t1x1;
                    FA(nx_, ix, 0, 3) = t1y3;
                    FA(ny_, ix, 0, 3) = -t1x3;
                    FA(nz_, ix, 0, 1) = 0.;
                    FA(nz_, ix, 0, 3) = 0.;
                    FA(ar, ix, 0, 1) = rr1 * wxm1[ix];
                    FA(ar, ix, 0, 3) = rr3 * wxm1[ix];
                }
            }
#undef FA
#undef V3
        }
        free(xr);
    }
}

/* cem_set_fc_ptr src/cem_common.F:214-283 with skpdat (src/nek5_connect11.F:1483-1528)
 * and eface=(4,2,1,3,5,6) (:1067-1072).  cemface is 0-based here. */
ORA_API void ora_set_fc_ptr(int ldim, int nx1, int nelt, int *cemface)
{
    int nx = nx1, ny = nx1, nz = (ldim == 3) ? nx1 : 1;
    int nxyz = nx * ny * nz, nxzf = nx1 * nz, nfaces = 2 * ldim;
    static const int eface[6] = {4, 2, 1, 3, 5, 6};
    int skp[6][6];
    /* Fortran SKPDAT(1..6, face) */
    skp[0][0] = 1; skp[0][1] = nx * (ny - 1) + 1; skp[0][2] = nx;
    skp[0][3] = 1; skp[0][4] = ny * (nz - 1) + 1; skp[0][5] = ny;
    skp[1][0] = 1 + (nx - 1); skp[1][1] = nx * (ny - 1) + 1 + (nx - 1); skp[1][2] = nx;
    skp[1][3] = 1; skp[1][4] = ny * (nz - 1) + 1; skp[1][5] = ny;
    skp[2][0] = 1; skp[2][1] = nx; skp[2][2] = 1;
    skp[2][3] = 1; skp[2][4] = ny * (nz - 1) + 1; skp[2][5] = ny;
    skp[3][0] = 1 + nx * (ny - 1); skp[3][1] = nx + nx * (ny - 1); skp[3][2] = 1;
    skp[3][3] = 1; skp[3][4] = ny * (nz - 1) + 1; skp[3][5] = ny;
    skp[4][0] = 1; skp[4][1] = nx; skp[4][2] = 1;
    skp[4][3] = 1; skp[4][4] = ny; skp[4][5] = 1;
    skp[5][0] = 1 + nx * ny * (nz - 1); skp[5][1] = nx + nx * ny * (nz - 1); skp[5][2] = 1;
    skp[5][3] = 1; skp[5][4] = ny; skp[5][5] = 1;
    for (int e = 1; e <= nelt; e++)
        for (int f = 1; f <= nfaces; f++) {
            int ef = eface[f - 1];
            int js1 = skp[f - 1][0], jf1 = skp[f - 1][1], jskip1 = skp[f - 1][2];
            int js2 = skp[f - 1][3], jf2 = skp[f - 1][4], jskip2 = skp[f - 1][5];
            int i = 0;
            for (int j2 = js2; j2 <= jf2; j2 += jskip2)
                for (int j1 = js1; j1 <= jf1; j1 += jskip1) {
                    i = i + 1;
                    int k = i + nxzf * (ef - 1) + nxzf * nfaces * (e - 1);
                    cemface[k - 1] = (j1 + nx1 * (j2 - 1) + nxyz * (e - 1)) - 1;
                }
        }
}

/* ------------------------------------------------------------------ */
/* gather-scatter on face ids: semantic of gs_setup/gs_op_fields        */
/* (src/jl/gs.c:1898-1907, 2074-2100): every set of entries sharing a   */
/* non-zero id is replaced by op(all of them).  Ids of 0 do not take    */
/* part.  Single process.                                               */
/* ------------------------------------------------------------------ */
typedef struct {
    int n;
    int ngroup;
    int *gptr; /* ngroup+1 */
    int *gind; /* members, ascending index within a group */
} ora_gs_t;

typedef struct {
    long long id;
    int idx;
} idpair_t;

static int cmp_idpair(const void *a, const void *b)
{
    const idpair_t *x = (const idpair_t *)a, *y = (const idpair_t *)b;
    if (x->id < y->id) return -1;
    if (x->id > y->id) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

ORA_API ora_gs_t *ora_gs_setup(const long long *id, int n)
{
    ora_gs_t *g = (ora_gs_t *)calloc(1, sizeof(ora_gs_t));
    idpair_t *p = (idpair_t *)malloc(sizeof(idpair_t) * (n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (id[i] != 0) {
            p[m].id = id[i];
            p[m].idx = i;
            m++;
        }
    qsort(p, m, sizeof(idpair_t), cmp_idpair);
    g->n = n;
    g->gptr = (int *)malloc(sizeof(int) * (m + 1));
    g->gind = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
    int ng = 0, k = 0;
    for (int i = 0; i < m;) {
        int j = i;
        while (j < m && p[j].id == p[i].id) j++;
        if (j - i > 1) {
            g->gptr[ng++] = k;
            for (int q = i; q < j; q++) g->gind[k++] = p[q].idx;
        }
        i = j;
    }
    g->gptr[ng] = k;
    g->ngroup = ng;
    free(p);
    return g;
}

ORA_API void ora_gs_free(ora_gs_t *g)
{
    if (!g) return;
    free(g->gptr);
    free(g->gind);
    free(g);
}

/* op: 1 add, 2 mul, 3 min, 4 max (src/jl/gs_defs.h ordering used by the Fortran API) */
ORA_API void ora_gs_op_fields(const ora_gs_t *g, double *u, int stride, int nf, int op)
{
#pragma omp parallel for
    for (int q = 0; q < g->ngroup; q++)
        for (int f = 0; f < nf; f++) {
            double *v = u + (size_t)stride * f;
            int b = g->gptr[q], e = g->gptr[q + 1];
            double s = v[g->gind[b]];
            for (int m = b + 1; m < e; m++) {
                double t = v[g->gind[m]];
                if (op == 1) s = s + t;
                else if (op == 2) s = s * t;
                else if (op == 3) s = (t < s) ? t : s;
                else s = (t > s) ? t : s;
            }
            for (int m = b; m < e; m++) v[g->gind[m]] = s;
        }
}

/* ------------------------------------------------------------------ */
/* Solver state: mirrors the COMMON blocks the hot path touches          */
/* ------------------------------------------------------------------ */
typedef void (*ora_userface_cb)(double tt, double *a1, double *a2, double *a3, double *a4,
                                double *a5, double *a6, void *ctx);

typedef struct ora_state {
    /* SIZE / DIMN */
    int ldim, nx1, nelt, imode; /* imode: 3 = 3D, 2 = TM, 1 = TE (src/INPUT) */
    int nxyz, nxzf, nfaces, npts, nxzfl;
    int ifupwind, ifcentral, ifpml, ifpec;
    /* TSTEP / RK5 */
    double dt, time, rktime;
    int rkstep, istep;
    double rk4a[5], rk4b[5], rk4c[6];
    /* DXYZ / WZ / GEOM */
    double *dxm1, *dxtm1, *w3mn;
    double *rxmn, *rymn, *rzmn, *sxmn, *symn, *szmn, *txmn, *tymn, *tzmn, *bmn;
    double *unxm, *unym, *unzm, *aream;
    /* INPUT */
    int *cemface;
    int ncemface;
    int *cempec;
    int ncempec;
    ora_gs_t *gsh_face;
    /* EMWAVE */
    double *hn, *en, *khn, *ken, *reshn, *resen; /* (npts,3) */
    double *fhn, *fen;                           /* (nxzfl,3) */
    double *srflx;                               /* 6*nxzfl */
    double *hbm1, *ebm1;
    double *Y_0, *Y_1, *Z_0, *Z_1;
    double *permittivity, *permeability;
    /* PML */
    int maxpml;
    int *pmlptr; /* 0-based element ids */
    double *pmlsigma, *pmlbn, *pmldn, *respmlbn, *respmldn, *respmlhn, *respmlen, *kpmlbn,
        *kpmldn;
    /* .usr callbacks (NULL = empty routine) */
    ora_userface_cb userinc;  /* (tt, fhx,fhy,fhz, fex,fey,fez) */
    ora_userface_cb usersrc;  /* (tt, reshx..reshz, resex..resez) */
    ora_userface_cb userfsrc; /* (tt, args exactly as the reference passes them) */
    void *ctx;
} ora_state;

/* maxwell_wght_curl src/cem_maxwell.F:1428-1539 */
static void maxwell_wght_curl(const ora_state *s, double *w1, double *w2, double *w3,
                              const double *u1, const double *u2, const double *u3)
{
    int nn = s->nx1 - 1, nxyz = s->nxyz;
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * nxyz * 9);
        double *u1r = buf, *u1s = u1r + nxyz, *u1t = u1s + nxyz, *u2r = u1t + nxyz,
               *u2s = u2r + nxyz, *u2t = u2s + nxyz, *u3r = u2t + nxyz, *u3s = u3r + nxyz,
               *u3t = u3s + nxyz;
#pragma omp for
        for (int e = 0; e < s->nelt; e++) {
            size_t j = (size_t)nxyz * e;
            if (s->ldim == 3) {
                local_grad3(u1r, u1s, u1t, u1 + j, nn, s->dxm1, s->dxtm1);
                local_grad3(u2r, u2s, u2t, u2 + j, nn, s->dxm1, s->dxtm1);
                local_grad3(u3r, u3s, u3t, u3 + j, nn, s->dxm1, s->dxtm1);
                for (int i = 0; i < nxyz; i++) {
                    size_t k = i + j;
                    double u1rw = u1r[i] * s->w3mn[i], u1sw = u1s[i] * s->w3mn[i],
                           u1tw = u1t[i] * s->w3mn[i];
                    double u2rw = u2r[i] * s->w3mn[i], u2sw = u2s[i] * s->w3mn[i],
                           u2tw = u2t[i] * s->w3mn[i];
                    double u3rw = u3r[i] * s->w3mn[i], u3sw = u3s[i] * s->w3mn[i],
                           u3tw = u3t[i] * s->w3mn[i];
                    double rx = s->rxmn[k], sx = s->sxmn[k], tx = s->txmn[k];
                    double ry = s->rymn[k], sy = s->symn[k], ty = s->tymn[k];
                    double rz = s->rzmn[k], sz = s->szmn[k], tz = s->tzmn[k];
                    w1[k] = u3rw * ry + u3sw * sy + u3tw * ty - u2rw * rz - u2sw * sz - u2tw * tz;
                    w2[k] = u1rw * rz + u1sw * sz + u1tw * tz - u3rw * rx - u3sw * sx - u3tw * tx;
                    w3[k] = u2rw * rx + u2sw * sx + u2tw * tx - u1rw * ry - u1sw * sy - u1tw * ty;
                }
            } else {
                local_grad2(u1r, u1s, u1 + j, nn, s->dxm1, s->dxtm1);
                local_grad2(u2r, u2s, u2 + j, nn, s->dxm1, s->dxtm1);
                local_grad2(u3r, u3s, u3 + j, nn, s->dxm1, s->dxtm1);
                for (int i = 0; i < nxyz; i++) {
                    size_t k = i + j;
                    double u1rw = u1r[i] * s->w3mn[i], u1sw = u1s[i] * s->w3mn[i];
                    double u2rw = u2r[i] * s->w3mn[i], u2sw = u2s[i] * s->w3mn[i];
                    double u3rw = u3r[i] * s->w3mn[i], u3sw = u3s[i] * s->w3mn[i];
                    double rx = s->rxmn[k], sx = s->sxmn[k], ry = s->rymn[k], sy = s->symn[k];
                    w1[k] = (u3rw * ry + u3sw * sy);
                    w2[k] = -(u3rw * rx + u3sw * sx);
                    w3[k] = (u2rw * rx + u2sw * sx - u1rw * ry - u1sw * sy);
                }
            }
        }
        free(buf);
    }
}

static void chsign(double *a, int n)
{
#pragma omp parallel for
    for (int i = 0; i < n; i++) a[i] = -a[i];
}

#define HN(c) (s->hn + (size_t)(c)*s->npts)
#define EN(c) (s->en + (size_t)(c)*s->npts)
#define RESHN(c) (s->reshn + (size_t)(c)*s->npts)
#define RESEN(c) (s->resen + (size_t)(c)*s->npts)
#define KHN(c) (s->khn + (size_t)(c)*s->npts)
#define KEN(c) (s->ken + (size_t)(c)*s->npts)
#define FHN(c) (s->fhn + (size_t)(c)*s->nxzfl)
#define FEN(c) (s->fen + (size_t)(c)*s->nxzfl)

/* cem_maxwell src/cem_maxwell.F:510-602 (non-dealiased, non-OpenACC branch :574-596) */
ORA_API void ora_cem_maxwell(ora_state *s)
{
    int npts = s->npts;
    if (s->imode == 3) {
        maxwell_wght_curl(s, RESEN(0), RESEN(1), RESEN(2), HN(0), HN(1), HN(2));
        maxwell_wght_curl(s, RESHN(0), RESHN(1), RESHN(2), EN(0), EN(1), EN(2));
        chsign(RESHN(0), npts);
        chsign(RESHN(1), npts);
        chsign(RESHN(2), npts);
    } else if (s->imode == 2) {
        maxwell_wght_curl(s, RESHN(0), RESHN(1), RESEN(2), HN(0), HN(1), EN(2));
        chsign(RESHN(0), npts);
        chsign(RESHN(1), npts);
    } else if (s->imode == 1) {
        maxwell_wght_curl(s, RESEN(0), RESEN(1), RESHN(2), EN(0), EN(1), HN(2));
        chsign(RESHN(2), npts);
    }
}

/* cem_maxwell_restrict_to_face src/cem_maxwell.F:604-652 */
ORA_API void ora_restrict_to_face(ora_state *s)
{
#pragma omp parallel for
    for (int j = 0; j < s->ncemface; j++) {
        int i = s->cemface[j];
        if (s->imode == 3) {
            FHN(0)[j] = HN(0)[i];
            FHN(1)[j] = HN(1)[i];
            FHN(2)[j] = HN(2)[i];
            FEN(0)[j] = EN(0)[i];
            FEN(1)[j] = EN(1)[i];
            FEN(2)[j] = EN(2)[i];
        } else if (s->imode == 2) {
            FHN(0)[j] = HN(0)[i];
            FHN(1)[j] = HN(1)[i];
            FHN(2)[j] = 0.0;
            FEN(0)[j] = 0.0;
            FEN(1)[j] = 0.0;
            FEN(2)[j] = EN(2)[i];
        } else {
            FHN(0)[j] = 0.0;
            FHN(1)[j] = 0.0;
            FHN(2)[j] = HN(2)[i];
            FEN(0)[j] = EN(0)[i];
            FEN(1)[j] = EN(1)[i];
            FEN(2)[j] = 0.0;
        }
    }
}

/* cem_maxwell_flux_pec src/cem_maxwell.F:1368-1426 */
static void flux_pec(ora_state *s)
{
    size_t k = s->nxzfl;
    double *f = s->srflx;
#pragma omp parallel for
    for (int j = 0; j < s->ncempec; j++) {
        int i = s->cempec[j];
        if (s->imode == 3) {
            f[0 * k + i] = 2.0 * f[0 * k + i];
            f[1 * k + i] = 2.0 * f[1 * k + i];
            f[2 * k + i] = 2.0 * f[2 * k + i];
            f[3 * k + i] = 0;
            f[4 * k + i] = 0;
            f[5 * k + i] = 0;
        } else if (s->imode == 2) {
            f[0 * k + i] = 2.0 * f[0 * k + i];
            f[1 * k + i] = 2.0 * f[1 * k + i];
            f[2 * k + i] = 0;
        } else {
            f[0 * k + i] = 0;
            f[1 * k + i] = 0;
            f[2 * k + i] = 2.0 * f[2 * k + i];
        }
    }
}

/* cem_maxwell_flux3d src/cem_maxwell.F:922-1002 */
static void flux3d(ora_state *s)
{
    size_t k = s->nxzfl;
    int n = s->nxzfl;
    double *f = s->srflx;
    double C0 = 0.0;
    if (s->ifcentral) C0 = 0.0;
    if (s->ifupwind) C0 = 1.0;
    const double *unx = s->unxm, *uny = s->unym, *unz = s->unzm;
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        f[0 * k + i] = -uny[i] * FEN(2)[i] + unz[i] * FEN(1)[i];
        f[1 * k + i] = -unz[i] * FEN(0)[i] + unx[i] * FEN(2)[i];
        f[2 * k + i] = -unx[i] * FEN(1)[i] + uny[i] * FEN(0)[i];
        f[3 * k + i] = -uny[i] * FHN(2)[i] + unz[i] * FHN(1)[i];
        f[4 * k + i] = -unz[i] * FHN(0)[i] + unx[i] * FHN(2)[i];
        f[5 * k + i] = -unx[i] * FHN(1)[i] + uny[i] * FHN(0)[i];
    }
    /* :958-959 -- note the argument aliasing: "H" args are slots 3..5, "E" args 0..2 */
    if (s->userfsrc)
        s->userfsrc(s->rktime, f + 3 * k, f + 4 * k, f + 5 * k, f + 0 * k, f + 1 * k, f + 2 * k,
                    s->ctx);
    ora_gs_op_fields(s->gsh_face, f, s->nxzfl, 6, 1);
    if (s->ifpec || s->ifpml) flux_pec(s);
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        double Y0 = s->Y_0[i], Y1 = s->Y_1[i], Z0 = s->Z_0[i], Z1 = s->Z_1[i];
        double Y02 = -0.5 / Y0 * Y1;
        double Z02 = 0.5 / Z0 * Z1;
        double C02Y = 0.5 / Y0 * C0;
        double C02Z = 0.5 / Z0 * C0;
        double fu1 = uny[i] * f[5 * k + i] - unz[i] * f[4 * k + i];
        double fu2 = unz[i] * f[3 * k + i] - unx[i] * f[5 * k + i];
        double fu3 = unx[i] * f[4 * k + i] - uny[i] * f[3 * k + i];
        double fw1 = uny[i] * f[2 * k + i] - unz[i] * f[1 * k + i];
        double fw2 = unz[i] * f[0 * k + i] - unx[i] * f[2 * k + i];
        double fw3 = unx[i] * f[1 * k + i] - uny[i] * f[0 * k + i];
        f[0 * k + i] = Y02 * f[0 * k + i] - C02Y * fu1;
        f[1 * k + i] = Y02 * f[1 * k + i] - C02Y * fu2;
        f[2 * k + i] = Y02 * f[2 * k + i] - C02Y * fu3;
        f[3 * k + i] = Z02 * f[3 * k + i] - C02Z * fw1;
        f[4 * k + i] = Z02 * f[4 * k + i] - C02Z * fw2;
        f[5 * k + i] = Z02 * f[5 * k + i] - C02Z * fw3;
    }
}

/* cem_maxwell_flux2d src/cem_maxwell.F:811-920 */
static void flux2d(ora_state *s)
{
    size_t k = s->nxzfl;
    int n = s->nxzfl;
    double *f = s->srflx;
    double C0 = 0.0;
    if (s->ifcentral) C0 = 0.0;
    if (s->ifupwind) C0 = 1.0;
    const double *unx = s->unxm, *uny = s->unym;
    if (s->imode == 2) { /* TM */
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            f[0 * k + i] = -uny[i] * FEN(2)[i];
            f[1 * k + i] = unx[i] * FEN(2)[i];
            f[2 * k + i] = -unx[i] * FHN(1)[i] + uny[i] * FHN(0)[i];
        }
        if (s->userfsrc)
            s->userfsrc(s->rktime, f + 3 * k, f + 4 * k, f + 2 * k, f + 0 * k, f + 1 * k,
                        f + 5 * k, s->ctx);
        ora_gs_op_fields(s->gsh_face, f, s->nxzfl, 3, 1);
        if (s->ifpml || s->ifpec) flux_pec(s);
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            double Y0 = s->Y_0[i], Y1 = s->Y_1[i], Z0 = s->Z_0[i], Z1 = s->Z_1[i];
            double fu1 = uny[i] * f[2 * k + i];
            double fu2 = -unx[i] * f[2 * k + i];
            double fw3 = unx[i] * f[1 * k + i] - uny[i] * f[0 * k + i];
            f[0 * k + i] = 0.5 / Y0 * (-Y1 * f[0 * k + i] - C0 * fu1);
            f[1 * k + i] = 0.5 / Y0 * (-Y1 * f[1 * k + i] - C0 * fu2);
            f[2 * k + i] = 0.5 / Z0 * (Z1 * f[2 * k + i] - C0 * fw3);
        }
    } else if (s->imode == 1) { /* TE */
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            f[0 * k + i] = -uny[i] * FHN(2)[i];
            f[1 * k + i] = unx[i] * FHN(2)[i];
            f[2 * k + i] = -unx[i] * FEN(1)[i] + uny[i] * FEN(0)[i];
        }
        if (s->userfsrc)
            s->userfsrc(s->rktime, f + 0 * k, f + 1 * k, f + 3 * k, f + 4 * k, f + 5 * k,
                        f + 2 * k, s->ctx);
        ora_gs_op_fields(s->gsh_face, f, s->nxzfl, 3, 1);
        if (s->ifpml || s->ifpec) flux_pec(s);
#pragma omp parallel for
        for (int i = 0; i < n; i++) {
            double Y0 = s->Y_0[i], Y1 = s->Y_1[i], Z0 = s->Z_0[i], Z1 = s->Z_1[i];
            double fw1 = uny[i] * f[2 * k + i];
            double fw2 = -unx[i] * f[2 * k + i];
            double fu3 = unx[i] * f[1 * k + i] - uny[i] * f[0 * k + i];
            f[0 * k + i] = 0.5 / Z0 * (Z1 * f[0 * k + i] - C0 * fw1);
            f[1 * k + i] = 0.5 / Z0 * (Z1 * f[1 * k + i] - C0 * fw2);
            f[2 * k + i] = 0.5 / Y0 * (-Y1 * f[2 * k + i] - C0 * fu3);
        }
    }
}

/* cem_maxwell_flux src/cem_maxwell.F:758-773 */
ORA_API void ora_flux(ora_state *s)
{
    if (s->ldim == 3) flux3d(s);
    else flux2d(s);
}

/* cem_maxwell_add_flux_to_res src/cem_maxwell.F:654-756 (CPU branch :725-752).
 * Sequential in j like the reference: edge/corner nodes receive their 2/3 face
 * contributions in ascending face-point order. */
ORA_API void ora_add_flux_to_res(ora_state *s)
{
    size_t k = s->nxzfl;
    const double *f = s->srflx;
    int nfp = s->nxzf * s->nfaces;
#pragma omp parallel for
    for (int e = 0; e < s->nelt; e++)
        for (int j = e * nfp; j < (e + 1) * nfp; j++) {
            int i = s->cemface[j];
            double a = s->aream[j];
            if (s->imode == 3) {
                RESHN(0)[i] = RESHN(0)[i] + a * f[0 * k + j];
                RESHN(1)[i] = RESHN(1)[i] + a * f[1 * k + j];
                RESHN(2)[i] = RESHN(2)[i] + a * f[2 * k + j];
                RESEN(0)[i] = RESEN(0)[i] + a * f[3 * k + j];
                RESEN(1)[i] = RESEN(1)[i] + a * f[4 * k + j];
                RESEN(2)[i] = RESEN(2)[i] + a * f[5 * k + j];
            } else if (s->imode == 2) {
                RESHN(0)[i] = RESHN(0)[i] + a * f[0 * k + j];
                RESHN(1)[i] = RESHN(1)[i] + a * f[1 * k + j];
                RESEN(2)[i] = RESEN(2)[i] + a * f[2 * k + j];
            } else {
                RESEN(0)[i] = RESEN(0)[i] + a * f[0 * k + j];
                RESEN(1)[i] = RESEN(1)[i] + a * f[1 * k + j];
                RESHN(2)[i] = RESHN(2)[i] + a * f[2 * k + j];
            }
        }
}

/* pml_step src/cem_maxwell_pml.F:508-592.  bm1 == bmn (flat copy, cem_maxwell.F:139). */
ORA_API void ora_pml_step(ora_state *s)
{
    int nxyz = s->nxyz;
    size_t np = s->npts;
#define P3(a, c) ((a) + (size_t)(c)*np)
#pragma omp parallel for
    for (int ie = 0; ie < s->maxpml; ie++) {
        int e = s->pmlptr[ie];
        for (int i = 0; i < nxyz; i++) {
            size_t j = i + (size_t)nxyz * e;
            double bm1 = s->bmn[j];
            double bm1inv = 1.0 / bm1;
            double sigx = P3(s->pmlsigma, 0)[j];
            double sigy = P3(s->pmlsigma, 1)[j];
            double sigz = P3(s->pmlsigma, 2)[j];
            double sigx_permitt = sigx / s->permittivity[j];
            double sigy_permitt = sigy / s->permittivity[j];
            double sigz_permitt = sigz / s->permittivity[j];
            double permeab = s->permeability[j];

            P3(s->respmlbn, 0)[j] = RESHN(0)[j] * bm1inv - sigy_permitt * P3(s->pmlbn, 0)[j];
            P3(s->respmlbn, 1)[j] = RESHN(1)[j] * bm1inv - sigz_permitt * P3(s->pmlbn, 1)[j];
            P3(s->respmlbn, 2)[j] = RESHN(2)[j] * bm1inv - sigx_permitt * P3(s->pmlbn, 2)[j];
            P3(s->respmldn, 0)[j] = RESEN(0)[j] * bm1inv - sigy_permitt * P3(s->pmldn, 0)[j];
            P3(s->respmldn, 1)[j] = RESEN(1)[j] * bm1inv - sigz_permitt * P3(s->pmldn, 1)[j];
            P3(s->respmldn, 2)[j] = RESEN(2)[j] * bm1inv - sigx_permitt * P3(s->pmldn, 2)[j];

            P3(s->respmlhn, 0)[j] = -sigy_permitt * P3(s->pmlbn, 0)[j] +
                                    sigx_permitt * P3(s->pmlbn, 0)[j] -
                                    sigz_permitt * permeab * HN(0)[j];
            P3(s->respmlhn, 1)[j] = -sigz_permitt * P3(s->pmlbn, 1)[j] +
                                    sigy_permitt * P3(s->pmlbn, 1)[j] -
                                    sigx_permitt * permeab * HN(1)[j];
            P3(s->respmlhn, 2)[j] = -sigx_permitt * P3(s->pmlbn, 2)[j] +
                                    sigz_permitt * P3(s->pmlbn, 2)[j] -
                                    sigy_permitt * permeab * HN(2)[j];

            P3(s->respmlen, 0)[j] = -sigy_permitt * P3(s->pmldn, 0)[j] +
                                    sigx_permitt * P3(s->pmldn, 0)[j] - sigz * EN(0)[j];
            P3(s->respmlen, 1)[j] = -sigz_permitt * P3(s->pmldn, 1)[j] +
                                    sigy_permitt * P3(s->pmldn, 1)[j] - sigx * EN(1)[j];
            P3(s->respmlen, 2)[j] = -sigx_permitt * P3(s->pmldn, 2)[j] +
                                    sigz_permitt * P3(s->pmldn, 2)[j] - sigy * EN(2)[j];

            RESHN(0)[j] = RESHN(0)[j] + P3(s->respmlhn, 0)[j] * bm1;
            RESHN(1)[j] = RESHN(1)[j] + P3(s->respmlhn, 1)[j] * bm1;
            RESHN(2)[j] = RESHN(2)[j] + P3(s->respmlhn, 2)[j] * bm1;
            RESEN(0)[j] = RESEN(0)[j] + P3(s->respmlen, 0)[j] * bm1;
            RESEN(1)[j] = RESEN(1)[j] + P3(s->respmlen, 1)[j] * bm1;
            RESEN(2)[j] = RESEN(2)[j] + P3(s->respmlen, 2)[j] * bm1;
        }
    }
}

/* cem_maxwell_invqmass src/cem_maxwell.F:1833-1904 */
ORA_API void ora_invqmass(ora_state *s)
{
#pragma omp parallel for
    for (int i = 0; i < s->npts; i++) {
        if (s->imode == 3) {
            RESHN(0)[i] = RESHN(0)[i] * s->hbm1[i];
            RESHN(1)[i] = RESHN(1)[i] * s->hbm1[i];
            RESHN(2)[i] = RESHN(2)[i] * s->hbm1[i];
            RESEN(0)[i] = RESEN(0)[i] * s->ebm1[i];
            RESEN(1)[i] = RESEN(1)[i] * s->ebm1[i];
            RESEN(2)[i] = RESEN(2)[i] * s->ebm1[i];
        } else if (s->imode == 2) {
            RESHN(0)[i] = RESHN(0)[i] * s->hbm1[i];
            RESHN(1)[i] = RESHN(1)[i] * s->hbm1[i];
            RESEN(2)[i] = RESEN(2)[i] * s->ebm1[i];
        } else {
            RESEN(0)[i] = RESEN(0)[i] * s->ebm1[i];
            RESEN(1)[i] = RESEN(1)[i] * s->ebm1[i];
            RESHN(2)[i] = RESHN(2)[i] * s->hbm1[i];
        }
    }
}

/* rk4_upd src/cem_common.F:18-76 (the 4-way unrolling does not change arithmetic) */
ORA_API void ora_rk4_upd(double *h, double *kh, const double *resh, double cb, double ca,
                         double dt, int n)
{
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        kh[i] = ca * kh[i] + dt * resh[i];
        h[i] = h[i] + cb * kh[i];
    }
}

/* rk_maxwell_ab src/cem_maxwell.F:1906-1968 */
ORA_API void ora_rk_maxwell_ab(ora_state *s, int ii /* 1-based */)
{
    double ca = s->rk4a[ii - 1], cb = s->rk4b[ii - 1], dt = s->dt;
    int npts = s->npts;
    if (s->imode == 3) {
        for (int c = 0; c < 3; c++) ora_rk4_upd(HN(c), KHN(c), RESHN(c), cb, ca, dt, npts);
        for (int c = 0; c < 3; c++) ora_rk4_upd(EN(c), KEN(c), RESEN(c), cb, ca, dt, npts);
    } else if (s->imode == 2) {
        ora_rk4_upd(HN(0), KHN(0), RESHN(0), cb, ca, dt, npts);
        ora_rk4_upd(HN(1), KHN(1), RESHN(1), cb, ca, dt, npts);
        ora_rk4_upd(EN(2), KEN(2), RESEN(2), cb, ca, dt, npts);
    } else {
        ora_rk4_upd(EN(0), KEN(0), RESEN(0), cb, ca, dt, npts);
        ora_rk4_upd(EN(1), KEN(1), RESEN(1), cb, ca, dt, npts);
        ora_rk4_upd(HN(2), KHN(2), RESHN(2), cb, ca, dt, npts);
    }
    if (s->ifpml) {
        size_t np = s->npts;
        for (int c = 0; c < 3; c++)
            ora_rk4_upd(s->pmlbn + c * np, s->kpmlbn + c * np, s->respmlbn + c * np, cb, ca, dt,
                        npts);
        for (int c = 0; c < 3; c++)
            ora_rk4_upd(s->pmldn + c * np, s->kpmldn + c * np, s->respmldn + c * np, cb, ca, dt,
                        npts);
    }
}

/* cem_maxwell_drude src/cem_maxwell.F:3095-3147.  jn,kjn,resjn: (npts,3); params: (npts,2);
 * dindex: 0-based node list.  Called from the user's usersrc like the reference. */
ORA_API void ora_cem_maxwell_drude(ora_state *s, double *jn, double *kjn, double *resjn,
                                   const double *params, const int *dindex, int n)
{
    size_t np = s->npts;
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        int j = dindex[i];
        double a = params[j], b = params[np + j];
        RESEN(0)[j] = RESEN(0)[j] - jn[j] * s->bmn[j];
        RESEN(1)[j] = RESEN(1)[j] - jn[np + j] * s->bmn[j];
        RESEN(2)[j] = RESEN(2)[j] - jn[2 * np + j] * s->bmn[j];
        resjn[j] = -a * jn[j] + b * EN(0)[j];
        resjn[np + j] = -a * jn[np + j] + b * EN(1)[j];
        resjn[2 * np + j] = -a * jn[2 * np + j] + b * EN(2)[j];
    }
    double ca = s->rk4a[s->rkstep - 1], cb = s->rk4b[s->rkstep - 1];
    for (int c = 0; c < 3; c++)
        ora_rk4_upd(jn + c * np, kjn + c * np, resjn + c * np, cb, ca, s->dt, s->npts);
}

/* cem_maxwell_lorentz src/cem_maxwell.F:3149-3211.  jn,kjn,resjn: (npts,3,2); params (npts,3) */
ORA_API void ora_cem_maxwell_lorentz(ora_state *s, double *jn, double *kjn, double *resjn,
                                     const double *params, const int *lindex, int n)
{
    size_t np = s->npts;
#define J(a, c, q) ((a)[(size_t)j + np * ((c) + 3 * (q))])
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        int j = lindex[i];
        double a = params[j], b = params[np + j], c = params[2 * np + j];
        RESEN(0)[j] = RESEN(0)[j] - J(jn, 0, 0) * s->bmn[j];
        RESEN(1)[j] = RESEN(1)[j] - J(jn, 1, 0) * s->bmn[j];
        RESEN(2)[j] = RESEN(2)[j] - J(jn, 2, 0) * s->bmn[j];
        J(resjn, 0, 0) = -a * J(jn, 0, 0) - b * J(jn, 0, 1) + c * EN(0)[j];
        J(resjn, 1, 0) = -a * J(jn, 1, 0) - b * J(jn, 1, 1) + c * EN(1)[j];
        J(resjn, 2, 0) = -a * J(jn, 2, 0) - b * J(jn, 2, 1) + c * EN(2)[j];
        J(resjn, 0, 1) = J(jn, 0, 0);
        J(resjn, 1, 1) = J(jn, 1, 0);
        J(resjn, 2, 1) = J(jn, 2, 0);
    }
#undef J
    double ca = s->rk4a[s->rkstep - 1], cb = s->rk4b[s->rkstep - 1];
    for (int q = 0; q < 6; q++)
        ora_rk4_upd(jn + q * np, kjn + q * np, resjn + q * np, cb, ca, s->dt, s->npts);
}

/* cem_3d_graphene_current / cem_te_graphene_current / cem_tm_graphene_current
 * src/cem_maxwell.F:2827-2931, 2933-3022, 3024-3093 (which one: s->imode, as the shipped
 * userfsrc of tests/3dgraphene and tests/2dgraphene selects it).  fjn,kfjn,resfjn: (nxzfl,3,6);
 * params: (nxzfl,12); gindex: 0-based face points; yconduc: own-side conductance restricted to
 * the faces (src/cem_maxwell.F:285, COMMON /EMWAVE/).  Called from the user's userfsrc like the
 * reference: slot 1 of fjn is the algebraic total surface current the caller then subtracts from
 * its -(n x H) face source. */
ORA_API void ora_cem_graphene_current(ora_state *s, double *fjn, double *kfjn, double *resfjn,
                                      const double *params, const double *yconduc,
                                      const int *gindex, int n)
{
    size_t nf = s->nxzfl;
    const double *unx = s->unxm, *uny = s->unym, *unz = s->unzm;
#define FJ(a, c, q) ((a)[(size_t)j + nf * ((c) + 3 * (q))]) /* (j, c+1, q+1) */
#define PAR(q) (params[(size_t)j + nf * (q)])
    const int c0 = s->imode == 2 ? 2 : 0, c1 = s->imode == 1 ? 2 : 3; /* active components */
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        int j = gindex[i];
        double a_d = PAR(0), b_d = PAR(1), b_cp1 = PAR(2), a_211 = PAR(3), a_221 = PAR(4),
               b_11 = PAR(5), b_21 = PAR(6), b_cp2 = PAR(7), a_212 = PAR(8), a_222 = PAR(9),
               b_12 = PAR(10), b_22 = PAR(11);
        double nH[3] = {0, 0, 0}, nEn[3] = {0, 0, 0};
        if (s->imode == 3) {
            nH[0] = -uny[j] * FHN(2)[j] + unz[j] * FHN(1)[j];
            nH[1] = unx[j] * FHN(2)[j] - unz[j] * FHN(0)[j];
            nH[2] = -unx[j] * FHN(1)[j] + uny[j] * FHN(0)[j];
            double ndotE = unx[j] * FEN(0)[j] + uny[j] * FEN(1)[j] + unz[j] * FEN(2)[j];
            nEn[0] = FEN(0)[j] - unx[j] * ndotE;
            nEn[1] = FEN(1)[j] - uny[j] * ndotE;
            nEn[2] = FEN(2)[j] - unz[j] * ndotE;
        } else if (s->imode == 1) { /* TE :2969-2976 */
            nH[0] = -uny[j] * FHN(2)[j];
            nH[1] = unx[j] * FHN(2)[j];
            nEn[0] = (uny[j] * uny[j]) * FEN(0)[j] - unx[j] * uny[j] * FEN(1)[j];
            nEn[1] = (unx[j] * unx[j]) * FEN(1)[j] - unx[j] * uny[j] * FEN(0)[j];
        } else { /* TM :3061-3065: n x (E x n) = E */
            nH[2] = -unx[j] * FHN(1)[j] + uny[j] * FHN(0)[j];
            nEn[2] = FEN(2)[j];
        }
        double Yfac = 0.5 / s->Y_0[j];
        double cpfac = b_cp1 + b_cp2;
        double jnfac = 1.0 - cpfac * Yfac;
        for (int c = c0; c < c1; c++) {
            double tmp = Yfac * (nH[c] + yconduc[j] * nEn[c]);
            FJ(fjn, c, 0) = (FJ(fjn, c, 1) + FJ(fjn, c, 2) + FJ(fjn, c, 4) - cpfac * tmp) / jnfac;
            double f = tmp - Yfac * FJ(fjn, c, 0);
            FJ(resfjn, c, 1) = -a_d * FJ(fjn, c, 1) + b_d * f;
            FJ(resfjn, c, 2) = FJ(fjn, c, 3) + b_11 * f;
            FJ(resfjn, c, 3) = -a_211 * FJ(fjn, c, 2) - a_221 * FJ(fjn, c, 3) + b_21 * f;
            FJ(resfjn, c, 4) = FJ(fjn, c, 5) + b_12 * f;
            FJ(resfjn, c, 5) = -a_212 * FJ(fjn, c, 4) - a_222 * FJ(fjn, c, 5) + b_22 * f;
        }
    }
#undef FJ
#undef PAR
    double ca = s->rk4a[s->rkstep - 1], cb = s->rk4b[s->rkstep - 1];
    for (int q = 1; q < 6; q++)
        for (int c = c0; c < c1; c++) {
            size_t o = nf * (c + 3 * q);
            ora_rk4_upd(fjn + o, kfjn + o, resfjn + o, cb, ca, s->dt, s->nxzfl);
        }
}

/* cem_maxwell_op src/cem_maxwell.F:484-508 */
ORA_API void ora_cem_maxwell_op(ora_state *s)
{
    ora_cem_maxwell(s);
    ora_restrict_to_face(s);
    if (s->userinc)
        s->userinc(s->rktime, FHN(0), FHN(1), FHN(2), FEN(0), FEN(1), FEN(2), s->ctx);
    ora_flux(s);
    ora_add_flux_to_res(s);
    if (s->ifpml) ora_pml_step(s);
    if (s->usersrc)
        s->usersrc(s->rktime, RESHN(0), RESHN(1), RESHN(2), RESEN(0), RESEN(1), RESEN(2),
                   s->ctx);
    ora_invqmass(s);
}

/* rk_storage src/cem_common.F:78-114 (ifrk45 branch) */
ORA_API void ora_rk_storage(ora_state *s)
{
    s->rk4a[0] = 0.0;
    s->rk4a[1] = -567301805773.0 / 1357537059087.0;
    s->rk4a[2] = -2404267990393.0 / 2016746695238.0;
    s->rk4a[3] = -3550918686646.0 / 2091501179385.0;
    s->rk4a[4] = -1275806237668.0 / 842570457699.0;
    s->rk4b[0] = 1432997174477.0 / 9575080441755.0;
    s->rk4b[1] = 5161836677717.0 / 13612068292357.0;
    s->rk4b[2] = 1720146321549.0 / 2090206949498.0;
    s->rk4b[3] = 3134564353537.0 / 4481467310338.0;
    s->rk4b[4] = 2277821191437.0 / 14882151754819.0;
    s->rk4c[0] = 0.0;
    s->rk4c[1] = 1432997174477.0 / 9575080441755.0;
    s->rk4c[2] = 2526269341429.0 / 6820363962896.0;
    s->rk4c[3] = 2006345519317.0 / 3224310063776.0;
    s->rk4c[4] = 2802321613138.0 / 2924317926251.0;
    s->rk4c[5] = 1.0;
}

/* cem_maxwell_op_rk src/cem_maxwell.F:327-345 with rk_c (src/cem_common.F:2-16).
 * The caller advances time (time_advancing_pde, src/cem_drive.F:618-654). */
ORA_API void ora_cem_maxwell_op_rk(ora_state *s)
{
    for (s->rkstep = 1; s->rkstep <= 5; s->rkstep++) {
        s->rktime = s->time + s->dt * s->rk4c[s->rkstep - 1];
        ora_cem_maxwell_op(s);
        ora_rk_maxwell_ab(s, s->rkstep);
    }
    s->rkstep = 5;
}

/* nsteps of time_advancing_pde (src/cem_drive.F:618-654) without userchk / output */
ORA_API void ora_advance(ora_state *s, int nsteps)
{
    for (int i = 0; i < nsteps; i++) {
        s->istep++;
        ora_cem_maxwell_op_rk(s);
        s->time = s->time + s->dt;
    }
}

/* filterq src/nek5_filter.F:92-144 applied to one field v(nxyz,nelt) with the n x n filter matrix f
 * (column-major, as build_new_filter returns it): v <- (F (x) F (x) F) v, every contraction in the
 * order of the reference's mxm calls (sum over the contracted index ascending). */
static void filterq_field(double *v, const double *f, int n, int nz, int nelt)
{
    int nxyz = n * n * nz;
#pragma omp parallel
    {
        double *w1 = (double *)malloc(sizeof(double) * nxyz);
        double *w2 = (double *)malloc(sizeof(double) * nxyz);
#pragma omp for
        for (int e = 0; e < nelt; e++) {
            double *ve = v + (size_t)nxyz * e;
            if (nz > 1) {
                /* w1 = F * v  (n x n)(n x n^2) */
                for (int jk = 0; jk < n * n; jk++)
                    for (int i = 0; i < n; i++) {
                        double sum = f[i] * ve[n * jk];
                        for (int m = 1; m < n; m++) sum = sum + f[i + n * m] * ve[m + n * jk];
                        w1[i + n * jk] = sum;
                    }
                /* per k: w2(:,:,k) = w1(:,:,k) * F^T */
                for (int k = 0; k < n; k++)
                    for (int j = 0; j < n; j++)
                        for (int i = 0; i < n; i++) {
                            const double *a = w1 + n * n * k;
                            double sum = a[i] * f[j];
                            for (int m = 1; m < n; m++) sum = sum + a[i + n * m] * f[j + n * m];
                            w2[i + n * j + n * n * k] = sum;
                        }
                /* w1 = w2 (n^2 x n) * F^T */
                for (int k = 0; k < n; k++)
                    for (int ij = 0; ij < n * n; ij++) {
                        double sum = w2[ij] * f[k];
                        for (int m = 1; m < n; m++) sum = sum + w2[ij + n * n * m] * f[k + n * m];
                        w1[ij + n * n * k] = sum;
                    }
            } else {
                /* w2 = F * v ; w1 = w2 * F^T */
                for (int j = 0; j < n; j++)
                    for (int i = 0; i < n; i++) {
                        double sum = f[i] * ve[n * j];
                        for (int m = 1; m < n; m++) sum = sum + f[i + n * m] * ve[m + n * j];
                        w2[i + n * j] = sum;
                    }
                for (int j = 0; j < n; j++)
                    for (int i = 0; i < n; i++) {
                        double sum = w2[i] * f[j];
                        for (int m = 1; m < n; m++) sum = sum + w2[i + n * m] * f[j + n * m];
                        w1[i + n * j] = sum;
                    }
            }
            for (int i = 0; i < nxyz; i++) ve[i] = w1[i];
        }
        free(w1);
        free(w2);
    }
}

/* q_filter src/nek5_filter.F:2-90 (MAXWELL branch): the six field components, E first */
ORA_API void ora_q_filter(ora_state *s, const double *intv)
{
    int nz = s->ldim == 3 ? s->nx1 : 1;
    for (int c = 0; c < 3; c++) filterq_field(EN(c), intv, s->nx1, nz, s->nelt);
    for (int c = 0; c < 3; c++) filterq_field(HN(c), intv, s->nx1, nz, s->nelt);
}

/* cem_error src/cem_common.F:1335-1355: l2 = sqrt(sum(err*bm1*err)/volvm1), linf = max|err| */
ORA_API void ora_cem_error(const double *u, const double *exact, double *error, int n,
                           const double *bm1, double volvm1, double *l2, double *linf)
{
    double sum = 0.0, mx = 0.0;
    for (int i = 0; i < n; i++) {
        error[i] = exact[i] - u[i];
        sum = sum + error[i] * bm1[i] * error[i];
        double a = fabs(error[i]);
        if (a > mx) mx = a;
    }
    double v = sum / volvm1;
    if (v > 0.0) v = sqrt(v);
    *l2 = v;
    *linf = mx;
}

/* get_dxmin src/nek5_courant.F:2-79 */
ORA_API double ora_get_dxmin(int ldim, int nx1, int nelt, const double *xm1, const double *ym1,
                             const double *zm1)
{
    int ny1 = nx1, nz1 = (ldim == 3) ? nx1 : 1;
    size_t nxyz = (size_t)nx1 * ny1 * nz1;
    double d2m = 1.e20;
#define X(a, i, j, k, e) a[(i) + nx1 * ((j) + ny1 * (k)) + nxyz * (e)]
    for (int e = 0; e < nelt; e++) {
        if (ldim == 3) {
            for (int k = 1; k < nz1 - 1; k++)
                for (int j = 1; j < ny1 - 1; j++)
                    for (int i = 1; i < nx1 - 1; i++) {
                        double dx, dy, dz, d2;
                        dx = X(xm1, i + 1, j, k, e) - X(xm1, i - 1, j, k, e);
                        dy = X(ym1, i + 1, j, k, e) - X(ym1, i - 1, j, k, e);
                        dz = X(zm1, i + 1, j, k, e) - X(zm1, i - 1, j, k, e);
                        d2 = dx * dx + dy * dy + dz * dz;
                        d2m = d2 < d2m ? d2 : d2m;
                        dx = X(xm1, i, j + 1, k, e) - X(xm1, i, j - 1, k, e);
                        dy = X(ym1, i, j + 1, k, e) - X(ym1, i, j - 1, k, e);
                        dz = X(zm1, i, j + 1, k, e) - X(zm1, i, j - 1, k, e);
                        d2 = dx * dx + dy * dy + dz * dz;
                        d2m = d2 < d2m ? d2 : d2m;
                        dx = X(xm1, i, j, k + 1, e) - X(xm1, i, j, k - 1, e);
                        dy = X(ym1, i, j, k + 1, e) - X(ym1, i, j, k - 1, e);
                        dz = X(zm1, i, j, k + 1, e) - X(zm1, i, j, k - 1, e);
                        d2 = dx * dx + dy * dy + dz * dz;
                        d2m = d2 < d2m ? d2 : d2m;
                    }
        } else {
            for (int j = 1; j < ny1 - 1; j++)
                for (int i = 1; i < nx1 - 1; i++) {
                    double dx, dy, d2;
                    dx = X(xm1, i + 1, j, 0, e) - X(xm1, i - 1, j, 0, e);
                    dy = X(ym1, i + 1, j, 0, e) - X(ym1, i - 1, j, 0, e);
                    d2 = dx * dx + dy * dy;
                    d2m = d2 < d2m ? d2 : d2m;
                    dx = X(xm1, i, j + 1, 0, e) - X(xm1, i, j - 1, 0, e);
                    dy = X(ym1, i, j + 1, 0, e) - X(ym1, i, j - 1, 0, e);
                    d2 = dx * dx + dy * dy;
                    d2m = d2 < d2m ? d2 : d2m;
                }
        }
    }
#undef X
    return sqrt(d2m) / 2.;
}

ORA_API int ora_state_size(void) { return (int)sizeof(ora_state); }

ORA_API int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
