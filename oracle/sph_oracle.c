/* CPU oracle for the pyticles SPH step hot path -- plain C restatement (scalar, one thread).
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs (through oracle/c_oracle.py).  Never linked into or
 * called from the product library.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks every entry point against the
 * golden vectors produced by the reference itself (tests/golden/, see make_golden.py) and
 * against the numpy restatement oracle/oracle.py.
 *
 * Reference semantics restated (file:line into /root/reference):
 *   oracle_build_pairs    VerletList.build            neighbour_list.py:160-189
 *                         minimum_image               neighbour_list.py:105-123
 *   oracle_separations    NeighbourList.separations   neighbour_list.py:63-83  (fp64)
 *   oracle_density_eos    spam_properties             properties.py:63-120
 *                         lucy_kernel                 spkernel.py:86-118
 *                         vdw/vdw_energy/vdw_temp     properties.py:38-49
 *   oracle_force          SpamForce.apply(_force)     forces.py:327-368 (2-D: :246-274)
 *   oracle_ponder_rebuild VerletList.ponder_rebuild   neighbour_list.py:225-234
 * The reference scans all n(n-1)/2 pairs; oracle_build_pairs prunes candidates with a
 * periodic cell grid first (a strict superset of the pairs the predicate can accept) and
 * then applies the reference's own predicate to the reference's own operands, emitting the
 * pairs in the reference's lexicographic i<j order.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared oracle/sph_oracle.c -o oracle/libsph_oracle.so -lm
 * (-ffp-contract=off: no FMA contraction, so every product and sum rounds as numpy's do.)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline double min_image1(double d, double L)
{                                            /* neighbour_list.py:111-122 */
    if (d > L / 2.) d = d - L;
    if (d < -L / 2.) d = d + L;
    return d;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

static inline long long cell_coord(double x, double inv_w, int nc)
{
    double f = floor(x * inv_w);
    long long c;
    if (!(f > -9.0e15 && f < 9.0e15)) return 0;      /* NaN / absurd: any cell will do */
    c = (long long)f % nc;
    if (c < 0) c += nc;
    return c;
}

/* Returns the number of pairs found (which may exceed cap; only the first cap are stored). */
long long oracle_build_pairs(int n, const double *r, const double *box, double thr,
                             int *iap, long long cap)
{
    int nc[3], d, i;
    double inv_w[3];
    double rl = sqrt(thr) * (1.0 + 1.0e-6);
    long long ncell = 1, nip = 0;
    int *cell_of, *start, *fill, *order, *cand;
    int ncand_cap = 1024;

    for (d = 0; d < 3; ++d) {
        double q = floor(box[d] / rl);
        nc[d] = (q >= 1.0 && q < 1024.0) ? (int)q : (q >= 1024.0 ? 1024 : 1);
        inv_w[d] = nc[d] / box[d];
        ncell *= nc[d];
    }
    cell_of = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    fill = (int *)calloc((size_t)ncell, sizeof(int));
    order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    cand = (int *)malloc(sizeof(int) * (size_t)ncand_cap);
    for (i = 0; i < n; ++i) {
        long long cx = cell_coord(r[3 * i + 0], inv_w[0], nc[0]);
        long long cy = cell_coord(r[3 * i + 1], inv_w[1], nc[1]);
        long long cz = cell_coord(r[3 * i + 2], inv_w[2], nc[2]);
        cell_of[i] = (int)((cz * nc[1] + cy) * nc[0] + cx);
        start[cell_of[i] + 1]++;
    }
    for (i = 0; i < ncell; ++i) start[i + 1] += start[i];
    for (i = 0; i < n; ++i) order[start[cell_of[i]] + fill[cell_of[i]]++] = i;

    for (i = 0; i < n; ++i) {
        int c = cell_of[i];
        int cx = c % nc[0], cy = (c / nc[0]) % nc[1], cz = c / (nc[0] * nc[1]);
        int ncand = 0, ox, oy, oz, k;
        int lox = nc[0] >= 3 ? -1 : 0, hix = nc[0] >= 3 ? 1 : nc[0] - 1;
        int loy = nc[1] >= 3 ? -1 : 0, hiy = nc[1] >= 3 ? 1 : nc[1] - 1;
        int loz = nc[2] >= 3 ? -1 : 0, hiz = nc[2] >= 3 ? 1 : nc[2] - 1;
        for (oz = loz; oz <= hiz; ++oz)
            for (oy = loy; oy <= hiy; ++oy)
                for (ox = lox; ox <= hix; ++ox) {
                    int ax = nc[0] >= 3 ? (cx + ox + nc[0]) % nc[0] : ox;
                    int ay = nc[1] >= 3 ? (cy + oy + nc[1]) % nc[1] : oy;
                    int az = nc[2] >= 3 ? (cz + oz + nc[2]) % nc[2] : oz;
                    int cc = (az * nc[1] + ay) * nc[0] + ax;
                    for (k = start[cc]; k < start[cc + 1]; ++k) {
                        int j = order[k];
                        double dx, dy, dz, rsq;
                        if (j <= i) continue;
                        dx = min_image1(r[3 * j + 0] - r[3 * i + 0], box[0]);   /* :170-176 */
                        dy = min_image1(r[3 * j + 1] - r[3 * i + 1], box[1]);
                        dz = min_image1(r[3 * j + 2] - r[3 * i + 2], box[2]);
                        rsq = dx * dx + dy * dy + dz * dz;                      /* :177 */
                        if (rsq < thr) {                                        /* :178 */
                            if (ncand == ncand_cap) {
                                ncand_cap *= 2;
                                cand = (int *)realloc(cand, sizeof(int) * (size_t)ncand_cap);
                            }
                            cand[ncand++] = j;
                        }
                    }
                }
        qsort(cand, (size_t)ncand, sizeof(int), cmp_int);
        for (k = 0; k < ncand; ++k) {
            if (nip < cap) {
                iap[2 * nip + 0] = i;
                iap[2 * nip + 1] = cand[k];
            }
            ++nip;
        }
    }
    free(cell_of); free(start); free(fill); free(order); free(cand);
    return nip;
}

void oracle_separations(long long nip, const int *iap, const double *r, const double *v,
                        const double *box, double *drij, double *rij, double *rsq, double *dv)
{
    long long k;
    for (k = 0; k < nip; ++k) {
        int i = iap[2 * k], j = iap[2 * k + 1], c;
        double d[3], s;
        for (c = 0; c < 3; ++c) {
            d[c] = min_image1(r[3 * j + c] - r[3 * i + c], box[c]);
            drij[3 * k + c] = d[c];
            dv[3 * k + c] = v[3 * j + c] - v[3 * i + c];
        }
        s = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        rsq[k] = s;
        rij[k] = sqrt(s);
    }
}

static void lucy3(double r, const double *dx, double h, double *w, double *dw)
{                                            /* spkernel.py:86-118, 3 columns */
    double q = 105. / (M_PI * 16. * pow(h, 3));
    if (r < 0) r = fabs(r);
    dw[0] = dw[1] = dw[2] = 0.0;
    if (r < h) {
        *w = q * (1 + 3. * r / h) * pow(1. - r / h, 3);
        if (r != 0) {
            double f = q * ((-12. / pow(h, 4)) * pow(r, 3) + (24. / pow(h, 3)) * pow(r, 2)
                            - (12. * r / pow(h, 2)));
            dw[0] = f * dx[0] / r;
            dw[1] = f * dx[1] / r;
            dw[2] = f * dx[2] / r;
        }
    } else {
        *w = 0.0;
    }
}

void oracle_density_eos(int n, long long nip, const int *iap, const double *rij,
                        const double *drij, const double *m, const double *h, const double *t,
                        double adash, double bdash, double kbdash,
                        double *rho, double *p, double *pco, double *u, double *tout,
                        double *wij, double *dwij)
{
    long long k;
    int i;
    double zero[3] = {0., 0., 0.}, zk, dum[3];
    if (n <= 0) return;
    lucy3(0.0, zero, h[0], &zk, dum);                       /* properties.py:76 */
    for (i = 0; i < n; ++i) rho[i] = zk;                    /* :77 */
    for (k = 0; k < nip; ++k) {
        int a = iap[2 * k], b = iap[2 * k + 1];
        double w, dw[3];
        lucy3(rij[k], drij + 3 * k, h[a], &w, dw);          /* :88 */
        if (wij) wij[k] = w;
        if (dwij) { dwij[3 * k] = dw[0]; dwij[3 * k + 1] = dw[1]; dwij[3 * k + 2] = dw[2]; }
        rho[a] += w * m[b];                                 /* :90 */
        rho[b] += w * m[a];                                 /* :91 */
    }
    for (i = 0; i < n; ++i) {
        p[i] = (rho[i] * kbdash * t[i]) / (1 - rho[i] * bdash);   /* :41 */
        pco[i] = -adash * rho[i] * rho[i];
        u[i] = t[i] * kbdash - adash * rho[i];                     /* :46, :119 */
        tout[i] = (u[i] + adash * rho[i]) / kbdash;                /* :49, :120 */
    }
}

void oracle_force(int n, long long nip, const int *iap, const double *rij, const double *dwij,
                  const double *dv, const double *m, const double *press, const double *rho,
                  double fcutoff, int dim, double *vdot, double *udot)
{
    long long k;
    double cutsq = fcutoff * fcutoff;
    (void)n;
    for (k = 0; k < nip; ++k) {
        int i, j;
        double ps, ax, ay, az, du;
        if (!(rij[k] * rij[k] <= cutsq)) continue;          /* forces.py:40 */
        i = iap[2 * k];
        j = iap[2 * k + 1];
        ps = press[i] / (rho[i] * rho[i]) + press[j] / (rho[j] * rho[j]);   /* :353 */
        ax = ps * dwij[3 * k];
        ay = ps * dwij[3 * k + 1];
        az = dim == 2 ? 0.0 : ps * dwij[3 * k + 2];
        vdot[3 * i] += ax; vdot[3 * i + 1] += ay; vdot[3 * i + 2] += az;
        vdot[3 * j] -= ax; vdot[3 * j + 1] -= ay; vdot[3 * j + 2] -= az;
        if (dim == 2) du = 0.5 * (ax * dv[3 * k] + ay * dv[3 * k + 1]);
        else du = 0.5 * (ax * dv[3 * k] + ay * dv[3 * k + 1] + az * dv[3 * k + 2]);   /* :366 */
        udot[i] += du * m[j];
        udot[j] += du * m[i];
    }
}

int oracle_ponder_rebuild(int n, const double *r_old, const double *r, double tol_sq)
{
    int i;
    double dsq = -INFINITY;
    for (i = 0; i < n; ++i) {
        double a = r_old[3 * i] - r[3 * i], b = r_old[3 * i + 1] - r[3 * i + 1],
               c = r_old[3 * i + 2] - r[3 * i + 2];
        double s = a * a + b * b + c * c;
        if (s > dsq) dsq = s;
    }
    return dsq > tol_sq;
}
