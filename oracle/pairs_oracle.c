/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement ("oracle") of the halotools pair-counting engines, in plain C,
 * strict IEEE-754 double arithmetic (build with -O2 -ffp-contract=off, no
 * fast-math).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this.  Parity status: PINNED — checked against
 * the unmodified reference engines built by oracle/build_ref.py and against the
 * golden vectors in tests/golden/ (generated through the reference's public API).
 *
 * Each routine follows the loop nest of the reference engine it names:
 *   cell1 -> window of neighbour cell2 (per-dimension unwrapped index, shift =
 *   -/+ period * PBCs when the index is <0 / >= ndivs2, index wrapped mod ndivs2)
 *   -> i in cell1 (x1tmp = x1 - shift) -> j in cell2 -> top-down bin scan.
 *
 *   npairs_3d        /root/reference/halotools/mock_observables/pair_counters/cpairs/npairs_3d_engine.pyx:103-182
 *   npairs_xy_z      .../cpairs/npairs_xy_z_engine.pyx:113-194
 *   npairs_s_mu      .../cpairs/npairs_s_mu_engine.pyx:120-234
 *   marked_npairs_3d .../marked_cpairs/marked_npairs_3d_engine.pyx:118-216
 *   weight functions .../marked_cpairs/marking_functions.pyx:14-217, custom_marking_func.pyx:12-16
 *   mean_delta_sigma /root/reference/halotools/mock_observables/surface_density/engines/mean_delta_sigma_engine.pyx:95-180
 *
 * The points are passed ALREADY SORTED by cell (x[idx_sorted]) together with the
 * cell offset tables, exactly what the reference engines build at entry
 * (npairs_3d_engine.pyx:58-67).  Mesh construction itself is restated with numpy
 * in oracle/mesh.py because its semantics ARE numpy's (float floor-division,
 * argsort, searchsorted).
 *
 * Threading: cell1 ranges are split over OpenMP threads (the reference splits
 * them over multiprocessing workers, npairs_3d.py:139-148); integer counts are
 * order independent; float sums are reduced per thread then added in thread
 * order (the reference adds per-worker partial sums the same way).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int ndivs1[3];     /* mesh1 num_{x,y,z}divs                     */
    int ndivs2[3];     /* mesh2 num_{x,y,z}divs                     */
    int cover[3];      /* ceil(search_length / mesh2 cell size)     */
    int pbc;           /* double_mesh._PBCs                         */
    double period[3];
} oracle_geom_t;

/* one dimension of the neighbour window of a cell1 index: wrapped cell2 index + shift */
typedef struct { int idx; double shift; } nbr_t;

static int fill_window(int i1, int per, int cover, int ndivs2, double period, int pbc, nbr_t *out)
{
    int lo = i1 * per - cover, hi = (i1 + 1) * per + cover, n = 0;
    for (int u = lo; u < hi; ++u) {
        double s = 0.0;
        if (u < 0) s = -period * pbc;
        else if (u >= ndivs2) s = period * pbc;
        int w = u % ndivs2;
        if (w < 0) w += ndivs2;           /* python modulo */
        out[n].idx = w;
        out[n].shift = s;
        ++n;
    }
    return n;
}

static int max_window(const oracle_geom_t *g)
{
    int m = 0;
    for (int d = 0; d < 3; ++d) {
        if (g->ndivs1[d] <= 0) continue;
        int w = g->ndivs2[d] / g->ndivs1[d] + 2 * g->cover[d];
        if (w > m) m = w;
    }
    return m;
}

/* ------------------------------------------------------------------ npairs_3d */
void oracle_npairs_3d(const oracle_geom_t *g,
                      const double *x1, const double *y1, const double *z1, const int64_t *off1,
                      const double *x2, const double *y2, const double *z2, const int64_t *off2,
                      const double *rbins, int nb, int64_t first_cell1, int64_t last_cell1,
                      int nthreads, int64_t *counts_out)
{
    double *rsq = (double *)malloc(sizeof(double) * nb);
    for (int k = 0; k < nb; ++k) rsq[k] = rbins[k] * rbins[k];
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    memset(counts_out, 0, sizeof(int64_t) * nb);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        int64_t *cnt = (int64_t *)calloc(nb, sizeof(int64_t));
        nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
#pragma omp for schedule(dynamic, 1)
        for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
            const int64_t a = off1[c1], b = off1[c1 + 1];
            if (b <= a) continue;
            const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
            const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
            const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
            const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
            const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
            const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
            for (int ax = 0; ax < nx; ++ax)
            for (int ay = 0; ay < ny; ++ay)
            for (int az = 0; az < nz; ++az) {
                const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
                const int64_t p = off2[c2], q = off2[c2 + 1];
                if (q <= p) continue;
                const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
                for (int64_t i = a; i < b; ++i) {
                    const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                    for (int64_t j = p; j < q; ++j) {
                        const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                        const double dsq = dx * dx + dy * dy + dz * dz;
                        int k = nb - 1;
                        while (dsq <= rsq[k]) { cnt[k] += 1; if (--k < 0) break; }
                    }
                }
            }
        }
#pragma omp critical
        for (int k = 0; k < nb; ++k) counts_out[k] += cnt[k];
        free(cnt); free(wx);
    }
    free(rsq);
}

/* ---------------------------------------------------------------- npairs_xy_z */
void oracle_npairs_xy_z(const oracle_geom_t *g,
                        const double *x1, const double *y1, const double *z1, const int64_t *off1,
                        const double *x2, const double *y2, const double *z2, const int64_t *off2,
                        const double *rp_bins, int nrp, const double *pi_bins, int npi,
                        int64_t first_cell1, int64_t last_cell1, int nthreads, int64_t *counts_out)
{
    double *rpsq = (double *)malloc(sizeof(double) * (nrp + npi)), *pisq = rpsq + nrp;
    for (int k = 0; k < nrp; ++k) rpsq[k] = rp_bins[k] * rp_bins[k];
    for (int k = 0; k < npi; ++k) pisq[k] = pi_bins[k] * pi_bins[k];
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    const int nc = nrp * npi;
    memset(counts_out, 0, sizeof(int64_t) * nc);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        int64_t *cnt = (int64_t *)calloc(nc, sizeof(int64_t));
        nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
#pragma omp for schedule(dynamic, 1)
        for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
            const int64_t a = off1[c1], b = off1[c1 + 1];
            if (b <= a) continue;
            const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
            const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
            const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
            const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
            const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
            const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
            for (int ax = 0; ax < nx; ++ax)
            for (int ay = 0; ay < ny; ++ay)
            for (int az = 0; az < nz; ++az) {
                const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
                const int64_t p = off2[c2], q = off2[c2 + 1];
                if (q <= p) continue;
                const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
                for (int64_t i = a; i < b; ++i) {
                    const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                    for (int64_t j = p; j < q; ++j) {
                        const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                        const double dxy_sq = dx * dx + dy * dy;
                        const double dz_sq = dz * dz;
                        int k = nrp - 1;
                        while (dxy_sq <= rpsq[k]) {
                            int q2 = npi - 1;
                            while (dz_sq <= pisq[q2]) { cnt[k * npi + q2] += 1; if (--q2 < 0) break; }
                            if (--k < 0) break;
                        }
                    }
                }
            }
        }
#pragma omp critical
        for (int k = 0; k < nc; ++k) counts_out[k] += cnt[k];
        free(cnt); free(wx);
    }
    free(rpsq);
}

/* ---------------------------------------------------------------- npairs_s_mu */
/* mu_bins are the already transformed sin(theta_los) edges (npairs_s_mu.py:174-175). */
void oracle_npairs_s_mu(const oracle_geom_t *g,
                        const double *x1, const double *y1, const double *z1, const int64_t *off1,
                        const double *x2, const double *y2, const double *z2, const int64_t *off2,
                        const double *s_bins, int ns, const double *mu_bins, int nmu,
                        int64_t first_cell1, int64_t last_cell1, int nthreads, int64_t *counts_out)
{
    double *ssq = (double *)malloc(sizeof(double) * (ns + nmu)), *musq = ssq + ns;
    double ssq_max = -INFINITY, musq_max = -INFINITY;
    for (int k = 0; k < ns; ++k) { ssq[k] = s_bins[k] * s_bins[k]; if (ssq[k] > ssq_max) ssq_max = ssq[k]; }
    for (int k = 0; k < nmu; ++k) { musq[k] = mu_bins[k] * mu_bins[k]; if (musq[k] > musq_max) musq_max = musq[k]; }
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    const int nc = ns * nmu;
    int64_t *diff = (int64_t *)calloc(nc, sizeof(int64_t));
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        int64_t *cnt = (int64_t *)calloc(nc, sizeof(int64_t));
        nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
#pragma omp for schedule(dynamic, 1)
        for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
            const int64_t a = off1[c1], b = off1[c1 + 1];
            if (b <= a) continue;
            const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
            const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
            const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
            const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
            const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
            const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
            for (int ax = 0; ax < nx; ++ax)
            for (int ay = 0; ay < ny; ++ay)
            for (int az = 0; az < nz; ++az) {
                const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
                const int64_t p = off2[c2], q = off2[c2 + 1];
                if (q <= p) continue;
                const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
                for (int64_t i = a; i < b; ++i) {
                    const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                    for (int64_t j = p; j < q; ++j) {
                        const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                        const double dxy_sq = dx * dx + dy * dy;
                        const double dz_sq = dz * dz;
                        const double sqr_s = dz_sq + dxy_sq;
                        if (sqr_s > ssq_max) continue;
                        double sqr_mu = 0.0;
                        if (sqr_s > 0.0) sqr_mu = dxy_sq / sqr_s;
                        if (sqr_mu > musq_max) continue;
                        int k = ns - 2;
                        while (k != -1) { if (sqr_s > ssq[k]) break; --k; }
                        int m = nmu - 2;
                        while (m != -1) { if (sqr_mu > musq[m]) break; --m; }
                        cnt[(k + 1) * nmu + (m + 1)] += 1;
                    }
                }
            }
        }
#pragma omp critical
        for (int k = 0; k < nc; ++k) diff[k] += cnt[k];
        free(cnt); free(wx);
    }
    /* 2-D inclusive prefix sum (npairs_s_mu_engine.pyx:232-234) */
    for (int k = 0; k < ns; ++k)
        for (int m = 0; m < nmu; ++m) {
            int64_t s = 0;
            for (int kk = 0; kk <= k; ++kk)
                for (int mm = 0; mm <= m; ++mm) s += diff[kk * nmu + mm];
            counts_out[k * nmu + m] = s;
        }
    free(diff); free(ssq);
}

/* ----------------------------------------------------------- weight functions */
static double pair_weight(int id, const double *w1, const double *w2)
{
    double d;
    switch (id) {
    case 0:  return w1[0] * w2[0];                               /* custom_func */
    case 1:  return w1[0] * w2[0];                               /* mweights */
    case 2:  return w1[0] + w2[0];                               /* sweights */
    case 3:  return (w1[0] == w2[0]) ? w1[1] * w2[1] : 0.0;      /* eqweights */
    case 4:  return (w1[0] != w2[0]) ? w1[1] * w2[1] : 0.0;      /* ineqweights */
    case 5:  return (w2[0] > w1[0]) ? w1[1] * w2[1] : 0.0;       /* gweights */
    case 6:  return (w2[0] < w1[0]) ? w1[1] * w2[1] : 0.0;       /* lweights */
    case 7:  return (w2[0] > (w1[0] + w1[1])) ? w2[1] : 0.0;     /* tgweights */
    case 8:  return (w2[0] < (w1[0] + w1[1])) ? w2[1] : 0.0;     /* tlweights (sic: '+', marking_functions.pyx:106) */
    case 9:  return (fabs(w1[0] - w2[0]) < w1[1]) ? w2[1] : 0.0; /* tweights */
    case 10: return (fabs(w1[0] - w2[0]) > w1[1]) ? w2[1] : 0.0; /* exweights */
    case 11: return (w2[0] > w1[0] * w1[1]) ? w2[1] : 0.0;       /* ratio_weights */
    case 12: return w1[0] * w2[0] * (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
    case 13: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]); return w1[0] * w2[0] * d * d;
    case 14: return w1[0] * w2[0] * (w1[1] * w2[1] + w1[2] * w2[2]);
    case 15: d = (w1[1] * w2[1] + w1[2] * w2[2]); return w1[0] * w2[0] * d * d;
    case 16: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
             return (w1[4] == w2[4]) ? w1[0] * w2[0] * d * d : 0.0;
    case 17: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
             return (w1[4] != w2[4]) ? w1[0] * w2[0] * d * d : 0.0;
    default: return NAN;
    }
}

/* ----------------------------------------------------------- marked_npairs_3d */
/* w1, w2: row-major (N, nw) weights in SORTED order. */
void oracle_marked_npairs_3d(const oracle_geom_t *g,
                             const double *x1, const double *y1, const double *z1, const int64_t *off1,
                             const double *x2, const double *y2, const double *z2, const int64_t *off2,
                             const double *w1, const double *w2, int nw, int wfunc_id,
                             const double *rbins, int nb, int64_t first_cell1, int64_t last_cell1,
                             int nthreads, double *counts_out)
{
    double *rsq = (double *)malloc(sizeof(double) * nb);
    for (int k = 0; k < nb; ++k) rsq[k] = rbins[k] * rbins[k];
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    if (nthreads < 1) nthreads = 1;
    double *partial = (double *)calloc((size_t)nthreads * nb, sizeof(double));
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        double *cnt = partial + (size_t)omp_get_thread_num() * nb;
#else
        double *cnt = partial;
#endif
        nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
#pragma omp for schedule(static)
        for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
            const int64_t a = off1[c1], b = off1[c1 + 1];
            if (b <= a) continue;
            const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
            const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
            const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
            const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
            const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
            const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
            for (int ax = 0; ax < nx; ++ax)
            for (int ay = 0; ay < ny; ++ay)
            for (int az = 0; az < nz; ++az) {
                const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
                const int64_t p = off2[c2], q = off2[c2 + 1];
                if (q <= p) continue;
                const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
                for (int64_t i = a; i < b; ++i) {
                    const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                    for (int64_t j = p; j < q; ++j) {
                        const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                        const double dsq = dx * dx + dy * dy + dz * dz;
                        const double w = pair_weight(wfunc_id, w1 + i * nw, w2 + j * nw);
                        int k = nb - 1;
                        while (dsq <= rsq[k]) { cnt[k] += w; if (--k < 0) break; }
                    }
                }
            }
        }
        free(wx);
    }
    for (int k = 0; k < nb; ++k) {
        double s = 0.0;
        for (int t = 0; t < nthreads; ++t) s += partial[(size_t)t * nb + k];
        counts_out[k] = s;
    }
    free(partial); free(rsq);
}

/* ----------------------------------------------------------- mean_delta_sigma */
/* 2-D mesh (cell id = ix*ny + iy); out is (n1, nrp-1) row-major in SORTED sample1
 * order, already divided by pi*(rp[k+1]^2 - rp[k]^2); rows of cells outside
 * [first_cell1, last_cell1) stay zero, like each reference worker's array.      */
void oracle_mean_delta_sigma(const oracle_geom_t *g,
                             const double *x1, const double *y1, const int64_t *off1, int64_t n1,
                             const double *x2, const double *y2, const double *m2, const int64_t *off2,
                             const double *rp_bins, int nrp, int64_t first_cell1, int64_t last_cell1,
                             int nthreads, double *out, double *absout)
{
    /* absout (optional, same shape as out): A_ik = the sum of the ABSOLUTE values of the terms accumulated into
     * out[i][k], normalised like out - the scale of the rounding error of the cancelling difference (SURVEY 8d
     * per-object parity gate).                                                                              */
    const int nbin = nrp - 1;
    double *rpsq = (double *)malloc(sizeof(double) * (nrp + nbin)), *dlog = rpsq + nrp;
    for (int k = 0; k < nrp; ++k) rpsq[k] = rp_bins[k] * rp_bins[k];
    for (int k = 0; k < nbin; ++k) dlog[k] = log(rp_bins[k + 1] / rp_bins[k]);
    const int ny1 = g->ndivs1[1], ny2 = g->ndivs2[1];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1;
    const int mw = max_window(g);
    memset(out, 0, sizeof(double) * (size_t)n1 * nbin);
    if (absout) memset(absout, 0, sizeof(double) * (size_t)n1 * nbin);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 2), *wy = wx + mw;
#pragma omp for schedule(dynamic, 1)
        for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
            const int64_t a = off1[c1], b = off1[c1 + 1];
            if (b <= a) continue;
            const int ix1 = (int)(c1 / ny1);
            const int iy1 = (int)(c1 - (int64_t)ix1 * ny1);
            const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
            const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
            for (int ax = 0; ax < nx; ++ax)
            for (int ay = 0; ay < ny; ++ay) {
                const int64_t c2 = (int64_t)wx[ax].idx * ny2 + wy[ay].idx;
                const int64_t p = off2[c2], q = off2[c2 + 1];
                if (q <= p) continue;
                const double sx = wx[ax].shift, sy = wy[ay].shift;
                for (int64_t i = a; i < b; ++i) {
                    const double xt = x1[i] - sx, yt = y1[i] - sy;
                    double *row = out + (size_t)i * nbin;
                    double *arow = absout ? absout + (size_t)i * nbin : NULL;
                    for (int64_t j = p; j < q; ++j) {
                        const double dx = xt - x2[j], dy = yt - y2[j];
                        const double dxy_sq = dx * dx + dy * dy;
                        const double m = m2[j];
                        int k = nbin - 1;
                        while (k >= 0 && dxy_sq <= rpsq[k + 1]) {
                            double t;
                            if (dxy_sq > rpsq[k]) { t = m * (1 - log(rpsq[k + 1] / dxy_sq)); row[k] -= t; }
                            else                  { t = m * 2 * dlog[k]; row[k] += t; }
                            if (arow) arow[k] += fabs(t);
                            --k;
                        }
                    }
                }
            }
        }
        free(wx);
    }
    for (int k = 0; k < nbin; ++k) {
        const double norm = M_PI * (rpsq[k + 1] - rpsq[k]);
        for (int64_t i = 0; i < n1; ++i) out[(size_t)i * nbin + k] /= norm;
        if (absout) for (int64_t i = 0; i < n1; ++i) absout[(size_t)i * nbin + k] /= norm;
    }
    free(rpsq);
}

/* ------------------------------------------------------------ the section-8(f) counters
 *   npairs_projected      .../cpairs/npairs_projected_engine.pyx:113-189
 *   npairs_per_object_3d  .../cpairs/npairs_per_object_3d_engine.pyx:118-213
 *   marked_npairs_xy_z    .../marked_cpairs/marked_npairs_xy_z_engine.pyx:126-225
 *   weighted_npairs_xy    /root/reference/halotools/mock_observables/surface_density/engines/weighted_npairs_xy_engine.pyx:95-175
 * One 3-D loop nest serves the first three (the reference repeats it verbatim in each engine); `mode` picks the
 * innermost statement.                                                                                        */
typedef struct {
    int mode;                 /* 0 projected, 1 per object, 2 marked (rp, pi) */
    const double *e0; int n0; /* squared rbins / rp_bins */
    const double *e1; int n1; /* squared pi_bins (mode 2) */
    double pi_max_sq;         /* mode 0 */
    const double *w1, *w2; int nw, wfunc;   /* mode 2: SORTED row-major weights */
} f_inner_t;

static void f_loop3(const oracle_geom_t *g,
                    const double *x1, const double *y1, const double *z1, const int64_t *off1,
                    const double *x2, const double *y2, const double *z2, const int64_t *off2,
                    int64_t first_cell1, int64_t last_cell1, const f_inner_t *in,
                    int64_t *icnt /* mode 0: [n0]; mode 1: (N1, n0) in sorted order */, double *fcnt /* mode 2: [n0*n1] */)
{
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
    for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
        const int64_t a = off1[c1], b = off1[c1 + 1];
        if (b <= a) continue;
        const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
        const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
        const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
        const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
        const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
        const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
        for (int ax = 0; ax < nx; ++ax)
        for (int ay = 0; ay < ny; ++ay)
        for (int az = 0; az < nz; ++az) {
            const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
            const int64_t p = off2[c2], q = off2[c2 + 1];
            if (q <= p) continue;
            const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
            for (int64_t i = a; i < b; ++i) {
                const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                for (int64_t j = p; j < q; ++j) {
                    const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                    if (in->mode == 0) {
                        /* npairs_projected_engine.pyx:180-189 */
                        const double dxy_sq = dx * dx + dy * dy;
                        const double dz_sq = dz * dz;
                        int k = in->n0 - 1;
                        while (dxy_sq <= in->e0[k]) {
                            if (dz_sq <= in->pi_max_sq) icnt[k] += 1;
                            if (--k < 0) break;
                        }
                    } else if (in->mode == 1) {
                        /* npairs_per_object_3d_engine.pyx:193-207 (inner counts folded into the row at once) */
                        const double dsq = dx * dx + dy * dy + dz * dz;
                        int k = in->n0 - 1;
                        while (dsq <= in->e0[k]) { icnt[i * in->n0 + k] += 1; if (--k < 0) break; }
                    } else {
                        /* marked_npairs_xy_z_engine.pyx:211-225 */
                        const double dxy_sq = dx * dx + dy * dy;
                        const double dz_sq = dz * dz;
                        const double w = pair_weight(in->wfunc, in->w1 + i * in->nw, in->w2 + j * in->nw);
                        int k = in->n0 - 1;
                        while (dxy_sq <= in->e0[k]) {
                            int gq = in->n1 - 1;
                            while (dz_sq <= in->e1[gq]) { fcnt[k * in->n1 + gq] += w; if (--gq < 0) break; }
                            if (--k < 0) break;
                        }
                    }
                }
            }
        }
    }
    free(wx);
}

static double *squares(const double *v, int n)
{
    double *s = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) s[k] = v[k] * v[k];
    return s;
}

void oracle_npairs_projected(const oracle_geom_t *g,
                             const double *x1, const double *y1, const double *z1, const int64_t *off1,
                             const double *x2, const double *y2, const double *z2, const int64_t *off2,
                             const double *rp_bins, int nrp, double pi_max,
                             int64_t first_cell1, int64_t last_cell1, int64_t *counts_out)
{
    f_inner_t in; memset(&in, 0, sizeof(in));
    double *e0 = squares(rp_bins, nrp);
    in.mode = 0; in.e0 = e0; in.n0 = nrp; in.pi_max_sq = pi_max * pi_max;
    memset(counts_out, 0, sizeof(int64_t) * nrp);
    f_loop3(g, x1, y1, z1, off1, x2, y2, z2, off2, first_cell1, last_cell1, &in, counts_out, NULL);
    free(e0);
}

/* counts_out: (n1, nb) row-major in SORTED sample1 order (the wrapper un-sorts, engine :209-213) */
void oracle_npairs_per_object_3d(const oracle_geom_t *g,
                                 const double *x1, const double *y1, const double *z1, const int64_t *off1, int64_t n1,
                                 const double *x2, const double *y2, const double *z2, const int64_t *off2,
                                 const double *rbins, int nb,
                                 int64_t first_cell1, int64_t last_cell1, int64_t *counts_out)
{
    f_inner_t in; memset(&in, 0, sizeof(in));
    double *e0 = squares(rbins, nb);
    in.mode = 1; in.e0 = e0; in.n0 = nb;
    memset(counts_out, 0, sizeof(int64_t) * (size_t)n1 * nb);
    f_loop3(g, x1, y1, z1, off1, x2, y2, z2, off2, first_cell1, last_cell1, &in, counts_out, NULL);
    free(e0);
}

void oracle_marked_npairs_xy_z(const oracle_geom_t *g,
                               const double *x1, const double *y1, const double *z1, const int64_t *off1,
                               const double *x2, const double *y2, const double *z2, const int64_t *off2,
                               const double *w1, const double *w2, int nw, int wfunc_id,
                               const double *rp_bins, int nrp, const double *pi_bins, int npi,
                               int64_t first_cell1, int64_t last_cell1, double *counts_out)
{
    f_inner_t in; memset(&in, 0, sizeof(in));
    double *e0 = squares(rp_bins, nrp), *e1 = squares(pi_bins, npi);
    in.mode = 2; in.e0 = e0; in.n0 = nrp; in.e1 = e1; in.n1 = npi;
    in.w1 = w1; in.w2 = w2; in.nw = nw; in.wfunc = wfunc_id;
    for (int k = 0; k < nrp * npi; ++k) counts_out[k] = 0.0;
    f_loop3(g, x1, y1, z1, off1, x2, y2, z2, off2, first_cell1, last_cell1, &in, NULL, counts_out);
    free(e0); free(e1);
}

/* 2-D mesh (cell id = ix*ny + iy), weighted_npairs_xy_engine.pyx:95-175 */
static void weighted_xy_loop(const oracle_geom_t *g,
                             const double *x1, const double *y1, const int64_t *off1,
                             const double *x2, const double *y2, const double *w2, const int64_t *off2,
                             const double *rp_bins, int nrp, int64_t first_cell1, int64_t last_cell1,
                             double *counts_out, double *rows_out);

void oracle_weighted_npairs_xy(const oracle_geom_t *g,
                               const double *x1, const double *y1, const int64_t *off1,
                               const double *x2, const double *y2, const double *w2, const int64_t *off2,
                               const double *rp_bins, int nrp, int64_t first_cell1, int64_t last_cell1,
                               double *counts_out)
{
    for (int k = 0; k < nrp; ++k) counts_out[k] = 0.0;
    weighted_xy_loop(g, x1, y1, off1, x2, y2, w2, off2, rp_bins, nrp, first_cell1, last_cell1, counts_out, NULL);
}

/* weighted_npairs_per_object_xy_engine.pyx:95-190: rows_out is (n1, nrp) in SORTED sample1 order (the wrapper un-sorts) */
void oracle_weighted_npairs_per_object_xy(const oracle_geom_t *g,
                                          const double *x1, const double *y1, const int64_t *off1, int64_t n1,
                                          const double *x2, const double *y2, const double *w2, const int64_t *off2,
                                          const double *rp_bins, int nrp, int64_t first_cell1, int64_t last_cell1,
                                          double *rows_out)
{
    for (int64_t k = 0; k < n1 * nrp; ++k) rows_out[k] = 0.0;
    weighted_xy_loop(g, x1, y1, off1, x2, y2, w2, off2, rp_bins, nrp, first_cell1, last_cell1, NULL, rows_out);
}

static void weighted_xy_loop(const oracle_geom_t *g,
                             const double *x1, const double *y1, const int64_t *off1,
                             const double *x2, const double *y2, const double *w2, const int64_t *off2,
                             const double *rp_bins, int nrp, int64_t first_cell1, int64_t last_cell1,
                             double *counts_out, double *rows_out)
{
    double *rpsq = squares(rp_bins, nrp);
    const int ny1 = g->ndivs1[1], ny2 = g->ndivs2[1];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1;
    const int mw = max_window(g);
    nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 2), *wy = wx + mw;
    for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
        const int64_t a = off1[c1], b = off1[c1 + 1];
        if (b <= a) continue;
        const int ix1 = (int)(c1 / ny1);
        const int iy1 = (int)(c1 - (int64_t)ix1 * ny1);
        const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
        const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
        for (int ax = 0; ax < nx; ++ax)
        for (int ay = 0; ay < ny; ++ay) {
            const int64_t c2 = (int64_t)wx[ax].idx * ny2 + wy[ay].idx;
            const int64_t p = off2[c2], q = off2[c2 + 1];
            if (q <= p) continue;
            const double sx = wx[ax].shift, sy = wy[ay].shift;
            for (int64_t i = a; i < b; ++i) {
                const double xt = x1[i] - sx, yt = y1[i] - sy;
                for (int64_t j = p; j < q; ++j) {
                    const double dx = xt - x2[j], dy = yt - y2[j];
                    const double dxy_sq = dx * dx + dy * dy;
                    const double w2tmp = w2[j];
                    int k = nrp - 1;
                    double *dst = rows_out ? rows_out + i * nrp : counts_out;
                    while (dxy_sq <= rpsq[k]) { dst[k] += w2tmp; if (--k < 0) break; }
                }
            }
        }
    }
    free(wx); free(rpsq);
}

/* ------------------------------------------------------------ jackknife counters (section 8(f) rank 3)
 *   npairs_jackknife_3d    .../cpairs/npairs_jackknife_3d_engine.pyx:120-233, jweight :237-291
 *   npairs_jackknife_xy_z  .../cpairs/npairs_jackknife_xy_z_engine.pyx:128-246
 * counts_out: (n_samples + 1, nb) or (n_samples + 1, nrp, npi) row-major.  npi = 0 selects the 3-D version.     */
static double jweight(int64_t j, int64_t j1, int64_t j2, double w1, double w2)
{
    if (j == 0) return w1 * w2;
    if ((j1 == j2) && (j1 == j)) return 0.0;
    if ((j1 != j) && (j2 != j)) return w1 * w2;
    return 0.5 * (w1 * w2);                /* (j1 != j2) & ((j1 == j) | (j2 == j)) */
}

void oracle_npairs_jackknife(const oracle_geom_t *g,
                             const double *x1, const double *y1, const double *z1, const int64_t *off1,
                             const double *x2, const double *y2, const double *z2, const int64_t *off2,
                             const double *w1, const double *w2, const int64_t *jt1, const int64_t *jt2, int n_samples,
                             const double *bins0, int n0, const double *bins1, int npi,
                             int64_t first_cell1, int64_t last_cell1, double *counts_out)
{
    double *e0 = squares(bins0, n0), *e1 = squares(bins1, npi);
    const int n1 = npi > 0 ? npi : 1;
    const int ny1 = g->ndivs1[1], nz1 = g->ndivs1[2];
    const int ny2 = g->ndivs2[1], nz2 = g->ndivs2[2];
    const int perx = g->ndivs2[0] / g->ndivs1[0], pery = ny2 / ny1, perz = nz2 / nz1;
    const int mw = max_window(g);
    nbr_t *wx = (nbr_t *)malloc(sizeof(nbr_t) * mw * 3), *wy = wx + mw, *wz = wy + mw;
    const size_t nh = (size_t)n0 * n1;
    for (size_t k = 0; k < (size_t)(n_samples + 1) * nh; ++k) counts_out[k] = 0.0;
    for (int64_t c1 = first_cell1; c1 < last_cell1; ++c1) {
        const int64_t a = off1[c1], b = off1[c1 + 1];
        if (b <= a) continue;
        const int ix1 = (int)(c1 / ((int64_t)ny1 * nz1));
        const int iy1 = (int)((c1 - (int64_t)ix1 * ny1 * nz1) / nz1);
        const int iz1 = (int)(c1 - (int64_t)ix1 * ny1 * nz1 - (int64_t)iy1 * nz1);
        const int nx = fill_window(ix1, perx, g->cover[0], g->ndivs2[0], g->period[0], g->pbc, wx);
        const int ny = fill_window(iy1, pery, g->cover[1], ny2, g->period[1], g->pbc, wy);
        const int nz = fill_window(iz1, perz, g->cover[2], nz2, g->period[2], g->pbc, wz);
        for (int ax = 0; ax < nx; ++ax)
        for (int ay = 0; ay < ny; ++ay)
        for (int az = 0; az < nz; ++az) {
            const int64_t c2 = (int64_t)wx[ax].idx * ny2 * nz2 + (int64_t)wy[ay].idx * nz2 + wz[az].idx;
            const int64_t p = off2[c2], q = off2[c2 + 1];
            if (q <= p) continue;
            const double sx = wx[ax].shift, sy = wy[ay].shift, sz = wz[az].shift;
            for (int64_t i = a; i < b; ++i) {
                const double xt = x1[i] - sx, yt = y1[i] - sy, zt = z1[i] - sz;
                for (int64_t j = p; j < q; ++j) {
                    const double dx = xt - x2[j], dy = yt - y2[j], dz = zt - z2[j];
                    if (npi == 0) {
                        const double dsq = dx * dx + dy * dy + dz * dz;
                        if (!(dsq <= e0[n0 - 1])) continue;          /* nothing would be added for any s */
                        for (int s = 0; s <= n_samples; ++s) {
                            const double w = jweight(s, jt1[i], jt2[j], w1[i], w2[j]);
                            int k = n0 - 1;
                            while (dsq <= e0[k]) { counts_out[(size_t)s * nh + k] += w; if (--k < 0) break; }
                        }
                    } else {
                        const double dxy_sq = dx * dx + dy * dy;
                        const double dz_sq = dz * dz;
                        if (!(dxy_sq <= e0[n0 - 1]) || !(dz_sq <= e1[npi - 1])) continue;
                        for (int s = 0; s <= n_samples; ++s) {
                            const double w = jweight(s, jt1[i], jt2[j], w1[i], w2[j]);
                            int k = n0 - 1;
                            while (dxy_sq <= e0[k]) {
                                int gq = npi - 1;
                                while (dz_sq <= e1[gq]) { counts_out[(size_t)s * nh + (size_t)k * npi + gq] += w; if (--gq < 0) break; }
                                if (--k < 0) break;
                            }
                        }
                    }
                }
            }
        }
    }
    free(wx); free(e0); free(e1);
}

/* ------------------------------------------------------------ brute force O(N^2)
 * restating pair_counters/pairs.py:17-84 (npairs): per-pair minimum-image distance,
 * used by the reference's own tests as ground truth on small inputs.            */
void oracle_brute_npairs_3d(const double *s1, int64_t n1, const double *s2, int64_t n2,
                            const double *rbins, int nb, const double *period /* NULL = none */,
                            int64_t *counts_out)
{
    memset(counts_out, 0, sizeof(int64_t) * nb);
    for (int64_t i = 0; i < n1; ++i)
        for (int64_t j = 0; j < n2; ++j) {
            double d2 = 0.0;
            for (int d = 0; d < 3; ++d) {
                double m = fabs(s1[3 * i + d] - s2[3 * j + d]);
                if (period) { double alt = period[d] - m; if (alt < m) m = alt; }
                d2 += m * m;
            }
            const double dist = sqrt(d2);
            for (int k = 0; k < nb; ++k) if (dist <= rbins[k]) counts_out[k] += 1;
        }
}
