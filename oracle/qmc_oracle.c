/* CPU oracle for the batched Monte Carlo hot path -- TEST INFRASTRUCTURE, never linked into the product.
 * See qmc_oracle.h and oracle/__init__.py.  Every function cites the reference (quansino v0.1.3,
 * paths relative to /root/reference) or the ASE 3.26.0 file whose published algorithm it restates.
 *
 * Deliberately simple: brute-force image enumeration, full-system recompute on every attempt -- exactly
 * what `atoms.get_potential_energy()` does for the reference after any change of the atoms.
 */
#include "qmc_oracle.h"

#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define KB 8.617330337217213e-05 /* ase.units.kB (CODATA 2014): _k / _e */
#define H_PLANCK 6.626070040e-34
#define E_CHARGE 1.6021766208e-19
#define N_AVOGADRO 6.022140857e23
#define TWO_PI 6.283185307179586 /* 2 * np.pi */

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) and the slot layout of oracle/philox.py
 * ---------------------------------------------------------------------------------------------- */
void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void qo_uniform_pair(uint64_t seed, uint64_t attempt, uint32_t replica, uint32_t block, double out[2]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)attempt, (uint32_t)(attempt >> 32), replica, block};
    uint32_t w[4];
    qo_philox4x32_10(ctr, key, w);
    out[0] = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
    out[1] = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) / 9007199254740992.0;
}

double qo_normal(uint64_t seed, uint64_t attempt, uint32_t replica, uint32_t m) {
    double u[2];
    qo_uniform_pair(seed, attempt, replica, 16u + m / 2u, u);
    double r = sqrt(-2.0 * log(1.0 - u[0]));
    double t = TWO_PI * u[1];
    return (m % 2u == 0u) ? r * cos(t) : r * sin(t);
}

/* ------------------------------------------------------------------------------------------------
 * Geometry helpers
 * ---------------------------------------------------------------------------------------------- */
static double det3(const double c[9]) {
    return c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
}

/* |det(cell)|: Atoms.get_volume / Cell.volume.  For the orthorhombic cells of the configs this is the plain
 * product (a*b)*c; numpy's det goes through exp(sum(log|u_kk|)) and may differ in the last ulps. */
static double volume(const double c[9]) {
    if (c[1] == 0 && c[2] == 0 && c[3] == 0 && c[5] == 0 && c[6] == 0 && c[7] == 0) return fabs((c[0] * c[4]) * c[8]);
    return fabs(det3(c));
}

static void inv3(const double c[9], double inv[9]) {
    double d = det3(c);
    inv[0] = (c[4] * c[8] - c[5] * c[7]) / d;
    inv[1] = (c[2] * c[7] - c[1] * c[8]) / d;
    inv[2] = (c[1] * c[5] - c[2] * c[4]) / d;
    inv[3] = (c[5] * c[6] - c[3] * c[8]) / d;
    inv[4] = (c[0] * c[8] - c[2] * c[6]) / d;
    inv[5] = (c[2] * c[3] - c[0] * c[5]) / d;
    inv[6] = (c[3] * c[7] - c[4] * c[6]) / d;
    inv[7] = (c[1] * c[6] - c[0] * c[7]) / d;
    inv[8] = (c[0] * c[4] - c[1] * c[3]) / d;
}

typedef struct {
    int nmax[3];
    int n_off;
    int *off;       /* [n_off][3] */
    double *shift;  /* [n_off][3] = off @ cell */
    int *wrap;      /* [n][3] integer cell shift of each atom (0 along non-periodic axes) */
} images_t;

/* ase/neighborlist.py (PrimitiveNeighborList.build): the number of image repeats along axis k follows from the
 * cell height along k; atoms are wrapped only to find the offsets, never moved. */
static int images_build(images_t *im, int n, const double *pos, const double cell[9], const int32_t pbc[3], double cutoff) {
    memset(im, 0, sizeof(*im));
    int any = pbc[0] || pbc[1] || pbc[2];
    im->wrap = (int *)calloc((size_t)(n > 0 ? n : 1) * 3, sizeof(int));
    if (!im->wrap) return -1;
    if (any) {
        double vol = fabs(det3(cell));
        if (!(vol > 0)) return -1;
        for (int k = 0; k < 3; ++k) {
            if (!pbc[k]) continue;
            const double *a = cell + 3 * ((k + 1) % 3), *b = cell + 3 * ((k + 2) % 3);
            double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
            double height = vol / sqrt(cx * cx + cy * cy + cz * cz);
            im->nmax[k] = (int)ceil(cutoff / height);
        }
        double inv[9];
        inv3(cell, inv);
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) {
                if (!pbc[k]) continue;
                /* fractional coordinate k = pos . inv[:, k] */
                double f = pos[3 * i] * inv[k] + pos[3 * i + 1] * inv[3 + k] + pos[3 * i + 2] * inv[6 + k];
                im->wrap[3 * i + k] = (int)floor(f);
            }
        }
    }
    im->n_off = (2 * im->nmax[0] + 1) * (2 * im->nmax[1] + 1) * (2 * im->nmax[2] + 1);
    im->off = (int *)malloc(sizeof(int) * 3 * (size_t)im->n_off);
    im->shift = (double *)malloc(sizeof(double) * 3 * (size_t)im->n_off);
    if (!im->off || !im->shift) return -1;
    int q = 0;
    for (int a = -im->nmax[0]; a <= im->nmax[0]; ++a)
        for (int b = -im->nmax[1]; b <= im->nmax[1]; ++b)
            for (int c = -im->nmax[2]; c <= im->nmax[2]; ++c) {
                im->off[3 * q] = a; im->off[3 * q + 1] = b; im->off[3 * q + 2] = c;
                for (int x = 0; x < 3; ++x)
                    im->shift[3 * q + x] = ((double)a * cell[x] + (double)b * cell[3 + x]) + (double)c * cell[6 + x];
                ++q;
            }
    return 0;
}

static void images_free(images_t *im) {
    free(im->off); free(im->shift); free(im->wrap);
    memset(im, 0, sizeof(*im));
}

/* distance vector of image q of atom j seen from atom i: positions[j] + T @ cell - positions[i] with the
 * integer offset T = off[q] + wrap[i] - wrap[j]  (ASE: positions[neighbors] + offsets @ cell - positions[a1]) */
static inline void image_vector(const images_t *im, const double *pos, const double cell[9], int i, int j, int q, double d[3]) {
    const int *wi = im->wrap + 3 * i, *wj = im->wrap + 3 * j;
    int t0 = wi[0] - wj[0], t1 = wi[1] - wj[1], t2 = wi[2] - wj[2];
    if ((t0 | t1 | t2) == 0) {
        for (int x = 0; x < 3; ++x) d[x] = (pos[3 * j + x] + im->shift[3 * q + x]) - pos[3 * i + x];
    } else {
        double a = (double)(im->off[3 * q] + t0), b = (double)(im->off[3 * q + 1] + t1), c = (double)(im->off[3 * q + 2] + t2);
        for (int x = 0; x < 3; ++x) {
            double s = (a * cell[x] + b * cell[3 + x]) + c * cell[6 + x];
            d[x] = (pos[3 * j + x] + s) - pos[3 * i + x];
        }
    }
}

/* Candidate list for large orthorhombic cells (test-size speed-up only: C5's 2048 atoms): per atom the (j, q) pairs that
 * can be inside the cutoff, in EXACTLY the order of the brute-force loops (j ascending, q ascending), found from a cheap
 * per-axis bound.  The evaluation loops apply the very same arithmetic and the very same `r < rc_list` test to them, so the
 * results are bit-identical to the brute-force enumeration (tests/test_oracle.py checks this).  NULL start = not built. */
typedef struct {
    int64_t *start; /* [n + 1] */
    int32_t *j, *q;
} pairlist_t;

static void pairlist_free(pairlist_t *pl) {
    free(pl->start); free(pl->j); free(pl->q);
    memset(pl, 0, sizeof(*pl));
}

static int pairlist_build(pairlist_t *pl, const images_t *im, int n, const double *pos, const double cell[9], double cutoff) {
    memset(pl, 0, sizeof(*pl));
    if (n < 256 || getenv("QO_NO_PAIRLIST")) return 0;
    for (int x = 0; x < 9; ++x)
        if (x % 4 != 0 && cell[x] != 0.0) return 0; /* diagonal cells only */
    const double rc = cutoff * (1.0 + 1e-9) + 1e-12, rc2 = rc * rc;
    const int w0 = 2 * im->nmax[0] + 1, w1 = 2 * im->nmax[1] + 1, w2 = 2 * im->nmax[2] + 1;
    (void)w0;
    size_t cap = (size_t)n * 128, cnt = 0;
    pl->start = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n + 1));
    pl->j = (int32_t *)malloc(sizeof(int32_t) * cap);
    pl->q = (int32_t *)malloc(sizeof(int32_t) * cap);
    if (!pl->start || !pl->j || !pl->q) { pairlist_free(pl); return -1; }
    for (int i = 0; i < n; ++i) {
        pl->start[i] = (int64_t)cnt;
        for (int j = 0; j < n; ++j) {
            int lo[3], hi[3];
            double u[3][8]; /* candidate components per axis (at most 2 nmax + 1 <= 7) */
            int ok = 1;
            for (int k = 0; k < 3 && ok; ++k) {
                const double L = cell[4 * k];
                const double base = (pos[3 * j + k] - pos[3 * i + k]) + (double)(im->wrap[3 * i + k] - im->wrap[3 * j + k]) * L;
                lo[k] = 1; hi[k] = 0;
                if (im->nmax[k] > 3) { ok = 0; break; }
                for (int a = -im->nmax[k]; a <= im->nmax[k]; ++a) {
                    const double v = base + (double)a * L;
                    u[k][a + im->nmax[k]] = v;
                    if (fabs(v) < rc) {
                        if (lo[k] > hi[k]) lo[k] = a;
                        hi[k] = a;
                    }
                }
                if (lo[k] > hi[k]) ok = 0;
            }
            if (!ok) {
                if (im->nmax[0] > 3 || im->nmax[1] > 3 || im->nmax[2] > 3) { pairlist_free(pl); return 0; }
                continue;
            }
            for (int a = lo[0]; a <= hi[0]; ++a)
                for (int b = lo[1]; b <= hi[1]; ++b)
                    for (int c = lo[2]; c <= hi[2]; ++c) {
                        const double x = u[0][a + im->nmax[0]], y = u[1][b + im->nmax[1]], z = u[2][c + im->nmax[2]];
                        if (!(x * x + y * y + z * z < rc2)) continue;
                        if (j == i && a == 0 && b == 0 && c == 0) continue;
                        if (cnt == cap) {
                            cap *= 2;
                            int32_t *nj = (int32_t *)realloc(pl->j, sizeof(int32_t) * cap), *nq = (int32_t *)realloc(pl->q, sizeof(int32_t) * cap);
                            if (!nj || !nq) { if (nj) pl->j = nj; if (nq) pl->q = nq; pairlist_free(pl); return -1; }
                            pl->j = nj; pl->q = nq;
                        }
                        pl->j[cnt] = j;
                        pl->q[cnt] = ((a + im->nmax[0]) * w1 + (b + im->nmax[1])) * w2 + (c + im->nmax[2]);
                        ++cnt;
                    }
        }
    }
    pl->start[n] = (int64_t)cnt;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ASE EMT, single species  (ase/calculators/emt.py: calculate, _get_neighbors, _calc_theta, _calc_dsigma1,
 * _calc_dsigma2, _calc_e_c_a2, _calc_efs_a1, _calc_fs_c_a2; energies -= E0 at the end)
 * ---------------------------------------------------------------------------------------------- */
static int64_t emt_eval(const qo_potential_t *p, int n, const double *pos, const double cell[9], const int32_t pbc[3],
                        double *energy, double *forces, double *sigma1_out) {
    images_t im;
    if (images_build(&im, n, pos, cell, pbc, p->rc_list) != 0) { images_free(&im); return -1; }
    pairlist_t pl;
    if (pairlist_build(&pl, &im, n, pos, cell, p->rc_list) != 0) { images_free(&im); return -1; }
    double *sig = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    double *deds = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    double *en = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    const double beta = p->beta;
    const double inv12gamma1 = 1.0 / (12.0 * p->gamma1);
    const double neghalfv0overgamma2 = -0.5 * p->V0 / p->gamma2;
    int64_t npairs = 0;
    /* pass 1: sigma1 and the pair part of E_AS */
    for (int i = 0; i < n; ++i) {
        double s1 = 0.0, s2 = 0.0;
        int cnt = 0;
        /* all (j, q) in ascending order -- or, for large orthorhombic cells, the candidates among them in the same order */
        const int64_t c_end = pl.start ? pl.start[i + 1] : (int64_t)n * im.n_off;
        for (int64_t c = pl.start ? pl.start[i] : 0; c < c_end; ++c) {
                const int j = pl.start ? pl.j[c] : (int)(c / im.n_off), q = pl.start ? pl.q[c] : (int)(c % im.n_off);
                if (j == i && im.off[3 * q] == 0 && im.off[3 * q + 1] == 0 && im.off[3 * q + 2] == 0) continue;
                double d[3];
                image_vector(&im, pos, cell, i, j, q, d);
                double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                if (!(r < p->rc_list)) continue; /* mask = r < self.rc_list */
                double w = 1.0 / (1.0 + exp(p->acut * (r - p->rc)));
                s1 += exp(-p->eta2 * (r - beta * p->s0)) * w;
                s2 += exp(-p->kappa * (r / beta - p->s0)) * w;
                ++cnt;
            }
        npairs += cnt;
        sig[i] = s1;
        if (cnt == 0) {
            en[i] = -p->E0; /* no neighbours: energies[a] stays 0, then -= E0 */
            deds[i] = 0.0;
            continue;
        }
        double ds = -1.0 * log(s1 * inv12gamma1) / (beta * p->eta2);
        double lmdsds = p->lambda * ds;
        double expneglmdds = exp(-1.0 * lmdsds);
        double e = p->E0 * (1.0 + lmdsds) * expneglmdds;
        double de = -1.0 * p->E0 * p->lambda * lmdsds * expneglmdds;
        double sixv0 = 6.0 * p->V0 * exp(-1.0 * p->kappa * ds);
        e += sixv0;
        de += -1.0 * p->kappa * sixv0;
        de /= -1.0 * beta * p->eta2 * s1;
        deds[i] = de;
        double es_sum = neghalfv0overgamma2 * s2;
        e += 0.5 * (es_sum + es_sum);
        en[i] = e - p->E0;
    }
    double etot = 0.0;
    for (int i = 0; i < n; ++i) etot += en[i];
    if (energy) *energy = etot;
    if (sigma1_out) memcpy(sigma1_out, sig, sizeof(double) * (size_t)n);
    if (forces) {
        for (int i = 0; i < n; ++i) {
            double f[3] = {0, 0, 0};
            const int64_t c_end = pl.start ? pl.start[i + 1] : (int64_t)n * im.n_off;
            for (int64_t c = pl.start ? pl.start[i] : 0; c < c_end; ++c) {
                    const int j = pl.start ? pl.j[c] : (int)(c / im.n_off), q = pl.start ? pl.q[c] : (int)(c % im.n_off);
                    if (j == i && im.off[3 * q] == 0 && im.off[3 * q + 1] == 0 && im.off[3 * q + 2] == 0) continue;
                    double d[3];
                    image_vector(&im, pos, cell, i, j, q, d);
                    double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    if (!(r < p->rc_list)) continue;
                    double w = 1.0 / (1.0 + exp(p->acut * (r - p->rc)));
                    double dwdroverw = p->acut * (w - 1.0);
                    double ds1 = exp(-p->eta2 * (r - beta * p->s0)) * w;
                    double ds2 = exp(-p->kappa * (r / beta - p->s0)) * w;
                    double invr = 1.0 / r;
                    double es = neghalfv0overgamma2 * ds2;
                    double dedr_pair = es * (dwdroverw - p->kappa / beta);
                    double dds1dr = ds1 * (dwdroverw - p->eta2);
                    double g = ((dedr_pair + dedr_pair) + (deds[i] * dds1dr + deds[j] * dds1dr)) * invr;
                    f[0] += g * d[0]; f[1] += g * d[1]; f[2] += g * d[2];
                }
            forces[3 * i] = f[0]; forces[3 * i + 1] = f[1]; forces[3 * i + 2] = f[2];
        }
    }
    free(sig); free(deds); free(en);
    pairlist_free(&pl);
    images_free(&im);
    return npairs;
}

/* ASE LennardJones.calculate, smooth=False (ase/calculators/lj.py) */
static int64_t lj_eval(const qo_potential_t *p, int n, const double *pos, const double cell[9], const int32_t pbc[3],
                       double *energy, double *forces) {
    images_t im;
    if (images_build(&im, n, pos, cell, pbc, p->lj_rc) != 0) { images_free(&im); return -1; }
    const double rc2 = p->lj_rc * p->lj_rc, s2 = p->sigma * p->sigma;
    int64_t npairs = 0;
    double etot = 0.0;
    for (int i = 0; i < n; ++i) {
        double e = 0.0, f[3] = {0, 0, 0};
        for (int j = 0; j < n; ++j)
            for (int q = 0; q < im.n_off; ++q) {
                if (j == i && im.off[3 * q] == 0 && im.off[3 * q + 1] == 0 && im.off[3 * q + 2] == 0) continue;
                double d[3];
                image_vector(&im, pos, cell, i, j, q, d);
                double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 > rc2) continue; /* c6[r2 > rc**2] = 0 */
                double c6 = (s2 / r2) * (s2 / r2) * (s2 / r2);
                double c12 = c6 * c6;
                e += 4 * p->epsilon * (c12 - c6) - p->lj_e0;
                if (forces) {
                    double g = -24 * p->epsilon * (2 * c12 - c6) / r2;
                    f[0] += g * d[0]; f[1] += g * d[1]; f[2] += g * d[2];
                }
                ++npairs;
            }
        etot += 0.5 * e;
        if (forces) { forces[3 * i] = f[0]; forces[3 * i + 1] = f[1]; forces[3 * i + 2] = f[2]; }
    }
    if (energy) *energy = etot;
    images_free(&im);
    return npairs;
}

int64_t qo_energy(const qo_potential_t *pot, int32_t n, const double *pos, const double *cell, const int32_t *pbc,
                  double *energy, double *forces, double *sigma1) {
    if (pot->kind == QO_POT_EMT) return emt_eval(pot, n, pos, cell, pbc, energy, forces, sigma1);
    if (pot->kind == QO_POT_LJ) return lj_eval(pot, n, pos, cell, pbc, energy, forces);
    return -1;
}

/* neighbour images of atom i inside the list cutoff; marks distinct neighbour atoms in seen[] */
static int count_neighbours(const qo_potential_t *pot, int n, const double *pos, const double cell[9], const int32_t pbc[3],
                            int i, char *seen) {
    double cutoff = pot->kind == QO_POT_EMT ? pot->rc_list : pot->lj_rc;
    images_t im;
    if (images_build(&im, n, pos, cell, pbc, cutoff) != 0) { images_free(&im); return 0; }
    int cnt = 0;
    for (int j = 0; j < n; ++j)
        for (int q = 0; q < im.n_off; ++q) {
            if (j == i && im.off[3 * q] == 0 && im.off[3 * q + 1] == 0 && im.off[3 * q + 2] == 0) continue;
            double d[3];
            image_vector(&im, pos, cell, i, j, q, d);
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            int in = pot->kind == QO_POT_EMT ? (sqrt(r2) < cutoff) : !(r2 > cutoff * cutoff);
            if (in) { ++cnt; if (seen) seen[j] = 1; }
        }
    images_free(&im);
    return cnt;
}

/* ------------------------------------------------------------------------------------------------
 * Criteria (src/quansino/mc/criteria.py)
 * ---------------------------------------------------------------------------------------------- */
double qo_debroglie_wavelength(double mass_amu, double temperature) {
    /* mc/criteria.py:260-266, evaluated left to right like the Python expression */
    double denom = 2 * 3.141592653589793 * mass_amu * KB * temperature / N_AVOGADRO * 1e-3 * E_CHARGE;
    return sqrt(H_PLANCK * H_PLANCK / denom) * 1e10;
}

/* returns criteria * prefactor of GrandCanonicalCriteria.evaluate (mc/criteria.py:243-278); sets *overflow */
static double gc_probability(double delta_e, int particle_delta, int n_ex, double acc_volume, double mass, double temperature,
                             double mu, int *overflow) {
    double vol = pow(acc_volume, (double)particle_delta);
    double factorial_term = 1.0;
    if (particle_delta > 0) {
        for (int i = n_ex + 1; i < n_ex + particle_delta + 1; ++i) factorial_term /= (double)i;
    } else if (particle_delta < 0) {
        for (int i = n_ex + particle_delta + 1; i < n_ex + 1; ++i) factorial_term *= (double)i;
    }
    double debroglie = pow(qo_debroglie_wavelength(mass, temperature), (double)(-3 * particle_delta));
    double prefactor = vol * factorial_term * debroglie;
    double exponential = ((double)particle_delta * mu - delta_e) / (temperature * KB);
    if (exponential > 709.782712893384) { if (overflow) *overflow = 1; return INFINITY; }
    return exp(exponential) * prefactor;
}

double qo_gc_acceptance(double delta_e, int32_t particle_delta, int32_t n_exchange, double accessible_volume, double mass_amu,
                        double temperature, double mu) {
    int ov = 0;
    return gc_probability(delta_e, particle_delta, n_exchange, accessible_volume, mass_amu, temperature, mu, &ov);
}

int qo_pt_swap_accept(double e_k, double e_k1, double t_k, double t_k1, double u) {
    /* replica exchange between temperatures t_k and t_k1: u < exp((1/kT_k - 1/kT_k1) (E_k - E_k1)) */
    double x = (1.0 / (t_k * KB) - 1.0 / (t_k1 * KB)) * (e_k - e_k1);
    if (x >= 0) return 1;
    return u < exp(x);
}

/* ------------------------------------------------------------------------------------------------
 * Label bookkeeping (src/quansino/moves/displacement.py:173-184, 226-253)
 * ---------------------------------------------------------------------------------------------- */
static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* sorted unique non-negative labels; returns count */
static int unique_labels(const int32_t *labels, int n, int *uniq) {
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (labels[i] >= 0) uniq[m++] = labels[i];
    qsort(uniq, (size_t)m, sizeof(int), cmp_int);
    int k = 0;
    for (int i = 0; i < m; ++i)
        if (k == 0 || uniq[k - 1] != uniq[i]) uniq[k++] = uniq[i];
    return k;
}

/* the single atom carrying `label`; -2 if the group has != 1 member (molecules: SURVEY 8f, unsupported) */
static int atom_of_label(const int32_t *labels, int n, int label) {
    int found = -1, cnt = 0;
    for (int i = 0; i < n; ++i)
        if (labels[i] == label) { found = i; ++cnt; }
    return cnt == 1 ? found : -2;
}

/* on_atoms_changed(added, removed) for one move, one added atom at the end and/or one removed index */
static void labels_on_atoms_changed(qo_move_t *mv, int n_before, int added, int removed_index, int *scratch) {
    int n = n_before;
    if (added) {
        int k = unique_labels(mv->labels, n, scratch);
        /* `self.default_label or (...)`: None and 0 are both falsy */
        int label = (mv->default_label != INT32_MIN && mv->default_label != 0) ? mv->default_label : (k ? scratch[k - 1] + 1 : 0);
        mv->labels[n++] = label;
    }
    if (removed_index >= 0) {
        for (int i = removed_index; i + 1 < n; ++i) mv->labels[i] = mv->labels[i + 1];
        --n;
    }
}

/* the same for a molecule (moves/displacement.py:244-253): ONE new label for all `n_added` atoms appended at the end, or the
 * atoms flagged in `gone` deleted with np.delete (order preserved) */
static void labels_on_molecule_changed(qo_move_t *mv, int n_before, int n_added, const unsigned char *gone, int *scratch) {
    if (n_added > 0) {
        int k = unique_labels(mv->labels, n_before, scratch);
        int label = (mv->default_label != INT32_MIN && mv->default_label != 0) ? mv->default_label : (k ? scratch[k - 1] + 1 : 0);
        for (int q = 0; q < n_added; ++q) mv->labels[n_before + q] = label;
    } else if (gone) {
        int o = 0;
        for (int i = 0; i < n_before; ++i)
            if (!gone[i]) mv->labels[o++] = mv->labels[i];
    }
}

/* ------------------------------------------------------------------------------------------------
 * One attempt of each move type.  All work on a TRIAL copy; commit on accept => revert is bit-exact.
 * ---------------------------------------------------------------------------------------------- */
static double kinetic_energy(int n, const double *mom, double mass) {
    /* Atoms.get_kinetic_energy: 0.5 * vdot(p, p / m) */
    double s = 0.0;
    for (int i = 0; i < 3 * n; ++i) s += mom[i] * (mom[i] / mass);
    return 0.5 * s;
}

/* Atoms.set_cell(new, scale_atoms=True): M = solve(old, new); positions = positions @ M.  For the diagonal
 * cells of the configs M is diagonal with M_kk = new_kk / old_kk. */
static void scale_positions(int n, double *pos, const double old_cell[9], const double new_cell[9]) {
    double inv[9], M[9];
    int diag = old_cell[1] == 0 && old_cell[2] == 0 && old_cell[3] == 0 && old_cell[5] == 0 && old_cell[6] == 0 && old_cell[7] == 0;
    if (diag) {
        memset(M, 0, sizeof(M));
        M[0] = new_cell[0] / old_cell[0]; M[4] = new_cell[4] / old_cell[4]; M[8] = new_cell[8] / old_cell[8];
        M[1] = new_cell[1] / old_cell[0]; M[2] = new_cell[2] / old_cell[0];
        M[3] = new_cell[3] / old_cell[4]; M[5] = new_cell[5] / old_cell[4];
        M[6] = new_cell[6] / old_cell[8]; M[7] = new_cell[7] / old_cell[8];
    } else {
        inv3(old_cell, inv);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = (inv[3 * r] * new_cell[c] + inv[3 * r + 1] * new_cell[3 + c]) + inv[3 * r + 2] * new_cell[6 + c];
    }
    for (int i = 0; i < n; ++i) {
        double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        for (int c = 0; c < 3; ++c) pos[3 * i + c] = (x * M[c] + y * M[3 + c]) + z * M[6 + c];
    }
}

/* exp of a SYMMETRIC 3 x 3 matrix (the reference: scipy.linalg.expm, operations/cell.py:96,141) by cyclic Jacobi rotations:
 * A = V diag(w) V^T  =>  exp(A) = V diag(exp(w)) V^T.  Deliberately a different method from the device's scaling-and-squaring
 * Taylor polynomial; tests/test_oracle.py pins it against scipy to 1e-15. */
static void expm_sym3(const double A[9], double E[9]) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = A[3 * i + j];
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { /* A <- A J */
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) { /* A <- J^T A */
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    double w[3] = {exp(a[0][0]), exp(a[1][1]), exp(a[2][2])};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) E[3 * i + j] = (v[i][0] * w[0] * v[j][0] + v[i][1] * w[1] * v[j][1]) + v[i][2] * w[2] * v[j][2];
}

void qo_expm_sym3(const double A[9], double E[9]) { expm_sym3(A, E); }

static void mat3_mul(const double a[9], const double b[9], double c[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}

/* deformation gradient of a CellMove (operations/cell.py:67-169) from the six uniform draws u[0..5] of the attempt */
static void deformation_gradient(const qo_move_t *mv, const double u[6], double F[9]) {
    if (mv->op == QO_OP_ISOTROPIC) {
        double s = exp(-mv->step + (mv->step - -mv->step) * u[0]);
        for (int q = 0; q < 9; ++q) F[q] = (q % 4 == 0) ? ((mv->axes == 0 || ((mv->axes >> (q / 4)) & 1)) ? s : 1.0) : 0.0;
        return;
    }
    double c[6], T[9], E[9];
    for (int q = 0; q < 6; ++q) c[q] = -mv->step + (mv->step - -mv->step) * u[q];
    double mean = mv->op == QO_OP_SHAPE ? (c[0] + c[1] + c[2]) / 3.0 : 0.0;
    T[0] = c[0] - mean; T[4] = c[1] - mean; T[8] = c[2] - mean;
    T[1] = T[3] = c[3]; T[2] = T[6] = c[4]; T[5] = T[7] = c[5];
    expm_sym3(T, E);
    const int mask = mv->axes == 0 ? 0x1FF : mv->axes; /* bit 3 i + j = mask[i][j] */
    for (int q = 0; q < 9; ++q) F[q] = ((mask >> q) & 1) ? E[q] : ((q % 4 == 0) ? 1.0 : 0.0);
}

/* the work term of the cell criteria: IsobaricCriteria p dV (mc/criteria.py:168-170), or IsotensionCriteria's elastic energy
 * (:196-208), literally: strain = 0.5 (inv(old^T) @ new^T @ old @ inv(old) - 1); p dV + V_old trace((stress - p) @ strain) with
 * the scalar p subtracted from EVERY component of the stress matrix (numpy broadcasting) */
static double cell_work(const qo_state_t *st, const double new_cell[9]) {
    double v_new = volume(new_cell), v_old = volume(st->cell);
    double work = st->pressure * (v_new - v_old);
    if (st->isotension) {
        double oT[9], nT[9], ioT[9], io[9], P[9], Q[9], S[9];
        for (int q = 0; q < 9; ++q) { oT[q] = st->cell[3 * (q % 3) + q / 3]; nT[q] = new_cell[3 * (q % 3) + q / 3]; }
        inv3(oT, ioT); inv3(st->cell, io);
        mat3_mul(ioT, nT, P); mat3_mul(P, st->cell, Q); mat3_mul(Q, io, S);
        double tr = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) tr += (st->external_stress[3 * i + j] - st->pressure) * (0.5 * (S[3 * j + i] - (i == j ? 1.0 : 0.0)));
        work += v_old * tr;
    }
    return work;
}

int qo_fbmc_step(const qo_potential_t *pot, int32_t n, double *pos, const double *cell, const int32_t *pbc, double mass,
                 double temperature, double delta, uint64_t seed, uint32_t replica, uint64_t step, double *energy,
                 double *gamma_max, int32_t *n_iter) {
    const int M = 3 * n;
    double *frc = (double *)malloc(sizeof(double) * (size_t)M), *gam = (double *)malloc(sizeof(double) * (size_t)M);
    double *den = (double *)malloc(sizeof(double) * (size_t)M), *zeta = (double *)malloc(sizeof(double) * (size_t)M);
    char *conv = (char *)calloc((size_t)M, 1);
    int rc = 0;
    if (!frc || !gam || !den || !zeta || !conv) { rc = -1; goto out; }
    double e0;
    if (qo_energy(pot, n, pos, cell, pbc, &e0, frc, NULL) < 0) { rc = -1; goto out; }
    const double gmaxv = 709.782712; /* ForceBias.gamma_max_value */
    double gm = 0.0;
    for (int m = 0; m < M; ++m) { /* calculate_gamma (mc/fbmc.py:93-108) */
        double g = (frc[m] * delta) / (2 * temperature * KB);
        g = g < -gmaxv ? -gmaxv : (g > gmaxv ? gmaxv : g);
        gam[m] = g;
        den[m] = exp(g) - exp(-g);
        if (fabs(g) > gm) gm = fabs(g);
    }
    int remaining = M, iter = 0;
    while (remaining > 0) { /* mc/fbmc.py:224-241 */
        int p = 0;
        for (int m = 0; m < M; ++m) {
            if (conv[m]) continue;
            double uz[2], ur[2];
            qo_uniform_pair(seed, step, replica, (uint32_t)(((2 * iter + 1) << 16) | (p >> 1)), uz);
            qo_uniform_pair(seed, step, replica, (uint32_t)(((2 * iter + 2) << 16) | (p >> 1)), ur);
            const double z = -1 + (1 - -1) * uz[p & 1], u = ur[p & 1];
            const double sg = z > 0 ? 1.0 : (z < 0 ? -1.0 : 0.0); /* np.sign */
            double pt = exp(sg * gam[m]) - exp(gam[m] * (2 * z - sg)); /* calculate_trial_probability (:254-273) */
            pt *= sg;
            const double ptrial = den[m] != 0 ? pt / den[m] : 1.0;
            zeta[m] = z;
            conv[m] = ptrial > u;
            ++p;
        }
        remaining = 0;
        for (int m = 0; m < M; ++m) remaining += !conv[m];
        ++iter;
        if (iter > 10000) { rc = -1; goto out; }
    }
    for (int m = 0; m < M; ++m) { /* displacement = zeta delta (min(m) / m)^0.25; momenta round trip (:243-251) */
        const double d = zeta[m] * delta * pow(mass / mass, 0.25);
        pos[m] = pos[m] + (mass * d) / mass;
    }
    if (qo_energy(pot, n, pos, cell, pbc, energy, NULL, NULL) < 0) { rc = -1; goto out; }
    if (gamma_max) *gamma_max = gm;
    if (n_iter) *n_iter = iter;
out:
    free(frc); free(gam); free(den); free(zeta); free(conv);
    return rc;
}

/* Philox block of (u_label, op draw 0) of drawing sub-move c >= 1 of a composite move; (op draw 1, op draw 2) sit in the next block.
 * Blocks 32 .. 63 serve sub-moves 1 .. 16; the attempt's extra draws start at block 64, so later sub-moves continue at 0x8000. */
static uint32_t submove_block(int c) { return c <= 16 ? (uint32_t)(32 + 2 * (c - 1)) : 0x8000u + (uint32_t)(2 * (c - 17)); }

int qo_mc_run(const qo_potential_t *pot, qo_state_t *st, qo_move_t *moves, int32_t n_moves, uint64_t seed, uint32_t replica,
              uint64_t attempt_base, int32_t n_attempts, qo_trace_t *trace, int64_t counters[4]) {
    const int nmax = st->n_max;
    int rc = 0;
    double *trial = (double *)malloc(sizeof(double) * 3 * (size_t)nmax);
    double *mom = (double *)malloc(sizeof(double) * 3 * (size_t)nmax);
    double *frc = (double *)malloc(sizeof(double) * 3 * (size_t)nmax);
    int *uniq = (int *)malloc(sizeof(int) * (size_t)(nmax + 1));
    char *seen = (char *)malloc((size_t)(nmax + 1));
    unsigned char *gone = (unsigned char *)calloc((size_t)(nmax + 1), 1); /* atoms of the molecule being deleted */
    double cdf[16];
    int64_t cnt[4] = {0, 0, 0, 0};
    if (!trial || !mom || !frc || !uniq || !seen || n_moves > 16 || n_moves < 1) { rc = -1; goto done; }

    /* validate_simulation (mc/canonical.py:114-122): prime last_potential_energy only if NaN */
    if (isnan(st->last_energy)) {
        int64_t np_ = qo_energy(pot, st->n_atoms, st->pos, st->cell, st->pbc, &st->last_energy, NULL, NULL);
        if (np_ < 0) { rc = -1; goto done; }
        cnt[0] += np_; cnt[3] += 1;
    }

    /* yield_moves (mc/core.py:340-349): probabilities normalised, numpy choice(p=) = searchsorted(cdf, u, 'right') */
    {
        /* a composite move is ONE MoveStorage: only the head of a chain carries a probability */
        double tot = 0.0, acc = 0.0;
        for (int m = 0; m < n_moves; ++m)
            if (m == 0 || !(moves[m - 1].chain & 1)) tot += moves[m].probability;
        for (int m = 0; m < n_moves; ++m) {
            if (m == 0 || !(moves[m - 1].chain & 1)) acc += moves[m].probability / tot;
            cdf[m] = acc;
        }
        double last = cdf[n_moves - 1];
        for (int m = 0; m < n_moves; ++m) cdf[m] /= last;
    }

    /* forced moves (mc/core.py:331-342): forced = repeat(available, minimum_count), placed at distinct cycle indices that are
     * drawn once per step -- here by a partial Fisher-Yates shuffle of arange(max_cycles) fed from blocks 4 + j / 2 of the
     * step's first attempt (include/quansino_b200.h) */
    int n_forced = 0, forced_move[16], forced_index[16];
    const int max_cycles = moves[0].max_cycles;
    for (int m = 0; m < n_moves; ++m)
        if (m == 0 || !(moves[m - 1].chain & 1))
            for (int q = 0; q < moves[m].minimum_count; ++q) {
                if (n_forced >= 16) { rc = -1; goto done; }
                forced_move[n_forced++] = m;
            }
    if (n_forced > 0 && (max_cycles < n_forced || attempt_base % (uint64_t)max_cycles != 0)) { rc = -1; goto done; }

    for (int k = 0; k < n_attempts; ++k) {
        const uint64_t attempt = attempt_base + (uint64_t)k;
        if (n_forced > 0 && k % max_cycles == 0) {
            int key[16], val[16], n_over = 0; /* sparse permutation: perm[key] = val, identity elsewhere */
            for (int j = 0; j < n_forced; ++j) {
                double u[2];
                qo_uniform_pair(seed, attempt, replica, (uint32_t)(4 + j / 2), u);
                const int t = j + (int)(u[j & 1] * (double)(max_cycles - j));
                int vt = t, vj = j, it = -1;
                for (int q = 0; q < n_over; ++q) {
                    if (key[q] == t) { vt = val[q]; it = q; }
                    if (key[q] == j) vj = val[q];
                }
                if (it >= 0) val[it] = vj; else { key[n_over] = t; val[n_over++] = vj; } /* perm[t] = perm[j] */
                forced_index[j] = vt;                                                    /* perm[j] = old perm[t] */
            }
        }
        double b0[2], b1[2], b2[2], b3[2];
        qo_uniform_pair(seed, attempt, replica, 0, b0);
        qo_uniform_pair(seed, attempt, replica, 1, b1);
        qo_uniform_pair(seed, attempt, replica, 2, b2);
        qo_uniform_pair(seed, attempt, replica, 3, b3);
        const double u_select = b0[0], u_label = b0[1], u_accept = b2[1], u_coin = b3[0];
        const double opd[3] = {b1[0], b1[1], b2[0]};
        int m = 0;
        while (m < n_moves - 1 && cdf[m] <= u_select) ++m;
        for (int j = 0; j < n_forced; ++j)
            if (forced_index[j] == k % max_cycles) m = forced_move[j]; /* a forced cycle: no selection draw */
        qo_move_t *mv = &moves[m];
        const int n = st->n_atoms;
        const double kT = st->temperature * KB;
        int accepted = -1, moved = -1;
        double delta = NAN;

        if (mv->type == QO_MOVE_DISPLACEMENT && (mv->repeat > 1 || (mv->chain & 3))) {
            /* Composite moves.  chain bit 0: the next entry belongs to the same move; chain bit 1: plain CompositeMove
             * (moves/composite.py:43-62: every sub-move is called on its own, labels may repeat) instead of
             * CompositeDisplacementMove (moves/displacement.py:392-431: every sub-move picks a label nobody displaced yet
             * in this composite -- rng.choice over the sorted set difference -- and sub-moves without candidates fail
             * silently).  A CellMove may close a plain chain (the docs' hybrid NPT move `DisplacementMove * N + CellMove`).
             * Then ONE criteria evaluation on the whole change (mc/core.py:369-377).
             * Draws: the first drawing sub-move uses the standard slots, drawing sub-move c >= 1 uses
             * block 32 + 2 (c - 1) = (u_label, op draw 0) and block 33 + 2 (c - 1) = (op draw 1, op draw 2); the cell draw
             * is op draw 0 if no displacement drew, else the first "extra" draw of the attempt (block 64, first double). */
            memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
            const int exclusive = !(mv->chain & 2);
            int n_moved = 0, c = 0, did_cell = 0; /* disp[]: label VALUES displaced so far (the reference compares labels, not atoms) */
            double new_cell[9];
            int *disp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
            if (!disp) { rc = -1; goto done; }
            for (int mm = m; mm < n_moves; ++mm) {
                qo_move_t *sm = &moves[mm];
                if (sm->type == QO_MOVE_CELL) {
                    if (exclusive || (sm->chain & 1)) { free(disp); rc = -1; goto done; } /* last entry of a plain chain only */
                    double bx[2];
                    qo_uniform_pair(seed, attempt, replica, 64, bx);
                    const double u_cell = c == 0 ? opd[0] : bx[0]; /* first extra draw of the attempt */
                    double sc = exp(-sm->step + (sm->step - -sm->step) * u_cell);
                    for (int q = 0; q < 9; ++q) /* F = eye * s * mask + eye * ~mask, F @ cell: row r scaled by F[r][r] */
                        new_cell[q] = ((sm->axes == 0 || ((sm->axes >> (q / 3)) & 1)) ? sc : 1.0) * st->cell[q];
                    scale_positions(n, trial, st->cell, new_cell);
                    did_cell = 1;
                    break;
                }
                if (sm->type != QO_MOVE_DISPLACEMENT) { free(disp); rc = -1; goto done; }
                const int reps = sm->repeat > 0 ? sm->repeat : 1;
                for (int rep_i = 0; rep_i < reps; ++rep_i) {
                    int nu = unique_labels(sm->labels, n, uniq);
                    int na = 0;
                    for (int q = 0; q < nu; ++q) { /* setdiff1d(unique_labels, displaced_labels): sorted, as unique_labels is */
                        int taken = 0;
                        if (exclusive)
                            for (int d = 0; d < n_moved; ++d) taken |= disp[d] == uniq[q];
                        if (!taken) uniq[na++] = uniq[q];
                    }
                    if (na == 0) continue; /* register_failure, nothing drawn */
                    double ul, od[3];
                    if (c == 0) { ul = u_label; od[0] = opd[0]; od[1] = opd[1]; od[2] = opd[2]; }
                    else {
                        double e0[2], e1[2];
                        qo_uniform_pair(seed, attempt, replica, submove_block(c), e0);
                        qo_uniform_pair(seed, attempt, replica, submove_block(c) + 1u, e1);
                        ul = e0[0]; od[0] = e0[1]; od[1] = e1[0]; od[2] = e1[1];
                    }
                    ++c;
                    const int label = uniq[(int)(ul * (double)na)];
                    int a = atom_of_label(sm->labels, n, label);
                    if (a < 0) { free(disp); rc = -2; goto done; }
                    double t[3];
                    if (sm->op == QO_OP_BALL) {
                        double r = 0.0 + (sm->step - 0.0) * od[0];
                        double phi = 0.0 + (TWO_PI - 0.0) * od[1];
                        double cc = -1.0 + (1.0 - -1.0) * od[2];
                        double s = sqrt(1 - cc * cc);
                        t[0] = r * s * cos(phi); t[1] = r * s * sin(phi); t[2] = r * cc;
                    } else if (sm->op == QO_OP_SPHERE) {
                        double phi = 0.0 + (TWO_PI - 0.0) * od[0];
                        double cc = -1.0 + (1.0 - -1.0) * od[1];
                        double s = sqrt(1 - cc * cc);
                        t[0] = sm->step * (s * cos(phi)); t[1] = sm->step * (s * sin(phi)); t[2] = sm->step * cc;
                    } else if (sm->op == QO_OP_BOX) {
                        for (int x = 0; x < 3; ++x) t[x] = -sm->step + (sm->step - -sm->step) * od[x];
                    } else { free(disp); rc = -1; goto done; }
                    for (int x = 0; x < 3; ++x) trial[3 * a + x] = trial[3 * a + x] + t[x];
                    disp[n_moved++] = label;
                    moved = a;
                }
                if (!(sm->chain & 1)) break;
            }
            free(disp);
            if (n_moved > 0 || did_cell) {
                double e_new;
                const double *cell_new = did_cell ? new_cell : st->cell;
                int64_t np_ = qo_energy(pot, n, trial, cell_new, st->pbc, &e_new, NULL, NULL);
                cnt[0] += np_; cnt[3] += 1;
                delta = e_new - st->last_energy;
                double x = -delta / kT;
                if (did_cell) { /* IsobaricCriteria (mc/criteria.py:160-172) */
                    double v_new = volume(new_cell), v_old = volume(st->cell);
                    x = -(delta + st->pressure * (v_new - v_old)) / kT + (double)(n + 1) * log(v_new / v_old);
                    moved = -1;
                }
                if (x > 709.782712893384) { rc = -4; goto done; }
                accepted = u_accept < exp(x);
                if (accepted) {
                    memcpy(st->pos, trial, sizeof(double) * 3 * (size_t)n);
                    if (did_cell) memcpy(st->cell, new_cell, sizeof(new_cell));
                    st->last_energy = e_new;
                }
            }
        } else if (mv->type == QO_MOVE_DISPLACEMENT) {
            int nu = unique_labels(mv->labels, n, uniq);
            if (nu > 0) {
                int label = uniq[(int)(u_label * (double)nu)];
                /* moving indices = where(labels == label) (moves/displacement.py:164): a label GROUP moves rigidly by one
                 * translation (translation[_moving_indices] = operation.calculate(context), :126-127) */
                int a = -1;
                for (int i = 0; i < n && a < 0; ++i)
                    if (mv->labels[i] == label) a = i;
                if (a < 0) { rc = -2; goto done; }
                double t[3];
                if (mv->op == QO_OP_BALL) { /* operations/displacement.py:146-153 */
                    double r = 0.0 + (mv->step - 0.0) * opd[0];
                    double phi = 0.0 + (TWO_PI - 0.0) * opd[1];
                    double c = -1.0 + (1.0 - -1.0) * opd[2];
                    double s = sqrt(1 - c * c);
                    t[0] = r * s * cos(phi); t[1] = r * s * sin(phi); t[2] = r * c;
                } else if (mv->op == QO_OP_SPHERE) { /* :103-110 */
                    double phi = 0.0 + (TWO_PI - 0.0) * opd[0];
                    double c = -1.0 + (1.0 - -1.0) * opd[1];
                    double s = sqrt(1 - c * c);
                    t[0] = mv->step * (s * cos(phi)); t[1] = mv->step * (s * sin(phi)); t[2] = mv->step * c;
                } else if (mv->op == QO_OP_BOX) { /* :67-68 */
                    for (int x = 0; x < 3; ++x) t[x] = -mv->step + (mv->step - -mv->step) * opd[x];
                } else if (mv->op == QO_OP_ROTATION) {
                    t[0] = t[1] = t[2] = 0.0; /* per-atom displacements below */
                } else if (mv->op == QO_OP_TRANSLATION || mv->op == QO_OP_TRANSLATION_ROTATION) {
                    /* Translation.calculate (operations/displacement.py:170-175): uniform(0, 1, (1, 3)) @ cell -
                     * positions[moving].mean(axis=0); the mean adds the rows in index order and divides by their number */
                    double mean[3] = {0, 0, 0};
                    int cntm = 0;
                    for (int i = 0; i < n; ++i)
                        if (mv->labels[i] == label) {
                            for (int x = 0; x < 3; ++x) mean[x] += st->pos[3 * i + x];
                            ++cntm;
                        }
                    for (int x = 0; x < 3; ++x) {
                        double fc = 0.0;
                        for (int q = 0; q < 3; ++q) fc += (0.0 + (1.0 - 0.0) * opd[q]) * st->cell[3 * q + x];
                        t[x] = fc - mean[x] / (double)cntm;
                    }
                } else { rc = -1; goto done; }
                /* CompositeOperation (operations/composite.py:58-73): np.sum of the operations' displacements, each drawing
                 * in turn.  Their draws are the attempt's extra draws e = 0, 1, ... (block 64 + e / 2, half e & 1). */
                for (int xo = 0, e = 0; xo < mv->n_extra_ops; ++xo) {
                    struct { int op; double step; } one = {mv->extra_op[xo], mv->extra_step[xo]}, *om = &one;
                    double od[3], t2[3];
                    const int nd = om->op == QO_OP_SPHERE ? 2 : 3;
                    for (int q = 0; q < nd; ++q, ++e) {
                        double bx[2];
                        qo_uniform_pair(seed, attempt, replica, (uint32_t)(64 + e / 2), bx);
                        od[q] = bx[e & 1];
                    }
                    if (om->op == QO_OP_BALL) {
                        double r = 0.0 + (om->step - 0.0) * od[0], phi = 0.0 + (TWO_PI - 0.0) * od[1], cc = -1.0 + (1.0 - -1.0) * od[2];
                        double sq = sqrt(1 - cc * cc);
                        t2[0] = r * sq * cos(phi); t2[1] = r * sq * sin(phi); t2[2] = r * cc;
                    } else if (om->op == QO_OP_SPHERE) {
                        double phi = 0.0 + (TWO_PI - 0.0) * od[0], cc = -1.0 + (1.0 - -1.0) * od[1];
                        double sq = sqrt(1 - cc * cc);
                        t2[0] = om->step * (sq * cos(phi)); t2[1] = om->step * (sq * sin(phi)); t2[2] = om->step * cc;
                    } else if (om->op == QO_OP_BOX) {
                        for (int x = 0; x < 3; ++x) t2[x] = -om->step + (om->step - -om->step) * od[x];
                    } else { rc = -1; goto done; }
                    for (int x = 0; x < 3; ++x) t[x] = t[x] + t2[x];
                }
                memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
                const int rotates = mv->op == QO_OP_ROTATION || mv->op == QO_OP_TRANSLATION_ROTATION;
                if (rotates) {
                    /* Rotation.calculate (operations/displacement.py:192-200): molecule.euler_rotate(phi, theta, psi, center="COM")
                     * with phi, theta, psi = uniform(0, 2 pi, 3); ASE reads the angles as DEGREES [ASE-recalled: ase/atoms.py
                     * euler_rotate, get_center_of_mass]; the move adds molecule.positions - positions[moving] */
                    double ang[3], com[3] = {0, 0, 0}, msum = 0.0;
                    double ua[3] = {opd[0], opd[1], opd[2]};
                    if (mv->op == QO_OP_TRANSLATION_ROTATION) /* TranslationRotation (:220-227): the translation drew first; the
                                                               * angles are the attempt's extra draws 0, 1, 2 */
                        for (int e = 0; e < 3; ++e) {
                            double bx[2];
                            qo_uniform_pair(seed, attempt, replica, (uint32_t)(64 + e / 2), bx);
                            ua[e] = bx[e & 1];
                        }
                    for (int q = 0; q < 3; ++q) ang[q] = (0.0 + (TWO_PI - 0.0) * ua[q]) * (3.141592653589793 / 180);
                    for (int i = 0; i < n; ++i)
                        if (mv->labels[i] == label) {
                            for (int x = 0; x < 3; ++x) com[x] += st->mass * st->pos[3 * i + x];
                            msum += st->mass;
                        }
                    for (int x = 0; x < 3; ++x) com[x] /= msum;
                    const double cphi = cos(ang[0]), sphi = sin(ang[0]), cth = cos(ang[1]), sth = sin(ang[1]), cpsi = cos(ang[2]), spsi = sin(ang[2]);
                    const double D[3][3] = {{cphi, sphi, 0.0}, {-sphi, cphi, 0.0}, {0.0, 0.0, 1.0}};
                    const double Cm[3][3] = {{1.0, 0.0, 0.0}, {0.0, cth, sth}, {0.0, -sth, cth}};
                    const double B[3][3] = {{cpsi, spsi, 0.0}, {-spsi, cpsi, 0.0}, {0.0, 0.0, 1.0}};
                    double CD[3][3], A[3][3];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) CD[i][j] = (Cm[i][0] * D[0][j] + Cm[i][1] * D[1][j]) + Cm[i][2] * D[2][j];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) A[i][j] = (B[i][0] * CD[0][j] + B[i][1] * CD[1][j]) + B[i][2] * CD[2][j];
                    for (int i = 0; i < n; ++i)
                        if (mv->labels[i] == label) {
                            double rcv[3], xn[3];
                            for (int x = 0; x < 3; ++x) rcv[x] = st->pos[3 * i + x] - com[x];
                            for (int x = 0; x < 3; ++x) xn[x] = ((A[x][0] * rcv[0] + A[x][1] * rcv[1]) + A[x][2] * rcv[2]) + com[x];
                            /* translation (1, 3) + rotation (k, 3), then positions[moving] += displacement */
                            for (int x = 0; x < 3; ++x)
                                trial[3 * i + x] = mv->op == QO_OP_ROTATION ? st->pos[3 * i + x] + (xn[x] - st->pos[3 * i + x])
                                                                            : st->pos[3 * i + x] + (t[x] + (xn[x] - st->pos[3 * i + x]));
                        }
                }
                for (int i = 0; i < n; ++i)
                    if (!rotates && mv->labels[i] == label)
                        for (int x = 0; x < 3; ++x) trial[3 * i + x] = st->pos[3 * i + x] + t[x];
                moved = a; /* first member of the group */
                double e_new;
                int64_t np_ = qo_energy(pot, n, trial, st->cell, st->pbc, &e_new, NULL, NULL);
                cnt[0] += np_; cnt[3] += 1;
                memset(seen, 0, (size_t)n);
                for (int i = 0; i < n; ++i)
                    if (mv->labels[i] == label) {
                        seen[i] = 1;
                        cnt[1] += count_neighbours(pot, n, st->pos, st->cell, st->pbc, i, seen);
                        cnt[1] += count_neighbours(pot, n, trial, st->cell, st->pbc, i, seen);
                    }
                for (int i = 0; i < n; ++i) cnt[2] += seen[i];
                delta = e_new - st->last_energy;
                double x = -delta / kT; /* mc/criteria.py:104-110 */
                if (x > 709.782712893384) { rc = -4; goto done; }
                accepted = u_accept < exp(x);
                if (accepted) {
                    memcpy(st->pos, trial, sizeof(double) * 3 * (size_t)n);
                    st->last_energy = e_new;
                }
            }
        } else if (mv->type == QO_MOVE_CELL) {
            /* operations/cell.py:67-169, moves/cell.py:73-84, mc/criteria.py:160-219.  Draws of Anisotropic / ShapeDeformation
             * (uniform(-max, max, size=6)): components 0..2 are the operation draws, 3..5 the attempt's extra draws 0..2 */
            double u6[6] = {opd[0], opd[1], opd[2], 0, 0, 0}, F[9], new_cell[9];
            if (mv->op != QO_OP_ISOTROPIC) {
                double e0[2], e1[2];
                qo_uniform_pair(seed, attempt, replica, 64, e0);
                qo_uniform_pair(seed, attempt, replica, 65, e1);
                u6[3] = e0[0]; u6[4] = e0[1]; u6[5] = e1[0];
            }
            deformation_gradient(mv, u6, F);
            if (mv->op == QO_OP_ISOTROPIC) { /* F is diagonal: row r scaled by F[r][r], exactly */
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) new_cell[3 * r + c] = F[4 * r] * st->cell[3 * r + c];
            } else {
                mat3_mul(F, st->cell, new_cell);
            }
            memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
            scale_positions(n, trial, st->cell, new_cell);
            double e_new;
            int64_t np_ = qo_energy(pot, n, trial, new_cell, st->pbc, &e_new, NULL, NULL);
            cnt[0] += np_; cnt[3] += 1; cnt[1] += np_; cnt[2] += n;
            delta = e_new - st->last_energy;
            double v_new = volume(new_cell), v_old = volume(st->cell);
            double x = -(delta + cell_work(st, new_cell)) / kT + (double)(n + 1) * log(v_new / v_old);
            if (x > 709.782712893384) { rc = -4; goto done; }
            accepted = u_accept < exp(x);
            if (accepted) {
                memcpy(st->pos, trial, sizeof(double) * 3 * (size_t)n);
                memcpy(st->cell, new_cell, sizeof(new_cell));
                st->last_energy = e_new;
            }
        } else if (mv->type == QO_MOVE_EXCHANGE) {
            /* moves/exchange.py:134-176 */
            int is_addition = u_coin < mv->bias_insert;
            int n_new = n, particle_delta = 0, removed = -1;
            const int n_mol = st->n_exchange_atoms > 1 ? st->n_exchange_atoms : 1;
            /* repeat = k > 1: CompositeExchangeMove of k single-atom ExchangeMoves holding equal labels (moves/exchange.py:237-316):
             * ONE coin for all of them (the composite's own bias_towards_insert); insertion = every sub-move places one atom with
             * its own Translation draws (sub-move 0: the standard operation slots, sub-move c >= 1: "extra" draws 3 (c - 1) + {0,
             * 1, 2} = block 64 + e / 2, half e & 1); deletion = every sub-move draws a label among those nobody deleted yet in this
             * composite (sub-move 0: u_label, drawing sub-move c >= 1: block 32 + 2 (c - 1), first double) and ALL atoms carrying
             * it go; particle_delta = +-(number of sub-moves that acted); GrandCanonical.save_state then gives the k added atoms
             * ONE new label (moves/displacement.py:244-249), so that later deletions take them out together. */
            const int k_sub = mv->repeat > 1 ? mv->repeat : 1;
            const int composite = k_sub > 1;
            /* (chain bit 3: the caller's labels hold GROUPS -- `where(labels == label)`, moves/exchange.py:127, takes out every atom of
             * the drawn label whatever the size of exchange_atoms) */
            const int molecular = composite || n_mol > 1 || mv->op == QO_OP_TRANSLATION_ROTATION || (mv->chain & 8);
            int n_gone = 0;
            if (composite && (n_mol > 1 || mv->op != QO_OP_TRANSLATION || !mv->labels)) { rc = -1; goto done; }
            if (composite && is_addition) {
                if (n + k_sub > nmax) { rc = -3; goto done; }
                memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
                const double *tmpl = st->n_exchange_atoms >= 1 ? st->exchange_mol : st->exchange_pos;
                for (int c = 0; c < k_sub; ++c) {
                    double f[3];
                    for (int j = 0; j < 3; ++j) {
                        if (c == 0) f[j] = opd[j];
                        else {
                            const int e = 3 * (c - 1) + j;
                            double bx[2];
                            qo_uniform_pair(seed, attempt, replica, (uint32_t)(64 + e / 2), bx);
                            f[j] = bx[e & 1];
                        }
                    }
                    for (int x = 0; x < 3; ++x) {
                        double fc = (f[0] * st->cell[x] + f[1] * st->cell[3 + x]) + f[2] * st->cell[6 + x];
                        trial[3 * (n + c) + x] = tmpl[x] + (fc - tmpl[x] / 1.0);
                    }
                }
                n_new = n + k_sub; particle_delta = k_sub; moved = n;
                memset(seen, 0, (size_t)n_new);
                for (int q = 0; q < k_sub; ++q) {
                    seen[n + q] = 1;
                    cnt[1] += count_neighbours(pot, n_new, trial, st->cell, st->pbc, n + q, seen);
                }
                for (int i = 0; i < n_new; ++i) cnt[2] += seen[i];
            } else if (composite) {
                int n_del = 0, c = 0;
                int deleted[16];
                memset(gone, 0, (size_t)n);
                memset(seen, 0, (size_t)n);
                for (int sub = 0; sub < k_sub; ++sub) {
                    int nu = unique_labels(mv->labels, n, uniq), na = 0;
                    for (int q = 0; q < nu; ++q) { /* setdiff1d(unique_labels, deleted_labels) */
                        int taken = 0;
                        for (int d = 0; d < n_del; ++d) taken |= deleted[d] == uniq[q];
                        if (!taken) uniq[na++] = uniq[q];
                    }
                    if (na == 0) continue;
                    double ul = u_label;
                    if (c > 0) {
                        double e0[2];
                        qo_uniform_pair(seed, attempt, replica, submove_block(c), e0);
                        ul = e0[0];
                    }
                    ++c;
                    const int label = uniq[(int)(ul * (double)na)];
                    if (n_del >= 16) { rc = -1; goto done; }
                    deleted[n_del++] = label;
                    for (int i = 0; i < n; ++i)
                        if (mv->labels[i] == label) {
                            if (removed < 0) removed = i;
                            gone[i] = 1;
                            ++n_gone;
                            cnt[1] += count_neighbours(pot, n, st->pos, st->cell, st->pbc, i, seen);
                        }
                }
                if (n_del > 0) {
                    for (int i = 0; i < n; ++i) cnt[2] += seen[i];
                    for (int i = 0, o = 0; i < n; ++i)
                        if (!gone[i]) { memcpy(trial + 3 * o, st->pos + 3 * i, sizeof(double) * 3); ++o; }
                    n_new = n - n_gone; particle_delta = -n_del; moved = removed;
                }
            } else if (is_addition) {
                if (n + n_mol > nmax) { rc = -3; goto done; }
                memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
                const double *tmpl = st->n_exchange_atoms >= 1 ? st->exchange_mol : st->exchange_pos;
                /* atoms.extend(exchange_atoms), then attempt_displacement on the new indices (moves/exchange.py:103-112).
                 * Translation (operations/displacement.py:173-175): uniform(0,1,(1,3)) @ cell - mean(pos[moving]) */
                double mean[3] = {0, 0, 0}, t[3];
                for (int q = 0; q < n_mol; ++q)
                    for (int x = 0; x < 3; ++x) mean[x] += tmpl[3 * q + x];
                for (int x = 0; x < 3; ++x) {
                    double fc = (opd[0] * st->cell[x] + opd[1] * st->cell[3 + x]) + opd[2] * st->cell[6 + x];
                    t[x] = fc - mean[x] / (double)n_mol;
                }
                if (mv->op == QO_OP_TRANSLATION_ROTATION) {
                    /* + Rotation.calculate (:192-200) of the still untranslated molecule; its angles are extra draws 0, 1, 2 */
                    double ua[3], ang[3], com[3] = {0, 0, 0}, msum = 0.0;
                    const double m1 = st->exchange_mass / (double)n_mol; /* single species */
                    for (int e = 0; e < 3; ++e) {
                        double bx[2];
                        qo_uniform_pair(seed, attempt, replica, (uint32_t)(64 + e / 2), bx);
                        ua[e] = bx[e & 1];
                    }
                    for (int q = 0; q < 3; ++q) ang[q] = (0.0 + (TWO_PI - 0.0) * ua[q]) * (3.141592653589793 / 180);
                    for (int q = 0; q < n_mol; ++q) {
                        for (int x = 0; x < 3; ++x) com[x] += m1 * tmpl[3 * q + x];
                        msum += m1;
                    }
                    for (int x = 0; x < 3; ++x) com[x] /= msum;
                    const double cphi = cos(ang[0]), sphi = sin(ang[0]), cth = cos(ang[1]), sth = sin(ang[1]), cpsi = cos(ang[2]), spsi = sin(ang[2]);
                    const double D[3][3] = {{cphi, sphi, 0.0}, {-sphi, cphi, 0.0}, {0.0, 0.0, 1.0}};
                    const double Cm[3][3] = {{1.0, 0.0, 0.0}, {0.0, cth, sth}, {0.0, -sth, cth}};
                    const double B[3][3] = {{cpsi, spsi, 0.0}, {-spsi, cpsi, 0.0}, {0.0, 0.0, 1.0}};
                    double CD[3][3], A[3][3];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) CD[i][j] = (Cm[i][0] * D[0][j] + Cm[i][1] * D[1][j]) + Cm[i][2] * D[2][j];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) A[i][j] = (B[i][0] * CD[0][j] + B[i][1] * CD[1][j]) + B[i][2] * CD[2][j];
                    for (int q = 0; q < n_mol; ++q) {
                        double rcv[3], xn[3];
                        for (int x = 0; x < 3; ++x) rcv[x] = tmpl[3 * q + x] - com[x];
                        for (int x = 0; x < 3; ++x) xn[x] = ((A[x][0] * rcv[0] + A[x][1] * rcv[1]) + A[x][2] * rcv[2]) + com[x];
                        for (int x = 0; x < 3; ++x) trial[3 * (n + q) + x] = tmpl[3 * q + x] + (t[x] + (xn[x] - tmpl[3 * q + x]));
                    }
                } else {
                    for (int q = 0; q < n_mol; ++q)
                        for (int x = 0; x < 3; ++x) trial[3 * (n + q) + x] = tmpl[3 * q + x] + t[x];
                }
                n_new = n + n_mol; particle_delta = 1; moved = n;
                memset(seen, 0, (size_t)n_new);
                for (int q = 0; q < n_mol; ++q) {
                    seen[n + q] = 1;
                    cnt[1] += count_neighbours(pot, n_new, trial, st->cell, st->pbc, n + q, seen);
                }
                for (int i = 0; i < n_new; ++i) cnt[2] += seen[i];
            } else {
                int nu = unique_labels(mv->labels, n, uniq);
                if (nu > 0) {
                    int label = uniq[(int)(u_label * (double)nu)];
                    int a = -1;
                    if (molecular) {
                        for (int i = 0; i < n && a < 0; ++i)
                            if (mv->labels[i] == label) a = i;
                    } else {
                        a = atom_of_label(mv->labels, n, label);
                    }
                    if (a < 0) { rc = -2; goto done; }
                    memset(seen, 0, (size_t)n);
                    /* indices = where(labels == label) (moves/exchange.py:129): the whole molecule goes */
                    memset(gone, 0, (size_t)n);
                    for (int i = 0; i < n; ++i)
                        if (molecular ? mv->labels[i] == label : i == a) {
                            gone[i] = 1;
                            ++n_gone;
                            cnt[1] += count_neighbours(pot, n, st->pos, st->cell, st->pbc, i, seen);
                        }
                    for (int i = 0; i < n; ++i) cnt[2] += seen[i];
                    for (int i = 0, o = 0; i < n; ++i)
                        if (!gone[i]) { memcpy(trial + 3 * o, st->pos + 3 * i, sizeof(double) * 3); ++o; }
                    n_new = n - n_gone; particle_delta = -1; removed = a; moved = a;
                }
            }
            if (particle_delta != 0) {
                double e_new;
                int64_t np_ = qo_energy(pot, n_new, trial, st->cell, st->pbc, &e_new, NULL, NULL);
                cnt[0] += np_; cnt[3] += 1;
                delta = e_new - st->last_energy;
                int ov = 0;
                double prob = gc_probability(delta, particle_delta, st->n_exchange, st->accessible_volume, st->exchange_mass,
                                             st->temperature, st->chemical_potential, &ov);
                if (ov) { rc = -4; goto done; }
                accepted = u_accept < prob;
                if (accepted) {
                    /* GrandCanonical.save_state (mc/gcmc.py:207-214): every move re-labels, then N_ex += delta */
                    for (int q = 0; q < n_moves; ++q)
                        if (moves[q].labels) {
                            if (molecular) labels_on_molecule_changed(&moves[q], n, particle_delta > 0 ? (composite ? k_sub : n_mol) : 0, particle_delta < 0 ? gone : NULL, uniq);
                            else labels_on_atoms_changed(&moves[q], n, particle_delta > 0, removed, uniq);
                        }
                    memcpy(st->pos, trial, sizeof(double) * 3 * (size_t)n_new);
                    if (st->momenta) {
                        if (particle_delta > 0) memset(st->momenta + 3 * n, 0, sizeof(double) * 3 * (size_t)(n_new - n));
                        else
                            for (int i = 0, o = 0; i < n; ++i)
                                if (!gone[i]) { memmove(st->momenta + 3 * o, st->momenta + 3 * i, sizeof(double) * 3); ++o; }
                    }
                    st->n_atoms = n_new;
                    st->n_exchange += particle_delta;
                    st->last_energy = e_new;
                }
            }
        } else if (mv->type == QO_MOVE_HMC) {
            /* moves/displacement.py:307-345, utils/dynamics.py:27-43, integrators/displacement.py:52-81,
             * mc/criteria.py:117-140, mc/contexts.py:186-198 */
            if (!st->momenta) { rc = -1; goto done; }
            const double mass = st->mass, dt = mv->dt;
            const double pscale = sqrt(mass * kT);
            for (int i = 0; i < 3 * n; ++i) mom[i] = qo_normal(seed, attempt, replica, (uint32_t)i) * pscale;
            for (int i = 0; i < 3 * n; ++i) mom[i] = mom[i] * 1.0; /* set_momenta(get_momenta() * scale), scale = 1.0 */
            st->last_kinetic = kinetic_energy(n, mom, mass);
            memcpy(trial, st->pos, sizeof(double) * 3 * (size_t)n);
            double e_pot;
            int64_t np_ = qo_energy(pot, n, trial, st->cell, st->pbc, &e_pot, frc, NULL);
            cnt[0] += np_; cnt[3] += 1; cnt[1] += np_; cnt[2] += n;
            for (int s = 0; s < mv->n_steps; ++s) {
                for (int i = 0; i < 3 * n; ++i) {
                    double new_mom = mom[i] + 0.5 * frc[i] * dt;
                    double x_old = trial[i];
                    double x_new = x_old + new_mom / mass * dt;
                    trial[i] = x_new;
                    mom[i] = (x_new - x_old) * mass / dt; /* apply_constraints=True re-derivation (:74-75) */
                }
                np_ = qo_energy(pot, n, trial, st->cell, st->pbc, &e_pot, frc, NULL);
                cnt[0] += np_; cnt[3] += 1; cnt[1] += np_; cnt[2] += n;
                for (int i = 0; i < 3 * n; ++i) mom[i] = mom[i] + 0.5 * frc[i] * dt;
            }
            double e_kin = kinetic_energy(n, mom, mass);
            delta = (e_pot + e_kin) - st->last_energy - st->last_kinetic;
            double x = -delta / kT;
            if (x > 709.782712893384) { rc = -4; goto done; }
            accepted = u_accept < exp(x);
            if (accepted) {
                memcpy(st->pos, trial, sizeof(double) * 3 * (size_t)n);
                memcpy(st->momenta, mom, sizeof(double) * 3 * (size_t)n);
                st->last_energy = e_pot;
                st->last_kinetic = e_kin;
            }
        } else { rc = -1; goto done; }

        if (trace) {
            if (trace->move) trace->move[k] = (int8_t)m;
            if (trace->accepted) trace->accepted[k] = (int8_t)accepted;
            if (trace->energy) trace->energy[k] = st->last_energy;
            if (trace->delta) trace->delta[k] = delta;
            if (trace->n_atoms) trace->n_atoms[k] = st->n_atoms;
            if (trace->moved) trace->moved[k] = moved;
        }
    }
done:
    if (counters) for (int i = 0; i < 4; ++i) counters[i] += cnt[i];
    free(trial); free(mom); free(frc); free(uniq); free(seen); free(gone);
    return rc;
}
