// Hamiltonian Monte Carlo: force + velocity-Verlet trajectory + acceptance, and parallel-tempering swaps.
//
// One attempt (reference file:line):
//   momenta     p = z sqrt(m kB T), z ~ N(0,1) in row-major (N,3) order            [utils/dynamics.py:27-43]
//   KE_0        0.5 sum p.p/m -> last_kinetic_energy                                [moves/displacement.py:334]
//   trajectory  F = forces; max_steps x { p += F dt/2; x += p/m dt; p = (x_new - x) m / dt (apply_constraints);
//               F = forces(x); p += F dt/2 }                                        [integrators/displacement.py:52-81]
//   criteria    u < exp(-((E_pot + E_kin) - last_pot - last_kin) / kT)              [mc/criteria.py:117-140]
//   commit      positions / momenta / energies kept on accept, untouched on reject  [mc/contexts.py:144-156,186-198]
//
// Organisation (config C5: 64 replicas x 2048 atoms, 8 per GPU on 8 GPUs): LPA lanes per atom (chosen so that the whole
// GPU is busy whatever the replica count), a Verlet neighbour list with skin in global memory ([slot][atom]: coalesced
// for LPA = 1, L2-resident), double-buffered positions, and four kernels per force evaluation:
//   build   (only when an atom moved more than skin / 2 since the last build: the kernel exits at once otherwise); one thread
//           per atom for batches of small replicas, up to 32 lanes per atom with ballot compaction and chunk bounding boxes for
//           large ones (hmc_bbox_kernel + hmc_build_split_kernel)
//   pass 1  sigma_1, pair energy, embedding energy and dE/dsigma_1 per atom          [emt.py _calc_e_c_a2]
//   pass 2  forces (recomputing the pair functions: streaming 16 B per list entry from pass 1 was measured and is no faster,
//           profiles/r1_m_hmc.md), fused with the
//           velocity-Verlet update of the atom and the displacement check            [emt.py _calc_fs_c_a2 / _calc_efs_a1]
//   finish  deterministic reductions (fixed-order sums of per-CTA partials), acceptance test, commit
// Forces are the analytic gradient of exactly the energy expression of the sweep kernel.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <mutex>
#include <string>
#include <vector>

#include "replica_warp.cuh"

namespace qmc {

constexpr int HMC_THREADS = 128;
constexpr double HMC_SKIN_DEFAULT = 0.7;  // Angstrom; measured on C5 (64 x 2048 atoms): 0.3 .. 1.0 -> best 0.6 - 0.7 (profiles/r1_m_hmc.md)
// Verlet-list skin: entries within cutoff + skin are listed, the list is rebuilt once an atom has moved skin / 2.  QMC_HMC_SKIN
// (environment, read at every call) overrides the default for tuning and tests.
__host__ inline double hmc_skin() {
    const char *e = std::getenv("QMC_HMC_SKIN");
    const double v = e ? std::atof(e) : 0.0;
    return v > 0.0 && v <= 4.0 ? v : HMC_SKIN_DEFAULT;
}

struct HmcWork {  // carved from the caller's workspace
    double *x[2];      // [R][3][N] double-buffered positions
    double *p;         // [R][3][N] momenta
    double *deds;      // [R][N]
    double *ref;       // [R][3][N] positions at the last list build
    double *partial;   // [R][NB][2] per-CTA partial sums (energy, kinetic)
    double *ke0;       // [R][NB]
    double *result;    // [R][2] accepted (energy, kinetic) of the attempt, published after every CTA has read the old values
    unsigned *list;    // [R][max_nb][N]
    int *nnb;          // [R][N]
    int *flags;        // [R] rebuild requested
    float *bbox;       // [R][ceil(N / 32)][6] centre and half-width of every 32-atom chunk (list build)
    unsigned long long *ctr;  // [2] attempt index, trace column of the trajectory in flight (advanced on the device: one CUDA graph
                              // serves every trajectory)
    int max_nb, nb_blocks;
};

struct HmcArgs {
    PotDev pot;
    qmc_batch_t b;
    HmcWork w;
    double dt;
    int n_steps;
    unsigned long long seed;
    int n_attempts;
    qmc_trace_t tr;
    double *forces_out;
    int lpa;   // lanes per atom (power of two <= 32)
    int one_wave;  // the batch fits the GPU as one wave of 1024 threads per SM: pass 2 runs in its 64-register instantiation
    double skin;  // Verlet-list skin (hmc_skin())
    int build_lanes;  // lanes per atom of the list build (1: hmc_build_kernel)
    int cur;   // position buffer holding the current positions
    int mode;  // pass 2: 0 first evaluation, 1 middle, 2 last, 3 forces only, 4 single evaluation of a zero-step trajectory
};

__host__ inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

__host__ inline int hmc_max_neighbours(const PotDev &pot, const qmc_batch_t &b) {
    // List slots per atom.  A fully periodic box shared by every replica is sized from its mean density (x 1.5); anything with
    // vacuum in it (slabs, clusters, per-replica cells) cannot be: the atoms sit at condensed-phase density whatever the box
    // volume says, so those are sized for the densest packing the potential supports (0.10 atoms / A^3 covers the EMT metals,
    // 1.4 / sigma^3 a compressed Lennard-Jones solid).
    const double rl = pot.cut + hmc_skin(), sphere = 4.18879 * rl * rl * rl;
    const double dense = pot.kind == QMC_POT_EMT ? 0.10 : 1.4 / (pot.sigma2 * std::sqrt(pot.sigma2));
    double est = dense * sphere * 1.3 + 24.0;
    const bool periodic = b.pbc[0] && b.pbc[1] && b.pbc[2];
    if (periodic && b.cells_uniform && !b.triclinic && b.cell_diag[0] > 0 && b.cell_diag[1] > 0 && b.cell_diag[2] > 0) {
        const double density = (double)b.n_max / (b.cell_diag[0] * b.cell_diag[1] * b.cell_diag[2]);
        est = density * sphere * 1.5 + 24.0;
    }
    int m = (int)est;
    m = (m + 7) / 8 * 8;
    return m < 32 ? 32 : (m > 1024 ? 1024 : m);
}

__host__ inline int64_t hmc_carve(const PotDev &pot, const qmc_batch_t &b, void *base, HmcWork *w) {
    const int64_t R = b.n_replicas, N = b.n_max;
    const int max_nb = hmc_max_neighbours(pot, b);
    const int nb_blocks = (int)((N * 32 + HMC_THREADS - 1) / HMC_THREADS);  // upper bound over all lanes-per-atom choices
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void *p = base ? (char *)base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    void *x0 = take(R * 3 * N * 8), *x1 = take(R * 3 * N * 8), *p = take(R * 3 * N * 8), *deds = take(R * N * 8), *ref = take(R * 3 * N * 8);
    void *partial = take(R * nb_blocks * 2 * 8), *ke0 = take(R * nb_blocks * 8), *list = take(R * (int64_t)max_nb * N * 4);
    void *nnb = take(R * N * 4), *flags = take(R * 4), *result = take(R * 2 * 8), *bbox = take(R * ((N + 31) / 32) * 6 * 4);
    void *ctr = take(16);
    if (w) {
        w->x[0] = (double *)x0; w->x[1] = (double *)x1; w->p = (double *)p; w->deds = (double *)deds; w->ref = (double *)ref;
        w->partial = (double *)partial; w->ke0 = (double *)ke0; w->list = (unsigned *)list; w->nnb = (int *)nnb; w->flags = (int *)flags; w->result = (double *)result; w->bbox = (float *)bbox; w->ctr = (unsigned long long *)ctr;
        w->max_nb = max_nb; w->nb_blocks = nb_blocks;
    }
    return off;
}

// deterministic block sum (fixed tree); result on every thread
__device__ inline double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[wid] = v;
    __syncthreads();
    double t = (int)(threadIdx.x & 31) < nw ? red[threadIdx.x & 31] : 0.0;
    return warp_sum(t);
}

struct CellDev {
    double L[3], invL[3], thr[3];
    bool per[3];
    // general (triclinic) cells, qmc_batch_t.triclinic: rows of H are the cell vectors, Hi = H^-1 (the column of a non-periodic axis
    // zeroed), thr = fractional threshold 1 - cut / height (csrc/replica_warp.cuh: TriView has the image rule)
    bool tri;
    double H[9], Hi[9];
};

template <bool TRI>
__device__ inline bool hmc_cell(const HmcArgs &a, int r, double cut, CellDev &C) {
    bool ok = true;
    C.tri = TRI;
    if (TRI) {
        double h[9];
        for (int q = 0; q < 9; ++q) h[q] = C.H[q] = a.b.cell[(size_t)r * 9 + q];
        const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]);
        const double d = 1.0 / det;
        const double inv[9] = {(h[4] * h[8] - h[5] * h[7]) * d, (h[2] * h[7] - h[1] * h[8]) * d, (h[1] * h[5] - h[2] * h[4]) * d,
                               (h[5] * h[6] - h[3] * h[8]) * d, (h[0] * h[8] - h[2] * h[6]) * d, (h[2] * h[3] - h[0] * h[5]) * d,
                               (h[3] * h[7] - h[4] * h[6]) * d, (h[1] * h[6] - h[0] * h[7]) * d, (h[0] * h[4] - h[1] * h[3]) * d};
        for (int k = 0; k < 3; ++k) {
            const double *p = h + 3 * ((k + 1) % 3), *q = h + 3 * ((k + 2) % 3);
            const double cx = p[1] * q[2] - p[2] * q[1], cy = p[2] * q[0] - p[0] * q[2], cz = p[0] * q[1] - p[1] * q[0];
            const double w = fabs(det) / sqrt(cx * cx + cy * cy + cz * cz);  // height of the cell along axis k
            C.per[k] = a.b.pbc[k] != 0;
            C.L[k] = w;
            C.invL[k] = 0.0;
            C.thr[k] = C.per[k] ? 1.0 - cut / w - 1e-9 : 1e300;
            for (int x = 0; x < 3; ++x) C.Hi[3 * x + k] = C.per[k] ? inv[3 * x + k] : 0.0;
            if (C.per[k] && !(w >= cut)) ok = false;
        }
        return ok;
    }
    for (int k = 0; k < 3; ++k) {
        const double L = a.b.cell[(size_t)r * 9 + 4 * k];
        C.L[k] = L;
        C.per[k] = a.b.pbc[k] != 0;
        C.invL[k] = C.per[k] ? 1.0 / L : 0.0;
        C.thr[k] = C.per[k] ? L - cut : 1e300;
        if (C.per[k] && !(L >= cut)) ok = false;
    }
    return ok;
}

// Candidate images of atom j seen from atom i for the list build: the image found by rounding (vector d0, integer shift k0) and,
// per flagged axis q (bit q of `avail`), the one shifted by -sgn[q] cell vectors.  Image c (a subset of `avail`) is
// d0 - sum_{q in c} sgn[q] a_q with the integer shift k0[q] - sgn[q].
__device__ inline void hmc_images(const CellDev &C, const double pi[3], const double pj[3], double d0[3], double k0[3], double sgn[3],
                                  unsigned &avail) {
    avail = 0;
    if (C.tri) {
        const double ex = pi[0] - pj[0], ey = pi[1] - pj[1], ez = pi[2] - pj[2];
        double fr[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double f = fma(ez, C.Hi[6 + q], fma(ey, C.Hi[3 + q], ex * C.Hi[q]));
            k0[q] = (f + RINT_MAGIC) - RINT_MAGIC;
            fr[q] = f - k0[q];
            sgn[q] = fr[q] > 0.0 ? -1.0 : 1.0;  // the fractional coordinate of d0 along q is -fr[q]
            if (fabs(fr[q]) > C.thr[q]) avail |= 1u << q;
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) d0[x] = fma(k0[2], C.H[6 + x], fma(k0[1], C.H[3 + x], fma(k0[0], C.H[x], pj[x]))) - pi[x];
        return;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (C.per[q]) {
            const double f = (pi[q] - pj[q]) * C.invL[q];
            k0[q] = (f + RINT_MAGIC) - RINT_MAGIC;
            d0[q] = fma(k0[q], C.L[q], pj[q]) - pi[q];
            if (fabs(d0[q]) > C.thr[q]) avail |= 1u << q;
        } else {
            k0[q] = 0.0;
            d0[q] = pj[q] - pi[q];
        }
        sgn[q] = d0[q] > 0.0 ? 1.0 : -1.0;
    }
}

__device__ inline void hmc_image_vector(const CellDev &C, const double d0[3], const double sgn[3], unsigned c, double d[3]) {
    if (C.tri) {
        const double m0 = (c & 1u) ? sgn[0] : 0.0, m1 = (c & 2u) ? sgn[1] : 0.0, m2 = (c & 4u) ? sgn[2] : 0.0;
#pragma unroll
        for (int x = 0; x < 3; ++x) d[x] = fma(-m2, C.H[6 + x], fma(-m1, C.H[3 + x], fma(-m0, C.H[x], d0[x])));
        return;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) d[q] = ((c >> q) & 1u) ? d0[q] - copysign(C.L[q], d0[q]) : d0[q];
}

__device__ inline double hmc_normal(PhiloxKey key, unsigned long long attempt, uint32_t replica, uint32_t m) {
    double d0, d1;
    uniform_pair(key, attempt, replica, 16u + m / 2u, d0, d1);
    const double rr = sqrt(-2.0 * log(1.0 - d0));
    const double t = TWO_PI * d1;
    return (m % 2u == 0u) ? rr * cos(t) : rr * sin(t);
}

// ---- init: working copies, momenta, KE_0 partials, rebuild request --------------------------------------------
__global__ void __launch_bounds__(HMC_THREADS) hmc_init_kernel(const __grid_constant__ HmcArgs a) {
    __shared__ double red[32];
    const int r = blockIdx.y, nmax = a.b.n_max, n = a.b.n_atoms[r];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = (size_t)r * 3 * nmax;
    double ke = 0.0;
    if (i < n) {
        const double mass = a.b.mass;
        const double pscale = sqrt(mass * (a.b.temperature[r] * KB));
        const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
        const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.w.x[0][base + (size_t)c * nmax + i] = a.b.pos[base + (size_t)c * nmax + i];
            const double p = hmc_normal(key, a.w.ctr[0], grep, 3u * i + c) * pscale;
            a.w.p[base + (size_t)c * nmax + i] = p;
            ke += p * (p / mass);
        }
    }
    const double s = block_sum(ke, red);
    if (threadIdx.x == 0) {
        a.w.ke0[(size_t)r * a.w.nb_blocks + blockIdx.x] = s;
        if (blockIdx.x == 0) a.w.flags[r] = 1;
    }
}

// ---- neighbour list (brute force over tiles staged in shared memory; only when requested) ----------------------
// (diagonal cells: the round-1 code, untouched)
__global__ void __launch_bounds__(HMC_THREADS) hmc_build_kernel_diag(const __grid_constant__ HmcArgs a) {
    __shared__ double sx[HMC_THREADS], sy[HMC_THREADS], sz[HMC_THREADS];
    const int r = blockIdx.y;
    if (a.w.flags[r] == 0) return;
    const int nmax = a.b.n_max, n = a.b.n_atoms[r];
    const double rl = a.pot.cut + a.skin, rl2 = rl * rl;
    CellDev C;
    if (!hmc_cell<false>(a, r, rl * (1.0 + 1e-9), C)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.b.status + r, (int)QMC_STATUS_CELL_TOO_SMALL);
        return;
    }
    const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax, *y = x + nmax, *z = y + nmax;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool own = i < n;
    const double xi = own ? x[i] : 0.0, yi = own ? y[i] : 0.0, zi = own ? z[i] : 0.0;
    unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    int k = 0, overflow = 0;
    for (int j0 = 0; j0 < n; j0 += HMC_THREADS) {
        __syncthreads();
        const int jl = j0 + threadIdx.x;
        if (jl < n) {
            sx[threadIdx.x] = x[jl];
            sy[threadIdx.x] = y[jl];
            sz[threadIdx.x] = z[jl];
        }
        __syncthreads();
        if (!own) continue;
        const int jn = min(HMC_THREADS, n - j0);
        for (int t = 0; t < jn; ++t) {
            const int j = j0 + t;
            if (j == i) continue;
            double d0[3], d1[3], k0[3];
            unsigned avail = 0;
            const double pj[3] = {sx[t], sy[t], sz[t]}, pi[3] = {xi, yi, zi};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (C.per[c]) {
                    const double q = (pi[c] - pj[c]) * C.invL[c];
                    k0[c] = (q + RINT_MAGIC) - RINT_MAGIC;
                    d0[c] = fma(k0[c], C.L[c], pj[c]) - pi[c];
                    d1[c] = d0[c] - copysign(C.L[c], d0[c]);
                    if (fabs(d0[c]) > C.thr[c]) avail |= 1u << c;
                } else {
                    k0[c] = 0.0;
                    d0[c] = d1[c] = pj[c] - pi[c];
                }
            }
            unsigned c = 0;
            while (true) {
                const double dx = (c & 1u) ? d1[0] : d0[0], dy = (c & 2u) ? d1[1] : d0[1], dz = (c & 4u) ? d1[2] : d0[2];
                if (dx * dx + dy * dy + dz * dz < rl2) {
                    // integer image shift per axis: nearest image k0, second image k0 -+ 1
                    int sh[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) sh[q] = (int)k0[q] - (((c >> q) & 1u) ? (d0[q] > 0.0 ? 1 : -1) : 0);
                    if (abs(sh[0]) > 15 || abs(sh[1]) > 15 || abs(sh[2]) > 15 || k >= a.w.max_nb) {
                        overflow = 1;
                    } else {
                        list[(size_t)k * nmax + i] = (unsigned)j | ((unsigned)(sh[0] + 16) << 16) | ((unsigned)(sh[1] + 16) << 21) | ((unsigned)(sh[2] + 16) << 26);
                        ++k;
                    }
                }
                if (c == avail) break;
                c = ((c | (~avail & 7u)) + 1u) & avail;
            }
        }
    }
    if (own) {
        a.w.nnb[(size_t)r * nmax + i] = k;
        double *ref = a.w.ref + (size_t)r * 3 * nmax;
        ref[i] = xi;
        ref[nmax + i] = yi;
        ref[2 * nmax + i] = zi;
        if (overflow) atomicOr(a.b.status + r, (int)QMC_STATUS_LIST_OVERFLOW);
    }
}

// (general cells: images through hmc_images / hmc_image_vector)
__global__ void __launch_bounds__(HMC_THREADS) hmc_build_kernel_tri(const __grid_constant__ HmcArgs a) {
    __shared__ double sx[HMC_THREADS], sy[HMC_THREADS], sz[HMC_THREADS];
    const int r = blockIdx.y;
    if (a.w.flags[r] == 0) return;
    const int nmax = a.b.n_max, n = a.b.n_atoms[r];
    const double rl = a.pot.cut + a.skin, rl2 = rl * rl;
    CellDev C;
    if (!hmc_cell<true>(a, r, rl * (1.0 + 1e-9), C)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.b.status + r, (int)QMC_STATUS_CELL_TOO_SMALL);
        return;
    }
    const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax, *y = x + nmax, *z = y + nmax;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool own = i < n;
    const double xi = own ? x[i] : 0.0, yi = own ? y[i] : 0.0, zi = own ? z[i] : 0.0;
    unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    int k = 0, overflow = 0;
    for (int j0 = 0; j0 < n; j0 += HMC_THREADS) {
        __syncthreads();
        const int jl = j0 + threadIdx.x;
        if (jl < n) {
            sx[threadIdx.x] = x[jl];
            sy[threadIdx.x] = y[jl];
            sz[threadIdx.x] = z[jl];
        }
        __syncthreads();
        if (!own) continue;
        const int jn = min(HMC_THREADS, n - j0);
        for (int t = 0; t < jn; ++t) {
            const int j = j0 + t;
            if (j == i) continue;
            double d0[3], k0[3], sgn[3];
            unsigned avail = 0;
            const double pj[3] = {sx[t], sy[t], sz[t]}, pi[3] = {xi, yi, zi};
            hmc_images(C, pi, pj, d0, k0, sgn, avail);
            unsigned c = 0;
            while (true) {
                double d[3];
                hmc_image_vector(C, d0, sgn, c, d);
                if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < rl2) {
                    // integer image shift per axis: the image found by rounding k0, the further image k0 - sgn
                    int sh[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) sh[q] = (int)k0[q] - (((c >> q) & 1u) ? (int)sgn[q] : 0);
                    if (abs(sh[0]) > 15 || abs(sh[1]) > 15 || abs(sh[2]) > 15 || k >= a.w.max_nb) {
                        overflow = 1;
                    } else {
                        list[(size_t)k * nmax + i] = (unsigned)j | ((unsigned)(sh[0] + 16) << 16) | ((unsigned)(sh[1] + 16) << 21) | ((unsigned)(sh[2] + 16) << 26);
                        ++k;
                    }
                }
                if (c == avail) break;
                c = ((c | (~avail & 7u)) + 1u) & avail;
            }
        }
    }
    if (own) {
        a.w.nnb[(size_t)r * nmax + i] = k;
        double *ref = a.w.ref + (size_t)r * 3 * nmax;
        ref[i] = xi;
        ref[nmax + i] = yi;
        ref[2 * nmax + i] = zi;
        if (overflow) atomicOr(a.b.status + r, (int)QMC_STATUS_LIST_OVERFLOW);
    }
}

// The same list built by G lanes per atom (G = a.build_lanes, a power of two <= 32): large replicas flag a rebuild one or a few
// at a time, and one thread per atom then walks all N candidates alone (0.2 ms for 2048 atoms, whatever the number of flagged
// replicas).  The G lanes of an atom take G consecutive candidates per round; hits are compacted in lane order with one ballot
// per round (and per further periodic image), so the entries come out in ascending candidate order: a fixed order, hence
// bit-reproducible sums downstream.
// (diagonal cells: the round-1 code, untouched)
__global__ void __launch_bounds__(HMC_THREADS) hmc_build_split_kernel_diag(const __grid_constant__ HmcArgs a) {
    __shared__ double sx[HMC_THREADS], sy[HMC_THREADS], sz[HMC_THREADS];
    for (int r = blockIdx.y; r < a.b.n_replicas; r += gridDim.y) {  // few grid rows: a launch without requests costs little
    if (a.w.flags[r] == 0) continue;
    const int nmax = a.b.n_max, n = a.b.n_atoms[r];
    const double rl = a.pot.cut + a.skin, rl2 = rl * rl;
    CellDev C;
    if (!hmc_cell<false>(a, r, rl * (1.0 + 1e-9), C)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.b.status + r, (int)QMC_STATUS_CELL_TOO_SMALL);
        continue;
    }
    const int G = a.build_lanes, lane = threadIdx.x & 31, sub = lane & (G - 1);
    // single-precision copy of the atom and the cell for the chunk test below (margins cover the rounding)
    const float rlf = (float)rl + 0.02f, rlf2 = rlf * rlf;
    const float *bbox = a.w.bbox + (size_t)r * ((nmax + 31) / 32) * 6;
    const unsigned gmask = (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << (lane - sub), below = gmask & ((1u << lane) - 1u);
    const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax, *y = x + nmax, *z = y + nmax;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool own = i < n;
    const double pi[3] = {own ? x[i] : 0.0, own ? y[i] : 0.0, own ? z[i] : 0.0};
    const float pif[3] = {(float)pi[0], (float)pi[1], (float)pi[2]}, Lf[3] = {(float)C.L[0], (float)C.L[1], (float)C.L[2]};
    const float iLf[3] = {(float)C.invL[0], (float)C.invL[1], (float)C.invL[2]};
    unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    int k = 0, overflow = 0;  // k: entries of the atom so far, the same on its G lanes
    for (int j0 = 0; j0 < n; j0 += HMC_THREADS) {
        __syncthreads();
        const int jl = j0 + threadIdx.x;
        if (jl < n) {
            sx[threadIdx.x] = x[jl];
            sy[threadIdx.x] = y[jl];
            sz[threadIdx.x] = z[jl];
        }
        __syncthreads();
        const int jn = min(HMC_THREADS, n - j0);
        for (int t0 = 0; t0 < jn; t0 += G) {
            const int t = t0 + sub, j = j0 + t;
            // chunk test: the 32 atoms around candidate j lie in a box (centre, half-width); when the nearest image of that box
            // is farther than the list radius no candidate of this round can be listed (second images are farther still)
            const float *bb = bbox + (size_t)((j0 + t0) >> 5) * 6;
            float gap2 = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float d = pif[q] - bb[q];
                if (C.per[q]) d -= Lf[q] * rintf(d * iLf[q]);
                const float gap = fabsf(d) - bb[3 + q];
                gap2 += gap > 0.0f ? gap * gap : 0.0f;
            }
            bool active = own && t < jn && j != i && !(gap2 > rlf2);
            double d0[3] = {0.0, 0.0, 0.0}, d1[3] = {0.0, 0.0, 0.0}, k0[3] = {0.0, 0.0, 0.0};
            unsigned avail = 0, c = 0;
            if (active) {
                const double pj[3] = {sx[t], sy[t], sz[t]};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (C.per[q]) {
                        const double f = (pi[q] - pj[q]) * C.invL[q];
                        k0[q] = (f + RINT_MAGIC) - RINT_MAGIC;
                        d0[q] = fma(k0[q], C.L[q], pj[q]) - pi[q];
                        d1[q] = d0[q] - copysign(C.L[q], d0[q]);
                        if (fabs(d0[q]) > C.thr[q]) avail |= 1u << q;
                    } else {
                        d0[q] = d1[q] = pj[q] - pi[q];
                    }
                }
            }
            while (__any_sync(FULL, active)) {  // nearest image first, then the further images a short cell edge allows
                bool hit = false;
                int sh[3] = {0, 0, 0};
                if (active) {
                    const double dx = (c & 1u) ? d1[0] : d0[0], dy = (c & 2u) ? d1[1] : d0[1], dz = (c & 4u) ? d1[2] : d0[2];
                    hit = dx * dx + dy * dy + dz * dz < rl2;
#pragma unroll
                    for (int q = 0; q < 3; ++q) sh[q] = (int)k0[q] - (((c >> q) & 1u) ? (d0[q] > 0.0 ? 1 : -1) : 0);
                    if (hit && (abs(sh[0]) > 15 || abs(sh[1]) > 15 || abs(sh[2]) > 15)) {
                        overflow = 1;
                        hit = false;
                    }
                }
                const unsigned hits = __ballot_sync(FULL, hit) & gmask;
                if (hit) {
                    const int slot = k + __popc(hits & below);
                    if (slot < a.w.max_nb)
                        list[(size_t)slot * nmax + i] = (unsigned)j | ((unsigned)(sh[0] + 16) << 16) | ((unsigned)(sh[1] + 16) << 21) | ((unsigned)(sh[2] + 16) << 26);
                    else
                        overflow = 1;
                }
                k += __popc(hits);
                if (active) {
                    if (c == avail) active = false;
                    else c = ((c | (~avail & 7u)) + 1u) & avail;
                }
            }
        }
    }
    if (own && sub == 0) {
        a.w.nnb[(size_t)r * nmax + i] = min(k, a.w.max_nb);
        double *ref = a.w.ref + (size_t)r * 3 * nmax;
        ref[i] = pi[0];
        ref[nmax + i] = pi[1];
        ref[2 * nmax + i] = pi[2];
    }
    if (overflow) atomicOr(a.b.status + r, (int)QMC_STATUS_LIST_OVERFLOW);
    }
}

// (general cells: images through hmc_images / hmc_image_vector)
__global__ void __launch_bounds__(HMC_THREADS) hmc_build_split_kernel_tri(const __grid_constant__ HmcArgs a) {
    __shared__ double sx[HMC_THREADS], sy[HMC_THREADS], sz[HMC_THREADS];
    for (int r = blockIdx.y; r < a.b.n_replicas; r += gridDim.y) {  // few grid rows: a launch without requests costs little
    if (a.w.flags[r] == 0) continue;
    const int nmax = a.b.n_max, n = a.b.n_atoms[r];
    const double rl = a.pot.cut + a.skin, rl2 = rl * rl;
    CellDev C;
    if (!hmc_cell<true>(a, r, rl * (1.0 + 1e-9), C)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.b.status + r, (int)QMC_STATUS_CELL_TOO_SMALL);
        continue;
    }
    const int G = a.build_lanes, lane = threadIdx.x & 31, sub = lane & (G - 1);
    // single-precision copy of the atom and the cell for the chunk test below (margins cover the rounding)
    const float rlf = (float)rl + 0.02f, rlf2 = rlf * rlf;
    const float *bbox = a.w.bbox + (size_t)r * ((nmax + 31) / 32) * 6;
    const unsigned gmask = (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << (lane - sub), below = gmask & ((1u << lane) - 1u);
    const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax, *y = x + nmax, *z = y + nmax;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool own = i < n;
    const double pi[3] = {own ? x[i] : 0.0, own ? y[i] : 0.0, own ? z[i] : 0.0};
    const float pif[3] = {(float)pi[0], (float)pi[1], (float)pi[2]}, Lf[3] = {(float)C.L[0], (float)C.L[1], (float)C.L[2]};
    const float iLf[3] = {(float)C.invL[0], (float)C.invL[1], (float)C.invL[2]};
    unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    int k = 0, overflow = 0;  // k: entries of the atom so far, the same on its G lanes
    for (int j0 = 0; j0 < n; j0 += HMC_THREADS) {
        __syncthreads();
        const int jl = j0 + threadIdx.x;
        if (jl < n) {
            sx[threadIdx.x] = x[jl];
            sy[threadIdx.x] = y[jl];
            sz[threadIdx.x] = z[jl];
        }
        __syncthreads();
        const int jn = min(HMC_THREADS, n - j0);
        for (int t0 = 0; t0 < jn; t0 += G) {
            const int t = t0 + sub, j = j0 + t;
            // chunk test: the 32 atoms around candidate j lie in a box (centre, half-width); when the nearest image of that box
            // is farther than the list radius no candidate of this round can be listed (second images are farther still)
            const float *bb = bbox + (size_t)((j0 + t0) >> 5) * 6;
            float gap2 = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float d = pif[q] - bb[q];
                if (C.per[q]) d -= Lf[q] * rintf(d * iLf[q]);
                const float gap = fabsf(d) - bb[3 + q];
                gap2 += gap > 0.0f ? gap * gap : 0.0f;
            }
            if (C.tri) gap2 = 0.0f;  // general cells: the per-axis wrap of the chunk test does not apply, every round is looked at
            bool active = own && t < jn && j != i && !(gap2 > rlf2);
            double d0[3] = {0.0, 0.0, 0.0}, sgn[3] = {1.0, 1.0, 1.0}, k0[3] = {0.0, 0.0, 0.0};
            unsigned avail = 0, c = 0;
            if (active) {
                const double pj[3] = {sx[t], sy[t], sz[t]};
                hmc_images(C, pi, pj, d0, k0, sgn, avail);
            }
            while (__any_sync(FULL, active)) {  // the image found by rounding first, then the further images a short cell allows
                bool hit = false;
                int sh[3] = {0, 0, 0};
                if (active) {
                    double d[3];
                    hmc_image_vector(C, d0, sgn, c, d);
                    hit = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < rl2;
#pragma unroll
                    for (int q = 0; q < 3; ++q) sh[q] = (int)k0[q] - (((c >> q) & 1u) ? (int)sgn[q] : 0);
                    if (hit && (abs(sh[0]) > 15 || abs(sh[1]) > 15 || abs(sh[2]) > 15)) {
                        overflow = 1;
                        hit = false;
                    }
                }
                const unsigned hits = __ballot_sync(FULL, hit) & gmask;
                if (hit) {
                    const int slot = k + __popc(hits & below);
                    if (slot < a.w.max_nb)
                        list[(size_t)slot * nmax + i] = (unsigned)j | ((unsigned)(sh[0] + 16) << 16) | ((unsigned)(sh[1] + 16) << 21) | ((unsigned)(sh[2] + 16) << 26);
                    else
                        overflow = 1;
                }
                k += __popc(hits);
                if (active) {
                    if (c == avail) active = false;
                    else c = ((c | (~avail & 7u)) + 1u) & avail;
                }
            }
        }
    }
    if (own && sub == 0) {
        a.w.nnb[(size_t)r * nmax + i] = min(k, a.w.max_nb);
        double *ref = a.w.ref + (size_t)r * 3 * nmax;
        ref[i] = pi[0];
        ref[nmax + i] = pi[1];
        ref[2 * nmax + i] = pi[2];
    }
    if (overflow) atomicOr(a.b.status + r, (int)QMC_STATUS_LIST_OVERFLOW);
    }
}

// centre and half-width (single precision, rounded outwards) of every 32-atom chunk of the replicas that requested a rebuild
__global__ void __launch_bounds__(HMC_THREADS) hmc_bbox_kernel(const __grid_constant__ HmcArgs a) {
    const int nmax = a.b.n_max, lane = threadIdx.x & 31;
    for (int r = blockIdx.y; r < a.b.n_replicas; r += gridDim.y) {
        if (a.w.flags[r] == 0) continue;
        const int n = a.b.n_atoms[r], j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j - lane >= nmax) continue;
        const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax;
        float *bb = a.w.bbox + ((size_t)r * ((nmax + 31) / 32) + (j >> 5)) * 6;
        for (int q = 0; q < 3; ++q) {
            double lo = j < n ? x[(size_t)q * nmax + j] : 1e300, hi = j < n ? lo : -1e300;
            for (int o = 16; o > 0; o >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
                hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
            }
            if (lane == 0) {
                const bool any = hi >= lo;
                const float c = any ? (float)(0.5 * (lo + hi)) : 0.0f;
                // half-width measured from the rounded centre, plus a margin for the single-precision arithmetic of the test
                bb[q] = c;
                bb[3 + q] = any ? (float)fmax(hi - (double)c, (double)c - lo) * 1.0001f + 1e-3f : 0.0f;
            }
        }
    }
}

// (the rebuild flag of a replica is cleared by pass 1, which runs after every CTA of the build kernels has seen it)
__global__ void hmc_set_ctr_kernel(unsigned long long *ctr, unsigned long long attempt, unsigned long long column) {
    ctr[0] = attempt;
    ctr[1] = column;
}

__global__ void hmc_bump_ctr_kernel(unsigned long long *ctr) {
    ctr[0] += 1ull;
    ctr[1] += 1ull;
}

// List entries are walked through a two-deep software pipeline: the entry word two iterations ahead and the neighbour's
// coordinates one iteration ahead are in flight while the pair functions of the current entry are evaluated (ncu before: 6.8 of
// 13 stall cycles per issue were long-scoreboard waits on the dependent pair list word -> gathered coordinates).
struct NbPipe {
    unsigned e0, e1;    // current entry, next entry
    double gx, gy, gz;  // coordinates of the current entry's atom
};

__device__ __forceinline__ void nb_start(NbPipe &P, const unsigned *lp, int nmax, int k, int lpa, int nn, const double *x, const double *y,
                                         const double *z) {
    P.e0 = k < nn ? lp[(size_t)k * nmax] : 0u;
    P.e1 = k + lpa < nn ? lp[(size_t)(k + lpa) * nmax] : 0u;
    const int j = (int)(P.e0 & 0xFFFFu);
    P.gx = x[j];
    P.gy = y[j];
    P.gz = z[j];
}

// image vector of the current entry for the atom at (xi, yi, zi); its atom index in j
template <bool TRI>
__device__ __forceinline__ void nb_vector(const NbPipe &P, const CellDev &C, double xi, double yi, double zi, int &j, double &dx, double &dy,
                                          double &dz) {
    const unsigned e = P.e0;
    j = (int)(e & 0xFFFFu);
    const int sx = (int)((e >> 16) & 31u) - 16, sy = (int)((e >> 21) & 31u) - 16, sz = (int)((e >> 26) & 31u) - 16;
    if (TRI) {  // general cell: the integer shift counts cell VECTORS
        dx = fma((double)sz, C.H[6], fma((double)sy, C.H[3], fma((double)sx, C.H[0], P.gx))) - xi;
        dy = fma((double)sz, C.H[7], fma((double)sy, C.H[4], fma((double)sx, C.H[1], P.gy))) - yi;
        dz = fma((double)sz, C.H[8], fma((double)sy, C.H[5], fma((double)sx, C.H[2], P.gz))) - zi;
        return;
    }
    dx = fma((double)sx, C.L[0], P.gx) - xi;
    dy = fma((double)sy, C.L[1], P.gy) - yi;
    dz = fma((double)sz, C.L[2], P.gz) - zi;
}

// issue the loads of the following iterations (entry word k + 2 lpa, coordinates of entry k + lpa)
__device__ __forceinline__ void nb_advance(NbPipe &P, const unsigned *lp, int nmax, int k, int lpa, int nn, const double *x, const double *y,
                                           const double *z) {
    const unsigned e2 = k + 2 * lpa < nn ? lp[(size_t)(k + 2 * lpa) * nmax] : 0u;
    P.e0 = P.e1;
    P.e1 = e2;
    const int j = (int)(P.e0 & 0xFFFFu);  // entry 0 (atom 0) past the end: a harmless load
    P.gx = x[j];
    P.gy = y[j];
    P.gz = z[j];
}

// ---- pass 1: sigma_1, energies, dE/dsigma_1 ----------------------------------------------------------------------
template <int POT, bool TRI>
__global__ void __launch_bounds__(HMC_THREADS) hmc_pass1_kernel(const __grid_constant__ HmcArgs a) {
    __shared__ double red[32];
    const int r = blockIdx.y, nmax = a.b.n_max, n = a.b.n_atoms[r];
    const int lpa = a.lpa, sub = threadIdx.x & (lpa - 1);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / lpa;
    const bool own = i < n;
    CellDev C;
    hmc_cell<TRI>(a, r, a.pot.cut_pre, C);
    const double *x = a.w.x[a.cur] + (size_t)r * 3 * nmax, *y = x + nmax, *z = y + nmax;
    const unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    double s1 = 0.0, s2 = 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.w.flags[r] = 0;  // the list build (previous kernels) has consumed the request
    if (own) {
        const double xi = x[i], yi = y[i], zi = z[i];
        const int nn = a.w.nnb[(size_t)r * nmax + i];
        NbPipe P;
        nb_start(P, list + i, nmax, sub, lpa, nn, x, y, z);
        for (int k = sub; k < nn; k += lpa) {
            int j;
            double dx, dy, dz;
            nb_vector<TRI>(P, C, xi, yi, zi, j, dx, dy, dz);
            nb_advance(P, list + i, nmax, k, lpa, nn, x, y, z);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (r2 < a.pot.cut_pre2) {
                if (POT == QMC_POT_EMT) {
                    double t1, t2;
                    emt_pair(a.pot, r2, t1, t2);
                    s1 += t1;
                    s2 += t2;
                } else {
                    s2 += lj_pair(a.pot, r2);
                }
            }
        }
    }
    for (int o = lpa >> 1; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
    }
    double e = 0.0;
    if (own && sub == 0) {
        if (POT == QMC_POT_EMT) {
            double de;
            const double ea = emt_embed_deds(a.pot, s1, de);
            a.w.deds[(size_t)r * nmax + i] = de;
            e = ea + a.pot.neghalfv0overgamma2 * s2;
        } else {
            e = 0.5 * s2;
        }
    }
    const double s = block_sum(e, red);
    if (threadIdx.x == 0) a.w.partial[((size_t)r * a.w.nb_blocks + blockIdx.x) * 2] = s;
}

// ---- pass 2: forces + velocity-Verlet update --------------------------------------------------------------------
// MINB = 8 bounds the kernel to 64 registers: 1024 resident threads per SM instead of 512, which lets a small batch (8 x 2048
// atoms x 4 lanes) run as ONE wave instead of one and a tail (profiles/r2_aj_hmc_small_batch.md); MINB = 5 (96 registers, 640
// threads) is the general choice for EMT in diagonal cells
template <int POT, bool TRI, int MINB = 1>
__global__ void __launch_bounds__(HMC_THREADS, MINB) hmc_pass2_kernel(const __grid_constant__ HmcArgs a) {
    __shared__ double red[32];
    const int r = blockIdx.y, nmax = a.b.n_max, n = a.b.n_atoms[r];
    const int lpa = a.lpa, sub = threadIdx.x & (lpa - 1);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / lpa;
    const bool own = i < n;
    CellDev C;
    hmc_cell<TRI>(a, r, a.pot.cut_pre, C);
    const size_t base = (size_t)r * 3 * nmax;
    const double *x = a.w.x[a.cur] + base, *y = x + nmax, *z = y + nmax;
    const unsigned *list = a.w.list + (size_t)r * a.w.max_nb * nmax;
    const double *deds = a.w.deds + (size_t)r * nmax;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0, xi = 0.0, yi = 0.0, zi = 0.0;
    if (own) {
        xi = x[i]; yi = y[i]; zi = z[i];
        const double da = POT == QMC_POT_EMT ? deds[i] : 0.0;
        const int nn = a.w.nnb[(size_t)r * nmax + i];
        NbPipe P;
        nb_start(P, list + i, nmax, sub, lpa, nn, x, y, z);
        double dj_next = POT == QMC_POT_EMT ? deds[P.e0 & 0xFFFFu] : 0.0;
        for (int k = sub; k < nn; k += lpa) {
            int j;
            double dx, dy, dz;
            nb_vector<TRI>(P, C, xi, yi, zi, j, dx, dy, dz);
            const double dj = dj_next;
            nb_advance(P, list + i, nmax, k, lpa, nn, x, y, z);
            if (POT == QMC_POT_EMT) dj_next = deds[P.e0 & 0xFFFFu];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (r2 < a.pot.cut_pre2) {
                double g = 0.0;
                if (POT == QMC_POT_EMT) {
                    const double rr = sqrt_fast(fmax(r2, 1e-200));
                    if (rr < a.pot.rc_list) {
                        const double w = rcp_fast(1.0 + exp_fast(fma(a.pot.acut, rr, a.pot.theta_c)));
                        const double dwdroverw = a.pot.acut * (w - 1.0);
                        const double ds1 = exp_fast(fma(a.pot.sig1_a, rr, a.pot.sig1_c)) * w;
                        const double ds2 = exp_fast(fma(a.pot.sig2_a, rr, a.pot.sig2_c)) * w;
                        const double es = a.pot.neghalfv0overgamma2 * ds2;
                        const double dedr_pair = es * (dwdroverw - a.pot.kappa_over_beta);
                        const double dds1dr = ds1 * (dwdroverw - a.pot.eta2);
                        g = ((dedr_pair + dedr_pair) + (da * dds1dr + dj * dds1dr)) * rcp_fast(rr);
                    }
                } else if (!(r2 > a.pot.lj_rc2)) {
                    const double q = a.pot.sigma2 * rcp_fast(r2);
                    const double c6 = q * q * q;
                    g = -6.0 * a.pot.eps4 * (2.0 * c6 * c6 - c6) * rcp_fast(r2);  // -24 eps (2 c12 - c6) / r2
                }
                f0 = fma(g, dx, f0);
                f1 = fma(g, dy, f1);
                f2 = fma(g, dz, f2);
            }
        }
    }
    for (int o = lpa >> 1; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(FULL, f0, o);
        f1 += __shfl_xor_sync(FULL, f1, o);
        f2 += __shfl_xor_sync(FULL, f2, o);
    }
    double ke = 0.0;
    if (own && sub == 0) {
        const double f[3] = {f0, f1, f2}, xo[3] = {xi, yi, zi};
        if (a.mode == 3) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.forces_out[base + (size_t)c * nmax + i] = f[c];
        } else {
            const double mass = a.b.mass, dt = a.dt;
            double *xn = a.w.x[a.cur ^ 1] + base;
            const double *ref = a.w.ref + base;
            double dsq = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double p = a.w.p[base + (size_t)c * nmax + i];
                const double half_kick = __dmul_rn(__dmul_rn(0.5, f[c]), dt);
                if (a.mode == 1 || a.mode == 2) p = p + half_kick;  // complete the previous step: p = p_half + 0.5 F dt
                if (a.mode == 0 || a.mode == 1) {                   // start the next one
                    const double nm = p + half_kick;
                    const double xnew = xo[c] + __dmul_rn(nm / mass, dt);
                    xn[(size_t)c * nmax + i] = xnew;
                    p = __dmul_rn(xnew - xo[c], mass) / dt;  // apply_constraints=True re-derivation (:74-75)
                    const double dd = xnew - ref[(size_t)c * nmax + i];
                    dsq = fma(dd, dd, dsq);
                } else {
                    ke += p * (p / mass);
                }
                a.w.p[base + (size_t)c * nmax + i] = p;
            }
            if (dsq > 0.25 * a.skin * a.skin) a.w.flags[r] = 1;
        }
    }
    if (a.mode == 2 || a.mode == 4) {
        const double s = block_sum(ke, red);
        if (threadIdx.x == 0) a.w.partial[((size_t)r * a.w.nb_blocks + blockIdx.x) * 2 + 1] = s;
    }
}

// ---- finish: fixed-order reductions, acceptance, commit ------------------------------------------------------------
__global__ void __launch_bounds__(HMC_THREADS) hmc_finish_kernel(const __grid_constant__ HmcArgs a, int nb1, int nb2) {
    const int r = blockIdx.y, nmax = a.b.n_max, n = a.b.n_atoms[r];
    // every CTA of the replica recomputes the same decision from the same partials in the same order
    double e_pot = 0.0, ke = 0.0, ke0 = 0.0;
    for (int b = 0; b < nb2; ++b) {
        e_pot += a.w.partial[((size_t)r * a.w.nb_blocks + b) * 2];
        ke += a.w.partial[((size_t)r * a.w.nb_blocks + b) * 2 + 1];
    }
    for (int b = 0; b < nb1; ++b) ke0 += a.w.ke0[(size_t)r * a.w.nb_blocks + b];
    const double e_kin = 0.5 * ke, last_kin = 0.5 * ke0;
    const double last_pot = a.b.energy[r];
    const double delta = (e_pot + e_kin) - last_pot - last_kin;
    const double kT = a.b.temperature[r] * KB;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    double o2, u_accept;
    uniform_pair(key, a.w.ctr[0], grep, 2, o2, u_accept);
    int status = 0;
    const bool acc = metropolis(u_accept, -delta / kT, status);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = (size_t)r * 3 * nmax;
    __syncthreads();
    if (acc && i < n) {
        const double *xf = a.w.x[a.cur] + base;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.b.pos[base + (size_t)c * nmax + i] = xf[(size_t)c * nmax + i];
            a.b.momenta[base + (size_t)c * nmax + i] = a.w.p[base + (size_t)c * nmax + i];
        }
    }
    // energy[r] / kinetic[r] are still being read by the other CTAs of this replica: results go to a mailbox, published by
    // hmc_publish_kernel afterwards
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double *out = a.w.result + (size_t)r * 2;
        out[0] = acc ? e_pot : last_pot;
        out[1] = acc ? e_kin : last_kin;
        unsigned long long *cnt = reinterpret_cast<unsigned long long *>(a.b.counters) + (size_t)r * QMC_MAX_MOVES * 3;
        cnt[0] += 1;
        if (acc) cnt[1] += 1;
        if (status) atomicOr(a.b.status + r, status);
        const size_t ti = (size_t)r * a.n_attempts + (size_t)a.w.ctr[1];
        if (a.tr.move) a.tr.move[ti] = 0;
        if (a.tr.accepted) a.tr.accepted[ti] = acc ? 1 : 0;
        if (a.tr.energy) a.tr.energy[ti] = acc ? e_pot : last_pot;
        if (a.tr.delta) a.tr.delta[ti] = delta;
        if (a.tr.moved) a.tr.moved[ti] = -1;
    }
}

__global__ void hmc_publish_kernel(const __grid_constant__ HmcArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.b.n_replicas) return;
    const double *out = a.w.result + (size_t)r * 2;
    a.b.energy[r] = out[0];
    a.b.kinetic[r] = out[1];
}

__global__ void hmc_energy_publish_kernel(const __grid_constant__ HmcArgs a, int nb2) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.b.n_replicas) return;
    double e = 0.0;
    for (int b = 0; b < nb2; ++b) e += a.w.partial[((size_t)r * a.w.nb_blocks + b) * 2];
    a.b.energy[r] = e;
}

// ---- host side ---------------------------------------------------------------------------------------------------
inline int hmc_pick_lpa(const qmc_batch_t &b) {
    if (const char *e = std::getenv("QMC_HMC_LPA")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 32 && (v & (v - 1)) == 0) return v;
    }
    const long long atoms = (long long)b.n_replicas * b.n_max;
    int lpa = 1;
    while (lpa < 32 && atoms * lpa < 148LL * 1024) lpa <<= 1;
    // small batches: with half as many lanes the whole batch is resident at once (1024 threads per SM at 64 registers) -- one wave
    // beats 1.7 waves of pass 1 and 2.3 of pass 2
    if (lpa >= 8 && atoms * (lpa / 2) <= 148LL * 1024 && !(std::getenv("QMC_HMC_ONE_WAVE") && std::atoi(std::getenv("QMC_HMC_ONE_WAVE")) == 0)) lpa /= 2;
    return lpa;
}

inline int hmc_one_wave(const qmc_batch_t &b, int lpa) {
    if (const char *e = std::getenv("QMC_HMC_ONE_WAVE")) return std::atoi(e) != 0;
    return lpa >= 4 && (long long)b.n_replicas * b.n_max * lpa <= 148LL * 1024;
}

// Lanes per atom of the list build: replicas of a few hundred atoms come in large batches (one thread per atom fills the GPU);
// large replicas rebuild one at a time, so the walk over the N candidates is split.
inline int hmc_pick_build_lanes(const qmc_batch_t &b) {
    if (const char *e = std::getenv("QMC_HMC_BUILD_LANES")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 32 && (v & (v - 1)) == 0) return v;
    }
    int g = 1;
    while (g < 32 && b.n_max >= 128 * g) g <<= 1;  // 128 atoms: 2, 256: 4, ... 2048 and more: 32
    return g;
}

inline int hmc_check(cudaError_t e, const char *what, char *err, size_t errn) {
    if (e == cudaSuccess) return QMC_OK;
    snprintf(err, errn, "%s failed: %s", what, cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA;
}

// one force evaluation at the positions in buffer a.cur: (build) + pass 1 + pass 2
inline void hmc_force_eval(HmcArgs &a, cudaStream_t s) {
    const int N = a.b.n_max, R = a.b.n_replicas;
    const dim3 g1((N + HMC_THREADS - 1) / HMC_THREADS, R), gl((N * a.lpa + HMC_THREADS - 1) / HMC_THREADS, R);
    if (a.build_lanes > 1) {
        const int rows = R < 8 ? R : 8;
        const dim3 gb((N * a.build_lanes + HMC_THREADS - 1) / HMC_THREADS, rows), gq((N + HMC_THREADS - 1) / HMC_THREADS, rows);
        hmc_bbox_kernel<<<gq, HMC_THREADS, 0, s>>>(a);
        if (a.b.triclinic) hmc_build_split_kernel_tri<<<gb, HMC_THREADS, 0, s>>>(a);
        else hmc_build_split_kernel_diag<<<gb, HMC_THREADS, 0, s>>>(a);
    } else {
        if (a.b.triclinic) hmc_build_kernel_tri<<<g1, HMC_THREADS, 0, s>>>(a);
        else hmc_build_kernel_diag<<<g1, HMC_THREADS, 0, s>>>(a);
    }
    if (a.b.triclinic) {  // general cells: the list shifts count cell vectors
        if (a.pot.kind == QMC_POT_EMT) {
            hmc_pass1_kernel<QMC_POT_EMT, true><<<gl, HMC_THREADS, 0, s>>>(a);
            hmc_pass2_kernel<QMC_POT_EMT, true><<<gl, HMC_THREADS, 0, s>>>(a);
        } else {
            hmc_pass1_kernel<QMC_POT_LJ, true><<<gl, HMC_THREADS, 0, s>>>(a);
            hmc_pass2_kernel<QMC_POT_LJ, true><<<gl, HMC_THREADS, 0, s>>>(a);
        }
    } else if (a.pot.kind == QMC_POT_EMT) {
        hmc_pass1_kernel<QMC_POT_EMT, false><<<gl, HMC_THREADS, 0, s>>>(a);
        // five CTAs per SM (96 registers instead of the unbounded 106, no spills): +1.6 % on C5 and its tempering loop, +1 % on
        // ForceBias against 4 / 6 / 7 / 8 CTAs in one call on one box (profiles/r2_av_hmc_pass_ncu.md)
        if (a.one_wave) hmc_pass2_kernel<QMC_POT_EMT, false, 8><<<gl, HMC_THREADS, 0, s>>>(a);
        else hmc_pass2_kernel<QMC_POT_EMT, false, 5><<<gl, HMC_THREADS, 0, s>>>(a);
    } else {
        hmc_pass1_kernel<QMC_POT_LJ, false><<<gl, HMC_THREADS, 0, s>>>(a);
        hmc_pass2_kernel<QMC_POT_LJ, false><<<gl, HMC_THREADS, 0, s>>>(a);
    }
}

inline int64_t hmc_workspace_bytes(const PotDev &pot, const qmc_batch_t &b) { return hmc_carve(pot, b, nullptr, nullptr); }

inline int hmc_forces(const PotDev &pot, const qmc_batch_t &b, double *forces, void *workspace, cudaStream_t s, char *err, size_t errn) {
    HmcArgs a;
    memset(&a, 0, sizeof(a));
    a.pot = pot;
    a.b = b;
    hmc_carve(pot, b, workspace, &a.w);
    a.forces_out = forces;
    a.lpa = hmc_pick_lpa(b);
    a.one_wave = hmc_one_wave(b, a.lpa);
    a.skin = hmc_skin();
    a.build_lanes = hmc_pick_build_lanes(b);
    a.cur = 0;
    a.mode = 3;
    const int N = b.n_max, R = b.n_replicas;
    int rc = hmc_check(cudaMemcpyAsync(a.w.x[0], b.pos, (size_t)R * 3 * N * 8, cudaMemcpyDeviceToDevice, s), "position copy", err, errn);
    if (rc) return rc;
    if ((rc = hmc_check(cudaMemsetAsync(a.w.flags, 1, (size_t)R * 4, s), "flag reset", err, errn))) return rc;
    hmc_force_eval(a, s);
    const int nb2 = (N * a.lpa + HMC_THREADS - 1) / HMC_THREADS;
    hmc_energy_publish_kernel<<<(R + 127) / 128, 128, 0, s>>>(a, nb2);
    return hmc_check(cudaGetLastError(), "force kernels", err, errn);
}

// One trajectory = init + (n_steps + 1) force evaluations (4 launches each) + finish + publish + counter bump: ~410 launches for
// Verlet 100.  The sequence never depends on the data (list rebuilds are decided on the device), so it is captured ONCE into a
// CUDA graph -- on a private capture stream, the caller's stream may be the legacy default stream -- and launched once per
// attempt; the attempt index lives in device memory (w.ctr) and is advanced by the graph itself.  Executables are cached by the
// exact kernel arguments (the parallel-tempering loop calls with one trajectory per round).  QMC_HMC_GRAPH=0 launches directly.
inline void hmc_trajectory_launches(HmcArgs a, int n_steps, int nb1, int nb2, cudaStream_t s) {
    const int R = a.b.n_replicas;
    const dim3 g1(nb1, R);
    a.cur = 0;
    hmc_init_kernel<<<g1, HMC_THREADS, 0, s>>>(a);
    for (int step = 0; step <= n_steps; ++step) {
        a.mode = n_steps == 0 ? 4 : (step == 0 ? 0 : (step == n_steps ? 2 : 1));
        hmc_force_eval(a, s);
        if (a.mode == 0 || a.mode == 1) a.cur ^= 1;
    }
    hmc_finish_kernel<<<g1, HMC_THREADS, 0, s>>>(a, nb1, nb2);
    hmc_publish_kernel<<<(R + 127) / 128, 128, 0, s>>>(a);
    hmc_bump_ctr_kernel<<<1, 1, 0, s>>>(a.w.ctr);
}

struct HmcGraphCache {
    struct Entry {
        std::string key;
        cudaGraphExec_t exec;
        int device;
    };
    std::mutex mu;
    std::vector<Entry> entries;  // most recently used last
    cudaStream_t capture_stream[16] = {};
};

inline HmcGraphCache &hmc_graph_cache() {
    static HmcGraphCache c;
    return c;
}

inline cudaGraphExec_t hmc_trajectory_graph(const HmcArgs &a, int n_steps, int nb1, int nb2) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    std::string key(reinterpret_cast<const char *>(&a), sizeof(a));
    key.append(reinterpret_cast<const char *>(&n_steps), sizeof(n_steps));
    key.append(reinterpret_cast<const char *>(&dev), sizeof(dev));
    HmcGraphCache &c = hmc_graph_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    for (size_t k = 0; k < c.entries.size(); ++k)
        if (c.entries[k].key == key) {
            HmcGraphCache::Entry e = c.entries[k];
            c.entries.erase(c.entries.begin() + k);
            c.entries.push_back(e);
            return e.exec;
        }
    if (!c.capture_stream[dev] && cudaStreamCreateWithFlags(&c.capture_stream[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    cudaStream_t cs = c.capture_stream[dev];
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    hmc_trajectory_launches(a, n_steps, nb1, nb2, cs);
    if (cudaStreamEndCapture(cs, &graph) != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        return nullptr;
    }
    cudaGraphDestroy(graph);
    if (c.entries.size() >= 8) {
        cudaGraphExecDestroy(c.entries.front().exec);
        c.entries.erase(c.entries.begin());
    }
    c.entries.push_back({key, exec, dev});
    return exec;
}

inline int hmc_run(const PotDev &pot, const qmc_batch_t &b, double dt, int n_steps, unsigned long long seed,
                   unsigned long long attempt_base, int n_attempts, const qmc_trace_t &tr, void *workspace, cudaStream_t s, char *err,
                   size_t errn) {
    HmcArgs a;
    memset(&a, 0, sizeof(a));
    a.pot = pot;
    a.b = b;
    hmc_carve(pot, b, workspace, &a.w);
    a.dt = dt;
    a.n_steps = n_steps;
    a.seed = seed;
    a.tr = tr;
    a.n_attempts = n_attempts;
    a.lpa = hmc_pick_lpa(b);
    a.one_wave = hmc_one_wave(b, a.lpa);
    a.skin = hmc_skin();
    a.build_lanes = hmc_pick_build_lanes(b);
    const int N = b.n_max;
    const int nb1 = (N + HMC_THREADS - 1) / HMC_THREADS, nb2 = (N * a.lpa + HMC_THREADS - 1) / HMC_THREADS;
    hmc_set_ctr_kernel<<<1, 1, 0, s>>>(a.w.ctr, attempt_base, 0ull);
    const char *env = std::getenv("QMC_HMC_GRAPH");
    cudaGraphExec_t exec = (env && std::atoi(env) == 0) ? nullptr : hmc_trajectory_graph(a, n_steps, nb1, nb2);
    for (int k = 0; k < n_attempts; ++k) {
        if (exec) {
            int rc = hmc_check(cudaGraphLaunch(exec, s), "trajectory graph launch", err, errn);
            if (rc) return rc;
        } else {
            hmc_trajectory_launches(a, n_steps, nb1, nb2, s);
        }
        int rc = hmc_check(cudaGetLastError(), "trajectory kernels", err, errn);
        if (rc) return rc;
    }
    return QMC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Parallel tempering: temperature-adjacent pairs (even / odd alternating) exchange TEMPERATURES.
// Every rank runs this on the same all-gathered arrays and the same Philox slot => identical decisions, no
// configuration ever crosses NVLink.  accept iff u < exp((1/kT_a - 1/kT_b) (E_a - E_b)).
// ---------------------------------------------------------------------------------------------------
__global__ void pt_swap_kernel(double *temps, const double *energies, int n, unsigned long long seed, unsigned long long round,
                               signed char *accepted) {
    extern __shared__ int by_rank[];  // replica holding the k-th lowest temperature
    const int t = threadIdx.x;
    if (t < n) {
        const double mine = temps[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double o = temps[j];
            rank += (o < mine) || (o == mine && j < t);
        }
        by_rank[rank] = t;
        if (accepted) accepted[t] = 0;
    }
    __syncthreads();
    const int k = 2 * t + (int)(round & 1ull);
    if (k + 1 < n) {  // pairs touch disjoint replicas: no ordering between threads is needed
        const int ra = by_rank[k], rb = by_rank[k + 1];
        const double ta = temps[ra], tb = temps[rb];
        const double x = (1.0 / (ta * KB) - 1.0 / (tb * KB)) * (energies[ra] - energies[rb]);
        double d0, u;
        const PhiloxKey key = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uniform_pair(key, round, (uint32_t)k, 3, d0, u);
        const bool acc = x >= 0.0 || u < exp(x);
        if (acc) {
            temps[ra] = tb;
            temps[rb] = ta;
            if (accepted) accepted[ra] = 1;
        }
    }
}

inline int pt_swap(double *temps, const double *energies, int n, unsigned long long seed, unsigned long long round,
                   signed char *accepted, cudaStream_t s, char *err, size_t errn) {
    if (n > 1024) {
        snprintf(err, errn, "parallel tempering over more than 1024 replicas is not supported");
        return QMC_ERR_UNSUPPORTED;
    }
    const int threads = (n + 31) / 32 * 32;
    pt_swap_kernel<<<1, threads, n * sizeof(int), s>>>(temps, energies, n, seed, round, accepted);
    return hmc_check(cudaGetLastError(), "pt_swap launch", err, errn);
}

}  // namespace qmc
