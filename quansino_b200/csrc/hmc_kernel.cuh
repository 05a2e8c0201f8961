// Hamiltonian Monte Carlo: fused force + velocity-Verlet trajectory + acceptance, and parallel-tempering swaps.
//
// One attempt (reference file:line):
//   momenta     p = z sqrt(m kB T), z ~ N(0,1) in row-major (N,3) order            [utils/dynamics.py:27-43]
//   KE_0        0.5 sum p.p/m -> last_kinetic_energy                                [moves/displacement.py:334]
//   trajectory  F = forces; max_steps x { p += F dt/2; x += p/m dt; p = (x_new - x) m / dt (apply_constraints);
//               F = forces(x); p += F dt/2 }                                        [integrators/displacement.py:52-81]
//   criteria    u < exp(-((E_pot + E_kin) - last_pot - last_kin) / kT)              [mc/criteria.py:117-140]
//   commit      positions / momenta / energies kept on accept, positions reloaded on reject
//                                                                                   [mc/contexts.py:144-156,186-198]
//
// Layout: one CTA per replica; positions, momenta, forces, sigma_1 and dE/dsigma_1 of the replica stay in shared
// memory for the whole launch (88 B per atom: 2048 atoms = 176 KB of the 227 KB a B200 CTA can own).  Forces
// are the analytic gradient of exactly the energy expression of the sweep kernel, two passes like ASE
// (sigma_1 first, then the dE/dsigma_1-weighted pair terms: emt.py _calc_e_c_a2 / _calc_fs_c_a2 / _calc_efs_a1).
#pragma once

#include "replica_warp.cuh"

namespace qmc {

constexpr int HMC_THREADS = 1024;

struct HmcArgs {
    PotDev pot;
    qmc_batch_t b;
    double dt;
    int n_steps;
    unsigned long long seed, attempt_base;
    int n_attempts;
    qmc_trace_t tr;
    double *forces_out;  // qmc_forces_full only
    int ns;
};

struct HmcSmem {
    double *x, *y, *z, *px, *py, *pz, *fx, *fy, *fz, *sig, *deds;
    double *red;  // [32] block reduction scratch
    double *bc;   // [8] broadcast scratch
};

__host__ __device__ inline int64_t hmc_smem_bytes(int n_max) {
    const int ns = round_up(n_max, 4);
    return (int64_t)ns * 8 * 11 + 48 * 8;
}

__device__ inline void hmc_carve(HmcSmem &S, unsigned char *base, int ns) {
    double *d = reinterpret_cast<double *>(base);
    S.x = d; d += ns; S.y = d; d += ns; S.z = d; d += ns;
    S.px = d; d += ns; S.py = d; d += ns; S.pz = d; d += ns;
    S.fx = d; d += ns; S.fy = d; d += ns; S.fz = d; d += ns;
    S.sig = d; d += ns; S.deds = d; d += ns;
    S.red = d; d += 32;
    S.bc = d;
}

// deterministic block sum (fixed tree); result on every thread
__device__ inline double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x & 31) < nw ? red[threadIdx.x & 31] : 0.0;
    t = warp_sum(t);
    return t;
}

struct CellDev {
    double L[3], invL[3], thr[3];
    bool per[3];
};

// Energy and forces of the replica in shared memory.  Returns the potential energy on every thread.
template <int POT>
__device__ inline double hmc_energy_forces(const PotDev &pot, const HmcSmem &S, const CellDev &C, int n) {
    double epart = 0.0;
    // pass 1: sigma_1, pair energy, embedding energy and its derivative
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const double xa = S.x[a], ya = S.y[a], za = S.z[a];
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == a) continue;
            double d0[3], d1[3];
            unsigned avail = 0;
            axis_images(xa, S.x[j], C.L[0], C.invL[0], C.thr[0], C.per[0], d0[0], d1[0], avail, 1u);
            axis_images(ya, S.y[j], C.L[1], C.invL[1], C.thr[1], C.per[1], d0[1], d1[1], avail, 2u);
            axis_images(za, S.z[j], C.L[2], C.invL[2], C.thr[2], C.per[2], d0[2], d1[2], avail, 4u);
            unsigned c = 0;
            while (true) {
                const double dx = (c & 1u) ? d1[0] : d0[0], dy = (c & 2u) ? d1[1] : d0[1], dz = (c & 4u) ? d1[2] : d0[2];
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < pot.cut_pre2) {
                    if (POT == QMC_POT_EMT) {
                        double t1, t2;
                        emt_pair(pot, r2, t1, t2);
                        s1 += t1;
                        s2 += t2;
                    } else {
                        s2 += lj_pair(pot, r2);
                    }
                }
                if (c == avail) break;
                c = ((c | (~avail & 7u)) + 1u) & avail;
            }
        }
        if (POT == QMC_POT_EMT) {
            double de;
            const double ea = emt_embed_deds(pot, s1, de);
            S.sig[a] = s1;
            S.deds[a] = de;
            epart += ea + pot.neghalfv0overgamma2 * s2;
        } else {
            epart += 0.5 * s2;
        }
    }
    const double energy = block_sum(epart, S.red);
    __syncthreads();
    // pass 2: forces
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const double xa = S.x[a], ya = S.y[a], za = S.z[a];
        const double da = POT == QMC_POT_EMT ? S.deds[a] : 0.0;
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == a) continue;
            double d0[3], d1[3];
            unsigned avail = 0;
            axis_images(xa, S.x[j], C.L[0], C.invL[0], C.thr[0], C.per[0], d0[0], d1[0], avail, 1u);
            axis_images(ya, S.y[j], C.L[1], C.invL[1], C.thr[1], C.per[1], d0[1], d1[1], avail, 2u);
            axis_images(za, S.z[j], C.L[2], C.invL[2], C.thr[2], C.per[2], d0[2], d1[2], avail, 4u);
            unsigned c = 0;
            while (true) {
                const double dx = (c & 1u) ? d1[0] : d0[0], dy = (c & 2u) ? d1[1] : d0[1], dz = (c & 4u) ? d1[2] : d0[2];
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < pot.cut_pre2) {
                    double g = 0.0;
                    if (POT == QMC_POT_EMT) {
                        const double r = sqrt(r2);
                        if (r < pot.rc_list) {
                            const double w = 1.0 / (1.0 + exp(pot.acut * (r - pot.rc)));
                            const double dwdroverw = pot.acut * (w - 1.0);
                            const double ds1 = exp(-pot.eta2 * (r - pot.beta_s0)) * w;
                            const double ds2 = exp(-pot.kappa * (r * pot.inv_beta - pot.s0)) * w;
                            const double es = pot.neghalfv0overgamma2 * ds2;
                            const double dedr_pair = es * (dwdroverw - pot.kappa_over_beta);
                            const double dds1dr = ds1 * (dwdroverw - pot.eta2);
                            g = ((dedr_pair + dedr_pair) + (da * dds1dr + S.deds[j] * dds1dr)) / r;
                        }
                    } else {
                        if (!(r2 > pot.lj_rc2)) {
                            const double q = pot.sigma2 / r2;
                            const double c6 = q * q * q;
                            g = -6.0 * pot.eps4 * (2.0 * c6 * c6 - c6) / r2;  // -24 eps (2 c12 - c6) / r2
                        }
                    }
                    f0 += g * dx;
                    f1 += g * dy;
                    f2 += g * dz;
                }
                if (c == avail) break;
                c = ((c | (~avail & 7u)) + 1u) & avail;
            }
        }
        S.fx[a] = f0;
        S.fy[a] = f1;
        S.fz[a] = f2;
    }
    __syncthreads();
    return energy;
}

__device__ inline double hmc_normal(PhiloxKey key, unsigned long long attempt, uint32_t replica, uint32_t m) {
    double d0, d1;
    uniform_pair(key, attempt, replica, 16u + m / 2u, d0, d1);
    const double rr = sqrt(-2.0 * log(1.0 - d0));
    const double t = TWO_PI * d1;
    return (m % 2u == 0u) ? rr * cos(t) : rr * sin(t);
}

template <int POT, bool FORCES_ONLY>
__global__ void __launch_bounds__(HMC_THREADS, 1) hmc_kernel(const __grid_constant__ HmcArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int r = blockIdx.x;
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max, n = a.b.n_atoms[r];
    HmcSmem S;
    hmc_carve(S, smem, a.ns);
    double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        S.x[j] = gx[j];
        S.y[j] = gy[j];
        S.z[j] = gz[j];
    }
    CellDev C;
    bool cell_ok = true;
    {
        const double cutp = pot.cut * (1.0 + 1e-9);
        for (int k = 0; k < 3; ++k) {
            const double L = a.b.cell[(size_t)r * 9 + 4 * k];
            C.L[k] = L;
            C.per[k] = a.b.pbc[k] != 0;
            C.invL[k] = C.per[k] ? 1.0 / L : 0.0;
            C.thr[k] = C.per[k] ? L - cutp : 1e300;
            if (C.per[k] && !(L >= cutp)) cell_ok = false;
        }
    }
    __syncthreads();
    if (!cell_ok) {
        if (threadIdx.x == 0) atomicOr(a.b.status + r, (int)QMC_STATUS_CELL_TOO_SMALL);
        return;
    }
    if (FORCES_ONLY) {
        const double e = hmc_energy_forces<POT>(pot, S, C, n);
        double *fo = a.forces_out + (size_t)r * 3 * nmax;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            fo[j] = S.fx[j];
            fo[nmax + j] = S.fy[j];
            fo[2 * nmax + j] = S.fz[j];
        }
        if (threadIdx.x == 0) a.b.energy[r] = e;
        return;
    }

    double *gpx = a.b.momenta + (size_t)r * 3 * nmax, *gpy = gpx + nmax, *gpz = gpy + nmax;
    const double mass = a.b.mass, dt = a.dt;
    const double temperature = a.b.temperature[r];
    const double kT = temperature * KB;
    const double pscale = sqrt(mass * kT);
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    double last_pot = a.b.energy[r];
    int status = 0;

    for (int k = 0; k < a.n_attempts; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        double ke0p = 0.0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double p0 = hmc_normal(key, attempt, grep, 3u * j) * pscale;
            const double p1 = hmc_normal(key, attempt, grep, 3u * j + 1u) * pscale;
            const double p2 = hmc_normal(key, attempt, grep, 3u * j + 2u) * pscale;
            S.px[j] = p0;
            S.py[j] = p1;
            S.pz[j] = p2;
            ke0p += p0 * (p0 / mass) + p1 * (p1 / mass) + p2 * (p2 / mass);
        }
        const double last_kin = 0.5 * block_sum(ke0p, S.red);
        double e_pot = hmc_energy_forces<POT>(pot, S, C, n);
        for (int s = 0; s < a.n_steps; ++s) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                // new_momenta = p + 0.5 F dt; x += new_momenta / m * dt; p = (x_new - x) * m / dt
                double nm, xo, xn;
                nm = S.px[j] + __dmul_rn(__dmul_rn(0.5, S.fx[j]), dt); xo = S.x[j]; xn = xo + __dmul_rn(nm / mass, dt); S.x[j] = xn; S.px[j] = __dmul_rn(xn - xo, mass) / dt;
                nm = S.py[j] + __dmul_rn(__dmul_rn(0.5, S.fy[j]), dt); xo = S.y[j]; xn = xo + __dmul_rn(nm / mass, dt); S.y[j] = xn; S.py[j] = __dmul_rn(xn - xo, mass) / dt;
                nm = S.pz[j] + __dmul_rn(__dmul_rn(0.5, S.fz[j]), dt); xo = S.z[j]; xn = xo + __dmul_rn(nm / mass, dt); S.z[j] = xn; S.pz[j] = __dmul_rn(xn - xo, mass) / dt;
            }
            __syncthreads();
            e_pot = hmc_energy_forces<POT>(pot, S, C, n);
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                S.px[j] = S.px[j] + __dmul_rn(__dmul_rn(0.5, S.fx[j]), dt);
                S.py[j] = S.py[j] + __dmul_rn(__dmul_rn(0.5, S.fy[j]), dt);
                S.pz[j] = S.pz[j] + __dmul_rn(__dmul_rn(0.5, S.fz[j]), dt);
            }
            __syncthreads();
        }
        double kep = 0.0;
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            kep += S.px[j] * (S.px[j] / mass) + S.py[j] * (S.py[j] / mass) + S.pz[j] * (S.pz[j] / mass);
        const double e_kin = 0.5 * block_sum(kep, S.red);
        const double delta = (e_pot + e_kin) - last_pot - last_kin;
        double u_accept, o2;
        uniform_pair(key, attempt, grep, 2, o2, u_accept);
        const double x = -delta / kT;
        if (x > EXP_OVERFLOW) status |= QMC_STATUS_EXP_OVERFLOW;
        const bool acc = u_accept < exp(x);
        if (acc) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                gx[j] = S.x[j]; gy[j] = S.y[j]; gz[j] = S.z[j];
                gpx[j] = S.px[j]; gpy[j] = S.py[j]; gpz[j] = S.pz[j];
            }
            last_pot = e_pot;
            if (threadIdx.x == 0) a.b.kinetic[r] = e_kin;
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                S.x[j] = gx[j]; S.y[j] = gy[j]; S.z[j] = gz[j];
            }
            if (threadIdx.x == 0) a.b.kinetic[r] = last_kin;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(a.b.counters) + (size_t)r * QMC_MAX_MOVES * 3;
            cnt[0] += 1;
            if (acc) cnt[1] += 1;
            const size_t ti = (size_t)r * a.n_attempts + k;
            if (a.tr.move) a.tr.move[ti] = 0;
            if (a.tr.accepted) a.tr.accepted[ti] = acc ? 1 : 0;
            if (a.tr.energy) a.tr.energy[ti] = last_pot;
            if (a.tr.delta) a.tr.delta[ti] = delta;
            if (a.tr.moved) a.tr.moved[ti] = -1;
        }
    }
    if (threadIdx.x == 0) {
        a.b.energy[r] = last_pot;
        if (status) atomicOr(a.b.status + r, status);
    }
}

inline int64_t hmc_workspace_bytes(const qmc_batch_t &b) {
    (void)b;
    return 256;  // the one-CTA-per-replica trajectory kernel keeps everything in shared memory
}

template <bool FORCES_ONLY>
inline int hmc_launch(const HmcArgs &a, cudaStream_t s, char *err, size_t errn) {
    const size_t smem = (size_t)hmc_smem_bytes(a.b.n_max);
    cudaError_t e;
    int dev = 0, smem_blk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if ((int64_t)smem > smem_blk) {
        snprintf(err, errn, "Hamiltonian replica of %d atoms needs %zu B of shared memory (> %d B per CTA)", a.b.n_max, smem, smem_blk);
        return QMC_ERR_UNSUPPORTED;
    }
    if (a.pot.kind == QMC_POT_EMT) {
        e = cudaFuncSetAttribute(hmc_kernel<QMC_POT_EMT, FORCES_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) hmc_kernel<QMC_POT_EMT, FORCES_ONLY><<<a.b.n_replicas, HMC_THREADS, smem, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(hmc_kernel<QMC_POT_LJ, FORCES_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) hmc_kernel<QMC_POT_LJ, FORCES_ONLY><<<a.b.n_replicas, HMC_THREADS, smem, s>>>(a);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(err, errn, "hmc launch failed: %s", cudaGetErrorString(e));
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA;
    }
    return QMC_OK;
}

inline int hmc_forces(const PotDev &pot, const qmc_batch_t &b, double *forces, cudaStream_t s, char *err, size_t errn) {
    HmcArgs a;
    memset(&a, 0, sizeof(a));
    a.pot = pot;
    a.b = b;
    a.forces_out = forces;
    a.ns = round_up(b.n_max, 4);
    return hmc_launch<true>(a, s, err, errn);
}

inline int hmc_run(const PotDev &pot, const qmc_batch_t &b, double dt, int n_steps, unsigned long long seed,
                   unsigned long long attempt_base, int n_attempts, const qmc_trace_t &tr, void *workspace, cudaStream_t s, char *err,
                   size_t errn) {
    (void)workspace;
    HmcArgs a;
    memset(&a, 0, sizeof(a));
    a.pot = pot;
    a.b = b;
    a.dt = dt;
    a.n_steps = n_steps;
    a.seed = seed;
    a.attempt_base = attempt_base;
    a.n_attempts = n_attempts;
    a.tr = tr;
    a.ns = round_up(b.n_max, 4);
    return hmc_launch<false>(a, s, err, errn);
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
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(err, errn, "pt_swap launch failed: %s", cudaGetErrorString(e));
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA;
    }
    return QMC_OK;
}

}  // namespace qmc
