// Force-bias Monte Carlo (src/quansino/mc/fbmc.py:202-273, https://doi.org/10.1063/1.4902136): every atom moves every step, each
// coordinate by zeta * delta with zeta in (-1, 1) drawn by REJECTION against the force-biased trial probability; there is no
// acceptance test.  One CTA per replica: the rejection loop of the reference redraws, call after call, exactly the coordinates
// that are still refused, in flattened (N, 3) order -- so element p of a call is the p-th unconverged coordinate, which takes a
// block-wide prefix sum per call.  Forces come from the neighbour-list kernels (hmc_kernel.cuh).
//
// Draws: element p of call c (zeta calls even, random calls odd) of step s = Philox(seed; s, global replica,
// ((c + 1) << 16) | (p >> 1)), half p & 1 (include/quansino_b200.h: qmc_fbmc_step).
#pragma once

#include "device_common.cuh"

namespace qmc {

constexpr int FBMC_THREADS = 256;
constexpr double FBMC_GAMMA_MAX = 709.782712;  // ForceBias.gamma_max_value

struct FbmcArgs {
    qmc_batch_t b;
    const double *forces;  // [R][3][n_max] at the current positions
    double delta;
    const double *delta_r;  // device [R] or nullptr: per-replica delta (AdaptiveForceBias, energy scheme)
    unsigned long long seed, step;
    double *gamma_max;  // [R] or null: max |gamma| of the step (the reference's "Gamma/GammaMax" logger field)
    int *n_calls;       // [R] or null: rejection rounds
};

// exclusive prefix sum of one int per thread over the CTA; `total` = the sum.  scratch: 32 ints of shared memory.
__device__ inline int block_exclusive_scan(int v, int *scratch, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const int s = scratch[w];
        if (w < wid) woff += s;
        tot += s;
    }
    total = tot;
    return woff + incl - v;
}

// dynamic shared memory: gamma [M], denominator [M], zeta [M] doubles + converged [M] bytes, M = 3 * n_max
__global__ void __launch_bounds__(FBMC_THREADS) fbmc_kernel(const __grid_constant__ FbmcArgs a) {
    extern __shared__ __align__(16) double fb_smem[];
    __shared__ int scratch[32];
    __shared__ double red[32];
    const int r = blockIdx.x, nmax = a.b.n_max, n = a.b.n_atoms[r], M = 3 * n;
    double *gam = fb_smem, *den = gam + 3 * nmax, *zeta = den + 3 * nmax;
    unsigned char *conv = reinterpret_cast<unsigned char *>(zeta + 3 * nmax);
    const size_t base = (size_t)r * 3 * nmax;
    const double kT2 = 2 * a.b.temperature[r] * KB;  // (2 * T) * kB, mc/fbmc.py:104
    const double delta = a.delta_r ? a.delta_r[r] : a.delta;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    // contiguous chunk of the flattened (N, 3) array per thread: element m = 3 i + c lives at forces[c][i]
    const int per = (M + FBMC_THREADS - 1) / FBMC_THREADS;
    const int m0 = min(M, (int)threadIdx.x * per), m1 = min(M, m0 + per);
    double gm = 0.0;
    for (int m = m0; m < m1; ++m) {  // calculate_gamma (mc/fbmc.py:93-108)
        const double f = a.forces[base + (size_t)(m % 3) * nmax + m / 3];
        double g = __ddiv_rn(__dmul_rn(f, delta), kT2);
        g = g < -FBMC_GAMMA_MAX ? -FBMC_GAMMA_MAX : (g > FBMC_GAMMA_MAX ? FBMC_GAMMA_MAX : g);
        gam[m] = g;
        den[m] = exp(g) - exp(-g);
        conv[m] = 0;
        gm = fmax(gm, fabs(g));
    }
    int calls = 0;
    for (;;) {  // mc/fbmc.py:224-241
        int mine = 0;
        for (int m = m0; m < m1; ++m) mine += conv[m] ? 0 : 1;
        int remaining;
        int p = block_exclusive_scan(mine, scratch, remaining);
        if (remaining == 0) break;
        for (int m = m0; m < m1; ++m) {
            if (conv[m]) continue;
            double z0, z1, u0, u1;
            uniform_pair(key, a.step, grep, (uint32_t)(((2 * calls + 1) << 16) | (p >> 1)), z0, z1);
            uniform_pair(key, a.step, grep, (uint32_t)(((2 * calls + 2) << 16) | (p >> 1)), u0, u1);
            const double z = __dadd_rn(-1.0, __dmul_rn(2.0, (p & 1) ? z1 : z0)), u = (p & 1) ? u1 : u0;
            const double sg = z > 0.0 ? 1.0 : (z < 0.0 ? -1.0 : 0.0);  // np.sign
            const double g = gam[m];
            // calculate_trial_probability (mc/fbmc.py:254-273)
            double pt = exp(__dmul_rn(sg, g)) - exp(__dmul_rn(g, __dsub_rn(__dmul_rn(2.0, z), sg)));
            pt = __dmul_rn(pt, sg);
            const double ptrial = den[m] != 0.0 ? __ddiv_rn(pt, den[m]) : 1.0;
            zeta[m] = z;
            conv[m] = ptrial > u ? 1 : 0;
            ++p;
        }
        ++calls;
        __syncthreads();
    }
    // displacement = zeta * delta * (min(m) / m)^0.25 (single species: 1); momenta round trip m * d / m (mc/fbmc.py:243-251)
    const double mass = a.b.mass;
    for (int m = m0; m < m1; ++m) {
        const double d = __dmul_rn(__dmul_rn(zeta[m], delta), 1.0);
        double *x = a.b.pos + base + (size_t)(m % 3) * nmax + m / 3;
        *x = __dadd_rn(*x, __ddiv_rn(__dmul_rn(mass, d), mass));
    }
    // max |gamma| of the replica
    gm = fmax(gm, 0.0);
    for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(FULL, gm, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) g = fmax(g, red[w]);
        if (a.gamma_max) a.gamma_max[r] = g;
        if (a.n_calls) a.n_calls[r] = calls;
    }
}

inline size_t fbmc_smem_bytes(int n_max) { return (size_t)3 * n_max * (3 * 8 + 1) + 16; }

inline int fbmc_launch(const FbmcArgs &a, cudaStream_t s, char *err, size_t errn) {
    const size_t smem = fbmc_smem_bytes(a.b.n_max);
    if (smem > 227 * 1024) {
        snprintf(err, errn, "force-bias step: %d atoms per replica need %zu bytes of shared memory", a.b.n_max, smem);
        return QMC_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(fbmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        fbmc_kernel<<<a.b.n_replicas, FBMC_THREADS, smem, s>>>(a);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) return QMC_OK;
    snprintf(err, errn, "force-bias kernel failed: %s", cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA;
}

}  // namespace qmc
