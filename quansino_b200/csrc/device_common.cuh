// Device-side building blocks shared by the sweep / energy / HMC kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quansino_b200.h"
#include "fastmath.cuh"

namespace qmc {

constexpr unsigned FULL = 0xffffffffu;
constexpr double KB = 8.617330337217213e-05;     // ase.units.kB (CODATA 2014)
constexpr double TWO_PI = 6.283185307179586;     // 2 * np.pi
constexpr double EXP_OVERFLOW = 709.782712893384;  // math.exp raises OverflowError above this
constexpr double RINT_MAGIC = 6755399441055744.0;  // 1.5 * 2^52: (q + M) - M == rint(q) for |q| < 2^51

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, slot layout of include/quansino_b200.h ("Random stream")
// ---------------------------------------------------------------------------------------------------
struct PhiloxKey {
    uint32_t k0, k1;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key,
                                              uint32_t (&out)[4]) {
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two doubles in [0, 1) with 53 random bits each: ((a >> 5) * 2^26 + (b >> 6)) * 2^-53 (all operations exact)
__device__ __forceinline__ void uniform_pair(PhiloxKey key, uint64_t attempt, uint32_t replica, uint32_t block, double &d0,
                                             double &d1) {
    uint32_t w[4];
    philox4x32_10((uint32_t)attempt, (uint32_t)(attempt >> 32), replica, block, key, w);
    d0 = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
    d1 = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;  // butterfly: every lane holds the same, order-deterministic sum
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------------------------------------------
// potentials: device constants derived once on the host from qmc_potential_t
// ---------------------------------------------------------------------------------------------------
struct PotDev {
    int kind;
    // neighbour cutoff, slightly widened for the distance prefilter; the exact test happens in the pair function
    double cut, cut_pre, cut_pre2;
    // EMT (ase/calculators/emt.py, single species)
    double acut, rc, rc_list, eta2, beta_s0, kappa, inv_beta, s0, lambda, E0, V0x6, inv12gamma1, neg_inv_beta_eta2;
    double neghalfv0overgamma2;
    double theta_c, sig1_a, sig1_c, sig2_a, sig2_c;  // exponent = a * r + c for theta, sigma_1, sigma_2
    // LJ (ase/calculators/lj.py, smooth=False)
    double sigma2, eps4, lj_rc2, lj_e0;
    // forces
    double kappa_over_beta;
};

// Exponential factors e1 (sigma_1), e2 (sigma_2) of one neighbour image at squared distance r2 > 0 and its cutoff weight
// w = theta(r) inside `r < rc_list` (_get_neighbors), 0 outside: the contributions are e1 * w and e2 * w (_calc_theta,
// _calc_dsigma1, _calc_dsigma2 with chi = 1).  Branch-free: 1 sqrt, 3 exp, 1 reciprocal.  Returns `r < rc_list`.
__device__ __forceinline__ bool emt_pair_w(const PotDev &p, double r2, double &e1, double &e2, double &w) {
    const double r = sqrt_fast(r2);
    // the three exponents are linear in r: one FMA each (constants folded on the host)
    double et;
    exp3_fast(fma(p.acut, r, p.theta_c), fma(p.sig1_a, r, p.sig1_c), fma(p.sig2_a, r, p.sig2_c), et, e1, e2);
    const double t = rcp_fast(1.0 + et);
    const bool in = r < p.rc_list;  // also discards whatever a padding entry (r2 = 1e6) produced
    w = in ? t : 0.0;
    return in;
}

// sigma_1 and sigma_2 contributions of one neighbour image at squared distance r2 (any r2 >= 0)
__device__ __forceinline__ void emt_pair(const PotDev &p, double r2, double &s1, double &s2) {
    double e1, e2, w;
    emt_pair_w(p, fmax(r2, 1e-200), e1, e2, w);
    s1 = e1 * w;
    s2 = e2 * w;
}

// E_c + second term of E_AS - E0 for one atom given its sigma_1 (_calc_e_c_a2 and the final `energies -= E0`);
// an atom without neighbours keeps energy 0 - E0.
__device__ __forceinline__ double emt_embed(const PotDev &p, double sigma1) {
    const double sg = sigma1 > 1e-14 ? sigma1 : 1.0;
    const double ds = log_fast(sg * p.inv12gamma1) * p.neg_inv_beta_eta2;
    const double x = p.lambda * ds;
    double ex, ek;
    exp2_fast(-x, -p.kappa * ds, ex, ek);
    const double e = p.E0 * (1.0 + x) * ex + p.V0x6 * ek - p.E0;
    return sigma1 > 1e-14 ? e : -p.E0;
}

// same, plus dE/dsigma_1 (self.deds of ASE)
__device__ __forceinline__ double emt_embed_deds(const PotDev &p, double sigma1, double &deds) {
    if (!(sigma1 > 1e-14)) {
        deds = 0.0;
        return -p.E0;
    }
    const double ds = log_fast(sigma1 * p.inv12gamma1) * p.neg_inv_beta_eta2;
    const double x = p.lambda * ds;
    const double ex = exp_fast(-x);
    const double z = p.V0x6 * exp_fast(-p.kappa * ds);
    // deds = (-E0 lambda x e^-x - kappa z) / (-beta eta2 sigma1)
    deds = (-p.E0 * p.lambda * x * ex - p.kappa * z) * p.neg_inv_beta_eta2 * rcp_fast(sigma1);
    return p.E0 * (1.0 + x) * ex + z - p.E0;
}

// pair energy of one neighbour image (LennardJones.calculate: c6[r2 > rc^2] = 0; e = 4 eps (c12 - c6) - e0 (c6 != 0));
// `in` = inside the cutoff
__device__ __forceinline__ double lj_pair(const PotDev &p, double r2, bool &in) {
    const double q = p.sigma2 * rcp_fast(fmax(r2, 1e-200));
    const double c6 = q * q * q;
    const double e = p.eps4 * (c6 * c6 - c6) - p.lj_e0;
    in = !(r2 > p.lj_rc2);
    return in ? e : 0.0;
}
__device__ __forceinline__ double lj_pair(const PotDev &p, double r2) {
    bool in;
    return lj_pair(p, r2, in);
}

// u < exp(x) without evaluating exp where the answer is already decided; flags the reference's OverflowError
__device__ __forceinline__ bool metropolis(double u, double x, int &status) {
    if (x > EXP_OVERFLOW) status |= QMC_STATUS_EXP_OVERFLOW;
    if (x >= 0.0) return true;  // exp(x) >= 1 > u
    if (!(x > -700.0)) return false;  // exp(x) < 1e-304: below every u > 0 (u == 0 has probability 2^-53; DESIGN.md ties)
    return u < exp_fast(x);
}

}  // namespace qmc
