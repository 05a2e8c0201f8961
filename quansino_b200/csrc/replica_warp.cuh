// One warp owns one replica: its state lives in shared memory for a whole launch.
//
// Building blocks used by the sweep kernel (local energy delta of one moved / inserted / deleted atom) and by
// the full-energy kernel (row-by-row evaluation with the SAME pair arithmetic, so that a local delta and the
// difference of two full energies agree to rounding).
//
// Work decomposition of a local delta for atom i (old position and/or new position):
//   scan     lanes stride over all other atoms j, enumerate the periodic images of j that can lie inside the
//            cutoff (nearest image, plus the second image along any axis where |d| > L - cutoff: needed for the
//            108-atom Cu cell, L = 10.83 A < 2 x 5.877 A), and COMPACT the hits into a warp-shared list with
//            ballot / popc -- so the transcendental-heavy pair functions below run with all 32 lanes busy.
//   flush    32 list entries per pass: theta, sigma_1 and sigma_2 terms; the signed sigma_1 terms are
//            scattered into dsig[j] without atomics (match_any groups duplicates of j, the group leader adds the
//            group's values in lane order => bit-reproducible); the sigma_2 terms only need a warp sum.
//   embed    the atoms whose sigma_1 changed (second compacted list) get their embedding energy re-evaluated.
#pragma once

#include "device_common.cuh"

namespace qmc {

constexpr unsigned SIDE_NEW = 0x8000u;  // list_j flag: entry belongs to the NEW position of the moved atom

// Per-warp view of the replica in shared memory.
struct Replica {
    double *x, *y, *z;  // [ns]
    double *sig;        // [ns]   sigma_1 cache (EMT)
    double *eat;        // [ns]   embedding-energy cache (EMT)
    double *dsig;       // [ns]   sigma_1 delta of the attempt in flight; all-zero between attempts
    double *lr;         // [cap]  list: r^2 of the hit, later reused for trial embedding energies
    uint16_t *lj;       // [cap]  list: neighbour index | SIDE_NEW
    uint16_t *l2;       // [ns]   list of atoms whose sigma_1 changed
    int n;              // atoms
    int cap;            // list capacity (multiple of 32, >= ns)
    double L[3], invL[3], thr[3];  // edge, 1/edge, edge - cutoff (second image exists iff |d| > thr)
    bool per[3];
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// shared-memory bytes of one replica
__host__ __device__ inline int64_t replica_smem_bytes(int kind, int n_max) {
    const int ns = round_up(n_max, 4);
    const int cap = round_up(ns > 208 ? ns : 208, 32);
    int64_t b = 0;
    b += 3 * (int64_t)ns * 8;                        // x y z
    if (kind == QMC_POT_EMT) b += 3 * (int64_t)ns * 8;  // sig eat dsig
    b += (int64_t)cap * 8;                           // lr
    b += (int64_t)cap * 2;                           // lj
    b += (int64_t)ns * 2;                            // l2
    return (b + 15) / 16 * 16;
}

__device__ inline void replica_carve(Replica &R, unsigned char *base, int kind, int n_max) {
    const int ns = round_up(n_max, 4);
    R.cap = round_up(ns > 208 ? ns : 208, 32);
    double *d = reinterpret_cast<double *>(base);
    R.x = d; d += ns;
    R.y = d; d += ns;
    R.z = d; d += ns;
    if (kind == QMC_POT_EMT) {
        R.sig = d; d += ns;
        R.eat = d; d += ns;
        R.dsig = d; d += ns;
    } else {
        R.sig = R.eat = R.dsig = nullptr;
    }
    R.lr = d; d += R.cap;
    R.lj = reinterpret_cast<uint16_t *>(d);
    R.l2 = R.lj + R.cap;
}

// returns false if a periodic edge is shorter than the cutoff (more than two images per axis would be needed)
__device__ inline bool replica_set_cell(Replica &R, const double cell_diag[3], const int pbc[3], double cut) {
    bool ok = true;
    const double cutp = cut * (1.0 + 1e-9);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R.L[k] = cell_diag[k];
        R.per[k] = pbc[k] != 0;
        R.invL[k] = R.per[k] ? 1.0 / cell_diag[k] : 0.0;
        R.thr[k] = R.per[k] ? cell_diag[k] - cutp : 1e300;
        if (R.per[k] && !(cell_diag[k] >= cutp)) ok = false;
    }
    return ok;
}

// nearest image d0 of (xj - xi) along one axis, and whether the second image d1 can be inside the cutoff
__device__ __forceinline__ void axis_images(double xi, double xj, double L, double invL, double thr, bool per, double &d0,
                                            double &d1, unsigned &extra, unsigned bit) {
    if (per) {
        const double q = (xi - xj) * invL;
        const double k0 = (q + RINT_MAGIC) - RINT_MAGIC;
        d0 = (xj + k0 * L) - xi;
        d1 = d0 - copysign(L, d0);
        extra |= (fabs(d0) > thr) ? bit : 0u;
    } else {
        d0 = xj - xi;
        d1 = d0;
    }
}

// ---------------------------------------------------------------------------------------------------
// list processing
// ---------------------------------------------------------------------------------------------------
struct DeltaAcc {
    double v;     // per-lane partial of the signed sigma_2 terms (EMT) or signed pair energies (LJ)
    double snew;  // per-lane partial of the sigma_1 terms of the NEW position (= sigma_1 of the moved atom)
    int pairs;    // per-lane count of evaluated list entries (statistics)
};

// mode 0: local delta (scatter signed sigma_1 into dsig); mode 1: one row of a full evaluation (sums only)
template <int POT, int MODE>
__device__ __forceinline__ void flush_list(const PotDev &pot, Replica &R, int &cnt, DeltaAcc &acc) {
    const int lane = lane_id();
    __syncwarp();
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e < cnt;
        const double r2 = act ? R.lr[e] : 1e300;
        const unsigned jj = act ? R.lj[e] : 0u;
        const bool isnew = (jj & SIDE_NEW) != 0;
        if (POT == QMC_POT_EMT) {
            double s1, s2;
            emt_pair(pot, r2, s1, s2);
            acc.v += isnew ? s2 : -s2;
            acc.snew += isnew ? s1 : 0.0;
            if (MODE == 0) {
                const double sg = isnew ? s1 : -s1;
                const unsigned j = jj & 0x7fffu;
                const unsigned key = act ? j : (0x10000u + lane);
                const unsigned grp = __match_any_sync(FULL, key);
                const int maxc = __reduce_max_sync(FULL, __popc(grp));
                double a = 0.0;
                unsigned rem = grp;
                for (int t = 0; t < maxc; ++t) {
                    const int src = rem ? (__ffs(rem) - 1) : lane;
                    const double val = __shfl_sync(FULL, sg, src);
                    if (rem) {
                        a += val;
                        rem &= rem - 1;
                    }
                }
                if (act && lane == (__ffs(grp) - 1)) R.dsig[j] += a;
                __syncwarp();
            }
        } else {
            const double ep = lj_pair(pot, r2);
            acc.v += isnew ? ep : -ep;
        }
        acc.pairs += act ? 1 : 0;
    }
    cnt = 0;
    __syncwarp();
}

// Scan all atoms j < n_scan (j != i_excl) against the old and/or new position of the moved atom and process the
// compacted hits.  MODE 0 also builds l2 (atoms whose sigma_1 changes) and returns its length.
template <int POT, int MODE>
__device__ __forceinline__ int scan_and_accumulate(const PotDev &pot, Replica &R, int n_scan, int i_excl, bool has_old,
                                                   bool has_new, const double po[3], const double pn[3], DeltaAcc &acc) {
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    int cnt = 0, n2 = 0;
    for (int base = 0; base < n_scan; base += 32) {
        const int j = base + lane;
        const bool valid = j < n_scan && j != i_excl;
        const double xj = valid ? R.x[j] : 0.0, yj = valid ? R.y[j] : 0.0, zj = valid ? R.z[j] : 0.0;
        bool emitted = false;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !has_old : !has_new) continue;  // warp-uniform
            const double *p = side == 0 ? po : pn;
            double d0x, d1x, d0y, d1y, d0z, d1z;
            unsigned avail = 0;
            axis_images(p[0], xj, R.L[0], R.invL[0], R.thr[0], R.per[0], d0x, d1x, avail, 1u);
            axis_images(p[1], yj, R.L[1], R.invL[1], R.thr[1], R.per[1], d0y, d1y, avail, 2u);
            axis_images(p[2], zj, R.L[2], R.invL[2], R.thr[2], R.per[2], d0z, d1z, avail, 4u);
            unsigned c = 0;
            bool more = valid;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? d1x : d0x, dy = (c & 2u) ? d1y : d0y, dz = (c & 4u) ? d1z : d0z;
                const double r2 = dx * dx + dy * dy + dz * dz;
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                if (b) {
                    if (cnt + 32 > R.cap) flush_list<POT, MODE>(pot, R, cnt, acc);
                    if (in) {
                        const int pos = cnt + __popc(b & lt);
                        R.lr[pos] = r2;
                        R.lj[pos] = (uint16_t)((unsigned)j | (side ? SIDE_NEW : 0u));
                    }
                    cnt += __popc(b);
                    emitted |= in;
                }
                if (more) {
                    if (c == avail) more = false;
                    else c = ((c | (~avail & 7u)) + 1u) & avail;
                }
            }
        }
        if (MODE == 0 && POT == QMC_POT_EMT) {
            const unsigned b2 = __ballot_sync(FULL, emitted);
            if (emitted) R.l2[n2 + __popc(b2 & lt)] = (uint16_t)j;
            n2 += __popc(b2);
        }
    }
    flush_list<POT, MODE>(pot, R, cnt, acc);
    return n2;
}

// Full evaluation of the replica in shared memory: fills sig / eat (EMT), returns the total energy on every lane.
// `pairs` (per-lane partial) counts the evaluated (atom, image) pairs.
template <int POT>
__device__ inline double replica_full_energy(const PotDev &pot, Replica &R, long long &pairs) {
    const int lane = lane_id();
    double epart = 0.0;  // lane partial of per-atom energies
    for (int a = 0; a < R.n; ++a) {
        const double pa[3] = {R.x[a], R.y[a], R.z[a]};
        DeltaAcc acc = {0.0, 0.0, 0};
        scan_and_accumulate<POT, 1>(pot, R, R.n, a, false, true, pa, pa, acc);
        pairs += acc.pairs;
        const double vs = warp_sum(acc.v);
        if (POT == QMC_POT_EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig[a] = s1;
                R.dsig[a] = vs;  // parked here until the embedding pass below; reset to zero there
            }
        } else {
            if (lane == (a & 31)) epart += 0.5 * vs;  // energies[ii] = 0.5 * pairwise_energies.sum()
        }
        __syncwarp();
    }
    if (POT == QMC_POT_EMT) {
        for (int a = lane; a < R.n; a += 32) {
            const double ea = emt_embed(pot, R.sig[a]);
            R.eat[a] = ea;
            // energies[a] = E_c + E_AS,2 - E0 + 0.5 * sum(es + eo), es = eo = neghalfv0overgamma2 * dsigma2
            epart += ea + pot.neghalfv0overgamma2 * R.dsig[a];
            R.dsig[a] = 0.0;
        }
        __syncwarp();
    }
    return warp_sum(epart);
}

}  // namespace qmc
