// One warp owns one replica: its state lives in shared memory for a whole launch.
//
// Building blocks used by the sweep kernel (local energy delta of one moved / inserted / deleted atom) and by the
// full-energy kernel (row-by-row evaluation with the SAME pair arithmetic, so that a local delta and the difference
// of two full energies agree to rounding).
//
// Local delta for atom i (old position and / or new position), EMT:
//   scan     lanes stride over the other atoms j; per side the NEAREST periodic image is tested against the cutoff
//            and the hits are COMPACTED into a warp-shared list (ballot / popc), so that the transcendental-heavy pair
//            function runs with all 32 lanes busy.  The owner lane of j remembers where its hits landed (pp[j]).
//            Second images (|d| > L - cutoff along an axis: needed for the 108-atom Cu cell, L = 10.83 A <
//            2 x 5.877 A) are queued as work items and appended to the list after all nearest images.
//   pairs    32 list entries per pass: theta, sigma_1 term, sigma_2 term.  The sigma_2 terms only need a warp sum;
//            the sigma_1 term is written back over the entry -- no scatter, no atomics.  Only the (few) second-image
//            entries add their term onto the slot of the nearest image of the same atom and side (match_any groups
//            duplicates, group leader adds in lane order => bit-reproducible).
//   embed    every atom whose sigma_1 changed (second compacted list l2) gathers its own old / new terms through
//            pp[j] and gets its embedding energy re-evaluated; trial energies go to the tail of the list buffer.
//   commit   on acceptance sigma_1 / embedding energy of the touched atoms are overwritten; a rejected trial is never
//            written anywhere.
#pragma once

#include "device_common.cuh"

namespace qmc {

constexpr int XCAP = 64;             // capacity for second-image list entries of one attempt
constexpr unsigned NOPOS = 0xFFFFu;  // pp half-word: no list entry

// Per-warp view of the replica in shared memory.
struct Replica {
    double *x, *y, *z;  // [ns]
    double *sig;        // [ns]    sigma_1 cache (EMT)
    double *eat;        // [ns]    embedding-energy cache (EMT)
    double *lr;         // [capl]  list: +-r^2 of a hit (sign = side: + new, - old), overwritten by its sigma_1 term;
                        //         tail [cape, capl) = work-item queue during the scan, trial energies afterwards
    uint32_t *pp;       // [ns]    list position of atom j's nearest-image hit: old | new << 16 (EMT delta)
    uint16_t *l2;       // [ns]    atoms whose sigma_1 changes (EMT delta)
    uint16_t *xj;       // [XCAP]  atom index of each second-image list entry (EMT delta)
    int n;              // atoms
    int ns;             // padded capacity
    int cape;           // list entry capacity (= capl - ns)
    double L[3], invL[3], thr[3];  // edge, 1/edge, edge - cutoff (a second image exists iff |d| > thr)
    bool per[3];
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int list_capacity(int ns) { return round_up(ns + 212, 32); }

// shared-memory bytes of one replica
__host__ __device__ inline int64_t replica_smem_bytes(int kind, int n_max) {
    const int ns = round_up(n_max, 4);
    int64_t b = 3 * (int64_t)ns * 8 + (int64_t)list_capacity(ns) * 8;
    if (kind == QMC_POT_EMT) b += 2 * (int64_t)ns * 8 + (int64_t)ns * 4 + (int64_t)ns * 2 + XCAP * 2;
    return (b + 15) / 16 * 16;
}

__device__ inline void replica_carve(Replica &R, unsigned char *base, int kind, int n_max) {
    const int ns = round_up(n_max, 4);
    const int capl = list_capacity(ns);
    R.ns = ns;
    R.cape = capl - ns;
    double *d = reinterpret_cast<double *>(base);
    R.x = d; d += ns;
    R.y = d; d += ns;
    R.z = d; d += ns;
    R.sig = R.eat = nullptr;
    if (kind == QMC_POT_EMT) {
        R.sig = d; d += ns;
        R.eat = d; d += ns;
    }
    R.lr = d; d += capl;
    R.pp = reinterpret_cast<uint32_t *>(d);
    R.l2 = reinterpret_cast<uint16_t *>(R.pp + ns);
    R.xj = R.l2 + ns;
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

// nearest image d0 of (xj - xi) along one axis; sets `bit` in `extra` if the second image d0 -+ L can be inside the cutoff
__device__ __forceinline__ double axis_nearest(double xi, double xj, double L, double invL, double thr, bool per, unsigned &extra,
                                               unsigned bit) {
    if (per) {
        const double q = (xi - xj) * invL;
        const double k0 = (q + RINT_MAGIC) - RINT_MAGIC;
        const double d0 = fma(k0, L, xj) - xi;
        extra |= (fabs(d0) > thr) ? bit : 0u;
        return d0;
    }
    return xj - xi;
}

// kept for the trajectory kernel: nearest and second image
__device__ __forceinline__ void axis_images(double xi, double xj, double L, double invL, double thr, bool per, double &d0,
                                            double &d1, unsigned &extra, unsigned bit) {
    d0 = axis_nearest(xi, xj, L, invL, thr, per, extra, bit);
    d1 = per ? d0 - copysign(L, d0) : d0;
}

struct DeltaAcc {
    double v;     // per-lane partial of the signed sigma_2 terms (EMT) or signed pair energies (LJ)
    double snew;  // per-lane partial of the sigma_1 terms of the NEW position (= sigma_1 of the moved atom)
    int pairs;    // per-lane count of evaluated list entries
};

// ---------------------------------------------------------------------------------------------------
// scan: fills the list.  MODE 0 = local delta (pp, l2, xj maintained for EMT), MODE 1 = one row of a full evaluation.
// Returns the number of list entries; n2 = entries of l2; xstart = first second-image entry.
// ---------------------------------------------------------------------------------------------------
template <int POT, int MODE>
__device__ __forceinline__ int scan_neighbours(const PotDev &pot, Replica &R, const int lane, const unsigned lt, int n_scan,
                                               int i_excl, bool has_old, bool has_new, const double po[3], const double pn[3],
                                               int &n2, int &xstart, int &status) {
    constexpr bool TRACK = (MODE == 0 && POT == QMC_POT_EMT);
    uint32_t *queue = reinterpret_cast<uint32_t *>(R.lr + R.cape);  // 2 * ns items fit in the tail (ns doubles)
    int cnt = 0, nq = 0;
    n2 = 0;
    for (int base = 0; base < n_scan; base += 32) {
        const int j = base + lane;
        const bool valid = j < n_scan && j != i_excl;
        const double xj = valid ? R.x[j] : 0.0, yj = valid ? R.y[j] : 0.0, zj = valid ? R.z[j] : 0.0;
        unsigned code = NOPOS | (NOPOS << 16);
        bool touched = false;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !has_old : !has_new) continue;  // warp-uniform
            const double *p = side == 0 ? po : pn;
            unsigned fl = 0;
            const double dx = axis_nearest(p[0], xj, R.L[0], R.invL[0], R.thr[0], R.per[0], fl, 1u);
            const double dy = axis_nearest(p[1], yj, R.L[1], R.invL[1], R.thr[1], R.per[1], fl, 2u);
            const double dz = axis_nearest(p[2], zj, R.L[2], R.invL[2], R.thr[2], R.per[2], fl, 4u);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = valid && r2 < pot.cut_pre2;
            const unsigned b = __ballot_sync(FULL, in);
            const int pos = cnt + __popc(b & lt);
            if (in && pos < R.cape) R.lr[pos] = side ? r2 : -r2;
            if (in) code = side ? ((code & 0xFFFFu) | ((unsigned)pos << 16)) : ((code & 0xFFFF0000u) | (unsigned)pos);
            cnt += __popc(b);
            touched |= in;
            // second images: only if the nearest image is inside (it is the closest of all images)
            const bool ex = in && fl != 0u;
            const unsigned bq = __ballot_sync(FULL, ex);
            if (ex) queue[nq + __popc(bq & lt)] = (unsigned)j | ((unsigned)side << 15) | (fl << 16);
            nq += __popc(bq);
        }
        if (TRACK) {
            if (valid) R.pp[j] = code;
            const unsigned b2 = __ballot_sync(FULL, touched);
            if (touched) R.l2[n2 + __popc(b2 & lt)] = (uint16_t)j;
            n2 += __popc(b2);
        }
    }
    xstart = cnt;
    // second images, appended after every nearest image
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const bool has = q < nq;
            const unsigned item = has ? queue[q] : 0u;
            const int j = item & 0x7FFFu;
            const unsigned side = (item >> 15) & 1u, fl = item >> 16;
            const double *p = side ? pn : po;
            unsigned dummy = 0;
            const double dx0 = axis_nearest(p[0], R.x[j], R.L[0], R.invL[0], R.thr[0], R.per[0], dummy, 1u);
            const double dy0 = axis_nearest(p[1], R.y[j], R.L[1], R.invL[1], R.thr[1], R.per[1], dummy, 2u);
            const double dz0 = axis_nearest(p[2], R.z[j], R.L[2], R.invL[2], R.thr[2], R.per[2], dummy, 4u);
            const double dx1 = dx0 - copysign(R.L[0], dx0), dy1 = dy0 - copysign(R.L[1], dy0), dz1 = dz0 - copysign(R.L[2], dz0);
            unsigned c = fl & (0u - fl);  // first non-empty subset of fl
            bool more = has;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? dx1 : dx0, dy = (c & 2u) ? dy1 : dy0, dz = (c & 4u) ? dz1 : dz0;
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = cnt + __popc(b & lt);
                if (in && pos < R.cape) {
                    R.lr[pos] = side ? r2 : -r2;
                    if (TRACK && pos - xstart < XCAP) R.xj[pos - xstart] = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
        }
    }
    if (cnt > R.cape || (TRACK && cnt - xstart > XCAP)) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = min(cnt, R.cape);
        if (TRACK && cnt - xstart > XCAP) cnt = xstart + XCAP;
    }
    __syncwarp();
    return cnt;
}

// pairs: evaluate the list.  MODE 0 leaves the sigma_1 term of every nearest-image entry in lr[e], with the terms of
// the second images of the same atom and side added onto it.
template <int POT, int MODE>
__device__ __forceinline__ void eval_list(const PotDev &pot, Replica &R, const int lane, int cnt, int xstart, DeltaAcc &acc) {
    constexpr bool TRACK = (MODE == 0 && POT == QMC_POT_EMT);
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e < cnt;
        const double r2s = act ? R.lr[e] : 1.0;
        const bool isnew = r2s > 0.0;
        const double r2 = fabs(r2s);
        if (POT == QMC_POT_EMT) {
            double s1, s2;
            emt_pair(pot, r2, s1, s2);
            s1 = act ? s1 : 0.0;
            s2 = act ? s2 : 0.0;
            acc.v += isnew ? s2 : -s2;
            acc.snew += isnew ? s1 : 0.0;
            if (TRACK) {
                if (e0 + 32 <= xstart) {  // nearest images only: every entry owns its slot
                    R.lr[e] = s1;
                } else {
                    const bool is_x = act && e >= xstart;
                    if (act && !is_x) R.lr[e] = s1;
                    __syncwarp();
                    unsigned tgt = 0x10000u + lane;
                    if (is_x) {
                        const unsigned code = R.pp[R.xj[e - xstart]];
                        tgt = isnew ? (code >> 16) : (code & 0xFFFFu);
                    }
                    const unsigned grp = __match_any_sync(FULL, tgt);
                    const int maxc = __reduce_max_sync(FULL, __popc(grp));
                    double a = 0.0;
                    unsigned rem = grp;
                    for (int t = 0; t < maxc; ++t) {
                        const int src = rem ? (__ffs(rem) - 1) : lane;
                        const double val = __shfl_sync(FULL, s1, src);
                        if (rem) {
                            a += val;
                            rem &= rem - 1;
                        }
                    }
                    if (is_x && lane == (__ffs(grp) - 1)) R.lr[tgt] += a;
                    __syncwarp();
                }
            }
        } else {
            const double ep = act ? lj_pair(pot, r2) : 0.0;
            acc.v += isnew ? ep : -ep;
        }
        acc.pairs += act ? 1 : 0;
    }
    __syncwarp();
}

// Full evaluation of the replica in shared memory: fills sig / eat (EMT), returns the total energy on every lane.
// `pairs` (per-lane partial) counts the evaluated (atom, image) pairs.
template <int POT>
__device__ inline double replica_full_energy(const PotDev &pot, Replica &R, const int lane, long long &pairs, int &status) {
    const unsigned lt = lanemask_lt();
    double epart = 0.0;  // lane partial of per-atom energies
    for (int a = 0; a < R.n; ++a) {
        const double pa[3] = {R.x[a], R.y[a], R.z[a]};
        DeltaAcc acc = {0.0, 0.0, 0};
        int n2, xstart;
        const int cnt = scan_neighbours<POT, 1>(pot, R, lane, lt, R.n, a, false, true, pa, pa, n2, xstart, status);
        eval_list<POT, 1>(pot, R, lane, cnt, xstart, acc);
        pairs += acc.pairs;
        const double vs = warp_sum(acc.v);
        if (POT == QMC_POT_EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig[a] = s1;
                R.eat[a] = vs;  // sigma_2 sum parked here until the embedding pass below
            }
        } else {
            if (lane == (a & 31)) epart += 0.5 * vs;  // energies[ii] = 0.5 * pairwise_energies.sum()
        }
        __syncwarp();
    }
    if (POT == QMC_POT_EMT) {
        for (int a = lane; a < R.n; a += 32) {
            const double ea = emt_embed(pot, R.sig[a]);
            // energies[a] = E_c + E_AS,2 - E0 + 0.5 * sum(es + eo), es = eo = neghalfv0overgamma2 * dsigma2
            epart += ea + pot.neghalfv0overgamma2 * R.eat[a];
            R.eat[a] = ea;
        }
        __syncwarp();
    }
    return warp_sum(epart);
}

}  // namespace qmc
