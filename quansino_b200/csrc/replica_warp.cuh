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
//
// The shared-memory layout is a compile-time function of (potential, NS = padded atom capacity): every array sits at a
// constant byte offset from the warp's base, so an access is one LDS / STS with an immediate offset.
#pragma once

#include "device_common.cuh"

namespace qmc {

constexpr int XCAP = 64;             // capacity for second-image list entries of one attempt (cells of >= 64 atoms)
constexpr unsigned NOPOS = 0xFFFFu;  // pp half-word: no list entry
constexpr unsigned NOITEM = 0xFFFFFFFFu;  // queue word: empty
constexpr int DRAW_AHEAD = 4;  // attempts whose random numbers one Philox round of the warp produces (4 lanes per attempt)
constexpr int SMEM_PER_SM = 232448;  // 227 KB usable per SM on sm_100
constexpr int SMEM_PER_CTA_MAX = 232448 - 1024;

__host__ __device__ constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }

// atom capacities the kernels are instantiated for
__host__ __device__ constexpr int capacity_class(int n_max) {
    return n_max <= 32 ? 32 : n_max <= 64 ? 64 : n_max <= 108 ? 108 : n_max <= 128 ? 128 : n_max <= 256 ? 256 : n_max <= 512 ? 512 : n_max <= 1024 ? 1024 : 0;
}

template <int POT, int NS>
struct Layout {
    static constexpr bool EMT = POT == QMC_POT_EMT;
    // list doubles: entries + tail of NS doubles.  EMT has 78 neighbour images per position at Cu density (2 x 78 + second images
    // of the other side ~ 180 entries per displacement; 212 leaves room for ~ 20 % compression); a Lennard-Jones solid with
    // the default cutoff 3 sigma has ~ 125 per position, and no sigma_1 / E_atom arrays to pay for
    static constexpr int CAPL = round_up(NS + (EMT ? 212 : 360), 32);
    static constexpr int CAPE = CAPL - NS;               // list entry capacity
    // offsets in doubles / bytes from the warp base
    static constexpr int X = 0, Y = NS, Z = 2 * NS;
    static constexpr int SIG = 3 * NS, EAT = 4 * NS;
    static constexpr int LR = EMT ? 5 * NS : 3 * NS;
    static constexpr int TAIL = LR + CAPE;
    static constexpr int CELLC = LR + CAPL;  // 9 doubles: L[3], invL[3], thr[3] (per-replica cells only)
    static constexpr int PP_B = (CELLC + 10) * 8;  // bytes
    static constexpr int L2_B = PP_B + (EMT ? NS * 4 : 0);
    static constexpr int XJ_B = L2_B + (EMT ? NS * 2 : 0);
    // second-image entries of one attempt: a 32-atom cell (L = 7.2 A) holds 31 nearest images but 78 neighbour images per
    // position, i.e. ~47 second images per side and ~94 for a displacement; from 64 atoms on, two dozen
    static constexpr int XCAPN = NS <= 64 ? 128 : XCAP;
    static constexpr int DRAW_B = round_up(XJ_B + (EMT ? XCAPN * 2 : 0), 16);  // DRAW_AHEAD attempts x 4 Philox blocks x 2 doubles
    static constexpr int CNT_B = DRAW_B + DRAW_AHEAD * 64;  // attempted / accepted / failed of this launch per move, u32
    static constexpr int DISP_B = CNT_B + QMC_MAX_MOVES * 3 * 4;  // bitmap of the atoms a composite move displaced so far
    static constexpr int FORCED_B = round_up(DISP_B + round_up(NS, 32) / 8, 4);  // cycle indices of this step's forced moves, i32
    static constexpr int LABW = round_up(NS, 32) / 32;               // words of one "movable atoms" bitmap
    static constexpr int LABBIT_B = FORCED_B + QMC_MAX_FORCED * 4;  // one bitmap per move-table entry (label >= 0, atom < n)
    static constexpr int BYTES = round_up(LABBIT_B + QMC_MAX_MOVES * LABW * 4, 16);
    // launch geometry: ONE wide CTA per SM holding as many replicas (warps) as shared memory and the 72-register budget allow.
    // Measured on B200 (4096 x 108-atom Cu EMT): 7 CTAs x 4 warps 3.10e8 attempts/s, 2 x 14 warps 3.35e8, 1 x 28 warps 3.68e8 --
    // the warp schedulers favour the warps of older CTAs, so replicas in separate CTAs finish 0.77 .. 1.39 ms into a 1.43 ms launch
    // and leave the SM under-occupied; the warps of one CTA progress evenly (profiles/r1_g_cta_width.md).
    static constexpr int fit = (SMEM_PER_SM - 1024) / BYTES;
#ifdef QMC_WARPS_PER_CTA
    static constexpr int WARPS = QMC_WARPS_PER_CTA;  // tuning experiments
    static constexpr int MINB = 28 / QMC_WARPS_PER_CTA > 0 ? 28 / QMC_WARPS_PER_CTA : 1;
#else
    static constexpr int WARPS = fit > 28 ? 28 : (fit < 1 ? 1 : fit);
    static constexpr int MINB = 1;
#endif
    static constexpr bool FITS = WARPS * BYTES <= SMEM_PER_CTA_MAX && fit > 0;
};

// Per-warp view of the replica: base pointer + runtime atom count.  Arrays are reached through the layout constants.
template <class LY>
struct Rep {
    static constexpr bool EMT = LY::EMT;
    static constexpr int CAPE = LY::CAPE;
    static constexpr int XCAPN = LY::XCAPN;
    static constexpr int QCAP = 2 * (LY::CAPL - LY::CAPE);  // work items the tail can hold
    double *b;  // warp base in shared memory
    int n;      // atoms
    __device__ __forceinline__ double &x(int j) const { return b[LY::X + j]; }
    __device__ __forceinline__ double &y(int j) const { return b[LY::Y + j]; }
    __device__ __forceinline__ double &z(int j) const { return b[LY::Z + j]; }
    __device__ __forceinline__ double &sig(int j) const { return b[LY::SIG + j]; }
    __device__ __forceinline__ double &eat(int j) const { return b[LY::EAT + j]; }
    __device__ __forceinline__ double &lr(int e) const { return b[LY::LR + e]; }
    __device__ __forceinline__ double &tail(int e) const { return b[LY::TAIL + e]; }
    __device__ __forceinline__ double &cellc(int k) const { return b[LY::CELLC + k]; }
    __device__ __forceinline__ uint32_t *queue() const { return reinterpret_cast<uint32_t *>(b + LY::TAIL); }
    __device__ __forceinline__ uint32_t &pp(int j) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::PP_B)[j]; }
    __device__ __forceinline__ uint16_t &l2(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::L2_B)[e]; }
    __device__ __forceinline__ uint16_t &xj(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::XJ_B)[e]; }
    __device__ __forceinline__ uint32_t *labbits(int m) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::LABBIT_B) + m * LY::LABW; }
    __device__ __forceinline__ int *forced() const { return reinterpret_cast<int *>(reinterpret_cast<char *>(b) + LY::FORCED_B); }
    __device__ __forceinline__ uint32_t *displaced() const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::DISP_B); }
    __device__ __forceinline__ uint32_t *counts() const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::CNT_B); }
    __device__ __forceinline__ double2 *draws() const { return reinterpret_cast<double2 *>(reinterpret_cast<char *>(b) + LY::DRAW_B); }
};

// cell constants of one axis: edge, 1 / edge, edge - cutoff' (a second image can be inside the cutoff iff |d| > thr)
struct Axis {
    double L, invL, thr;
};


// The cell as the kernels see it.  UNIFORM: every replica shares one constant cell, values come straight from the kernel
// arguments (constant bank operands, no registers); otherwise per-replica values are kept in shared memory.
struct CellArgs {
    double L[3], invL[3], thr[3];
    int per[3];
};

__host__ __device__ inline bool cell_constants(const double diag[3], const int pbc[3], double cut_pre, CellArgs &c) {
    bool ok = true;
    for (int k = 0; k < 3; ++k) {
        c.per[k] = pbc[k] != 0;
        c.L[k] = diag[k];
        c.invL[k] = c.per[k] ? 1.0 / diag[k] : 0.0;
        c.thr[k] = c.per[k] ? diag[k] - cut_pre : 1e300;
        if (c.per[k] && !(diag[k] >= cut_pre)) ok = false;
    }
    return ok;
}

template <class LY, bool UNIFORM>
struct CellView {
    static constexpr bool TRI = false;
    const CellArgs *u;  // kernel-argument cell (uniform) -- also carries the pbc flags
    double *s;          // shared-memory cell constants of this replica (per-replica cells)
    __device__ __forceinline__ Axis axis(int k) const {
        if (UNIFORM) return Axis{u->L[k], u->invL[k], u->thr[k]};
        return Axis{s[k], s[3 + k], s[6 + k]};
    }
    __device__ __forceinline__ bool per(int k) const { return u->per[k] != 0; }
};

// General (triclinic) cell of one replica, constants in shared memory (csrc/sweep_tri.cuh: store_cell_tri):
//   h[0..8]   H, row k = cell vector a_k (positions = fractional @ H)
//   t[0..8]   H^-1 (fractional_k = sum_x d_x * t[3 x + k]); the column of a non-periodic axis is zero
//   t[9..11]  thr_k = 1 - cutoff' / w_k (w_k = height of the cell along axis k), 1e300 for a non-periodic axis
// Image rule: with every periodic height w_k >= cutoff', an image of j inside the cutoff has fractional components |f_k| <=
// cutoff' / w_k <= 1 (f_k = r . b_k, |b_k| = 1 / w_k), so with fr = f - rint(f) the candidates are fr_k and, where |fr_k| > thr_k,
// fr_k - sign(fr_k): at most 2 x 2 x 2 images, the same subset enumeration as for diagonal cells -- but the image found by rounding
// is NOT always the closest one, so a flagged lane first finds the closest of its candidates (`nearest`); that image plays the
// part the nearest image plays in a diagonal cell (it owns the slot, the others are looked at only if it is inside the cutoff).
struct TriView {
    static constexpr bool TRI = true;
    const double *h;  // H[9]
    const double *t;  // H^-1[9], thr[3]
    __device__ __forceinline__ bool per(int k) const { return t[9 + k] < 1e299; }
    // image of d = xj - xi found by rounding the fractional coordinates; sg[k] = sign of the reduced fractional coordinate;
    // returns bit k set if the image shifted by -sg[k] a_k can be inside the cutoff as well
    __device__ __forceinline__ unsigned reduce(double dx, double dy, double dz, double d0[3], double sg[3]) const {
        const double f0 = fma(dz, t[6], fma(dy, t[3], dx * t[0]));
        const double f1 = fma(dz, t[7], fma(dy, t[4], dx * t[1]));
        const double f2 = fma(dz, t[8], fma(dy, t[5], dx * t[2]));
        const double k0 = (f0 + RINT_MAGIC) - RINT_MAGIC, k1 = (f1 + RINT_MAGIC) - RINT_MAGIC, k2 = (f2 + RINT_MAGIC) - RINT_MAGIC;
        d0[0] = fma(-k2, h[6], fma(-k1, h[3], fma(-k0, h[0], dx)));
        d0[1] = fma(-k2, h[7], fma(-k1, h[4], fma(-k0, h[1], dy)));
        d0[2] = fma(-k2, h[8], fma(-k1, h[5], fma(-k0, h[2], dz)));
        const double r0 = f0 - k0, r1 = f1 - k1, r2 = f2 - k2;
        sg[0] = copysign(1.0, r0);
        sg[1] = copysign(1.0, r1);
        sg[2] = copysign(1.0, r2);
        return (fabs(r0) > t[9] ? 1u : 0u) | (fabs(r1) > t[10] ? 2u : 0u) | (fabs(r2) > t[11] ? 4u : 0u);
    }
    // r^2 of the CLOSEST image among the candidates (subset `best` of the flagged axes `fl`): flagged lanes walk their <= 8 subsets
    __device__ __forceinline__ double nearest(double dx, double dy, double dz, double d0[3], double sg[3], unsigned &fl, unsigned &best) const {
        fl = reduce(dx, dy, dz, d0, sg);
        double rb = fma(d0[2], d0[2], fma(d0[1], d0[1], d0[0] * d0[0]));
        best = 0u;
        if (fl != 0u) {
            unsigned c = fl & (0u - fl);
            while (true) {
                double d[3];
                shifted(d0, sg, c, d);
                const double r2 = fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]));
                if (r2 < rb) {
                    rb = r2;
                    best = c;
                }
                if (c == fl) break;
                c = ((c | (~fl & 7u)) + 1u) & fl;
            }
        }
        return rb;
    }
    // the image of subset c (bit k: shifted by -sg[k] a_k)
    __device__ __forceinline__ void shifted(const double d0[3], const double sg[3], unsigned c, double d[3]) const {
        const double m0 = (c & 1u) ? sg[0] : 0.0, m1 = (c & 2u) ? sg[1] : 0.0, m2 = (c & 4u) ? sg[2] : 0.0;
        d[0] = fma(-m2, h[6], fma(-m1, h[3], fma(-m0, h[0], d0[0])));
        d[1] = fma(-m2, h[7], fma(-m1, h[4], fma(-m0, h[1], d0[1])));
        d[2] = fma(-m2, h[8], fma(-m1, h[5], fma(-m0, h[2], d0[2])));
    }
};

// nearest image d0 of (xj - xi) along one axis (4 FP64 operations, branch-free); sets `bit` in `extra` if the second image
// d0 -+ L can be inside the cutoff.  Positions are never wrapped (moves/displacement.py:109-141), so |xj - xi| may span
// several cells.  A non-periodic axis has invL = 0 and thr = 1e300 (cell_constants): k = 0, d0 = d exactly, never flagged.
__device__ __forceinline__ double axis_nearest(double xi, double xj, const Axis &A, bool /*per*/, unsigned &extra, unsigned bit) {
    const double d = xj - xi;
    const double k = fma(d, A.invL, RINT_MAGIC) - RINT_MAGIC;
    const double d0 = fma(-k, A.L, d);
    extra |= (fabs(d0) > A.thr) ? bit : 0u;
    return d0;
}

// the same with a boolean "a second image may be inside" that accumulates over axes (one chained compare per axis)
__device__ __forceinline__ double axis_nearest_any(double xi, double xj, const Axis &A, bool &extra) {
    const double d = xj - xi;
    const double k = fma(d, A.invL, RINT_MAGIC) - RINT_MAGIC;
    const double d0 = fma(-k, A.L, d);
    extra = extra || fabs(d0) > A.thr;
    return d0;
}

template <class CELL>
__device__ __forceinline__ Axis axis_of(const CELL &C, int k) {
    if constexpr (CELL::TRI) return Axis{0.0, 0.0, 0.0};
    else return C.axis(k);
}

struct DeltaAcc {
    double v;     // per-lane partial of the signed sigma_2 terms (EMT) or signed pair energies (LJ)
    double snew;  // per-lane partial of the sigma_1 terms of the NEW position (= sigma_1 of the moved atom)
    int pairs;    // per-lane count of evaluated list entries
};

// One-sided scan in a general cell (full-energy rows, and the local delta of an inserted / deleted atom): every image of every other
// atom inside the cutoff, signed r^2 into the list.  The CLOSEST image of an atom comes first (for the local delta it owns the
// atom's slot, pp / l2); the further images are looked at only if that one is inside, and follow behind all closest images.
template <class RV, bool TRACK, class CELL>
__device__ __forceinline__ int scan_neighbours_tri(const PotDev &pot, const RV &R, const CELL &C, const int lane, const unsigned lt, int n_scan,
                                                   int i_excl, bool side, const double p[3], int &n2, int &xstart, int &status,
                                                   int tile_first, int tile_step) {
    uint32_t *queue = R.queue();
    int cnt = 0, nq = 0;
    for (int base = tile_first * 32; base < n_scan; base += 32 * tile_step) {
        const int j = base + lane;
        const bool valid = j < n_scan && j != i_excl;
        const int jc = valid ? j : 0;
        double d0[3], sg[3];
        unsigned fl, best;
        const double r2 = C.nearest(R.x(jc) - p[0], R.y(jc) - p[1], R.z(jc) - p[2], d0, sg, fl, best);
        const bool in = valid && r2 < pot.cut_pre2;
        const unsigned b = __ballot_sync(FULL, in);
        const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
        if (in) {
            R.lr(pos) = side ? r2 : -r2;
            if (TRACK) R.l2(pos) = (uint16_t)j;
        }
        if (TRACK && valid) R.pp(j) = !in ? (NOPOS | (NOPOS << 16)) : side ? (NOPOS | ((unsigned)pos << 16)) : ((NOPOS << 16) | (unsigned)pos);
        cnt += __popc(b);
        const bool ex = in && fl != 0u;
        const unsigned bq = __ballot_sync(FULL, ex);
        if (ex) queue[min(nq + __popc(bq & lt), RV::QCAP - 1)] = (unsigned)j;
        nq += __popc(bq);
    }
    if (cnt > RV::CAPE) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = RV::CAPE;
    }
    n2 = cnt;
    xstart = cnt;
    if (nq > RV::QCAP) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        nq = RV::QCAP;
    }
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const bool has = q < nq;
            const int j = has ? (int)queue[q] : 0;
            double d0[3], sg[3];
            unsigned fl, best;
            C.nearest(R.x(j) - p[0], R.y(j) - p[1], R.z(j) - p[2], d0, sg, fl, best);
            unsigned c = 0u;  // every subset of the flagged axes except the closest image
            bool more = has && fl != 0u;
            while (__any_sync(FULL, more)) {
                double d[3];
                C.shifted(d0, sg, c, d);
                const double r2 = fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]));
                const bool in = more && c != best && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
                if (in) {
                    R.lr(pos) = side ? r2 : -r2;
                    if (TRACK) R.xj(min(pos - xstart, RV::XCAPN - 1)) = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
        }
    }
    if (cnt > RV::CAPE || (TRACK && cnt - xstart > RV::XCAPN)) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = min(cnt, RV::CAPE);
        if (TRACK && cnt - xstart > RV::XCAPN) cnt = xstart + RV::XCAPN;
    }
    __syncwarp();
    return cnt;
}

// ---------------------------------------------------------------------------------------------------
// One-sided scan (an inserted atom or a row of a full evaluation: new side; a deleted atom: old side): fills the list with the
// signed r^2 (+ new, - old) of every neighbour image inside the cutoff.  MODE 0 = local delta (pp, l2, xj maintained for EMT:
// list slot e belongs to touched atom l2[e]), MODE 1 = one row of a full evaluation.  Exactly one of has_old / has_new is
// set (displacements, which have both, go through scan_both below).  Returns the number of list entries; n2 = touched atoms;
// xstart = first second-image entry.
// ---------------------------------------------------------------------------------------------------
template <class RV, int MODE, class CELL>
__device__ __forceinline__ int scan_neighbours(const PotDev &pot, const RV &R, const CELL &C, const int lane, const unsigned lt,
                                               int n_scan, int i_excl, bool has_old, bool has_new, const double po[3],
                                               const double pn[3], int &n2, int &xstart, int &status, int tile_first = 0,
                                               int tile_step = 1) {
    constexpr bool TRACK = (MODE == 0 && RV::EMT);
    uint32_t *queue = R.queue();  // one word per flagged neighbour: at most n - 1 <= QCAP
    if constexpr (CELL::TRI) {
        return scan_neighbours_tri<RV, TRACK>(pot, R, C, lane, lt, n_scan, i_excl, has_new, has_new ? pn : po, n2, xstart, status, tile_first, tile_step);
    } else {
    const bool p0 = C.per(0), p1 = C.per(1), p2 = C.per(2);
    const Axis A0 = C.axis(0), A1 = C.axis(1), A2 = C.axis(2);
    const bool side = has_new;  // true: new position (positive r^2), false: old position (negative)
    const double px = side ? pn[0] : po[0], py = side ? pn[1] : po[1], pz = side ? pn[2] : po[2];
    int cnt = 0, nq = 0;
    for (int base = tile_first * 32; base < n_scan; base += 32 * tile_step) {
        const int j = base + lane;
        const bool valid = j < n_scan && j != i_excl;
        const int jc = valid ? j : 0;
        bool f = false;
        const double dx = axis_nearest_any(px, R.x(jc), A0, f);
        const double dy = axis_nearest_any(py, R.y(jc), A1, f);
        const double dz = axis_nearest_any(pz, R.z(jc), A2, f);
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const bool in = valid && r2 < pot.cut_pre2;
        const unsigned b = __ballot_sync(FULL, in);
        const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
        if (in) {
            R.lr(pos) = side ? r2 : -r2;
            if (TRACK) R.l2(pos) = (uint16_t)j;  // slot e of the nearest images belongs to touched atom l2[e]
        }
        if (TRACK && valid) R.pp(j) = !in ? (NOPOS | (NOPOS << 16)) : side ? (NOPOS | ((unsigned)pos << 16)) : ((NOPOS << 16) | (unsigned)pos);
        cnt += __popc(b);
        // second images: only if the nearest image is inside (it is the closest of all images)
        const bool ex = in && f;
        const unsigned bq = __ballot_sync(FULL, ex);
        if (ex) queue[min(nq + __popc(bq & lt), RV::QCAP - 1)] = (unsigned)j;
        nq += __popc(bq);
    }
    if (cnt > RV::CAPE) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = RV::CAPE;
    }
    n2 = cnt;
    xstart = cnt;
    if (nq > RV::QCAP) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        nq = RV::QCAP;
    }
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const bool has = q < nq;
            const int j = has ? (int)queue[q] : 0;
            unsigned fl = 0;
            const double dx0 = axis_nearest(px, R.x(j), A0, p0, fl, 1u);
            const double dy0 = axis_nearest(py, R.y(j), A1, p1, fl, 2u);
            const double dz0 = axis_nearest(pz, R.z(j), A2, p2, fl, 4u);
            const double dx1 = dx0 - copysign(A0.L, dx0), dy1 = dy0 - copysign(A1.L, dy0), dz1 = dz0 - copysign(A2.L, dz0);
            unsigned c = fl & (0u - fl);  // first non-empty subset of fl
            bool more = has && fl != 0u;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? dx1 : dx0, dy = (c & 2u) ? dy1 : dy0, dz = (c & 4u) ? dz1 : dz0;
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
                if (in) {
                    R.lr(pos) = side ? r2 : -r2;
                    if (TRACK) R.xj(min(pos - xstart, RV::XCAPN - 1)) = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
        }
    }
    if (cnt > RV::CAPE || (TRACK && cnt - xstart > RV::XCAPN)) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = min(cnt, RV::CAPE);
        if (TRACK && cnt - xstart > RV::XCAPN) cnt = xstart + RV::XCAPN;
    }
    __syncwarp();
    return cnt;
    }
}

// Two-sided scan for a single-atom displacement (old and new position of atom i_excl): every neighbour that is inside the cutoff on
// EITHER side gets an aligned PAIR of list slots, lr[pos] = -r2_old, lr[pos + 1] = +r2_new (a side that is outside keeps its
// r2 >= cut_pre2 and evaluates to zero), so that one ballot / popc compaction serves both sides, the list of touched atoms
// is simply l2[pos / 2] = j and the gather reads both terms with one 16-byte load.  The queue of second-image candidates
// holds j only; the image flags are recomputed when the queue is expanded.
// the two-sided scan in a general cell (see scan_both below for the slot layout)
template <class RV, class CELL>
__device__ __forceinline__ int scan_both_tri(const PotDev &pot, const RV &R, const CELL &C, const int lane, const unsigned lt, int n_scan,
                                         int i_excl, const double po[3], const double pn[3], int &n2, int &xstart, int &status,
                                         int tile_first, int tile_step) {
    constexpr bool TRACK = RV::EMT;
    static_assert(RV::CAPE % 2 == 0, "paired list slots need an even capacity");
    uint32_t *queue = R.queue();
    const bool p0 = C.per(0), p1 = C.per(1), p2 = C.per(2);
    const auto A0 = axis_of(C, 0), A1 = axis_of(C, 1), A2 = axis_of(C, 2);  // (unused for general cells)
    (void)p0; (void)p1; (void)p2; (void)A0; (void)A1; (void)A2;
    int cnt = 0, nq = 0;
    // one tile: 32 candidate atoms, measured (load -> nearest image -> r^2 -> compare) and filed (ballot -> slot)
    struct Tile {
        int j;
        bool valid, any, ex_o, ex_n;
        double neg_r2o, r2n;
    };
    auto measure = [&](const int base) {
        Tile t;
        t.j = base + lane;
        t.valid = t.j < n_scan && t.j != i_excl;
        const int jc = t.valid ? t.j : 0;
        const double xj = R.x(jc), yj = R.y(jc), zj = R.z(jc);
        if constexpr (CELL::TRI) {
            // general cells: the slot pair holds the CLOSEST image of each side (found among the <= 8 candidates), so that -- as for
            // diagonal cells -- further images need looking at only if that one is inside, and their sigma_1 terms have a slot
            double d0[3], sg[3];
            unsigned flo, fln, best;
            const double r2o = C.nearest(xj - po[0], yj - po[1], zj - po[2], d0, sg, flo, best);
            t.r2n = C.nearest(xj - pn[0], yj - pn[1], zj - pn[2], d0, sg, fln, best);
            const bool in_o = t.valid && r2o < pot.cut_pre2, in_n = t.valid && t.r2n < pot.cut_pre2;
            t.any = in_o || in_n;
            t.neg_r2o = __hiloint2double(__double2hiint(r2o) ^ (int)0x80000000, __double2loint(r2o));
            t.ex_o = in_o && flo != 0u;
            t.ex_n = in_n && fln != 0u;
            return t;
        } else {
        bool fo = false, fn = false;
        const double dxo = axis_nearest_any(po[0], xj, A0, fo);
        const double dyo = axis_nearest_any(po[1], yj, A1, fo);
        const double dzo = axis_nearest_any(po[2], zj, A2, fo);
        const double dxn = axis_nearest_any(pn[0], xj, A0, fn);
        const double dyn = axis_nearest_any(pn[1], yj, A1, fn);
        const double dzn = axis_nearest_any(pn[2], zj, A2, fn);
        const double r2o = fma(dzo, dzo, fma(dyo, dyo, dxo * dxo));
        t.r2n = fma(dzn, dzn, fma(dyn, dyn, dxn * dxn));
        const bool in_o = t.valid && r2o < pot.cut_pre2, in_n = t.valid && t.r2n < pot.cut_pre2;
        t.any = in_o || in_n;
        t.neg_r2o = __hiloint2double(__double2hiint(r2o) ^ (int)0x80000000, __double2loint(r2o));  // sign flip, integer pipe
        // second images: only if the nearest image is inside (it is the closest of all images)
        t.ex_o = in_o && fo;
        t.ex_n = in_n && fn;
        return t;
        }
    };
    auto file = [&](const Tile &t) {
        const unsigned b = __ballot_sync(FULL, t.any);
        const int pos = min(cnt + 2 * __popc(b & lt), RV::CAPE - 2);
        if (t.any) {
            *reinterpret_cast<double2 *>(&R.lr(pos)) = make_double2(t.neg_r2o, t.r2n);
            if (TRACK) R.l2(pos >> 1) = (uint16_t)t.j;
        }
        if (TRACK && t.valid) R.pp(t.j) = t.any ? ((unsigned)pos | ((unsigned)(pos + 1) << 16)) : (NOPOS | (NOPOS << 16));
        cnt += 2 * __popc(b);
        // A flagged lane pushes two queue words (old side, new side; NOITEM for a side that is not flagged): at most
        // 2 (n - 1) < QCAP words in total.
        const unsigned bq = __ballot_sync(FULL, t.ex_o || t.ex_n);
        if (t.ex_o || t.ex_n)
            *reinterpret_cast<uint2 *>(queue + nq + 2 * __popc(bq & lt)) = make_uint2(t.ex_o ? (unsigned)t.j : NOITEM, t.ex_n ? ((unsigned)t.j | 0x8000u) : NOITEM);
        nq += 2 * __popc(bq);
    };
    // (two tiles in flight per iteration were measured: -2.4 % on the same box, profiles/r2_a_rows.md; one tile at a time)
    for (int base = tile_first * 32; base < n_scan; base += 32 * tile_step) file(measure(base));
    if (cnt > RV::CAPE) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = RV::CAPE;
    }
    n2 = cnt >> 1;
    xstart = cnt;
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const unsigned item = q < nq ? queue[q] : NOITEM;
            const bool has = item != NOITEM;
            const int j = has ? (int)(item & 0x7FFFu) : 0;
            const bool side = (item & 0x8000u) != 0u;  // false: old position, true: new position
            if constexpr (CELL::TRI) {
                double d0[3], sg[3];
                unsigned fl, best;
                C.nearest(R.x(j) - (side ? pn[0] : po[0]), R.y(j) - (side ? pn[1] : po[1]), R.z(j) - (side ? pn[2] : po[2]), d0, sg, fl, best);
                unsigned c = 0u;  // every subset of the flagged axes (the empty one included) except the closest image, which has its slot
                bool more = has && fl != 0u;
                while (__any_sync(FULL, more)) {
                    double d[3];
                    C.shifted(d0, sg, c, d);
                    const double r2 = fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]));
                    const bool in = more && c != best && r2 < pot.cut_pre2;
                    const unsigned b = __ballot_sync(FULL, in);
                    const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
                    if (in) {
                        R.lr(pos) = side ? r2 : -r2;
                        if (TRACK) R.xj(min(pos - xstart, RV::XCAPN - 1)) = (uint16_t)j;
                    }
                    cnt += __popc(b);
                    if (more) {
                        if (c == fl) more = false;
                        else c = ((c | (~fl & 7u)) + 1u) & fl;
                    }
                }
                continue;
            } else {
            unsigned fl = 0;
            const double dx0 = axis_nearest(side ? pn[0] : po[0], R.x(j), A0, p0, fl, 1u);
            const double dy0 = axis_nearest(side ? pn[1] : po[1], R.y(j), A1, p1, fl, 2u);
            const double dz0 = axis_nearest(side ? pn[2] : po[2], R.z(j), A2, p2, fl, 4u);
            const double dx1 = dx0 - copysign(A0.L, dx0), dy1 = dy0 - copysign(A1.L, dy0), dz1 = dz0 - copysign(A2.L, dz0);
            unsigned c = fl & (0u - fl);  // first non-empty subset of fl
            bool more = has && fl != 0u;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? dx1 : dx0, dy = (c & 2u) ? dy1 : dy0, dz = (c & 4u) ? dz1 : dz0;
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
                if (in) {
                    R.lr(pos) = side ? r2 : -r2;
                    if (TRACK) R.xj(min(pos - xstart, RV::XCAPN - 1)) = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
            }
        }
    }
    if (cnt > RV::CAPE || (TRACK && cnt - xstart > RV::XCAPN)) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = min(cnt, RV::CAPE);
        if (TRACK && cnt - xstart > RV::XCAPN) cnt = xstart + RV::XCAPN;
    }
    __syncwarp();
    return cnt;
}

template <class RV, class CELL>
__device__ __forceinline__ int scan_both(const PotDev &pot, const RV &R, const CELL &C, const int lane, const unsigned lt, int n_scan,
                                         int i_excl, const double po[3], const double pn[3], int &n2, int &xstart, int &status,
                                         int tile_first = 0, int tile_step = 1) {
    if constexpr (CELL::TRI) {  // general cells: csrc/replica_warp.cuh scan_both_tri (the closest of <= 8 images owns the slot pair)
        return scan_both_tri(pot, R, C, lane, lt, n_scan, i_excl, po, pn, n2, xstart, status, tile_first, tile_step);
    } else {
    // (diagonal cells: the code below is kept instruction for instruction -- it is the inner loop of the headline kernel)
    constexpr bool TRACK = RV::EMT;
    static_assert(RV::CAPE % 2 == 0, "paired list slots need an even capacity");
    uint32_t *queue = R.queue();
    const bool p0 = C.per(0), p1 = C.per(1), p2 = C.per(2);
    const Axis A0 = C.axis(0), A1 = C.axis(1), A2 = C.axis(2);
    int cnt = 0, nq = 0;
    for (int base = tile_first * 32; base < n_scan; base += 32 * tile_step) {
        const int j = base + lane;
        const bool valid = j < n_scan && j != i_excl;
        const int jc = valid ? j : 0;
        const double xj = R.x(jc), yj = R.y(jc), zj = R.z(jc);
        bool fo = false, fn = false;
        const double dxo = axis_nearest_any(po[0], xj, A0, fo);
        const double dyo = axis_nearest_any(po[1], yj, A1, fo);
        const double dzo = axis_nearest_any(po[2], zj, A2, fo);
        const double dxn = axis_nearest_any(pn[0], xj, A0, fn);
        const double dyn = axis_nearest_any(pn[1], yj, A1, fn);
        const double dzn = axis_nearest_any(pn[2], zj, A2, fn);
        const double r2o = fma(dzo, dzo, fma(dyo, dyo, dxo * dxo));
        const double r2n = fma(dzn, dzn, fma(dyn, dyn, dxn * dxn));
        const bool in_o = valid && r2o < pot.cut_pre2, in_n = valid && r2n < pot.cut_pre2;
        const bool any = in_o || in_n;
        const unsigned b = __ballot_sync(FULL, any);
        const int pos = min(cnt + 2 * __popc(b & lt), RV::CAPE - 2);
        if (any) {
            const double neg_r2o = __hiloint2double(__double2hiint(r2o) ^ (int)0x80000000, __double2loint(r2o));  // sign flip, integer pipe
            *reinterpret_cast<double2 *>(&R.lr(pos)) = make_double2(neg_r2o, r2n);
            if (TRACK) R.l2(pos >> 1) = (uint16_t)j;
        }
        if (TRACK && valid) R.pp(j) = any ? ((unsigned)pos | ((unsigned)(pos + 1) << 16)) : (NOPOS | (NOPOS << 16));
        cnt += 2 * __popc(b);
        // second images: only if the nearest image is inside (it is the closest of all images).  A flagged lane pushes two
        // queue words (old side, new side; NOITEM for a side that is not flagged): at most 2 (n - 1) < QCAP words in total.
        const bool exo = in_o && fo, exn = in_n && fn;
        const unsigned bq = __ballot_sync(FULL, exo || exn);
        if (exo || exn)
            *reinterpret_cast<uint2 *>(queue + nq + 2 * __popc(bq & lt)) = make_uint2(exo ? (unsigned)j : NOITEM, exn ? ((unsigned)j | 0x8000u) : NOITEM);
        nq += 2 * __popc(bq);
    }
    if (cnt > RV::CAPE) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = RV::CAPE;
    }
    n2 = cnt >> 1;
    xstart = cnt;
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const unsigned item = q < nq ? queue[q] : NOITEM;
            const bool has = item != NOITEM;
            const int j = has ? (int)(item & 0x7FFFu) : 0;
            const bool side = (item & 0x8000u) != 0u;  // false: old position, true: new position
            unsigned fl = 0;
            const double dx0 = axis_nearest(side ? pn[0] : po[0], R.x(j), A0, p0, fl, 1u);
            const double dy0 = axis_nearest(side ? pn[1] : po[1], R.y(j), A1, p1, fl, 2u);
            const double dz0 = axis_nearest(side ? pn[2] : po[2], R.z(j), A2, p2, fl, 4u);
            const double dx1 = dx0 - copysign(A0.L, dx0), dy1 = dy0 - copysign(A1.L, dy0), dz1 = dz0 - copysign(A2.L, dz0);
            unsigned c = fl & (0u - fl);  // first non-empty subset of fl
            bool more = has && fl != 0u;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? dx1 : dx0, dy = (c & 2u) ? dy1 : dy0, dz = (c & 4u) ? dz1 : dz0;
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = min(cnt + __popc(b & lt), RV::CAPE - 1);
                if (in) {
                    R.lr(pos) = side ? r2 : -r2;
                    if (TRACK) R.xj(min(pos - xstart, RV::XCAPN - 1)) = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
        }
    }
    if (cnt > RV::CAPE || (TRACK && cnt - xstart > RV::XCAPN)) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = min(cnt, RV::CAPE);
        if (TRACK && cnt - xstart > RV::XCAPN) cnt = xstart + RV::XCAPN;
    }
    __syncwarp();
    return cnt;
    }
}

// pairs: evaluate the list.  MODE 0 leaves the sigma_1 term of every nearest-image entry in lr[e], with the terms of
// the second images of the same atom and side added onto it.
template <class RV, int MODE>
__device__ __forceinline__ void eval_pass(const PotDev &pot, const RV &R, const int lane, int e0, int cnt, int xstart,
                                          DeltaAcc &acc) {
    constexpr bool TRACK = (MODE == 0 && RV::EMT);
    const int e = e0 + lane;
    const bool act = e < cnt;
    const double r2s = act ? R.lr(e) : 1.0e6;  // padding lanes sit far outside every cutoff: their terms come out as zero
    const bool isnew = __double2hiint(r2s) >= 0;
    const double r2 = __hiloint2double(__double2hiint(r2s) & 0x7fffffff, __double2loint(r2s));  // |r2s| on the integer pipe
    if (RV::EMT) {
        double e1, e2, w;
        const bool in = emt_pair_w(pot, r2, e1, e2, w);
        const double s1 = e1 * w;
        acc.v = fma(e2, copysign(w, r2s), acc.v);
        if (isnew) acc.snew += s1;
        if (in) ++acc.pairs;
        if (TRACK) {
            if (e0 + 32 <= xstart) {  // nearest images only: every entry owns its slot
                R.lr(e) = s1;
            } else {
                const bool is_x = act && e >= xstart;
                if (act && !is_x) R.lr(e) = s1;
                __syncwarp();
                unsigned tgt = 0x10000u + lane;
                if (is_x) {
                    const unsigned code = R.pp(R.xj(e - xstart));
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
                if (is_x && lane == (__ffs(grp) - 1)) R.lr(tgt) += a;
                __syncwarp();
            }
        }
    } else {
        bool in;
        const double e = lj_pair(pot, r2, in);
        acc.v = fma(copysign(1.0, r2s), e, acc.v);
        acc.pairs += (act && in) ? 1 : 0;
    }
}

template <class RV, int MODE>
__device__ __forceinline__ void eval_list(const PotDev &pot, const RV &R, const int lane, int cnt, int xstart, DeltaAcc &acc) {
    for (int e0 = 0; e0 < cnt; e0 += 32) eval_pass<RV, MODE>(pot, R, lane, e0, cnt, xstart, acc);
    __syncwarp();
}

// Full evaluation of the replica in shared memory: fills sig / eat (EMT), returns the total energy on every lane;
// `atom_e` (global, may be null) receives the per-atom energies.
// `pairs` (per-lane partial) counts the evaluated (atom, image) pairs.
template <class LY, class CELL>
__device__ inline double replica_full_energy(const PotDev &pot, const Rep<LY> &R, const CELL &C, const int lane, long long &pairs,
                                             int &status, double *atom_e = nullptr) {
    const unsigned lt = lanemask_lt();
    double epart = 0.0;  // lane partial of per-atom energies
    for (int a = 0; a < R.n; ++a) {
        const double pa[3] = {R.x(a), R.y(a), R.z(a)};
        DeltaAcc acc = {0.0, 0.0, 0};
        int n2, xstart;
        const int cnt = scan_neighbours<Rep<LY>, 1>(pot, R, C, lane, lt, R.n, a, false, true, pa, pa, n2, xstart, status);
        eval_list<Rep<LY>, 1>(pot, R, lane, cnt, xstart, acc);
        pairs += acc.pairs;
        const double vs = warp_sum(acc.v);
        if (LY::EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig(a) = s1;
                R.eat(a) = vs;  // sigma_2 sum parked here until the embedding pass below
            }
        } else {
            if (lane == (a & 31)) {
                epart += 0.5 * vs;  // energies[ii] = 0.5 * pairwise_energies.sum()
                if (atom_e) atom_e[a] = 0.5 * vs;
            }
        }
        __syncwarp();
    }
    if (LY::EMT) {
        for (int a = lane; a < R.n; a += 32) {
            const double ea = emt_embed(pot, R.sig(a));
            // energies[a] = E_c + E_AS,2 - E0 + 0.5 * sum(es + eo), es = eo = neghalfv0overgamma2 * dsigma2
            const double e_a = ea + pot.neghalfv0overgamma2 * R.eat(a);
            epart += e_a;
            if (atom_e) atom_e[a] = e_a;  // results["energies"] of ASE's calculator
            R.eat(a) = ea;
        }
        __syncwarp();
    }
    return warp_sum(epart);
}

}  // namespace qmc
