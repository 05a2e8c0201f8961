// The Monte Carlo sweep kernel: n_attempts cycles of `MonteCarlo.step` (src/quansino/mc/core.py:353-384) for
// every replica of the batch, one warp per replica, replica state resident in shared memory.
//
// Per attempt, for one replica (reference file:line in brackets):
//   draws      Philox blocks 0..3 of (attempt, global replica), four attempts at a time       [mc/core.py:344-349 and below]
//   select     searchsorted(cdf, u_select, 'right') over the move table; forced moves         [mc/core.py:331-349]
//              (minimum_count) placed once per step
//   propose    DisplacementMove + Ball / Box / Sphere, label groups move rigidly               [moves/displacement.py:143-171,
//                                                                                               operations/displacement.py:67-153]
//              CompositeDisplacementMove / CompositeMove chains (+ closing CellMove)            [moves/displacement.py:392-431,
//                                                                                               moves/composite.py:43-62]
//              CellMove + IsotropicDeformation (axis mask)                                      [moves/cell.py:57-94, operations/cell.py:153-169]
//              ExchangeMove + Translation                                                       [moves/exchange.py:134-176]
//   energy     LOCAL delta of the moved / inserted / deleted atom(s) (the reference recomputes the whole system
//              through ASE: mc/criteria.py:104-106); full recompute only for cell moves
//   criteria   Canonical / Isobaric / GrandCanonical                                            [mc/criteria.py:89-110,146-172,226-278]
//   commit     in place on accept; nothing to undo on reject (composites restore a snapshot)    [mc/contexts.py:144-156,350-368]
#pragma once

#include "replica_warp.cuh"

namespace qmc {

struct MoveDev {
    int type, op;
    double step, bias;
    int default_label;
    int repeat, chain;  // composite displacement moves: this entry `repeat` times, then on to the next entry if `chain`
    int axes;           // cell moves: bit d set = axis d is deformed (never 0 here)
    int n_extra_ops, extra_op[2];  // CompositeOperation: further displacement operations, translations added in list order
    double extra_step[2];
    int *labels;  // device [R][n_max] or nullptr (= arange, no exchange)
};

struct SweepArgs {
    PotDev pot;
    CellArgs cell;  // pbc flags; cell constants when every replica shares one fixed cell
    qmc_batch_t b;
    MoveDev mv[QMC_MAX_MOVES];
    double cdf[QMC_MAX_MOVES];
    int n_moves;
    int n_forced, max_cycles;             // forced moves (minimum_count): how many per step, cycles per step
    int forced_move[QMC_MAX_FORCED];      // table index of forced move j (heads repeated minimum_count times, table order)
    int need_coin;  // any exchange move present
    unsigned long long seed, attempt_base;
    int n_attempts;
    int has_trace;
    double *atom_energies;  // energy kernel only: device [R][n_max] per-atom energies, or nullptr
    qmc_trace_t tr;
};

// --- label bookkeeping (moves/displacement.py:173-184, 226-253); labels live in global memory -------------------
__device__ inline int count_labelled(const int *lab, int n) {
    int c = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane_id();
        c += __popc(__ballot_sync(FULL, j < n && __ldcg(lab + min(j, n - 1)) >= 0));
    }
    return c;
}

// index of the idx-th atom (in index order) with a non-negative label; labels are strictly increasing with the atom
// index among the non-negative ones (validated on the host), so this IS unique_labels[idx]'s atom.
__device__ inline int select_labelled(const int *lab, int n, int idx) {
    int run = 0, found = -1;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane_id();
        const unsigned b = __ballot_sync(FULL, j < n && __ldcg(lab + min(j, n - 1)) >= 0);
        const int c = __popc(b);
        if (found < 0 && idx < run + c) found = base + (int)__fns(b, 0, idx - run + 1);
        run += c;
    }
    return found;
}

// CompositeDisplacementMove (moves/displacement.py:408-416): candidates = labelled atoms that no earlier sub-move of this
// attempt displaced (`displaced` = bitmap over atom indices; one atom per label, so this is setdiff1d on the labels)
__device__ inline int count_available(const int *lab, int n, const uint32_t *displaced) {
    int c = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane_id();
        const unsigned b = __ballot_sync(FULL, j < n && (!lab || __ldcg(lab + min(j, n - 1)) >= 0));
        c += __popc(b & ~displaced[base >> 5]);
    }
    return c;
}

__device__ inline int select_available(const int *lab, int n, const uint32_t *displaced, int idx) {
    int run = 0, found = -1;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane_id();
        const unsigned b = __ballot_sync(FULL, j < n && (!lab || __ldcg(lab + min(j, n - 1)) >= 0)) & ~displaced[base >> 5];
        const int c = __popc(b);
        if (found < 0 && idx < run + c) found = base + (int)__fns(b, 0, idx - run + 1);
        run += c;
    }
    return found;
}

// --- the same selections from a bitmap of the movable atoms kept in shared memory (one per move-table entry): the label VALUES
// only matter to the caller (and to max + 1 on insertion); with one atom per label, increasing with the atom index,
// unique_labels[idx] is the idx-th set bit.  The bitmaps are rebuilt from the global label tables when a replica is loaded and
// after every accepted exchange, so an attempt reads no label from global memory.
template <class RV>
__device__ inline void rebuild_movable(const RV &R, const MoveDev *mv, int n_moves, int r, int nmax, int words) {
    const int lane = lane_id();
    for (int q = 0; q < n_moves; ++q) {
        const int *lab = mv[q].labels ? mv[q].labels + (size_t)r * nmax : nullptr;
        uint32_t *bits = R.labbits(q);
        for (int w = 0; w < words; ++w) {
            const int j = w * 32 + lane;
            const unsigned b = __ballot_sync(FULL, j < R.n && (!lab || __ldcg(lab + min(j, nmax - 1)) >= 0));
            if (lane == 0) bits[w] = b;
        }
    }
    __syncwarp();
}

__device__ inline int count_bits(const uint32_t *bits, const uint32_t *without, int words) {
    const int lane = lane_id();
    const unsigned w = lane < words ? (bits[lane] & ~(without ? without[lane] : 0u)) : 0u;
    return __reduce_add_sync(FULL, __popc(w));
}

// atom index of the idx-th set bit (idx < count_bits)
__device__ inline int select_bit(const uint32_t *bits, const uint32_t *without, int words, int idx) {
    const int lane = lane_id();
    const unsigned w = lane < words ? (bits[lane] & ~(without ? without[lane] : 0u)) : 0u;
    const int c = __popc(w);
    int incl = c;  // inclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int excl = incl - c;
    const bool mine = idx >= excl && idx < incl;
    const unsigned who = __ballot_sync(FULL, mine);
    const int found = mine ? lane * 32 + (int)__fns(w, 0, idx - excl + 1) : 0;
    return who ? __shfl_sync(FULL, found, __ffs(who) - 1) : -1;
}

__device__ inline int max_label(const int *lab, int n) {
    int m = -1;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane_id();
        const int v = (j < n) ? __ldcg(lab + j) : -1;
        m = max(m, __reduce_max_sync(FULL, v));
    }
    return m;
}

// a label table shared by several entries of the move table (qmc_sweep_host uploads equal host arrays once; the sub-moves of a
// composite share theirs) is re-labelled ONCE per accepted exchange
template <class MV>
__device__ inline bool table_first_use(const MV *mv, int q) {
    for (int p = 0; p < q; ++p)
        if (mv[p].labels == mv[q].labels) return false;
    return true;
}

// on_atoms_changed for one move: append one label and / or remove index `removed`
__device__ inline void labels_changed(int *lab, int n, bool added, int removed, int default_label) {
    const int lane = lane_id();
    if (added) {
        const int mx = max_label(lab, n);
        // `self.default_label or (max(unique_labels) + 1 if len(unique_labels) else 0)`: None and 0 are falsy
        const int label = (default_label != QMC_LABEL_NONE && default_label != 0) ? default_label : (mx >= 0 ? mx + 1 : 0);
        if (lane == 0) lab[n] = label;
        n += 1;
        __syncwarp();
    }
    if (removed >= 0) {
        for (int base = removed; base < n - 1; base += 32) {
            const int j = base + lane;
            int t = 0;
            if (j < n - 1) t = __ldcg(lab + j + 1);
            __syncwarp();
            if (j < n - 1) lab[j] = t;
            __syncwarp();
        }
    }
}

__device__ inline void shift_down(double *arr, int removed, int n) {
    const int lane = lane_id();
    for (int base = removed; base < n - 1; base += 32) {
        const int j = base + lane;
        double t = 0.0;
        if (j < n - 1) t = arr[j + 1];
        __syncwarp();
        if (j < n - 1) arr[j] = t;
        __syncwarp();
    }
}


// --- run-coded label groups (qmc_move_t.chain bit 3): non-negative labels never decrease with the atom index, a group is a
// contiguous run of equal labels, the k-th unique label (np.unique sorts) is the k-th run ---------------------------------
__device__ inline bool run_head(const int *lab, int n, int j) {
    if (j >= n) return false;
    const int v = __ldcg(lab + j);
    return v >= 0 && (j == 0 || __ldcg(lab + j - 1) != v);
}

__device__ inline int run_count(const int *lab, int n) {
    int c = 0;
    for (int base = 0; base < n; base += 32) c += __popc(__ballot_sync(FULL, run_head(lab, n, base + lane_id())));
    return c;
}

// first atom and length of the idx-th run (idx < run_count)
__device__ inline void run_select(const int *lab, int n, int idx, int &first, int &len) {
    int seen = 0;
    first = -1;
    for (int base = 0; base < n; base += 32) {
        const unsigned b = __ballot_sync(FULL, run_head(lab, n, base + lane_id()));
        const int c = __popc(b);
        if (first < 0 && idx < seen + c) first = base + (int)__fns(b, 0, idx - seen + 1);
        seen += c;
    }
    len = 0;
    if (first < 0) return;
    const int v = __ldcg(lab + first);
    for (int base = first; base < n; base += 32) {  // equal labels directly behind the head
        const int j = base + lane_id();
        const unsigned same = __ballot_sync(FULL, j < n && __ldcg(lab + min(j, n - 1)) == v);
        const int run = same == FULL ? 32 : __ffs(~same) - 1;
        len += run;
        if (run < 32) break;
    }
}

// on_atoms_changed for a molecule (moves/displacement.py:226-253): `added` atoms get ONE new label behind the last atom, or the
// `removed` atoms starting at `first` disappear (order preserved)
__device__ inline void labels_molecule_changed(int *lab, int n, int added, int first, int removed, int default_label) {
    const int lane = lane_id();
    if (added > 0) {
        const int mx = max_label(lab, n);
        const int label = (default_label != QMC_LABEL_NONE && default_label != 0) ? default_label : (mx >= 0 ? mx + 1 : 0);
        if (lane < added) lab[n + lane] = label;
        __syncwarp();
    } else if (removed > 0) {
        for (int base = first; base < n - removed; base += 32) {
            const int j = base + lane;
            int t = 0;
            if (j < n - removed) t = __ldcg(lab + j + removed);
            __syncwarp();
            if (j < n - removed) lab[j] = t;
            __syncwarp();
        }
    }
}

// z-x-z Euler matrix A = B (C D) of ASE's euler_rotate for the angles 2 pi u[q] read as DEGREES (operations/displacement.py:192-200);
// products and sums in numpy's order, without contraction
__device__ inline void euler_matrix(const double u3[3], double A[9]) {
    double sn[3], cs[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) sincos_2pi(__dmul_rn(__dmul_rn(TWO_PI, u3[q]), 3.141592653589793 / 180), sn[q], cs[q]);
    const double Dm[9] = {cs[0], sn[0], 0.0, -sn[0], cs[0], 0.0, 0.0, 0.0, 1.0};
    const double Cm[9] = {1.0, 0.0, 0.0, 0.0, cs[1], sn[1], 0.0, -sn[1], cs[1]};
    const double Bm[9] = {cs[2], sn[2], 0.0, -sn[2], cs[2], 0.0, 0.0, 0.0, 1.0};
    double CD[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int i3 = q / 3 * 3, j3 = q % 3;
        CD[q] = __dadd_rn(__dadd_rn(__dmul_rn(Cm[i3], Dm[j3]), __dmul_rn(Cm[i3 + 1], Dm[3 + j3])), __dmul_rn(Cm[i3 + 2], Dm[6 + j3]));
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int i3 = q / 3 * 3, j3 = q % 3;
        A[q] = __dadd_rn(__dadd_rn(__dmul_rn(Bm[i3], CD[j3]), __dmul_rn(Bm[i3 + 1], CD[3 + j3])), __dmul_rn(Bm[i3 + 2], CD[6 + j3]));
    }
}

// --- local energy delta -----------------------------------------------------------------------------------------
struct Trial {
    double dE;     // energy difference of the whole system
    double sig_i;  // sigma_1 of the moved atom at its new position (EMT)
    double eat_i;  // its embedding energy there
    int n2;        // atoms whose sigma_1 changed: l2[0..n2); their trial embedding energies sit in lr[cape + e]
    int pairs;     // pair evaluations done (per-lane partial)
};

// sigma_1 change of touched atom j, gathered from the list through pp[j] (terms of second images were added onto the
// nearest-image slots by eval_list)
template <class RV>
__device__ __forceinline__ double gathered_dsigma(const RV &R, int j) {
    const unsigned code = R.pp(j);
    const unsigned po = code & 0xFFFFu, pn = code >> 16;
    const double so = po != NOPOS ? R.lr(po) : 0.0;
    const double sn = pn != NOPOS ? R.lr(pn) : 0.0;
    return sn - so;
}

template <class LY, class CELL>
__device__ __forceinline__ Trial local_delta(const PotDev &pot, const Rep<LY> &R, const CELL &C, const int lane, const unsigned lt,
                                             int n_scan, int i, bool has_old, bool has_new, const double po[3], const double pn[3],
                                             int &status, int skip_when_adding = -1) {
    // `skip_when_adding`: an insertion normally scans every stored atom (the new one is not stored yet); a relocated atom is
    // still stored at its old place and must not see itself
    Trial T;
    DeltaAcc acc = {0.0, 0.0, 0};
    int xstart;
    const int cnt = scan_neighbours<Rep<LY>, 0>(pot, R, C, lane, lt, n_scan, has_old ? i : skip_when_adding, has_old, has_new, po, pn, T.n2, xstart, status);
    eval_list<Rep<LY>, 0>(pot, R, lane, cnt, xstart, acc);
    const double vs = warp_sum(acc.v);
    T.pairs = acc.pairs;
    T.sig_i = 0.0;
    T.eat_i = 0.0;
    if (!LY::EMT) {
        T.dE = vs;  // E = 1/2 sum_i sum_j e_ij: the moved atom's row and column change together
        return T;
    }
    if (has_new) T.sig_i = warp_sum(acc.snew);
    double part = 0.0, e_i_lane = 0.0;
    const int total = T.n2 + (has_new ? 1 : 0);
    for (int e = lane; e < total; e += 32) {
        const bool nb = e < T.n2;
        const int j = nb ? R.l2(e) : 0;
        const double sn = nb ? (R.sig(j) + gathered_dsigma(R, j)) : T.sig_i;
        const double eold = nb ? R.eat(j) : (has_old ? R.eat(i) : 0.0);
        const double en = emt_embed(pot, sn);
        part += en - eold;
        if (nb) R.tail(e) = en;
        else e_i_lane = en;
    }
    if (has_new) T.eat_i = __shfl_sync(FULL, e_i_lane, T.n2 & 31);
    if (!has_new && lane == 0) part -= R.eat(i);  // deletion: the atom's own embedding energy disappears
    // pair part: rows (i, j) and (j, i) change together: 2 * 0.5 * (es + eo) = 2 * neghalfv0overgamma2 * dsigma2
    T.dE = warp_sum(part) + 2.0 * pot.neghalfv0overgamma2 * vs;
    __syncwarp();
    return T;
}

// commit the trial values of the touched atoms (accepted moves only; a rejected trial leaves no trace)
template <class LY>
__device__ __forceinline__ void commit_trial(const Rep<LY> &R, const int lane, const Trial &T) {
    if constexpr (LY::EMT) {
        for (int e = lane; e < T.n2; e += 32) {
            const int j = R.l2(e);
            R.sig(j) = R.sig(j) + gathered_dsigma(R, j);
            R.eat(j) = R.tail(e);
        }
        __syncwarp();
    }
}

// Displacement of one atom (both sides present): paired list slots (scan_both).  Touched atom l2[e] owns the slots 2e (old term)
// and 2e + 1 (new term); after the embedding pass they hold its trial sigma_1 and trial embedding energy.
template <class LY, class CELL>
__device__ __forceinline__ Trial local_delta_move(const PotDev &pot, const Rep<LY> &R, const CELL &C, const int lane, const unsigned lt,
                                                  int n_scan, int i, const double po[3], const double pn[3], int &status) {
    Trial T;
    DeltaAcc acc = {0.0, 0.0, 0};
    int xstart;
    const int cnt = scan_both<Rep<LY>>(pot, R, C, lane, lt, n_scan, i, po, pn, T.n2, xstart, status);
    eval_list<Rep<LY>, 0>(pot, R, lane, cnt, xstart, acc);
    const double vs = warp_sum(acc.v);
    T.pairs = acc.pairs;
    T.sig_i = 0.0;
    T.eat_i = 0.0;
    if (!LY::EMT) {
        T.dE = vs;  // E = 1/2 sum_i sum_j e_ij: the moved atom's row and column change together
        return T;
    }
    T.sig_i = warp_sum(acc.snew);
    double part = 0.0, e_i_lane = 0.0;
    double2 *slots = reinterpret_cast<double2 *>(&R.lr(0));
    // one item = one touched atom (or, item n2, the moved atom): log + 2 exp (two items in flight per lane were measured: slower,
    // profiles/r2_a_rows.md)
    for (int e = lane; e <= T.n2; e += 32) {
        const bool nb = e < T.n2;
        const int j = nb ? R.l2(e) : i;
        const double2 t = nb ? slots[e] : make_double2(0.0, 0.0);
        const double sn = nb ? R.sig(j) + (t.y - t.x) : T.sig_i;
        const double en = emt_embed(pot, sn);
        part += en - R.eat(j);
        if (nb) slots[e] = make_double2(sn, en);
        else e_i_lane = en;
    }
    T.eat_i = __shfl_sync(FULL, e_i_lane, T.n2 & 31);
    // pair part: rows (i, j) and (j, i) change together: 2 * 0.5 * (es + eo) = 2 * neghalfv0overgamma2 * dsigma2
    T.dE = warp_sum(part) + 2.0 * pot.neghalfv0overgamma2 * vs;
    __syncwarp();
    return T;
}

template <class LY>
__device__ __forceinline__ void commit_move(const Rep<LY> &R, const int lane, const Trial &T) {
    if constexpr (LY::EMT) {
        const double2 *slots = reinterpret_cast<const double2 *>(&R.lr(0));
        for (int e = lane; e < T.n2; e += 32) {
            const int j = R.l2(e);
            const double2 t = slots[e];
            R.sig(j) = t.x;
            R.eat(j) = t.y;
        }
        __syncwarp();
    }
}

// GrandCanonicalCriteria.evaluate (mc/criteria.py:243-278): criteria * prefactor
__device__ inline double gc_probability(double dE, int delta, int n_ex, double acc_volume, double mass, double temperature,
                                        double mu, bool &overflow) {
    // particle_delta = +-1 for plain exchange moves; a CompositeExchangeMove changes several particles at once (mc/criteria.py:246-262)
    double vol = delta > 0 ? acc_volume : 1.0 / acc_volume;
    double factorial_term = delta > 0 ? 1.0 / (double)(n_ex + 1) : (double)n_ex;
    if (delta > 1 || delta < -1) {
        vol = pow(acc_volume, (double)delta);
        factorial_term = 1.0;
        if (delta > 0)
            for (int i = n_ex + 1; i < n_ex + delta + 1; ++i) factorial_term /= (double)i;
        else
            for (int i = n_ex + delta + 1; i < n_ex + 1; ++i) factorial_term *= (double)i;
    }
    const double denom = 2 * 3.141592653589793 * mass * KB * temperature / 6.022140857e23 * 1e-3 * 1.6021766208e-19;
    const double lam = sqrt(6.626070040e-34 * 6.626070040e-34 / denom) * 1e10;
    const double debroglie = pow(lam, (double)(-3 * delta));
    const double prefactor = vol * factorial_term * debroglie;
    const double exponential = ((double)delta * mu - dE) / (temperature * KB);
    overflow = exponential > EXP_OVERFLOW;
    return exp(exponential) * prefactor;
}

// feature bits of a sweep-kernel instantiation: code for move types that are not in the table is compiled out, which
// keeps the canonical fast path inside its register budget
constexpr int F_LABELS = 1;    // label tables in global memory (not arange)
constexpr int F_CELL = 2;      // cell moves
constexpr int F_EXCHANGE = 4;  // exchange moves
constexpr int F_CELLVAR = 8;   // per-replica cells (kept in shared memory); otherwise one constant cell for all
constexpr int F_COMPOSITE = 16;  // composite displacement moves (chains of move-table entries)

// Philox block of the (u_label, op draw 0) pair of drawing sub-move c >= 1 of a composite move; (op draw 1, op draw 2) sit in the
// next block.  Sub-moves 1 .. 16 use blocks 32 .. 63 (the layout of the golden traces); the extra draws start at block 64, so
// sub-moves 17 and up continue at 0x8000 (`DisplacementMove * N + CellMove` with N = len(atoms) is the reference's own example).
__host__ __device__ __forceinline__ uint32_t submove_block(int c) {
    return c <= 16 ? (uint32_t)(32 + 2 * (c - 1)) : 0x8000u + (uint32_t)(2 * (c - 17));
}

// translation of one displacement sub-move from its three operation draws (operations/displacement.py:67-153)
__device__ __forceinline__ void propose_translation(int op, double step, double o0, double o1, double o2, double t[3]) {
    if (op == QMC_OP_BALL) {
        const double rr = step * o0, phi = TWO_PI * o1, c = -1.0 + 2.0 * o2;
        const double s = sqrt_fast(__dsub_rn(1.0, __dmul_rn(c, c)) + 1e-300);  // + 1e-300: exact no-op unless c = +-1 (keeps rsqrt finite)
        double sp, cp;
        sincos_2pi(phi, sp, cp);
        t[0] = __dmul_rn(__dmul_rn(rr, s), cp);
        t[1] = __dmul_rn(__dmul_rn(rr, s), sp);
        t[2] = __dmul_rn(rr, c);
    } else if (op == QMC_OP_SPHERE) {
        const double phi = TWO_PI * o0, c = -1.0 + 2.0 * o1;
        const double s = sqrt_fast(__dsub_rn(1.0, __dmul_rn(c, c)) + 1e-300);  // + 1e-300: exact no-op unless c = +-1 (keeps rsqrt finite)
        double sp, cp;
        sincos_2pi(phi, sp, cp);
        t[0] = __dmul_rn(step, __dmul_rn(s, cp));
        t[1] = __dmul_rn(step, __dmul_rn(s, sp));
        t[2] = __dmul_rn(step, c);
    } else {  // QMC_OP_BOX
        const double w = step - -step;
        t[0] = __dadd_rn(-step, __dmul_rn(w, o0));
        t[1] = __dadd_rn(-step, __dmul_rn(w, o1));
        t[2] = __dadd_rn(-step, __dmul_rn(w, o2));
    }
}

// the random numbers of one attempt: lanes 0..3 run one Philox block each, the six / seven doubles are broadcast
struct Draws {
    double u_select, u_label, o0, o1, o2, u_accept, u_coin, u_swap;
};

__device__ __forceinline__ Draws draw_attempt(PhiloxKey key, unsigned long long attempt, uint32_t replica, const int lane,
                                              bool need_coin) {
    double d0, d1;
    uniform_pair(key, attempt, replica, (uint32_t)(lane & 3), d0, d1);
    Draws D;
    D.u_select = __shfl_sync(FULL, d0, 0);
    D.u_label = __shfl_sync(FULL, d1, 0);
    D.o0 = __shfl_sync(FULL, d0, 1);
    D.o1 = __shfl_sync(FULL, d1, 1);
    D.o2 = __shfl_sync(FULL, d0, 2);
    D.u_accept = __shfl_sync(FULL, d1, 2);
    D.u_coin = need_coin ? __shfl_sync(FULL, d0, 3) : 0.0;
    D.u_swap = 0.0;
    return D;
}

// The same numbers, DRAW_AHEAD attempts at a time: lane 4 a + b runs Philox block b of attempt k0 + a and parks its two doubles
// in shared memory; every attempt then reads its seven doubles with broadcast loads.  One Philox evaluation per
// DRAW_AHEAD attempts instead of one per attempt; the values are identical (keyed by attempt index, not by call order).
template <class RV>
__device__ __forceinline__ void draw_ahead(const RV &R, PhiloxKey key, unsigned long long attempt0, uint32_t replica, const int lane) {
    double d0, d1;
    uniform_pair(key, attempt0 + (unsigned long long)((lane >> 2) & (DRAW_AHEAD - 1)), replica, (uint32_t)(lane & 3), d0, d1);
    __syncwarp();
    if (lane < 4 * DRAW_AHEAD) R.draws()[lane] = make_double2(d0, d1);
    __syncwarp();
}

template <class RV>
__device__ __forceinline__ Draws read_draws(const RV &R, int a, bool need_coin) {
    const double2 *d = R.draws() + 4 * a;
    const double2 b0 = d[0], b1 = d[1], b2 = d[2];
    Draws D;
    D.u_select = b0.x;
    D.u_label = b0.y;
    D.o0 = b1.x;
    D.o1 = b1.y;
    D.o2 = b2.x;
    D.u_accept = b2.y;
    D.u_coin = need_coin ? d[3].x : 0.0;
    D.u_swap = need_coin ? d[3].y : 0.0;
    return D;
}

// per-replica cell constants in shared memory (F_CELLVAR); returns false if an edge is below the cutoff
template <class RV>
__device__ __forceinline__ bool store_cell(const RV &R, const double diag[3], const int pbc[3], double cut_pre, int lane) {
    CellArgs c;
    const bool ok = cell_constants(diag, pbc, cut_pre, c);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.cellc(k) = c.L[k];
            R.cellc(3 + k) = c.invL[k];
            R.cellc(6 + k) = c.thr[k];
        }
    }
    __syncwarp();
    return ok;
}

template <int POT, int NS, int FEAT>
__global__ void __launch_bounds__(Layout<POT, NS>::WARPS * 32, Layout<POT, NS>::MINB) sweep_kernel(const __grid_constant__ SweepArgs a) {
    using LY = Layout<POT, NS>;
    constexpr int WARPS = LY::WARPS;
    constexpr bool UNIFORM = (FEAT & F_CELLVAR) == 0;
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    Rep<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    {  // keep the (32-bit shared-window) base in a register: otherwise it is re-derived from the thread id at every other use
        unsigned sb = (unsigned)__cvta_generic_to_shared(R.b);  // (S2R x 2 + 4 integer instructions each time; +0.8 % measured)
        asm volatile("" : "+r"(sb));
        R.b = reinterpret_cast<double *>(__cvta_shared_to_generic((size_t)sb));
    }
    const CellView<LY, UNIFORM> C{&a.cell, R.b + LY::CELLC};
    const int r = blockIdx.x * WARPS + warp;  // warp w of the grid owns replica w for the whole launch
    if (r >= a.b.n_replicas) return;
    const int k_begin = 0, k_end = a.n_attempts;
    // ---- load the replica ----
    R.n = __ldcg(a.b.n_atoms + r);
    double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    double *gs = POT == QMC_POT_EMT ? a.b.sigma1 + (size_t)r * nmax : nullptr;
    double *ge = POT == QMC_POT_EMT ? a.b.eatom + (size_t)r * nmax : nullptr;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? __ldcg(gx + j) : 0.0;
        R.y(j) = in ? __ldcg(gy + j) : 0.0;
        R.z(j) = in ? __ldcg(gz + j) : 0.0;
        if (POT == QMC_POT_EMT) {
            R.sig(j) = in ? __ldcg(gs + j) : 0.0;
            R.eat(j) = in ? __ldcg(ge + j) : 0.0;
        }
    }
    int status = 0;
    double cell[3] = {0.0, 0.0, 0.0};
    if (!UNIFORM) {
        cell[0] = __ldcg(a.b.cell + (size_t)r * 9 + 0);
        cell[1] = __ldcg(a.b.cell + (size_t)r * 9 + 4);
        cell[2] = __ldcg(a.b.cell + (size_t)r * 9 + 8);
        if (!store_cell(R, cell, a.b.pbc, pot.cut_pre, lane)) status |= QMC_STATUS_CELL_TOO_SMALL;
    }
    double E = __ldcg(a.b.energy + r);
    const double temperature = a.b.temperature[r];
    const double kT = temperature * KB;
    const double inv_kT = 1.0 / kT;  // the acceptance exponents multiply by this (<= 1 ulp from the reference's division)
    if (lane < QMC_MAX_MOVES * 3) R.counts()[lane] = 0u;
    int n_ex = (FEAT & F_EXCHANGE) ? __ldcg(a.b.n_exchange + r) : 0;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    long long n_pairs = 0;
    __syncwarp();
    if (FEAT & F_LABELS) rebuild_movable(R, a.mv, a.n_moves, r, nmax, LY::LABW);

    for (int k = k_begin; k < k_end; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        const int ahead = (k - k_begin) & (DRAW_AHEAD - 1);
        if (ahead == 0) draw_ahead(R, key, attempt, grep, lane);
        const Draws D = read_draws(R, ahead, (FEAT & (F_EXCHANGE | F_COMPOSITE)) != 0);

        int m = 0;
        while (m < a.n_moves - 1 && a.cdf[m] <= D.u_select) ++m;
        if ((FEAT & F_COMPOSITE) && a.n_forced > 0) {
            // forced moves (mc/core.py:331-342): at the start of a step draw n_forced distinct cycle indices -- a partial
            // Fisher-Yates shuffle of arange(max_cycles), lane q holding override q of the sparse permutation -- then a cycle
            // whose index was drawn performs its forced move instead of the selected one
            const int c = k % a.max_cycles;
            if (c == 0) {
                double d0, d1;
                uniform_pair(key, attempt, grep, 4u + (uint32_t)(lane >> 1), d0, d1);
                const double u_mine = (lane & 1) ? d1 : d0;
                int okey = -1, oval = 0, n_over = 0;
                for (int j = 0; j < a.n_forced; ++j) {
                    const double u = __shfl_sync(FULL, u_mine, j);
                    const int t = j + (int)(u * (double)(a.max_cycles - j));
                    const unsigned bt = __ballot_sync(FULL, lane < n_over && okey == t);
                    const unsigned bj = __ballot_sync(FULL, lane < n_over && okey == j);
                    const int vt = bt ? __shfl_sync(FULL, oval, __ffs(bt) - 1) : t;
                    const int vj = bj ? __shfl_sync(FULL, oval, __ffs(bj) - 1) : j;
                    if (bt) {
                        if (lane == __ffs(bt) - 1) oval = vj;  // perm[t] = perm[j]
                    } else {
                        if (lane == n_over) {
                            okey = t;
                            oval = vj;
                        }
                        ++n_over;
                    }
                    if (lane == 0) R.forced()[j] = vt;  // perm[j] = old perm[t]
                }
                __syncwarp();
            }
            const unsigned bf = __ballot_sync(FULL, lane < a.n_forced && R.forced()[lane] == c);
            if (bf) m = a.forced_move[__ffs(bf) - 1];
        }
        const MoveDev &mv = a.mv[m];
        const int *lab = ((FEAT & F_LABELS) && mv.labels) ? mv.labels + (size_t)r * nmax : nullptr;

        int accepted = -1, moved = -1;
        double delta = __longlong_as_double(0x7ff8000000000000LL);  // NaN

        if ((FEAT & F_COMPOSITE) && mv.type == QMC_MOVE_DISPLACEMENT && (mv.chain & 12)) {
            // Label groups (moves/displacement.py:164-167, 126-127): every atom whose label equals the selected one moves by the
            // SAME translation; one criteria evaluation.  `lab` holds dense group ranks (the library's caller normalises them in the
            // order of the sorted unique labels), so unique_labels[idx] is rank idx.  Same snapshot / commit / restore scheme as the
            // composite moves below.
            // (chain bit 3: the groups are runs of equal labels -- tables with a molecular exchange move -- and the members of the
            // selected group are the atoms [g_first, g_first + g_len); bit 2: `lab` holds dense ranks, members have lab == g)
            const bool runs = (mv.chain & 8) != 0;
            const int ng = !lab ? 0 : runs ? run_count(lab, R.n) : max_label(lab, R.n) + 1;
            if (ng > 0) {
                const int g = (int)(D.u_label * (double)ng);
                int g_first = 0, g_len = 0;
                if (runs) run_select(lab, R.n, g, g_first, g_len);
                auto member = [&](const int j) {
                    return runs ? (j >= g_first && j < g_first + g_len) : (j < R.n && __ldcg(lab + min(j, R.n - 1)) == g);
                };
                const bool rot = mv.op == QMC_OP_ROTATION || mv.op == QMC_OP_TRANSLATION_ROTATION;
                const bool relocate = mv.op == QMC_OP_TRANSLATION || mv.op == QMC_OP_TRANSLATION_ROTATION;
                double t[3] = {0.0, 0.0, 0.0}, tt[3] = {0.0, 0.0, 0.0}, A[9], com[3] = {0.0, 0.0, 0.0};
                if (relocate) {
                    // Translation.calculate (operations/displacement.py:170-175): uniform(0, 1, (1, 3)) @ cell -
                    // positions[moving].mean(axis=0); the mean adds the rows in index order and divides by their number
                    double sum[3] = {0.0, 0.0, 0.0};
                    int members_n = 0;
                    for (int base = 0; base < R.n; base += 32) {
                        const int j = base + lane;
                        unsigned members = __ballot_sync(FULL, member(j));
                        while (members) {
                            const int i = base + __ffs(members) - 1;
                            members &= members - 1;
                            sum[0] = __dadd_rn(sum[0], R.x(i));
                            sum[1] = __dadd_rn(sum[1], R.y(i));
                            sum[2] = __dadd_rn(sum[2], R.z(i));
                            ++members_n;
                        }
                    }
                    const double f[3] = {D.o0, D.o1, D.o2};
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        tt[q] = __dsub_rn(__dmul_rn(f[q], UNIFORM ? a.cell.L[q] : cell[q]), __ddiv_rn(sum[q], (double)max(members_n, 1)));
                }
                if (!rot) {
                    if (relocate) {
                        t[0] = tt[0];
                        t[1] = tt[1];
                        t[2] = tt[2];
                    } else {
                        propose_translation(mv.op, mv.step, D.o0, D.o1, D.o2, t);
                    }
                } else {
                    // Rotation.calculate (operations/displacement.py:192-200): euler_rotate(phi, theta, psi, center="COM") with
                    // phi, theta, psi = uniform(0, 2 pi, 3) -- ASE reads them as DEGREES -- z-x-z, A = B (C D); products and sums in
                    // numpy's order, without contraction
                    double ms = 0.0;
                    for (int base = 0; base < R.n; base += 32) {
                        const int j = base + lane;
                        unsigned members = __ballot_sync(FULL, member(j));
                        while (members) {
                            const int i = base + __ffs(members) - 1;
                            members &= members - 1;
                            com[0] = __dadd_rn(com[0], __dmul_rn(a.b.mass, R.x(i)));
                            com[1] = __dadd_rn(com[1], __dmul_rn(a.b.mass, R.y(i)));
                            com[2] = __dadd_rn(com[2], __dmul_rn(a.b.mass, R.z(i)));
                            ms = __dadd_rn(ms, a.b.mass);
                        }
                    }
                    double sn[3], cs[3];
                    double u3[3] = {D.o0, D.o1, D.o2};
                    if (relocate) {  // TranslationRotation (:220-227): the translation drew first; the angles are extra draws 0, 1, 2
                        double d0, d1, d2, d3;
                        uniform_pair(key, attempt, grep, 64u, d0, d1);
                        uniform_pair(key, attempt, grep, 65u, d2, d3);
                        u3[0] = d0;
                        u3[1] = d1;
                        u3[2] = d2;
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        com[q] = __ddiv_rn(com[q], ms);
                        sincos_2pi(__dmul_rn(__dmul_rn(TWO_PI, u3[q]), 3.141592653589793 / 180), sn[q], cs[q]);
                    }
                    const double Dm[9] = {cs[0], sn[0], 0.0, -sn[0], cs[0], 0.0, 0.0, 0.0, 1.0};
                    const double Cm[9] = {1.0, 0.0, 0.0, 0.0, cs[1], sn[1], 0.0, -sn[1], cs[1]};
                    const double Bm[9] = {cs[2], sn[2], 0.0, -sn[2], cs[2], 0.0, 0.0, 0.0, 1.0};
                    double CD[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int i3 = q / 3 * 3, j3 = q % 3;
                        CD[q] = __dadd_rn(__dadd_rn(__dmul_rn(Cm[i3], Dm[j3]), __dmul_rn(Cm[i3 + 1], Dm[3 + j3])), __dmul_rn(Cm[i3 + 2], Dm[6 + j3]));
                    }
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int i3 = q / 3 * 3, j3 = q % 3;
                        A[q] = __dadd_rn(__dadd_rn(__dmul_rn(Bm[i3], CD[j3]), __dmul_rn(Bm[i3 + 1], CD[3 + j3])), __dmul_rn(Bm[i3 + 2], CD[6 + j3]));
                    }
                }
                for (int j = lane; j < R.n; j += 32) {
                    gx[j] = R.x(j);
                    gy[j] = R.y(j);
                    gz[j] = R.z(j);
                    if (POT == QMC_POT_EMT) {
                        gs[j] = R.sig(j);
                        ge[j] = R.eat(j);
                    }
                }
                __syncwarp();
                double de_total = 0.0;
                int n_moved = 0;
                for (int base = 0; base < R.n; base += 32) {
                    const int j = base + lane;
                    unsigned members = __ballot_sync(FULL, member(j));
                    while (members) {
                        const int i = base + __ffs(members) - 1;
                        members &= members - 1;
                        const double po[3] = {R.x(i), R.y(i), R.z(i)};
                        if (rot) {  // molecule.positions - positions[moving]: A (x - com) + com - x
                            const double rc[3] = {__dsub_rn(po[0], com[0]), __dsub_rn(po[1], com[1]), __dsub_rn(po[2], com[2])};
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const double ar = __dadd_rn(__dadd_rn(__dmul_rn(A[3 * q], rc[0]), __dmul_rn(A[3 * q + 1], rc[1])), __dmul_rn(A[3 * q + 2], rc[2]));
                                t[q] = __dsub_rn(__dadd_rn(ar, com[q]), po[q]);
                                if (relocate) t[q] = __dadd_rn(tt[q], t[q]);  // translation (1, 3) + rotation (k, 3)
                            }
                        }
                        const double pn[3] = {po[0] + t[0], po[1] + t[1], po[2] + t[2]};
                        Trial T;
                        if (relocate) {
                            // old and new surroundings are unrelated (twice the entries of a local move, beyond the paired list):
                            // take the atom out, then put it in -- the two one-sided deltas of an exchange move
                            const Trial T1 = local_delta<LY>(pot, R, C, lane, lt, R.n, i, true, false, po, pn, status);
                            commit_trial<LY>(R, lane, T1);
                            T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, false, true, po, pn, status, i);
                            commit_trial<LY>(R, lane, T);
                            n_pairs += T1.pairs + T.pairs;
                            de_total += T1.dE + T.dE;
                        } else {
                            T = local_delta_move<LY>(pot, R, C, lane, lt, R.n, i, po, pn, status);
                            n_pairs += T.pairs;
                            de_total += T.dE;
                            commit_move<LY>(R, lane, T);
                        }
                        if (lane == 0) {
                            R.x(i) = pn[0];
                            R.y(i) = pn[1];
                            R.z(i) = pn[2];
                            if (POT == QMC_POT_EMT) {
                                R.sig(i) = T.sig_i;
                                R.eat(i) = T.eat_i;
                            }
                        }
                        __syncwarp();
                        if (n_moved == 0) moved = i;
                        ++n_moved;
                    }
                }
                if (n_moved > 0) {
                    delta = de_total;
                    const bool acc = metropolis(D.u_accept, -delta * inv_kT, status);
                    if (acc) {
                        E += delta;
                    } else {
                        for (int j = lane; j < R.n; j += 32) {
                            R.x(j) = gx[j];
                            R.y(j) = gy[j];
                            R.z(j) = gz[j];
                            if (POT == QMC_POT_EMT) {
                                R.sig(j) = gs[j];
                                R.eat(j) = ge[j];
                            }
                        }
                    }
                    __syncwarp();
                    accepted = acc ? 1 : 0;
                }
            }
        } else if ((FEAT & F_COMPOSITE) && mv.type == QMC_MOVE_DISPLACEMENT && (mv.repeat > 1 || (mv.chain & 3))) {
            // Composite moves + one criteria evaluation (mc/core.py:369-377).  chain bit 1 clear: CompositeDisplacementMove
            // (moves/displacement.py:392-431: no label displaced twice); set: plain CompositeMove (moves/composite.py:43-62:
            // independent sub-moves, optionally closed by a CellMove -- the hybrid NPT move).  Snapshot the last accepted state
            // in global memory, displace sub-move by sub-move with the local delta of each (committed at once: the next sub-move
            // sees the moved atoms), deform + recompute if a CellMove closes the chain, test once, restore on reject.
            for (int j = lane; j < R.n; j += 32) {
                gx[j] = R.x(j);
                gy[j] = R.y(j);
                gz[j] = R.z(j);
                if (POT == QMC_POT_EMT) {
                    gs[j] = R.sig(j);
                    ge[j] = R.eat(j);
                }
            }
            const bool exclusive = !(mv.chain & 2);
            uint32_t *disp = R.displaced();
            if (lane < (NS + 31) / 32) disp[lane] = 0u;
            __syncwarp();
            double de_total = 0.0, e_new = E, ncell[3] = {cell[0], cell[1], cell[2]};
            int n_moved = 0, n_drawn = 0;
            bool did_cell = false, cell_ok = true;
            for (int mm = m;; ++mm) {
                const MoveDev &sm = a.mv[mm];
                if ((FEAT & F_CELL) && sm.type == QMC_MOVE_CELL) {  // closes a plain chain (validated by the library)
                    double u_cell = D.o0;  // no displacement drew: op draw 0; else the attempt's first extra draw (block 64)
                    if (n_drawn > 0) {
                        double d0, d1;
                        uniform_pair(key, attempt, grep, 64u, d0, d1);
                        u_cell = d0;
                    }
                    const double s = exp(__dadd_rn(-sm.step, __dmul_rn(sm.step - -sm.step, u_cell)));
                    double scale[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        ncell[d] = (((sm.axes >> d) & 1) ? s : 1.0) * cell[d];
                        scale[d] = ncell[d] / cell[d];
                    }
                    for (int j = lane; j < R.n; j += 32) {
                        R.x(j) = R.x(j) * scale[0];
                        R.y(j) = R.y(j) * scale[1];
                        R.z(j) = R.z(j) * scale[2];
                    }
                    __syncwarp();
                    cell_ok = store_cell(R, ncell, a.b.pbc, pot.cut_pre, lane);
                    if (!cell_ok) status |= QMC_STATUS_CELL_TOO_SMALL;
                    long long fp = 0;
                    e_new = cell_ok ? replica_full_energy<LY>(pot, R, C, lane, fp, status) : E;
                    n_pairs += fp;
                    did_cell = true;
                    break;
                }
                const int *slab = ((FEAT & F_LABELS) && sm.labels) ? sm.labels + (size_t)r * nmax : nullptr;
                const int reps = sm.repeat > 0 ? sm.repeat : 1;
                for (int q = 0; q < reps; ++q) {
                    const int na = (FEAT & F_LABELS) ? count_bits(R.labbits(mm), disp, LY::LABW) : count_available(slab, R.n, disp);  // `disp` stays empty for a plain CompositeMove
                    if (na == 0) continue;  // register_failure: nothing drawn
                    double ul = D.u_label, o0 = D.o0, o1 = D.o1, o2 = D.o2;
                    if (n_drawn > 0) {  // drawing sub-move c >= 1: blocks submove_block(c) and submove_block(c) + 1
                        double d0, d1;
                        uniform_pair(key, attempt, grep, submove_block(n_drawn) + (uint32_t)(lane & 1), d0, d1);
                        ul = __shfl_sync(FULL, d0, 0);
                        o0 = __shfl_sync(FULL, d1, 0);
                        o1 = __shfl_sync(FULL, d0, 1);
                        o2 = __shfl_sync(FULL, d1, 1);
                    }
                    ++n_drawn;
                    const int i = (FEAT & F_LABELS) ? select_bit(R.labbits(mm), disp, LY::LABW, (int)(ul * (double)na))
                                                    : select_available(slab, R.n, disp, (int)(ul * (double)na));
                    double t[3];
                    propose_translation(sm.op, sm.step, o0, o1, o2, t);
                    const double po[3] = {R.x(i), R.y(i), R.z(i)};
                    const double pn[3] = {po[0] + t[0], po[1] + t[1], po[2] + t[2]};
                    const Trial T = local_delta_move<LY>(pot, R, C, lane, lt, R.n, i, po, pn, status);
                    n_pairs += T.pairs;
                    de_total += T.dE;
                    commit_move<LY>(R, lane, T);
                    if (lane == 0) {
                        R.x(i) = pn[0];
                        R.y(i) = pn[1];
                        R.z(i) = pn[2];
                        if (POT == QMC_POT_EMT) {
                            R.sig(i) = T.sig_i;
                            R.eat(i) = T.eat_i;
                        }
                        if (exclusive) disp[i >> 5] |= 1u << (i & 31);
                    }
                    __syncwarp();
                    moved = i;
                    ++n_moved;
                }
                if (!(sm.chain & 1) || mm + 1 >= a.n_moves) break;
            }
            if (n_moved > 0 || did_cell) {
                bool acc;
                if (did_cell) {  // IsobaricCriteria on the whole change (mc/criteria.py:160-172)
                    delta = e_new - E;
                    const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
                    const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
                    acc = cell_ok && metropolis(D.u_accept, x, status);
                    moved = -1;
                } else {
                    delta = de_total;
                    acc = metropolis(D.u_accept, -delta * inv_kT, status);
                }
                if (acc) {
                    E = did_cell ? e_new : E + delta;
                    cell[0] = ncell[0];
                    cell[1] = ncell[1];
                    cell[2] = ncell[2];
                } else {
                    for (int j = lane; j < R.n; j += 32) {
                        R.x(j) = gx[j];
                        R.y(j) = gy[j];
                        R.z(j) = gz[j];
                        if (POT == QMC_POT_EMT) {
                            R.sig(j) = gs[j];
                            R.eat(j) = ge[j];
                        }
                    }
                    if (did_cell) store_cell(R, cell, a.b.pbc, pot.cut_pre, lane);
                }
                __syncwarp();
                accepted = acc ? 1 : 0;
            }
        } else if (mv.type == QMC_MOVE_DISPLACEMENT) {
            const int nu = (FEAT & F_LABELS) ? count_bits(R.labbits(m), nullptr, LY::LABW) : R.n;
            if (nu > 0) {
                const int idx = (int)(D.u_label * (double)nu);
                const int i = (FEAT & F_LABELS) ? select_bit(R.labbits(m), nullptr, LY::LABW, idx) : idx;
                double t[3];
                propose_translation(mv.op, mv.step, D.o0, D.o1, D.o2, t);
                if ((FEAT & F_COMPOSITE) && mv.n_extra_ops > 0) {
                    // CompositeOperation (operations/composite.py:58-73): np.sum of the operations' displacements; the further
                    // operations draw the attempt's extra draws e = 0, 1, ... = block 64 + e / 2, half e & 1
                    for (int xo = 0, e = 0; xo < mv.n_extra_ops; ++xo) {
                        double d0, d1;
                        uniform_pair(key, attempt, grep, 64u + (uint32_t)((e + lane) >> 1), d0, d1);
                        const double mine = ((e + lane) & 1) ? d1 : d0;  // lane j holds extra draw e + j
                        const double x0 = __shfl_sync(FULL, mine, 0), x1 = __shfl_sync(FULL, mine, 1), x2 = __shfl_sync(FULL, mine, 2);
                        double t2[3];
                        propose_translation(mv.extra_op[xo], mv.extra_step[xo], x0, x1, x2, t2);
                        t[0] = __dadd_rn(t[0], t2[0]);
                        t[1] = __dadd_rn(t[1], t2[1]);
                        t[2] = __dadd_rn(t[2], t2[2]);
                        e += mv.extra_op[xo] == QMC_OP_SPHERE ? 2 : 3;
                    }
                }
                const double po[3] = {R.x(i), R.y(i), R.z(i)};
                const double pn[3] = {po[0] + t[0], po[1] + t[1], po[2] + t[2]};
                const Trial T = local_delta_move<LY>(pot, R, C, lane, lt, R.n, i, po, pn, status);
                n_pairs += T.pairs;
                delta = T.dE;
                const bool acc = metropolis(D.u_accept, -delta * inv_kT, status);
                if (acc) {
                    commit_move<LY>(R, lane, T);
                    if (lane == 0) {
                        R.x(i) = pn[0];
                        R.y(i) = pn[1];
                        R.z(i) = pn[2];
                        if (POT == QMC_POT_EMT) {
                            R.sig(i) = T.sig_i;
                            R.eat(i) = T.eat_i;
                        }
                    }
                    E += delta;
                }
                __syncwarp();
                accepted = acc ? 1 : 0;
                moved = i;
            }
        } else if ((FEAT & F_CELL) && mv.type == QMC_MOVE_CELL) {
            // snapshot the last accepted state in global memory, deform in place, recompute everything
            for (int j = lane; j < R.n; j += 32) {
                gx[j] = R.x(j);
                gy[j] = R.y(j);
                gz[j] = R.z(j);
                if (POT == QMC_POT_EMT) {
                    gs[j] = R.sig(j);
                    ge[j] = R.eat(j);
                }
            }
            const double s = exp(__dadd_rn(-mv.step, __dmul_rn(mv.step - -mv.step, D.o0)));
            double ncell[3], scale[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                ncell[d] = (((mv.axes >> d) & 1) ? s : 1.0) * cell[d];  // F @ cell, F = eye * s * mask + eye * ~mask
                scale[d] = ncell[d] / cell[d];                          // solve(old_cell, new_cell), diagonal
            }
            for (int j = lane; j < R.n; j += 32) {
                R.x(j) = R.x(j) * scale[0];
                R.y(j) = R.y(j) * scale[1];
                R.z(j) = R.z(j) * scale[2];
            }
            __syncwarp();
            const bool cell_ok = store_cell(R, ncell, a.b.pbc, pot.cut_pre, lane);
            if (!cell_ok) status |= QMC_STATUS_CELL_TOO_SMALL;
            long long fp = 0;
            const double e_new = cell_ok ? replica_full_energy<LY>(pot, R, C, lane, fp, status) : E;
            n_pairs += fp;
            delta = e_new - E;
            const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
            const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
            const bool acc = cell_ok && metropolis(D.u_accept, x, status);
            if (acc) {
                E = e_new;
                cell[0] = ncell[0];
                cell[1] = ncell[1];
                cell[2] = ncell[2];
            } else {
                for (int j = lane; j < R.n; j += 32) {
                    R.x(j) = gx[j];
                    R.y(j) = gy[j];
                    R.z(j) = gz[j];
                    if (POT == QMC_POT_EMT) {
                        R.sig(j) = gs[j];
                        R.eat(j) = ge[j];
                    }
                }
                store_cell(R, cell, a.b.pbc, pot.cut_pre, lane);
            }
            __syncwarp();
            accepted = acc ? 1 : 0;
        } else if ((FEAT & F_EXCHANGE) && (FEAT & F_COMPOSITE) && mv.type == QMC_MOVE_EXCHANGE && (mv.chain & 8)) {
            // Molecular exchange (moves/exchange.py:88-176 with an `exchange_atoms` of several atoms, or any exchange placed by
            // TranslationRotation): insertion = the molecule's atoms appended one after the other, each with the local delta of a
            // single-atom insertion and committed at once (the next one sees it); deletion = the atoms of the selected label (a
            // run of equal labels) taken out from the top down.  ONE criteria evaluation on the summed energy change, one
            // particle; the snapshot in global memory is restored on reject.
            //
            // repeat = k > 1: CompositeExchangeMove of k single-atom ExchangeMoves holding equal labels (moves/exchange.py:237-316).
            // One coin; insertion = k atoms, each placed with its own Translation draws (sub-move 0: the standard operation slots,
            // sub-move c >= 1: extra draws 3 (c - 1) + {0, 1, 2}); deletion = every sub-move draws a label among those nobody
            // deleted yet in this composite (sub-move 0: u_label, drawing sub-move c >= 1: block 32 + 2 (c - 1), first double) and
            // its whole run goes; particle_delta = +-(sub-moves that acted); an accepted insertion gives the k atoms ONE label.
            const bool is_add = D.u_coin < mv.bias;
            const int k_mol = a.b.n_exchange_atoms > 1 ? a.b.n_exchange_atoms : 1;
            const int k_sub = mv.repeat > 1 ? min(mv.repeat, QMC_MAX_SUBMOVES) : 1;
            const bool comp = k_sub > 1;
            const int k_ins = comp ? k_sub : k_mol;  // atoms appended by an insertion
            const double *tmpl = a.b.n_exchange_atoms >= 1 ? a.b.exchange_mol : a.b.exchange_pos;
            const int n_before = R.n;
            int pd = 0, first = -1, len = 0;
            int n_del = 0, del_ord[QMC_MAX_SUBMOVES];  // composite deletion: ordinals of the chosen runs, ascending
            if (is_add) {
                if (R.n + k_ins > nmax) status |= QMC_STATUS_CAPACITY;
                else pd = comp ? k_sub : 1;
            } else if (comp) {
                const int nu = lab ? run_count(lab, R.n) : 0;
                int c = 0;
                for (int sub = 0; sub < k_sub; ++sub) {
                    const int na = nu - n_del;  // setdiff1d(unique_labels, deleted_labels)
                    if (na <= 0) break;
                    double ul = D.u_label, unused;
                    if (c > 0) uniform_pair(key, attempt, grep, submove_block(c), ul, unused);
                    ++c;
                    int ord = (int)(ul * (double)na);
                    int at = 0;  // the ord-th run that is still available; keep del_ord sorted
                    while (at < n_del && del_ord[at] <= ord) {
                        ++ord;
                        ++at;
                    }
                    for (int q = n_del; q > at; --q) del_ord[q] = del_ord[q - 1];
                    del_ord[at] = ord;
                    ++n_del;
                    if (sub == 0) run_select(lab, R.n, ord, first, len);  // `moved` of the trace: first atom of the first label drawn
                }
                pd = -n_del;
            } else {
                const int nu = lab ? run_count(lab, R.n) : 0;
                if (nu > 0) {
                    run_select(lab, R.n, (int)(D.u_label * (double)nu), first, len);
                    pd = -1;
                }
            }
            const int moved_first = first;
            if (pd != 0) {
                for (int j = lane; j < n_before; j += 32) {  // snapshot of the last accepted state
                    gx[j] = R.x(j);
                    gy[j] = R.y(j);
                    gz[j] = R.z(j);
                    if (POT == QMC_POT_EMT) {
                        gs[j] = R.sig(j);
                        ge[j] = R.eat(j);
                    }
                }
                __syncwarp();
                double de_total = 0.0;
                const double po0[3] = {0.0, 0.0, 0.0};
                if (pd > 0) {
                    // Translation.calculate (operations/displacement.py:170-175): uniform(0,1,(1,3)) @ cell - mean(molecule)
                    double mean[3] = {0.0, 0.0, 0.0}, t[3], A[9], com[3] = {0.0, 0.0, 0.0};
                    for (int q = 0; q < k_mol; ++q)
#pragma unroll
                        for (int d = 0; d < 3; ++d) mean[d] = __dadd_rn(mean[d], tmpl[3 * q + d]);
                    const double f[3] = {D.o0, D.o1, D.o2};
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        t[d] = __dsub_rn(__dmul_rn(f[d], UNIFORM ? a.cell.L[d] : cell[d]), __ddiv_rn(mean[d], (double)k_mol));
                    const bool rot = mv.op == QMC_OP_TRANSLATION_ROTATION;
                    if (rot) {  // + Rotation.calculate (:192-200) of the still untranslated molecule; angles = extra draws 0, 1, 2
                        double u3[3], d3;
                        uniform_pair(key, attempt, grep, 64u, u3[0], u3[1]);
                        uniform_pair(key, attempt, grep, 65u, u3[2], d3);
                        euler_matrix(u3, A);
                        const double m1 = __ddiv_rn(a.b.exchange_mass, (double)k_mol);  // single species
                        double ms = 0.0;
                        for (int q = 0; q < k_mol; ++q) {
#pragma unroll
                            for (int d = 0; d < 3; ++d) com[d] = __dadd_rn(com[d], __dmul_rn(m1, tmpl[3 * q + d]));
                            ms = __dadd_rn(ms, m1);
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d) com[d] = __ddiv_rn(com[d], ms);
                    }
                    for (int q = 0; q < k_ins; ++q) {
                        const int qt = comp ? 0 : q;
                        const double tq[3] = {tmpl[3 * qt], tmpl[3 * qt + 1], tmpl[3 * qt + 2]};
                        double pn[3];
                        if (comp) {  // sub-move q of a composite: its own fractional coordinates
                            double fq[3] = {D.o0, D.o1, D.o2};
                            if (q > 0) {
                                const unsigned e0 = 3u * (unsigned)(q - 1);
                                double lo, hi;
#pragma unroll
                                for (int d = 0; d < 3; ++d) {
                                    uniform_pair(key, attempt, grep, 64u + (e0 + d) / 2u, lo, hi);
                                    fq[d] = ((e0 + d) & 1u) ? hi : lo;
                                }
                            }
#pragma unroll
                            for (int d = 0; d < 3; ++d)
                                pn[d] = __dadd_rn(tq[d], __dsub_rn(__dmul_rn(fq[d], UNIFORM ? a.cell.L[d] : cell[d]), __ddiv_rn(tq[d], 1.0)));
                        } else if (rot) {
                            const double rc[3] = {__dsub_rn(tq[0], com[0]), __dsub_rn(tq[1], com[1]), __dsub_rn(tq[2], com[2])};
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                const double xn = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(A[3 * d], rc[0]), __dmul_rn(A[3 * d + 1], rc[1])), __dmul_rn(A[3 * d + 2], rc[2])), com[d]);
                                pn[d] = __dadd_rn(tq[d], __dadd_rn(t[d], __dsub_rn(xn, tq[d])));
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < 3; ++d) pn[d] = __dadd_rn(tq[d], t[d]);
                        }
                        const int i = R.n;
                        const Trial T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, false, true, po0, pn, status);
                        commit_trial<LY>(R, lane, T);
                        if (lane == 0) {
                            R.x(i) = pn[0];
                            R.y(i) = pn[1];
                            R.z(i) = pn[2];
                            if (POT == QMC_POT_EMT) {
                                R.sig(i) = T.sig_i;
                                R.eat(i) = T.eat_i;
                            }
                        }
                        R.n += 1;
                        n_pairs += T.pairs;
                        de_total += T.dE;
                        __syncwarp();
                    }
                    moved = n_before;
                } else {
                  for (int run = comp ? n_del - 1 : 0; run >= 0; --run) {  // composite: the chosen runs, highest first
                    if (comp) run_select(lab, n_before, del_ord[run], first, len);
                    for (int q = len - 1; q >= 0; --q) {  // top down: the indices below stay valid
                        const int i = first + q;
                        const double po[3] = {R.x(i), R.y(i), R.z(i)};
                        const Trial T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, true, false, po, po0, status);
                        commit_trial<LY>(R, lane, T);
                        shift_down(&R.x(0), i, R.n);
                        shift_down(&R.y(0), i, R.n);
                        shift_down(&R.z(0), i, R.n);
                        if (POT == QMC_POT_EMT) {
                            shift_down(&R.sig(0), i, R.n);
                            shift_down(&R.eat(0), i, R.n);
                        }
                        R.n -= 1;
                        n_pairs += T.pairs;
                        de_total += T.dE;
                        __syncwarp();
                    }
                  }
                    moved = moved_first;
                }
                delta = de_total;
                bool ov;
                const double prob = gc_probability(delta, pd, n_ex, a.b.accessible_volume, a.b.exchange_mass, temperature,
                                                   a.b.chemical_potential, ov);
                if (ov) status |= QMC_STATUS_EXP_OVERFLOW;
                const bool acc = D.u_accept < prob;
                if (acc) {
                    // GrandCanonical.save_state (mc/gcmc.py:207-214): every registered move re-labels
                    auto first_use = [&](int q) { return table_first_use(a.mv, q); };
                    int *labw = mv.labels + (size_t)r * nmax;
                    if (comp && pd < 0) {
                        // the runs were located in the exchange move's labels; every table holds the same runs (run-coded tables are
                        // re-labelled together from the start), so the positions are valid for all of them
                        int n_cur = n_before;
                        for (int run = n_del - 1; run >= 0; --run) {
                            run_select(lab, n_before, del_ord[run], first, len);  // `lab` itself is edited last (below)
                            for (int q = 0; q < a.n_moves; ++q)
                                if (a.mv[q].labels && a.mv[q].labels != mv.labels && first_use(q))
                                    labels_molecule_changed(a.mv[q].labels + (size_t)r * nmax, n_cur, 0, first, len, a.mv[q].default_label);
                            n_cur -= len;
                        }
                        n_cur = n_before;
                        for (int run = n_del - 1; run >= 0; --run) {
                            run_select(lab, n_cur, del_ord[run], first, len);  // lower ordinals are untouched by the deletions above them
                            labels_molecule_changed(labw, n_cur, 0, first, len, mv.default_label);
                            n_cur -= len;
                        }
                    } else
                    for (int q = 0; q < a.n_moves; ++q)
                        if (a.mv[q].labels && first_use(q))
                            labels_molecule_changed(a.mv[q].labels + (size_t)r * nmax, n_before, pd > 0 ? k_ins : 0, first, pd < 0 ? len : 0,
                                                    a.mv[q].default_label);
                    n_ex += pd;
                    E += delta;
                    __syncwarp();
                    if (FEAT & F_LABELS) rebuild_movable(R, a.mv, a.n_moves, r, nmax, LY::LABW);
                } else {
                    R.n = n_before;
                    for (int j = lane; j < n_before; j += 32) {
                        R.x(j) = gx[j];
                        R.y(j) = gy[j];
                        R.z(j) = gz[j];
                        if (POT == QMC_POT_EMT) {
                            R.sig(j) = gs[j];
                            R.eat(j) = ge[j];
                        }
                    }
                }
                __syncwarp();
                accepted = acc ? 1 : 0;
            }
        } else if ((FEAT & F_EXCHANGE) && mv.type == QMC_MOVE_EXCHANGE) {
            const bool is_add = D.u_coin < mv.bias;
            int pd = 0, i = -1;
            Trial T;
            double pn[3] = {0, 0, 0}, po[3] = {0, 0, 0};
            if (is_add) {
                if (R.n + 1 > nmax) {
                    status |= QMC_STATUS_CAPACITY;
                } else {
                    // Translation: uniform(0,1,(1,3)) @ cell - mean(positions[moving]); diagonal cell
                    const double f[3] = {D.o0, D.o1, D.o2};
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double fc = __dmul_rn(f[d], UNIFORM ? a.cell.L[d] : cell[d]);
                        pn[d] = __dadd_rn(a.b.exchange_pos[d], __dsub_rn(fc, a.b.exchange_pos[d]));
                    }
                    i = R.n;
                    T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, false, true, po, pn, status);
                    pd = 1;
                }
            } else {
                const int nu = (FEAT & F_LABELS) ? count_bits(R.labbits(m), nullptr, LY::LABW) : R.n;
                if (nu > 0) {
                    const int idx = (int)(D.u_label * (double)nu);
                    i = (FEAT & F_LABELS) ? select_bit(R.labbits(m), nullptr, LY::LABW, idx) : idx;
                    po[0] = R.x(i);
                    po[1] = R.y(i);
                    po[2] = R.z(i);
                    T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, true, false, po, pn, status);
                    pd = -1;
                }
            }
            if (pd != 0) {
                n_pairs += T.pairs;
                delta = T.dE;
                bool ov;
                const double prob = gc_probability(delta, pd, n_ex, a.b.accessible_volume, a.b.exchange_mass, temperature,
                                                   a.b.chemical_potential, ov);
                if (ov) status |= QMC_STATUS_EXP_OVERFLOW;
                const bool acc = D.u_accept < prob;
                if (acc) {
                    commit_trial<LY>(R, lane, T);
                    // GrandCanonical.save_state (mc/gcmc.py:207-214): every registered move re-labels
                    for (int q = 0; q < a.n_moves; ++q)
                        if (a.mv[q].labels && table_first_use(a.mv, q))
                            labels_changed(a.mv[q].labels + (size_t)r * nmax, R.n, pd > 0, pd < 0 ? i : -1, a.mv[q].default_label);
                    if (pd > 0) {
                        if (lane == 0) {
                            R.x(i) = pn[0];
                            R.y(i) = pn[1];
                            R.z(i) = pn[2];
                            if (POT == QMC_POT_EMT) {
                                R.sig(i) = T.sig_i;
                                R.eat(i) = T.eat_i;
                            }
                        }
                        R.n += 1;
                    } else {
                        shift_down(&R.x(0), i, R.n);
                        shift_down(&R.y(0), i, R.n);
                        shift_down(&R.z(0), i, R.n);
                        if (POT == QMC_POT_EMT) {
                            shift_down(&R.sig(0), i, R.n);
                            shift_down(&R.eat(0), i, R.n);
                        }
                        R.n -= 1;
                    }
                    n_ex += pd;
                    E += delta;
                    __syncwarp();
                    if (FEAT & F_LABELS) rebuild_movable(R, a.mv, a.n_moves, r, nmax, LY::LABW);  // the tables changed
                }
                __syncwarp();
                accepted = acc ? 1 : 0;
                moved = i;
            }
        }

        if (lane == 0) {
            uint32_t *cnt = R.counts() + m * 3;  // flushed to the global counters when the replica is stored
            cnt[0] += 1u;
            if (accepted == 1) cnt[1] += 1u;
            if (accepted < 0) cnt[2] += 1u;
            if (a.has_trace) {
                const size_t ti = (size_t)r * a.n_attempts + k;
                if (a.tr.move) a.tr.move[ti] = (int8_t)m;
                if (a.tr.accepted) a.tr.accepted[ti] = (int8_t)accepted;
                if (a.tr.energy) a.tr.energy[ti] = E;
                if (a.tr.delta) a.tr.delta[ti] = delta;
                if (a.tr.moved) a.tr.moved[ti] = moved;
            }
        }
    }

    // ---- store the replica ----
    __syncwarp();
    for (int j = lane; j < R.n; j += 32) {
        gx[j] = R.x(j);
        gy[j] = R.y(j);
        gz[j] = R.z(j);
        if (POT == QMC_POT_EMT) {
            gs[j] = R.sig(j);
            ge[j] = R.eat(j);
        }
    }
    if (lane < QMC_MAX_MOVES * 3) {
        const uint32_t c = R.counts()[lane];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(a.b.counters) + (size_t)r * QMC_MAX_MOVES * 3 + lane, (unsigned long long)c);
    }
    if (lane == 0) {
        a.b.n_atoms[r] = R.n;
        a.b.energy[r] = E;
        if (FEAT & F_CELL) {
            a.b.cell[(size_t)r * 9 + 0] = cell[0];
            a.b.cell[(size_t)r * 9 + 4] = cell[1];
            a.b.cell[(size_t)r * 9 + 8] = cell[2];
        }
        if (FEAT & F_EXCHANGE) a.b.n_exchange[r] = n_ex;
        if (status) atomicOr(a.b.status + r, status);
    }
    if (a.b.pair_evals) {
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        if (lane == 0) a.b.pair_evals[r] += n_pairs;
    }
}

// validate_simulation: full energy of every replica, caches primed (per-replica cells always)
template <int POT, int NS>
__global__ void __launch_bounds__(Layout<POT, NS>::WARPS * 32) energy_kernel(const __grid_constant__ SweepArgs a, long long *pair_count) {
    using LY = Layout<POT, NS>;
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * LY::WARPS + warp;
    if (r >= a.b.n_replicas) return;
    const int nmax = a.b.n_max;
    Rep<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    const CellView<LY, false> C{&a.cell, R.b + LY::CELLC};
    R.n = a.b.n_atoms[r];
    const double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? gx[j] : 0.0;
        R.y(j) = in ? gy[j] : 0.0;
        R.z(j) = in ? gz[j] : 0.0;
    }
    const double cell[3] = {a.b.cell[(size_t)r * 9 + 0], a.b.cell[(size_t)r * 9 + 4], a.b.cell[(size_t)r * 9 + 8]};
    const bool ok = store_cell(R, cell, a.b.pbc, a.pot.cut_pre, lane);
    long long pairs = 0;
    int status = 0;
    const double E = ok ? replica_full_energy<LY>(a.pot, R, C, lane, pairs, status, a.atom_energies ? a.atom_energies + (size_t)r * nmax : nullptr)
                        : __longlong_as_double(0x7ff8000000000000LL);
    if (POT == QMC_POT_EMT) {
        for (int j = lane; j < R.n; j += 32) {
            a.b.sigma1[(size_t)r * nmax + j] = R.sig(j);
            a.b.eatom[(size_t)r * nmax + j] = R.eat(j);
        }
    }
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(FULL, pairs, o);
    if (lane == 0) {
        a.b.energy[r] = E;
        if (pair_count) pair_count[r] = pairs;
        if (!ok) status |= QMC_STATUS_CELL_TOO_SMALL;
        if (status) atomicOr(a.b.status + r, status);
    }
}

}  // namespace qmc
