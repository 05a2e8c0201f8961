// Canonical fast path with a PAIR-TERM CACHE in HBM (EMT, displacement moves only, one fixed cell, labels = arange).
//
// The plain sweep kernel evaluates, per attempt, the pair functions of the moved atom at its OLD position (to subtract
// them) and at its NEW position -- ~78 + 78 evaluations of 3 exp + sqrt + 1/x each.  The old-position terms were all
// computed before: they are exactly the terms of the last accepted position.  180 GB of HBM buys them back:
//
//     cache[r][i][j] = (sum over images of the sigma_1 term, same for the sigma_2 term) of the pair (i, j)   16 B
//
// i.e. 108 x 108 x 16 B = 187 KB per replica, 764 MB for 4096 replicas.  Per attempt the kernel then
//   1. prefetches row i (1.7 KB) into L2 while it draws and proposes,
//   2. scans and evaluates the NEW position only (half the scan, 3 list passes instead of 5),
//   3. owner pass: lane j reads cache[i][j], forms d(sigma_1)_j and d(sigma_2)_j, parks the new row in a per-replica
//      `pending` buffer (L2-resident) and compacts the touched atoms,
//   4. embedding of the touched atoms (as before),
//   5. on acceptance copies `pending` into row i and column i of the cache (fire-and-forget stores).
// The cached old terms are bit-identical to a recomputation (same arithmetic on the same positions), so the result
// differs from the plain kernel only by summation order.
#pragma once

#include "sweep_kernel.cuh"

namespace qmc {

constexpr int XCAPC = 32;

template <int NS>
struct LayoutC {
    static constexpr int CAPE = round_up(NS + XCAPC, 4);  // new-side hits: <= NS - 1 nearest images + XCAPC second images
    static constexpr int X = 0, Y = NS, Z = 2 * NS, SIG = 3 * NS, EAT = 4 * NS;
    static constexpr int LR = 5 * NS;          // [CAPE] r^2 -> sigma_1 term; later trial embedding energies (compact index)
    static constexpr int LS2 = LR + CAPE;      // [CAPE] sigma_2 term; later trial sigma_1 (compact index)
    static constexpr int TAIL = LS2 + CAPE;    // [NS] work-item queue during the scan
    static constexpr int PP_B = (TAIL + NS) * 8;
    static constexpr int L2_B = PP_B + NS * 2;
    static constexpr int XJ_B = L2_B + NS * 2;
    static constexpr int BYTES = round_up(XJ_B + XCAPC * 2, 16);
    static constexpr int fit = (SMEM_PER_SM - 1024) / BYTES;
    static constexpr int WARPS = fit > 28 ? 28 : (fit < 1 ? 1 : fit);  // one wide CTA per SM (see Layout in replica_warp.cuh)
    static constexpr int MINB = 1;
    static constexpr bool FITS = WARPS * BYTES <= SMEM_PER_CTA_MAX;
};

template <class LY>
struct RepC {
    double *b;
    int n;
    __device__ __forceinline__ double &x(int j) const { return b[LY::X + j]; }
    __device__ __forceinline__ double &y(int j) const { return b[LY::Y + j]; }
    __device__ __forceinline__ double &z(int j) const { return b[LY::Z + j]; }
    __device__ __forceinline__ double &sig(int j) const { return b[LY::SIG + j]; }
    __device__ __forceinline__ double &eat(int j) const { return b[LY::EAT + j]; }
    __device__ __forceinline__ double &lr(int e) const { return b[LY::LR + e]; }
    __device__ __forceinline__ double &ls2(int e) const { return b[LY::LS2 + e]; }
    __device__ __forceinline__ uint32_t *queue() const { return reinterpret_cast<uint32_t *>(b + LY::TAIL); }
    __device__ __forceinline__ uint16_t &pp(int j) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::PP_B)[j]; }
    __device__ __forceinline__ uint16_t &l2(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::L2_B)[e]; }
    __device__ __forceinline__ uint16_t &xj(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::XJ_B)[e]; }
};

// Scan the atoms j < n (j != i) against position p and evaluate the pair functions; afterwards lr[pp[j]] / ls2[pp[j]] hold
// the image-summed sigma_1 / sigma_2 terms of every atom j inside the cutoff (pp[j] = NOPOS outside).
template <class LY>
__device__ __forceinline__ int new_side_terms(const PotDev &pot, const RepC<LY> &R, const CellArgs &cell, const int lane,
                                              const unsigned lt, int i_excl, const double p[3], int &status) {
    const bool p0 = cell.per[0] != 0, p1 = cell.per[1] != 0, p2 = cell.per[2] != 0;
    const Axis A0{cell.L[0], cell.invL[0], cell.thr[0]}, A1{cell.L[1], cell.invL[1], cell.thr[1]}, A2{cell.L[2], cell.invL[2], cell.thr[2]};
    uint32_t *queue = R.queue();
    int cnt = 0, nq = 0;
    for (int base = 0; base < R.n; base += 32) {
        const int j = base + lane;
        const bool valid = j < R.n && j != i_excl;
        const int jc = valid ? j : 0;
        unsigned fl = 0;
        const double dx = axis_nearest(p[0], R.x(jc), A0, p0, fl, 1u);
        const double dy = axis_nearest(p[1], R.y(jc), A1, p1, fl, 2u);
        const double dz = axis_nearest(p[2], R.z(jc), A2, p2, fl, 4u);
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const bool in = valid && r2 < pot.cut_pre2;
        const unsigned b = __ballot_sync(FULL, in);
        const int pos = cnt + __popc(b & lt);  // < NS <= CAPE
        if (in) R.lr(pos) = r2;
        if (j < R.n) R.pp(j) = in ? (uint16_t)pos : (uint16_t)NOPOS;
        cnt += __popc(b);
        const bool ex = in && fl != 0u;
        const unsigned bq = __ballot_sync(FULL, ex);
        if (ex) queue[nq + __popc(bq & lt)] = (unsigned)j | (fl << 16);  // <= NS items fit in the tail
        nq += __popc(bq);
    }
    const int xstart = cnt;
    if (nq > 0) {
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const bool has = q < nq;
            const unsigned item = has ? queue[q] : 0u;
            const int j = item & 0xFFFFu;
            const unsigned fl = item >> 16;
            unsigned dummy = 0;
            const double dx0 = axis_nearest(p[0], R.x(j), A0, p0, dummy, 1u);
            const double dy0 = axis_nearest(p[1], R.y(j), A1, p1, dummy, 2u);
            const double dz0 = axis_nearest(p[2], R.z(j), A2, p2, dummy, 4u);
            const double dx1 = dx0 - copysign(A0.L, dx0), dy1 = dy0 - copysign(A1.L, dy0), dz1 = dz0 - copysign(A2.L, dz0);
            unsigned c = fl & (0u - fl);
            bool more = has;
            while (__any_sync(FULL, more)) {
                const double dx = (c & 1u) ? dx1 : dx0, dy = (c & 2u) ? dy1 : dy0, dz = (c & 4u) ? dz1 : dz0;
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = more && r2 < pot.cut_pre2;
                const unsigned b = __ballot_sync(FULL, in);
                const int pos = min(cnt + __popc(b & lt), LY::CAPE - 1);
                if (in) {
                    R.lr(pos) = r2;
                    R.xj(min(pos - xstart, XCAPC - 1)) = (uint16_t)j;
                }
                cnt += __popc(b);
                if (more) {
                    if (c == fl) more = false;
                    else c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
        }
    }
    if (cnt - xstart > XCAPC) {
        status |= QMC_STATUS_LIST_OVERFLOW;
        cnt = xstart + XCAPC;
    }
    __syncwarp();
    // pair functions, 32 entries per pass; second images add both terms onto the nearest-image slot of their atom
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e < cnt;
        const double r2 = act ? R.lr(e) : 1.0e6;
        double s1, s2;
        emt_pair(pot, r2, s1, s2);
        if (e0 + 32 <= xstart) {
            R.lr(e) = s1;
            R.ls2(e) = s2;
        } else {
            const bool is_x = act && e >= xstart;
            if (act && !is_x) {
                R.lr(e) = s1;
                R.ls2(e) = s2;
            }
            __syncwarp();
            unsigned tgt = 0x10000u + lane;
            if (is_x) tgt = R.pp(R.xj(e - xstart));
            const unsigned grp = __match_any_sync(FULL, tgt);
            const int maxc = __reduce_max_sync(FULL, __popc(grp));
            double a1 = 0.0, a2 = 0.0;
            unsigned rem = grp;
            for (int t = 0; t < maxc; ++t) {
                const int src = rem ? (__ffs(rem) - 1) : lane;
                const double v1 = __shfl_sync(FULL, s1, src), v2 = __shfl_sync(FULL, s2, src);
                if (rem) {
                    a1 += v1;
                    a2 += v2;
                    rem &= rem - 1;
                }
            }
            if (is_x && lane == (__ffs(grp) - 1)) {
                R.lr(tgt) += a1;
                R.ls2(tgt) += a2;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    return cnt;
}

template <int NS>
__global__ void __launch_bounds__(LayoutC<NS>::WARPS * 32, LayoutC<NS>::MINB) sweep_cached_kernel(const __grid_constant__ SweepArgs a) {
    using LY = LayoutC<NS>;
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int r = blockIdx.x * LY::WARPS + warp;
    if (r >= a.b.n_replicas) return;
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    RepC<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    R.n = a.b.n_atoms[r];
    double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    double *gs = a.b.sigma1 + (size_t)r * nmax, *ge = a.b.eatom + (size_t)r * nmax;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? gx[j] : 0.0;
        R.y(j) = in ? gy[j] : 0.0;
        R.z(j) = in ? gz[j] : 0.0;
        R.sig(j) = in ? gs[j] : 0.0;
        R.eat(j) = in ? ge[j] : 0.0;
    }
    double2 *cache = reinterpret_cast<double2 *>(a.b.pair_cache) + (size_t)r * nmax * nmax;
    double2 *pending = reinterpret_cast<double2 *>(a.b.pair_pending) + (size_t)r * nmax;
    int status = 0;
    double E = a.b.energy[r];
    const double kT = a.b.temperature[r] * KB;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    const MoveDev &mv = a.mv[0];
    long long n_pairs = 0;
    __syncwarp();

    for (int k = 0; k < a.n_attempts; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        const Draws D = draw_attempt(key, attempt, grep, lane, false);
        int accepted = -1, moved = -1;
        double delta = __longlong_as_double(0x7ff8000000000000LL);
        if (R.n > 0) {
            const int i = (int)(D.u_label * (double)R.n);
            double2 *row = cache + (size_t)i * nmax;
            // pull the row towards L2 now; it is consumed after the scan and the pair passes
            if (lane * 8 < R.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + lane * 8));
            double t[3];
            if (mv.op == QMC_OP_BALL) {
                const double rr = mv.step * D.o0, phi = TWO_PI * D.o1, c = -1.0 + 2.0 * D.o2;
                const double s = sqrt_fast(fmax(__dsub_rn(1.0, __dmul_rn(c, c)), 1e-300));
                double sp, cp;
                sincos_2pi(phi, sp, cp);
                t[0] = __dmul_rn(__dmul_rn(rr, s), cp);
                t[1] = __dmul_rn(__dmul_rn(rr, s), sp);
                t[2] = __dmul_rn(rr, c);
            } else if (mv.op == QMC_OP_SPHERE) {
                const double phi = TWO_PI * D.o0, c = -1.0 + 2.0 * D.o1;
                const double s = sqrt_fast(fmax(__dsub_rn(1.0, __dmul_rn(c, c)), 1e-300));
                double sp, cp;
                sincos_2pi(phi, sp, cp);
                t[0] = __dmul_rn(mv.step, __dmul_rn(s, cp));
                t[1] = __dmul_rn(mv.step, __dmul_rn(s, sp));
                t[2] = __dmul_rn(mv.step, c);
            } else {
                const double w = mv.step - -mv.step;
                t[0] = __dadd_rn(-mv.step, __dmul_rn(w, D.o0));
                t[1] = __dadd_rn(-mv.step, __dmul_rn(w, D.o1));
                t[2] = __dadd_rn(-mv.step, __dmul_rn(w, D.o2));
            }
            const double pn[3] = {R.x(i) + t[0], R.y(i) + t[1], R.z(i) + t[2]};
            const int cnt = new_side_terms<LY>(pot, R, a.cell, lane, lt, i, pn, status);
            n_pairs += cnt;  // warp-uniform count; reduced differently below
            // ---- owner pass: old terms from the cache, new terms from the list ----
            // all row loads are issued before the first use: one exposed memory latency per attempt instead of one per tile
            constexpr int TILES = (NS + 31) / 32;
            double2 old_terms[TILES];
#pragma unroll
            for (int tq = 0; tq < TILES; ++tq) {
                const int j = tq * 32 + lane;
                old_terms[tq] = (j < R.n && j != i) ? row[j] : make_double2(0.0, 0.0);
            }
            double d2 = 0.0, s1new = 0.0;
            int n2 = 0;
#pragma unroll
            for (int tq = 0; tq < TILES; ++tq) {
                const int j = tq * 32 + lane;
                if (tq * 32 < R.n) {  // warp-uniform
                    const bool valid = j < R.n && j != i;
                    const double2 o = old_terms[tq];
                    double2 nw = make_double2(0.0, 0.0);
                    if (valid) {
                        const unsigned pos = R.pp(j);
                        if (pos != NOPOS) nw = make_double2(R.lr(pos), R.ls2(pos));
                    }
                    if (j < R.n) pending[j] = nw;
                    d2 += nw.y - o.y;
                    s1new += nw.x;
                    const bool touched = valid && (o.x != 0.0 || nw.x != 0.0);
                    const unsigned b2 = __ballot_sync(FULL, touched);
                    const int kk = n2 + __popc(b2 & lt);
                    if (touched) {
                        R.l2(kk) = (uint16_t)j;
                        R.b[LY::TAIL + kk] = nw.x - o.x;  // d(sigma_1), compact index
                    }
                    n2 += __popc(b2);
                }
            }
            __syncwarp();
            const double vs = warp_sum(d2);
            const double sig_i = warp_sum(s1new);
            // ---- embedding of the touched atoms (+ the moved atom as one more work item) ----
            double part = 0.0, e_i_lane = 0.0;
            const int total = n2 + 1;
            for (int e = lane; e < total; e += 32) {
                const bool nb = e < n2;
                const int j = nb ? R.l2(e) : i;
                const double sn = nb ? (R.sig(j) + R.b[LY::TAIL + e]) : sig_i;
                const double en = emt_embed(pot, sn);
                part += en - R.eat(j);
                if (nb) {
                    R.lr(e) = en;   // the list is dead after the owner pass: trial energies ...
                    R.ls2(e) = sn;  // ... and trial sigma_1, by compact index
                } else {
                    e_i_lane = en;
                }
            }
            const double eat_i = __shfl_sync(FULL, e_i_lane, n2 & 31);
            delta = warp_sum(part) + 2.0 * pot.neghalfv0overgamma2 * vs;
            __syncwarp();
            const bool acc = metropolis(D.u_accept, -delta / kT, status);
            if (acc) {
                for (int e = lane; e < n2; e += 32) {
                    const int j = R.l2(e);
                    R.sig(j) = R.ls2(e);
                    R.eat(j) = R.lr(e);
                }
                if (lane == 0) {
                    R.x(i) = pn[0];
                    R.y(i) = pn[1];
                    R.z(i) = pn[2];
                    R.sig(i) = sig_i;
                    R.eat(i) = eat_i;
                }
                // row i and column i of the cache take the parked new terms
                double2 parked[TILES];
#pragma unroll
                for (int tq = 0; tq < TILES; ++tq) {
                    const int j = tq * 32 + lane;
                    parked[tq] = j < R.n ? pending[j] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int tq = 0; tq < TILES; ++tq) {
                    const int j = tq * 32 + lane;
                    if (j < R.n) {
                        row[j] = parked[tq];
                        cache[(size_t)j * nmax + i] = parked[tq];
                    }
                }
                E += delta;
            }
            __syncwarp();
            accepted = acc ? 1 : 0;
            moved = i;
        }
        if (lane == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(a.b.counters) + ((size_t)r * QMC_MAX_MOVES) * 3;
            atomicAdd(cnt + 0, 1ull);
            if (accepted == 1) atomicAdd(cnt + 1, 1ull);
            if (accepted < 0) atomicAdd(cnt + 2, 1ull);
            if (a.has_trace) {
                const size_t ti = (size_t)r * a.n_attempts + k;
                if (a.tr.move) a.tr.move[ti] = 0;
                if (a.tr.accepted) a.tr.accepted[ti] = (int8_t)accepted;
                if (a.tr.energy) a.tr.energy[ti] = E;
                if (a.tr.delta) a.tr.delta[ti] = delta;
                if (a.tr.moved) a.tr.moved[ti] = moved;
            }
        }
    }

    __syncwarp();
    for (int j = lane; j < R.n; j += 32) {
        gx[j] = R.x(j);
        gy[j] = R.y(j);
        gz[j] = R.z(j);
        gs[j] = R.sig(j);
        ge[j] = R.eat(j);
    }
    if (lane == 0) {
        a.b.energy[r] = E;
        if (status) atomicOr(a.b.status + r, status);
        if (a.b.pair_evals) a.b.pair_evals[r] += n_pairs;
    }
}

// Full energy of every replica AND the pair-term cache (validate_simulation for the cached path).
template <int NS>
__global__ void __launch_bounds__(LayoutC<NS>::WARPS * 32) energy_cached_kernel(const __grid_constant__ SweepArgs a, long long *pair_count) {
    using LY = LayoutC<NS>;
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int r = blockIdx.x * LY::WARPS + warp;
    if (r >= a.b.n_replicas) return;
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    RepC<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    R.n = a.b.n_atoms[r];
    const double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? gx[j] : 0.0;
        R.y(j) = in ? gy[j] : 0.0;
        R.z(j) = in ? gz[j] : 0.0;
    }
    __syncwarp();
    double2 *cache = reinterpret_cast<double2 *>(a.b.pair_cache) + (size_t)r * nmax * nmax;
    int status = 0;
    long long pairs = 0;
    for (int i = 0; i < R.n; ++i) {
        const double p[3] = {R.x(i), R.y(i), R.z(i)};
        pairs += new_side_terms<LY>(pot, R, a.cell, lane, lt, i, p, status);
        double s1 = 0.0, s2 = 0.0;
        for (int base = 0; base < nmax; base += 32) {
            const int j = base + lane;
            double2 nw = make_double2(0.0, 0.0);
            if (j < R.n && j != i) {
                const unsigned pos = R.pp(j);
                if (pos != NOPOS) nw = make_double2(R.lr(pos), R.ls2(pos));
            }
            if (j < nmax) cache[(size_t)i * nmax + j] = nw;
            s1 += nw.x;
            s2 += nw.y;
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        __syncwarp();
        if (lane == 0) {
            R.sig(i) = s1;
            R.eat(i) = s2;  // parked until the embedding pass
        }
        __syncwarp();
    }
    double epart = 0.0;
    for (int j = lane; j < R.n; j += 32) {
        const double ea = emt_embed(pot, R.sig(j));
        epart += ea + pot.neghalfv0overgamma2 * R.eat(j);
        R.eat(j) = ea;
        a.b.sigma1[(size_t)r * nmax + j] = R.sig(j);
        a.b.eatom[(size_t)r * nmax + j] = ea;
    }
    const double E = warp_sum(epart);
    if (lane == 0) {
        a.b.energy[r] = E;
        if (pair_count) pair_count[r] = pairs;
        if (status) atomicOr(a.b.status + r, status);
    }
}

}  // namespace qmc
