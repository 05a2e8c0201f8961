// CTA-per-replica variant of the sweep kernel: BW warps cooperate on ONE replica.
//
// One warp per replica (sweep_kernel.cuh) fills a B200 only when the batch holds >= ~4000 replicas; batches such as
// config C3 (1024 replicas of 500 atoms per GPU) leave 7 warps per SM and run latency-bound.  Here the atom tiles of a
// replica are dealt round-robin to BW warps: every warp scans its own tiles, compacts its hits into its OWN list
// region, evaluates them, and re-evaluates the embedding energy of the touched atoms it owns -- all warp-local, with
// the same helpers and the same arithmetic as the warp kernel.  Only three block barriers per attempt remain: the sums
// (sigma_2, sigma_1 of the moved atom), the energy delta, and the commit.  All cross-warp sums run in fixed warp order,
// so results are bit-reproducible (and, summation order aside, equal to the warp kernel's).
#pragma once

#include "sweep_kernel.cuh"

namespace qmc {

// barrier over the BW warps of ONE replica (several replicas share a CTA): named barrier 1 + replica slot
__device__ __forceinline__ void rep_sync() {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 7)), "r"(128) : "memory");
}

constexpr int BW = 4;      // warps per replica
constexpr int XCAPB = 32;  // second-image entries per warp region

template <int POT, int NS>
struct LayoutB {
    static constexpr bool EMT = POT == QMC_POT_EMT;
    static constexpr int TILES = (NS + 31) / 32;
    static constexpr int NSW = ((TILES + BW - 1) / BW) * 32;  // atoms owned by one warp (tiles dealt round-robin)
    static constexpr int CAPE = EMT ? 112 : 176;              // list entries per warp region; a full row (78 images for EMT at Cu
                                                              // density, ~125 for a Lennard-Jones solid with rc = 3 sigma) must fit, too
    static constexpr int TAILN = NSW + 8;                     // work-item queue / trial energies of the owned atoms
    static constexpr int X = 0, Y = NS, Z = 2 * NS, SIG = 3 * NS, EAT = 4 * NS;
    static constexpr int CELLC = EMT ? 5 * NS : 3 * NS;  // 10 doubles
    static constexpr int RED = CELLC + 10;               // BW * 4 doubles
    static constexpr int PP_B = (RED + BW * 4) * 8;      // bytes
    static constexpr int DRAW_B = round_up(PP_B + (EMT ? NS * 4 : 0), 16);  // the draws of DRAW_AHEAD attempts, shared by the BW warps
    static constexpr int REG0_B = DRAW_B + DRAW_AHEAD * 64;
    // one warp region: lr [CAPE + TAILN] doubles, l2 [NSW + 8] u16, xj [XCAPB] u16
    static constexpr int L2_OFF = (CAPE + TAILN) * 8;
    static constexpr int XJ_OFF = L2_OFF + (EMT ? (NSW + 8) * 2 : 0);
    static constexpr int REG_B = round_up(XJ_OFF + (EMT ? XCAPB * 2 : 0), 16);  // list regions are 16-byte aligned (paired slots)
    static constexpr int BYTES = round_up(REG0_B + BW * REG_B, 16);
    // one wide CTA per SM holding REPS replicas of BW warps each (the warps of one CTA progress evenly, see Layout)
    static constexpr int fit = (SMEM_PER_SM - 1024) / BYTES;
    static constexpr int REPS = fit > 7 ? 7 : (fit < 1 ? 1 : fit);  // <= 7: 28 warps keep the 72-register budget
    static constexpr int MINB = 1;
    static constexpr bool FITS = BYTES <= SMEM_PER_CTA_MAX && fit > 0;
};

template <class LY>
struct RepB {
    static constexpr bool EMT = LY::EMT;
    static constexpr int CAPE = LY::CAPE;
    static constexpr int XCAPN = XCAPB;
    static constexpr int QCAP = 2 * LY::TAILN;  // work items the tail can hold
    double *b;   // replica base in shared memory
    double *wb;  // this warp's list region
    int n;
    __device__ __forceinline__ double &x(int j) const { return b[LY::X + j]; }
    __device__ __forceinline__ double &y(int j) const { return b[LY::Y + j]; }
    __device__ __forceinline__ double &z(int j) const { return b[LY::Z + j]; }
    __device__ __forceinline__ double &sig(int j) const { return b[LY::SIG + j]; }
    __device__ __forceinline__ double &eat(int j) const { return b[LY::EAT + j]; }
    __device__ __forceinline__ double &cellc(int k) const { return b[LY::CELLC + k]; }
    __device__ __forceinline__ double &red(int k) const { return b[LY::RED + k]; }
    __device__ __forceinline__ uint32_t &pp(int j) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::PP_B)[j]; }
    __device__ __forceinline__ double2 *draws() const { return reinterpret_cast<double2 *>(reinterpret_cast<char *>(b) + LY::DRAW_B); }
    __device__ __forceinline__ double &lr(int e) const { return wb[e]; }
    __device__ __forceinline__ double &tail(int e) const { return wb[LY::CAPE + e]; }
    __device__ __forceinline__ uint32_t *queue() const { return reinterpret_cast<uint32_t *>(wb + LY::CAPE); }
    __device__ __forceinline__ uint16_t &l2(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(wb) + LY::L2_OFF)[e]; }
    __device__ __forceinline__ uint16_t &xj(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(wb) + LY::XJ_OFF)[e]; }
};

// sum of one value per warp, in warp order; every thread gets the same result.  Slot k of red[warp][4].
template <class RV>
__device__ __forceinline__ void block_put(const RV &R, int warp, int lane, int k, double v) {
    if (lane == 0) R.red(warp * 4 + k) = v;
}
template <class RV>
__device__ __forceinline__ double block_get(const RV &R, int k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < BW; ++w) s += R.red(w * 4 + k);
    return s;
}

template <class LY>
__device__ __forceinline__ bool store_cell_block(const RepB<LY> &R, const double diag[3], const int pbc[3], double cut_pre) {
    CellArgs c;
    const bool ok = cell_constants(diag, pbc, cut_pre, c);
    rep_sync();
    if ((threadIdx.x & 127) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.cellc(k) = c.L[k];
            R.cellc(3 + k) = c.invL[k];
            R.cellc(6 + k) = c.thr[k];
        }
    }
    rep_sync();
    return ok;
}

// full energy with the rows dealt to the warps; fills sig / eat; result on every thread
template <class LY, class CELL>
__device__ inline double block_full_energy(const PotDev &pot, const RepB<LY> &R, const CELL &C, int warp, int lane, unsigned lt,
                                           long long &pairs, int &status) {
    double erow = 0.0;  // LJ: the rows of this warp (lane-uniform)
    for (int a = warp; a < R.n; a += BW) {
        const double pa[3] = {R.x(a), R.y(a), R.z(a)};
        DeltaAcc acc = {0.0, 0.0, 0};
        int n2, xstart;
        const int cnt = scan_neighbours<RepB<LY>, 1>(pot, R, C, lane, lt, R.n, a, false, true, pa, pa, n2, xstart, status);
        eval_list<RepB<LY>, 1>(pot, R, lane, cnt, xstart, acc);
        pairs += acc.pairs;
        const double vs = warp_sum(acc.v);
        if (LY::EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig(a) = s1;
                R.eat(a) = vs;  // sigma_2 sum parked until the embedding pass
            }
        } else {
            erow += 0.5 * vs;  // energies[ii] = 0.5 * pairwise_energies.sum()
        }
        __syncwarp();
    }
    rep_sync();
    double ew = erow;
    if (LY::EMT) {
        double epart = 0.0;
        for (int a = threadIdx.x & 127; a < R.n; a += BW * 32) {
            const double ea = emt_embed(pot, R.sig(a));
            epart += ea + pot.neghalfv0overgamma2 * R.eat(a);
            R.eat(a) = ea;
        }
        ew = warp_sum(epart);
    }
    block_put(R, warp, lane, 2, ew);
    rep_sync();
    const double e = block_get(R, 2);
    rep_sync();
    return e;
}

template <int POT, int NS, int FEAT>
__global__ void __launch_bounds__(LayoutB<POT, NS>::REPS * BW * 32, 1) sweep_block_kernel(const __grid_constant__ SweepArgs a) {
    using LY = LayoutB<POT, NS>;
    using RV = RepB<LY>;
    constexpr bool UNIFORM = (FEAT & F_CELLVAR) == 0;
    extern __shared__ __align__(16) double smem_d[];
    const int slot = threadIdx.x >> 7;              // replica slot inside the CTA
    const int tid = threadIdx.x & 127;              // thread inside the replica's group of BW warps
    const int warp = tid >> 5, lane = tid & 31;
    const unsigned lt = lanemask_lt();
    const int r = blockIdx.x * LY::REPS + slot;
    if (r >= a.b.n_replicas) return;  // the whole group leaves; barriers are per group
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    RV R;
    R.b = smem_d + slot * (LY::BYTES / 8);
    R.wb = reinterpret_cast<double *>(reinterpret_cast<char *>(R.b) + LY::REG0_B + warp * LY::REG_B);
    const CellView<LY, UNIFORM> C{&a.cell, R.b + LY::CELLC};

    // ---- load the replica ----
    R.n = a.b.n_atoms[r];
    double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    double *gs = POT == QMC_POT_EMT ? a.b.sigma1 + (size_t)r * nmax : nullptr;
    double *ge = POT == QMC_POT_EMT ? a.b.eatom + (size_t)r * nmax : nullptr;
    for (int j = tid; j < nmax; j += BW * 32) {
        const bool in = j < R.n;
        R.x(j) = in ? gx[j] : 0.0;
        R.y(j) = in ? gy[j] : 0.0;
        R.z(j) = in ? gz[j] : 0.0;
        if (POT == QMC_POT_EMT) {
            R.sig(j) = in ? gs[j] : 0.0;
            R.eat(j) = in ? ge[j] : 0.0;
        }
    }
    int status = 0;
    double cell[3] = {0.0, 0.0, 0.0};
    if (!UNIFORM) {
        cell[0] = a.b.cell[(size_t)r * 9 + 0];
        cell[1] = a.b.cell[(size_t)r * 9 + 4];
        cell[2] = a.b.cell[(size_t)r * 9 + 8];
        if (!store_cell_block(R, cell, a.b.pbc, pot.cut_pre)) status |= QMC_STATUS_CELL_TOO_SMALL;
    }
    double E = a.b.energy[r];
    const double temperature = a.b.temperature[r];
    const double kT = temperature * KB;
    const double inv_kT = 1.0 / kT;
    int n_ex = (FEAT & F_EXCHANGE) ? a.b.n_exchange[r] : 0;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    long long n_pairs = 0;
    rep_sync();

    for (int k = 0; k < a.n_attempts; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        // the draws of DRAW_AHEAD attempts at a time, into one region shared by the BW warps: every warp writes the same values;
        // the barrier keeps a warp that ran ahead through attempts without barriers (failed moves) from overwriting them early
        const int ahead = k & (DRAW_AHEAD - 1);
        if (ahead == 0) {
            rep_sync();
            draw_ahead(R, key, attempt, grep, lane);
        }
        const Draws D = read_draws(R, ahead, (FEAT & F_EXCHANGE) != 0);
        int m = 0;
        while (m < a.n_moves - 1 && a.cdf[m] <= D.u_select) ++m;
        const MoveDev &mv = a.mv[m];
        const int *lab = ((FEAT & F_LABELS) && mv.labels) ? mv.labels + (size_t)r * nmax : nullptr;
        int accepted = -1, moved = -1;
        double delta = __longlong_as_double(0x7ff8000000000000LL);

        const bool is_disp = mv.type == QMC_MOVE_DISPLACEMENT;
        const bool is_exch = (FEAT & F_EXCHANGE) && mv.type == QMC_MOVE_EXCHANGE;
        if (is_disp || is_exch) {
            // ---- proposal (identical on every warp) ----
            bool has_old = false, has_new = false;
            int i = -1, pd = 0;
            double po[3] = {0, 0, 0}, pn[3] = {0, 0, 0};
            if (is_disp) {
                const int nu = lab ? count_labelled(lab, R.n) : R.n;
                if (nu > 0) {
                    const int idx = (int)(D.u_label * (double)nu);
                    i = lab ? select_labelled(lab, R.n, idx) : idx;
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
                    po[0] = R.x(i); po[1] = R.y(i); po[2] = R.z(i);
                    pn[0] = po[0] + t[0]; pn[1] = po[1] + t[1]; pn[2] = po[2] + t[2];
                    has_old = has_new = true;
                }
            } else {
                if (D.u_coin < mv.bias) {
                    if (R.n + 1 > nmax) {
                        status |= QMC_STATUS_CAPACITY;
                    } else {
                        const double f[3] = {D.o0, D.o1, D.o2};
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const double fc = __dmul_rn(f[d], UNIFORM ? a.cell.L[d] : cell[d]);
                            pn[d] = __dadd_rn(a.b.exchange_pos[d], __dsub_rn(fc, a.b.exchange_pos[d]));
                        }
                        i = R.n;
                        has_new = true;
                        pd = 1;
                    }
                } else {
                    const int nu = lab ? count_labelled(lab, R.n) : R.n;
                    if (nu > 0) {
                        const int idx = (int)(D.u_label * (double)nu);
                        i = lab ? select_labelled(lab, R.n, idx) : idx;
                        po[0] = R.x(i); po[1] = R.y(i); po[2] = R.z(i);
                        has_old = true;
                        pd = -1;
                    }
                }
            }
            const bool paired = has_old && has_new;  // displacement: old / new terms of a neighbour in adjacent slots (scan_both)
            if (has_old || has_new) {  // block-uniform
                // ---- phase 1: own tiles: scan + pair functions ----
                DeltaAcc acc = {0.0, 0.0, 0};
                int n2, xstart;
                const int cnt = paired ? scan_both<RV>(pot, R, C, lane, lt, R.n, i, po, pn, n2, xstart, status, warp, BW)
                                       : scan_neighbours<RV, 0>(pot, R, C, lane, lt, R.n, has_old ? i : -1, has_old, has_new, po, pn, n2,
                                                                xstart, status, warp, BW);
                eval_list<RV, 0>(pot, R, lane, cnt, xstart, acc);
                n_pairs += acc.pairs;
                block_put(R, warp, lane, 0, warp_sum(acc.v));
                if (POT == QMC_POT_EMT) block_put(R, warp, lane, 1, warp_sum(acc.snew));
                rep_sync();  // B1
                const double vs = block_get(R, 0);
                double sig_i = 0.0, eat_i = 0.0;
                if (POT == QMC_POT_EMT) {
                    sig_i = has_new ? block_get(R, 1) : 0.0;
                    // ---- phase 2: embedding of the touched atoms this warp owns (+ the moved atom on warp 0) ----
                    double part = 0.0, e_i_lane = 0.0;
                    const bool extra = has_new && warp == 0;
                    const int total = n2 + (extra ? 1 : 0);
                    double2 *slots = reinterpret_cast<double2 *>(&R.lr(0));
                    for (int e = lane; e < total; e += 32) {
                        const bool nb = e < n2;
                        const int j = nb ? R.l2(e) : 0;
                        double sn;
                        if (paired) {
                            const double2 t = nb ? slots[e] : make_double2(0.0, 0.0);
                            sn = nb ? R.sig(j) + (t.y - t.x) : sig_i;
                        } else {
                            sn = nb ? (R.sig(j) + gathered_dsigma(R, j)) : sig_i;
                        }
                        const double eold = nb ? R.eat(j) : (has_old ? R.eat(i) : 0.0);
                        const double en = emt_embed(pot, sn);
                        part += en - eold;
                        if (nb) {
                            if (paired) slots[e] = make_double2(sn, en);  // trial (sigma_1, E_atom) back into the pair of slots
                            else R.tail(e) = en;
                        } else {
                            e_i_lane = en;
                        }
                    }
                    if (extra) eat_i = __shfl_sync(FULL, e_i_lane, n2 & 31);
                    if (!has_new && warp == 0 && lane == 0) part -= R.eat(i);
                    block_put(R, warp, lane, 2, warp_sum(part));
                    rep_sync();  // B2
                    delta = block_get(R, 2) + 2.0 * pot.neghalfv0overgamma2 * vs;
                } else {
                    delta = vs;
                }
                bool acc_ok;
                if (is_disp) {
                    acc_ok = metropolis(D.u_accept, -delta * inv_kT, status);
                } else {
                    bool ov;
                    const double prob = gc_probability(delta, pd, n_ex, a.b.accessible_volume, a.b.exchange_mass, temperature,
                                                       a.b.chemical_potential, ov);
                    if (ov) status |= QMC_STATUS_EXP_OVERFLOW;
                    acc_ok = D.u_accept < prob;
                }
                if (acc_ok) {
                    if (POT == QMC_POT_EMT) {
                        const double2 *slots = reinterpret_cast<const double2 *>(&R.lr(0));
                        for (int e = lane; e < n2; e += 32) {
                            const int j = R.l2(e);
                            if (paired) {
                                const double2 t = slots[e];
                                R.sig(j) = t.x;
                                R.eat(j) = t.y;
                            } else {
                                R.sig(j) = R.sig(j) + gathered_dsigma(R, j);
                                R.eat(j) = R.tail(e);
                            }
                        }
                    }
                    if (has_new && warp == 0 && lane == 0) {
                        R.x(i) = pn[0];
                        R.y(i) = pn[1];
                        R.z(i) = pn[2];
                        if (POT == QMC_POT_EMT) {
                            R.sig(i) = sig_i;
                            R.eat(i) = eat_i;
                        }
                    }
                    E += delta;
                    if (is_exch) {
                        rep_sync();
                        if (warp == 0) {
                            for (int q = 0; q < a.n_moves; ++q)
                                if (a.mv[q].labels && table_first_use(a.mv, q))
                                    labels_changed(a.mv[q].labels + (size_t)r * nmax, R.n, pd > 0, pd < 0 ? i : -1, a.mv[q].default_label);
                            if (pd < 0) {
                                shift_down(&R.x(0), i, R.n);
                                shift_down(&R.y(0), i, R.n);
                                shift_down(&R.z(0), i, R.n);
                                if (POT == QMC_POT_EMT) {
                                    shift_down(&R.sig(0), i, R.n);
                                    shift_down(&R.eat(0), i, R.n);
                                }
                            }
                        }
                        R.n += pd;
                        n_ex += pd;
                    }
                }
                rep_sync();  // B3: the committed state is visible to every warp before the next scan
                accepted = acc_ok ? 1 : 0;
                moved = i;
            }
        } else if ((FEAT & F_CELL) && mv.type == QMC_MOVE_CELL) {
            for (int j = tid; j < R.n; j += BW * 32) {  // snapshot of the last accepted state
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
                ncell[d] = (((mv.axes >> d) & 1) ? s : 1.0) * cell[d];
                scale[d] = ncell[d] / cell[d];
            }
            for (int j = tid; j < R.n; j += BW * 32) {
                R.x(j) = R.x(j) * scale[0];
                R.y(j) = R.y(j) * scale[1];
                R.z(j) = R.z(j) * scale[2];
            }
            const bool cell_ok = store_cell_block(R, ncell, a.b.pbc, pot.cut_pre);
            if (!cell_ok) status |= QMC_STATUS_CELL_TOO_SMALL;
            long long fp = 0;
            const double e_new = cell_ok ? block_full_energy<LY>(pot, R, C, warp, lane, lt, fp, status) : E;
            n_pairs += fp;
            delta = e_new - E;
            const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
            const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
            const bool acc_ok = cell_ok && metropolis(D.u_accept, x, status);
            if (acc_ok) {
                E = e_new;
                cell[0] = ncell[0];
                cell[1] = ncell[1];
                cell[2] = ncell[2];
            } else {
                rep_sync();
                for (int j = tid; j < R.n; j += BW * 32) {
                    R.x(j) = gx[j];
                    R.y(j) = gy[j];
                    R.z(j) = gz[j];
                    if (POT == QMC_POT_EMT) {
                        R.sig(j) = gs[j];
                        R.eat(j) = ge[j];
                    }
                }
                store_cell_block(R, cell, a.b.pbc, pot.cut_pre);
            }
            rep_sync();
            accepted = acc_ok ? 1 : 0;
        }

        if (tid == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(a.b.counters) + ((size_t)r * QMC_MAX_MOVES + m) * 3;
            atomicAdd(cnt + 0, 1ull);
            if (accepted == 1) atomicAdd(cnt + 1, 1ull);
            if (accepted < 0) atomicAdd(cnt + 2, 1ull);
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
    rep_sync();
    for (int j = tid; j < R.n; j += BW * 32) {
        gx[j] = R.x(j);
        gy[j] = R.y(j);
        gz[j] = R.z(j);
        if (POT == QMC_POT_EMT) {
            gs[j] = R.sig(j);
            ge[j] = R.eat(j);
        }
    }
    if (tid == 0) {
        a.b.n_atoms[r] = R.n;
        a.b.energy[r] = E;
        if (FEAT & F_CELL) {
            a.b.cell[(size_t)r * 9 + 0] = cell[0];
            a.b.cell[(size_t)r * 9 + 4] = cell[1];
            a.b.cell[(size_t)r * 9 + 8] = cell[2];
        }
        if (FEAT & F_EXCHANGE) a.b.n_exchange[r] = n_ex;
    }
    if (status && lane == 0) atomicOr(a.b.status + r, status);
    if (a.b.pair_evals) {
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.b.pair_evals) + r, (unsigned long long)n_pairs);
    }
}

}  // namespace qmc
