// Row-based Monte Carlo sweep kernel: the same attempts as sweep_kernel.cuh (replica state resident in shared memory), but
// the neighbours of the moved atom come from a per-atom NEIGHBOUR ROW instead of a scan over all atoms.  The reference side of
// this is ASE's NeighborList behind `atoms.get_potential_energy()` (src/quansino/mc/criteria.py:104-106); the attempt
// semantics are src/quansino/mc/core.py:353-384 as in sweep_kernel.cuh.  NW warps cooperate on one replica: NW = 1 for batches
// of thousands of replicas, NW = 4 for batches like C3 (1024 replicas of 500 atoms per GPU) that would otherwise leave seven
// latency-bound warps per SM.
//
//   rows      global memory, [replica][atom][ROWW] u32.  Word 0 = number of entries, words 2..7 = the atom's position when the
//             rows were built (3 doubles), words 8.. = entries (atom j | signed 7-bit image shift per axis << 10 / 17 / 24;
//             0xFFFFFFFF = padding): every periodic image of another atom within cutoff + skin, ordered by atom, the images
//             of one atom next to each other and never across a 32-entry boundary (the unit of work of a warp).
//             The row of the atom that the NEXT attempt moves is fetched with cp.async (16 B per lane) into a double buffer
//             in shared memory while the current attempt computes: the draws are produced one group of four attempts ahead,
//             so the next atom is known.
//   validity  distances are judged in the metric of the cell the rows were built in (cell moves scale that metric):
//             with q = max_k L_build,k / L_now,k the slack is S = skin - (q - 1) cutoff.  Every committed atom stays within
//             S / 2 of its build position and a TRIAL position is evaluated only within S / 2 of it, so an image that is
//             inside the cutoff of the old or the trial position is in the row.  Otherwise the replica rebuilds its rows
//             (brute force, one thread per row, all images with shifts k0 - 1, k0, k0 + 1 per axis) and carries on.
//   attempt   row pass (the 32-entry blocks of the row dealt to the NW warps): decode (j, shift), image vector
//             = (x_j - x_i) - shift * L for the old and the trial position, entries inside the cutoff on EITHER side compacted
//             into aligned slot pairs (-r2_old, +r2_new) -- no nearest-image arithmetic, no second-image queue.  Pair passes as
//             in replica_warp.cuh (the sigma_1 term overwrites the slot).  Embedding: the first entry of an atom sums the terms
//             of its further images, so the list of touched atoms needs no bookkeeping; the moved atom is one more item.
//             With NW > 1: three barriers per attempt (sums, energy delta, commit) and fixed-order sums over the warps.
//   cell move the full recompute walks the rows (one row per atom, prefetched, atoms dealt to the warps) instead of N x N
//             distance tests.
//   registers the launch's scalar state (energy, 1 / kT, status, counters, metric) lives in shared memory: at the 72 registers
//             that 28 resident warps allow, anything else alive across the list passes ends up in local memory, which misses L1
//             half of the time next to 227 KB of shared memory (profiles/r2_a_rows.md).
#pragma once

#include "sweep_kernel.cuh"

namespace qmc {

constexpr int ROW_HDR = 8;                 // header words of a row
constexpr int ROW_SHIFT_MAX = 63;          // |image shift| per axis an entry can hold
constexpr uint32_t ROW_PAD = 0xFFFFFFFFu;  // padding entry

template <int POT, int NS, int NW>
struct LayoutR {
    static constexpr bool EMT = POT == QMC_POT_EMT;
    // row words (header + entries), a multiple of 4 (16-byte cp.async chunks).  Cu EMT: 78 images inside the cutoff, ~88 with the
    // default skin; Lennard-Jones with a 3 sigma cutoff: ~125, ~160 with the skin.  Classes above 128 atoms have shared memory to
    // spare: their rows also hold the next shell that a trial compression of a cell move pulls inside (Cu at -5 % per axis: 134)
    static constexpr int ROWW = EMT ? (NS <= 128 ? 112 : 160) : 208;
    static constexpr int CAPS = ROWW - ROW_HDR;               // entries per row
    static constexpr int BLOCKS = (CAPS + 31) / 32;           // 32-entry blocks of a row
    // slot pairs of one warp's region: its share of the row's blocks -- and, as single doubles, a WHOLE row (full evaluation)
    static constexpr int SHARE = ((BLOCKS + NW - 1) / NW) * 32;
    static constexpr int CAPW = NW == 1 ? CAPS : round_up(SHARE > (CAPS + 1) / 2 ? SHARE : (CAPS + 1) / 2, 8);
    // ---- shared by the warps of the replica (doubles) ----
    static constexpr int X = 0, Y = NS, Z = 2 * NS;
    static constexpr int SIG = 3 * NS, EAT = 4 * NS;
    static constexpr int CELLC = (EMT ? 5 * NS : 3 * NS) + (NS & 1);  // 9 doubles (16-byte aligned)
    static constexpr int PS = CELLC + 10;   // q[3], bound^2, largest committed displacement^2, build cell[3], trial position[3] + its
                                            // squared displacement, E, 1 / kT
    static constexpr int RED = PS + 14;     // NW x 4 doubles: per-warp partial sums
    static constexpr int MVT = RED + NW * 4;  // move table: cdf[QMC_MAX_MOVES], step[QMC_MAX_MOVES], then (type, op, axes, -) ints per move
    static constexpr int PROP = MVT + 4 * QMC_MAX_MOVES;  // NW > 1: three proposals, 8 doubles each (translation[3], u_accept, o0, then 4 ints + pad)
    static constexpr int SROW = PROP + (NW > 1 ? 24 : 0);  // NW > 1: three row buffers shared by the warps of the replica
    static constexpr int DRAW_B = (SROW + (NW > 1 ? 3 * ROWW / 2 : 0)) * 8;  // 2 groups x DRAW_AHEAD attempts x 4 blocks x double2
    static constexpr int CNT_B = DRAW_B + 2 * DRAW_AHEAD * 64;        // counters + status, pair evaluations, next move, next atom
    static constexpr int LABW = round_up(NS, 32) / 32;
    static constexpr int LABBIT_B = CNT_B + (QMC_MAX_MOVES * 3 + 4) * 4;
    static constexpr int REG0_B = round_up(LABBIT_B + QMC_MAX_MOVES * LABW * 4, 16);
    // ---- one region per warp: slot pairs (double2), atom index per pair (+ sentinel), two row buffers ----
    static constexpr int L2_OFF = CAPW * 16;
    static constexpr int ROW_OFF = round_up(L2_OFF + (CAPW + 2) * 2, 16);
    static constexpr int NROWBUF = NW == 1 ? 2 : 1;  // (NW > 1: displacement rows come through the shared buffers; the full
                                                      // evaluation refills its single buffer right after the row pass)
    static constexpr int REG_B = ROW_OFF + NROWBUF * ROWW * 4;
    static constexpr int BYTES = round_up(REG0_B + NW * REG_B, 16);
    // launch geometry: one wide CTA per SM, as many replicas as shared memory and the 72-register budget (28 warps) allow
    static constexpr int fit = (SMEM_PER_SM - 1024) / BYTES;
    static constexpr int REPS = fit > 28 / NW ? 28 / NW : (fit < 1 ? 1 : fit);
    static constexpr int THREADS = REPS * NW * 32;
    static constexpr bool FITS = REPS * BYTES <= SMEM_PER_CTA_MAX && fit > 0;
};

template <class LY>
struct RepR {
    static constexpr bool EMT = LY::EMT;
    double *b;   // replica base in shared memory
    double *wb;  // this warp's region
    int n;
    __device__ __forceinline__ double &x(int j) const { return b[LY::X + j]; }
    __device__ __forceinline__ double &y(int j) const { return b[LY::Y + j]; }
    __device__ __forceinline__ double &z(int j) const { return b[LY::Z + j]; }
    __device__ __forceinline__ double &sig(int j) const { return b[LY::SIG + j]; }
    __device__ __forceinline__ double &eat(int j) const { return b[LY::EAT + j]; }
    __device__ __forceinline__ double &cellc(int k) const { return b[LY::CELLC + k]; }
    __device__ __forceinline__ double &ps(int k) const { return b[LY::PS + k]; }
    __device__ __forceinline__ double &red(int k) const { return b[LY::RED + k]; }
    // the move table, copied from the kernel arguments: indexing the ARGUMENTS with a run-time move index makes the compiler
    // keep a private copy of the table in local memory (measured: 110 instructions and a dozen local loads per attempt and warp)
    __device__ __forceinline__ double &mv_cdf(int m) const { return b[LY::MVT + m]; }
    __device__ __forceinline__ double &mv_step(int m) const { return b[LY::MVT + QMC_MAX_MOVES + m]; }
    __device__ __forceinline__ int4 &mv_info(int m) const { return reinterpret_cast<int4 *>(b + LY::MVT + 2 * QMC_MAX_MOVES)[m]; }
    __device__ __forceinline__ double *prop(int which) const { return b + LY::PROP + 8 * which; }
    __device__ __forceinline__ uint32_t *srow(int which) const { return reinterpret_cast<uint32_t *>(b + LY::SROW) + which * LY::ROWW; }
    __device__ __forceinline__ double2 *draws() const { return reinterpret_cast<double2 *>(reinterpret_cast<char *>(b) + LY::DRAW_B); }
    __device__ __forceinline__ uint32_t *counts() const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::CNT_B); }
    __device__ __forceinline__ uint32_t *labbits(int m) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::LABBIT_B) + m * LY::LABW; }
    __device__ __forceinline__ double &sl(int e) const { return wb[e]; }  // flat view of the slot pairs
    __device__ __forceinline__ double2 *slots() const { return reinterpret_cast<double2 *>(wb); }
    __device__ __forceinline__ uint16_t &l2(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(wb) + LY::L2_OFF)[e]; }
    __device__ __forceinline__ uint32_t *rowbuf(int which) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(wb) + LY::ROW_OFF) + which * LY::ROWW; }
};

// barrier over the NW warps of ONE replica (several replicas share a CTA): named barrier 1 + replica slot
template <int NW>
__device__ __forceinline__ void rep_sync_r() {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / (32 * NW))), "r"(32 * NW) : "memory");
}

// sum of one value per warp in warp order (every thread gets the same result); slot k of red[warp][4]
template <int NW, class RV>
__device__ __forceinline__ double rep_sum(const RV &R, const int warp, const int lane, const int k, const double v) {
    if (NW == 1) return v;
    if (lane == 0) R.red(warp * 4 + k) = v;
    rep_sync_r<NW>();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += R.red(w * 4 + k);
    return s;
}

// ---- row transport ---------------------------------------------------------------------------------------------------
template <int ROWW>
__device__ __forceinline__ void row_prefetch(uint32_t *dst_smem, const uint32_t *src, const int lane) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
#pragma unroll
    for (int c = lane; c < ROWW / 4; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * c), "l"(src + 4 * c) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void row_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

__device__ __forceinline__ void row_decode(uint32_t w, int &j, double &sx, double &sy, double &sz) {
    j = (int)(w & 0x3FFu);
    sx = (double)((int)(w << 15) >> 25);
    sy = (double)((int)(w << 8) >> 25);
    sz = (double)((int)(w << 1) >> 25);
}

__device__ __forceinline__ uint32_t row_encode(int j, int sx, int sy, int sz) {
    return (unsigned)j | ((unsigned)(sx & 0x7F) << 10) | ((unsigned)(sy & 0x7F) << 17) | ((unsigned)(sz & 0x7F) << 24);
}

// ---- build -----------------------------------------------------------------------------------------------------------
// Rows of every atom of the replica, from the positions and cell in shared memory: all images within cutb with image shifts
// k0 + {-1, 0, +1} per axis (k0 = nearest image; complete while cutoff + skin <= 1.5 L, and self-images do not occur while
// cutoff + skin <= L -- checked by the caller).  THREAD-PER-ROW: thread t of round g builds the row of atom 32 NW g + t by
// itself, walking all atoms j (broadcast loads), so the build needs no cross-lane traffic and its iterations are independent:
// it has to be quick even when a single warp of the SM runs it (the rest of the launch waits for the replica that rebuilds).
// Inlined: as a function call it costs the hot loop ~45 local-memory instructions per attempt (registers saved around the call).
template <class LY, int NW, class CELL>
__device__ __forceinline__ int build_rows(const RepR<LY> R, const CELL C, const int tid, uint32_t *grows, const double cutb) {
    int status = 0;
    const Axis A0 = C.axis(0), A1 = C.axis(1), A2 = C.axis(2);
    const double cutb2 = cutb * cutb;
    // a second image along an axis can be inside only if |d0| > L - cutb
    const double t0 = C.per(0) ? A0.L - cutb : 1e300, t1 = C.per(1) ? A1.L - cutb : 1e300, t2 = C.per(2) ? A2.L - cutb : 1e300;
    for (int a0 = 0; a0 < R.n; a0 += 32 * NW) {
        const int a = a0 + tid;
        const bool mine = a < R.n;
        const int ac = mine ? a : 0;
        const double xa = R.x(ac), ya = R.y(ac), za = R.z(ac);
        uint32_t *row = grows + (size_t)ac * LY::ROWW + ROW_HDR;
        int cnt = 0;
        for (int j = 0; j < R.n; ++j) {
            const double dx = R.x(j) - xa, dy = R.y(j) - ya, dz = R.z(j) - za;
            const double kx = fma(dx, A0.invL, RINT_MAGIC) - RINT_MAGIC, ky = fma(dy, A1.invL, RINT_MAGIC) - RINT_MAGIC,
                         kz = fma(dz, A2.invL, RINT_MAGIC) - RINT_MAGIC;
            const double dx0 = fma(-kx, A0.L, dx), dy0 = fma(-ky, A1.L, dy), dz0 = fma(-kz, A2.L, dz);
            const unsigned fl = (fabs(dx0) > t0 ? 1u : 0u) | (fabs(dy0) > t1 ? 2u : 0u) | (fabs(dz0) > t2 ? 4u : 0u);
            if (!mine || j == a) continue;
            if (fl == 0u) {  // the common case: only the nearest image can be inside
                if (fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0)) < cutb2) {
                    const int sx = __double2int_rn(kx), sy = __double2int_rn(ky), sz = __double2int_rn(kz);
                    if (max(max(abs(sx), abs(sy)), abs(sz)) > ROW_SHIFT_MAX) status |= QMC_STATUS_ROW_SHIFT;
                    if (cnt < LY::CAPS) row[cnt] = row_encode(j, sx, sy, sz);
                    ++cnt;
                }
                continue;
            }
            // the further image along a flagged axis lies on the other side: shift k0 + sign(d0)
            const double ex = copysign(1.0, dx0), ey = copysign(1.0, dy0), ez = copysign(1.0, dz0);
            const double dx1 = fma(-(kx + ex), A0.L, dx), dy1 = fma(-(ky + ey), A1.L, dy), dz1 = fma(-(kz + ez), A2.L, dz);
            const int ikx = __double2int_rn(kx), iky = __double2int_rn(ky), ikz = __double2int_rn(kz);
            int m = 0;  // images of this atom inside: counted first, so that the group does not straddle a 32-entry boundary
            for (unsigned c = 0;;) {
                const double ux = (c & 1u) ? dx1 : dx0, uy = (c & 2u) ? dy1 : dy0, uz = (c & 4u) ? dz1 : dz0;
                if (fma(uz, uz, fma(uy, uy, ux * ux)) < cutb2) ++m;
                if (c == fl) break;
                c = ((c | (~fl & 7u)) + 1u) & fl;  // next subset of fl
            }
            if (m == 0) continue;
            if ((cnt & 31) + m > 32)
                for (const int to = (cnt + 31) & ~31; cnt < to; ++cnt)
                    if (cnt < LY::CAPS) row[cnt] = ROW_PAD;
            for (unsigned c = 0;;) {  // subsets of the flagged axes, the nearest image (empty subset) first
                const double ux = (c & 1u) ? dx1 : dx0, uy = (c & 2u) ? dy1 : dy0, uz = (c & 4u) ? dz1 : dz0;
                if (fma(uz, uz, fma(uy, uy, ux * ux)) < cutb2) {
                    const int sx = ikx + ((c & 1u) ? (int)ex : 0), sy = iky + ((c & 2u) ? (int)ey : 0), sz = ikz + ((c & 4u) ? (int)ez : 0);
                    if (max(max(abs(sx), abs(sy)), abs(sz)) > ROW_SHIFT_MAX) status |= QMC_STATUS_ROW_SHIFT;
                    if (cnt < LY::CAPS) row[cnt] = row_encode(j, sx, sy, sz);
                    ++cnt;
                }
                if (c == fl) break;
                c = ((c | (~fl & 7u)) + 1u) & fl;
            }
        }
        if (cnt > LY::CAPS) {
            status |= QMC_STATUS_LIST_OVERFLOW;
            cnt = LY::CAPS;
        }
        if (mine) {
            row[0 - ROW_HDR] = (uint32_t)cnt;
            row[1 - ROW_HDR] = 0u;
            double *ref = reinterpret_cast<double *>(row + 2 - ROW_HDR);
            ref[0] = xa;
            ref[1] = ya;
            ref[2] = za;
        }
    }
    status = __reduce_or_sync(FULL, status);
    __threadfence_block();
    rep_sync_r<NW>();
    return status;
}

// ---- evaluation of a list of signed r^2 values ---------------------------------------------------------------------------
// TRACK: the sigma_1 term of every entry overwrites its slot (local delta); otherwise sums only (one row of a full evaluation)
template <class RV, bool TRACK>
__device__ __forceinline__ void eval_slots(const PotDev &pot, const RV &R, const int lane, const int cnt, DeltaAcc &acc) {
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e < cnt;
        const double r2s = act ? R.sl(e) : 1.0e6;  // padding lanes sit far outside every cutoff: their terms come out as zero
        const double r2 = __hiloint2double(__double2hiint(r2s) & 0x7fffffff, __double2loint(r2s));
        if (RV::EMT) {
            double e1, e2, w;
            const bool in = emt_pair_w(pot, r2, e1, e2, w);
            const double s1 = e1 * w;
            acc.v = fma(e2, copysign(w, r2s), acc.v);
            if (__double2hiint(r2s) >= 0) acc.snew += s1;
            if (in) ++acc.pairs;
            if (TRACK && act) R.sl(e) = s1;
        } else {
            bool in;
            const double en = lj_pair(pot, r2, in);
            acc.v = fma(copysign(1.0, r2s), en, acc.v);
            acc.pairs += (act && in) ? 1 : 0;
        }
    }
    __syncwarp();
}

// Local energy delta of the displacement po -> pn of atom i from its row (in this warp's buffer).  Every thread of the replica
// returns the same dE / sig_i; n2 is the number of slot pairs THIS warp owns; eat_i is valid on warp 0.
template <class LY, int NW, class CELL>
__device__ __forceinline__ Trial local_delta_rows(const PotDev &pot, const RepR<LY> &R, const CELL &C, const int warp, const int lane,
                                                  const unsigned lt, const uint32_t *row, const int i, const double po[3],
                                                  const double pn[3]) {
    Trial T;
    const double Lx = C.axis(0).L, Ly = C.axis(1).L, Lz = C.axis(2).L;
    const int len = (int)row[0];
    double2 *slots = R.slots();
    int cnt = 0;
    for (int e0 = 32 * warp; e0 < len; e0 += 32 * NW) {  // this warp's blocks of the row
        const int e = e0 + lane;
        const uint32_t w = e < len ? row[ROW_HDR + e] : ROW_PAD;
        const bool valid = (int)w >= 0;
        int j;
        double sx, sy, sz;
        row_decode(valid ? w : 0u, j, sx, sy, sz);
        const double xj = R.x(j), yj = R.y(j), zj = R.z(j);
        const double dxo = fma(-sx, Lx, xj - po[0]), dyo = fma(-sy, Ly, yj - po[1]), dzo = fma(-sz, Lz, zj - po[2]);
        const double dxn = fma(-sx, Lx, xj - pn[0]), dyn = fma(-sy, Ly, yj - pn[1]), dzn = fma(-sz, Lz, zj - pn[2]);
        const double r2o = fma(dzo, dzo, fma(dyo, dyo, dxo * dxo));
        const double r2n = fma(dzn, dzn, fma(dyn, dyn, dxn * dxn));
        const bool any = valid && (r2o < pot.cut_pre2 || r2n < pot.cut_pre2);
        const unsigned b = __ballot_sync(FULL, any);
        const int pos = cnt + __popc(b & lt);  // <= the warp's share of the row <= CAPW by construction
        if (any) {
            const double neg_r2o = __hiloint2double(__double2hiint(r2o) ^ (int)0x80000000, __double2loint(r2o));
            slots[pos] = make_double2(neg_r2o, r2n);
            if (LY::EMT) R.l2(pos) = (uint16_t)j;
        }
        cnt += __popc(b);
    }
    if (LY::EMT && lane == 0) R.l2(cnt) = (uint16_t)0xFFFFu;  // ends the run of the last atom
    __syncwarp();
    DeltaAcc acc = {0.0, 0.0, 0};
    eval_slots<RepR<LY>, true>(pot, R, lane, 2 * cnt, acc);
    T.pairs = acc.pairs;
    T.n2 = cnt;
    T.sig_i = 0.0;
    T.eat_i = 0.0;
    double vs = warp_sum(acc.v);
    if (!LY::EMT) {
        T.dE = rep_sum<NW>(R, warp, lane, 0, vs);
        return T;
    }
    double s1 = warp_sum(acc.snew);
    double2 *const sl2 = slots;
    // embedding of the touched atoms this warp owns: the first entry of an atom sums the terms of its further images
    auto embed_items = [&](const int total, double &part, double &e_i_lane) {
        for (int e = lane; e < total; e += 32) {
            const bool nb = e < cnt;
            const int j = nb ? (int)R.l2(e) : i;
            const bool head = !nb || e == 0 || (int)R.l2(e - 1) != j;
            if (head) {
                double2 t = nb ? sl2[e] : make_double2(0.0, 0.0);
                if (nb)
                    for (int q = e + 1; (int)R.l2(q) == j; ++q) {  // further images of the same atom
                        const double2 u = sl2[q];
                        t.x += u.x;
                        t.y += u.y;
                    }
                const double sn = nb ? R.sig(j) + (t.y - t.x) : T.sig_i;
                const double en = emt_embed(pot, sn);
                part += en - R.eat(j);
                if (nb) sl2[e] = make_double2(sn, en);
                else e_i_lane = en;
            }
        }
    };
    if (NW > 1) {  // both sums behind one barrier
        if (lane == 0) {
            R.red(warp * 4 + 0) = vs;
            R.red(warp * 4 + 1) = s1;
        }
        rep_sync_r<NW>();  // B1
        vs = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            vs += R.red(w * 4 + 0);
            s1 += R.red(w * 4 + 1);
        }
    }
    T.sig_i = s1;
    double part = 0.0, e_i_lane = 0.0;
    embed_items(cnt + (warp == 0 ? 1 : 0), part, e_i_lane);  // the moved atom is one more item of warp 0
    if (warp == 0) T.eat_i = __shfl_sync(FULL, e_i_lane, cnt & 31);
    const double psum = rep_sum<NW>(R, warp, lane, 2, warp_sum(part));  // B2
    T.dE = psum + 2.0 * pot.neghalfv0overgamma2 * vs;
    // (measured and rejected for NW = 4: evaluating the neighbours' embeddings before the exchange of the sums and the moved atom's
    // embedding on every warp saves one barrier but costs 5 % -- 1.42e8 instead of 1.49e8 attempts/s on C3)
    __syncwarp();
    return T;
}

template <class LY>
__device__ __forceinline__ void commit_rows(const RepR<LY> &R, const int lane, const Trial &T) {
    if constexpr (LY::EMT) {
        const double2 *slots = R.slots();
        for (int e = lane; e < T.n2; e += 32) {
            const int j = (int)R.l2(e);
            if (e == 0 || (int)R.l2(e - 1) != j) {
                const double2 t = slots[e];
                R.sig(j) = t.x;
                R.eat(j) = t.y;
            }
        }
        __syncwarp();
    }
}

// Full evaluation of the replica walking the rows (cell moves), atoms dealt to the warps: fills sig / eat (EMT), returns the
// total energy on every thread and the pair evaluations of this warp.
struct FullEnergy {
    double e;
    unsigned pairs;
};

template <class LY, int NW, class CELL>
__device__ __noinline__ FullEnergy full_energy_rows(const PotDev &pot, const RepR<LY> R, const CELL C, const int warp, const int lane,
                                                    const unsigned lt, const uint32_t *grows) {
    unsigned pairs = 0;
    const double Lx = C.axis(0).L, Ly = C.axis(1).L, Lz = C.axis(2).L;
    double epart = 0.0;
    constexpr bool TWO = LY::NROWBUF == 2;
    if (warp < R.n) row_prefetch<LY::ROWW>(R.rowbuf(0), grows + (size_t)warp * LY::ROWW, lane);
    int which = 0;
    for (int a = warp; a < R.n; a += NW, which ^= (TWO ? 1 : 0)) {
        row_wait();
        if (TWO && a + NW < R.n) row_prefetch<LY::ROWW>(R.rowbuf(which ^ 1), grows + (size_t)(a + NW) * LY::ROWW, lane);
        const uint32_t *row = R.rowbuf(which);
        const double xa = R.x(a), ya = R.y(a), za = R.z(a);
        const int len = (int)row[0];
        int cnt = 0;
        for (int e0 = 0; e0 < len; e0 += 32) {
            const int e = e0 + lane;
            const uint32_t w = e < len ? row[ROW_HDR + e] : ROW_PAD;
            const bool valid = (int)w >= 0;
            int j;
            double sx, sy, sz;
            row_decode(valid ? w : 0u, j, sx, sy, sz);
            const double dx = fma(-sx, Lx, R.x(j) - xa), dy = fma(-sy, Ly, R.y(j) - ya), dz = fma(-sz, Lz, R.z(j) - za);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = valid && r2 < pot.cut_pre2;
            const unsigned b = __ballot_sync(FULL, in);
            if (in) R.sl(cnt + __popc(b & lt)) = r2;  // <= len <= CAPS <= 2 CAPW doubles
            cnt += __popc(b);
        }
        __syncwarp();
        if (!TWO && a + NW < R.n) row_prefetch<LY::ROWW>(R.rowbuf(0), grows + (size_t)(a + NW) * LY::ROWW, lane);  // the buffer is free again
        __syncwarp();
        DeltaAcc acc = {0.0, 0.0, 0};
        eval_slots<RepR<LY>, false>(pot, R, lane, cnt, acc);
        pairs += (unsigned)acc.pairs;
        const double vs = warp_sum(acc.v);
        if (LY::EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig(a) = s1;
                R.eat(a) = vs;  // sigma_2 sum parked here until the embedding pass below
            }
        } else if (lane == 0) {
            epart += 0.5 * vs;
        }
        __syncwarp();
    }
    rep_sync_r<NW>();
    if (LY::EMT) {
        for (int a = warp * 32 + lane; a < R.n; a += 32 * NW) {
            const double ea = emt_embed(pot, R.sig(a));
            epart += ea + pot.neghalfv0overgamma2 * R.eat(a);
            R.eat(a) = ea;
        }
    }
    const double e = rep_sum<NW>(R, warp, lane, 3, warp_sum(epart));
    rep_sync_r<NW>();
    return FullEnergy{e, pairs};
}

// the draws of DRAW_AHEAD attempts into group `g` of the double-buffered draw area (every warp writes the same values)
template <class RV>
__device__ __forceinline__ void draw_group(const RV &R, PhiloxKey key, unsigned long long attempt0, uint32_t replica, const int lane, int g) {
    double d0, d1;
    uniform_pair(key, attempt0 + (unsigned long long)((lane >> 2) & (DRAW_AHEAD - 1)), replica, (uint32_t)(lane & 3), d0, d1);
    __syncwarp();
    if (lane < 4 * DRAW_AHEAD) R.draws()[g * 4 * DRAW_AHEAD + lane] = make_double2(d0, d1);
    __syncwarp();
}

// Hides a value from common-subexpression elimination: addresses derived from the replica index are recomputed where they are
// used (two integer instructions) instead of living in registers for the whole launch.
__device__ __forceinline__ int opaque(int v) {
    asm volatile("" : "+r"(v));
    return v;
}

// copies between the replica in shared memory and its home in global memory (load / store / snapshot of a cell move)
template <int POT, int NW, class RV>
__device__ __noinline__ void replica_to_global(const qmc_batch_t &b, const RV R, const int r, const int tid) {
    const int nmax = b.n_max;
    double *gx = b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    rep_sync_r<NW>();
    for (int j = tid; j < R.n; j += 32 * NW) {
        gx[j] = R.x(j);
        gy[j] = R.y(j);
        gz[j] = R.z(j);
        if (POT == QMC_POT_EMT) {
            b.sigma1[(size_t)r * nmax + j] = R.sig(j);
            b.eatom[(size_t)r * nmax + j] = R.eat(j);
        }
    }
    rep_sync_r<NW>();
}

template <int POT, int NW, class RV>
__device__ __noinline__ void replica_from_global(const qmc_batch_t &b, const RV R, const int r, const int tid) {
    const int nmax = b.n_max;
    const double *gx = b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    rep_sync_r<NW>();
    for (int j = tid; j < nmax; j += 32 * NW) {
        const bool in = j < R.n;
        R.x(j) = in ? __ldcg(gx + j) : 0.0;
        R.y(j) = in ? __ldcg(gy + j) : 0.0;
        R.z(j) = in ? __ldcg(gz + j) : 0.0;
        if (POT == QMC_POT_EMT) {
            R.sig(j) = in ? __ldcg(b.sigma1 + (size_t)r * nmax + j) : 0.0;
            R.eat(j) = in ? __ldcg(b.eatom + (size_t)r * nmax + j) : 0.0;
        }
    }
    rep_sync_r<NW>();
}

// per-replica cell constants in shared memory, written by one thread of the replica
template <int NW, class RV>
__device__ __forceinline__ bool store_cell_r(const RV &R, const double diag[3], const int pbc[3], double cut_pre, const int tid) {
    CellArgs c;
    const bool ok = cell_constants(diag, pbc, cut_pre, c);
    rep_sync_r<NW>();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.cellc(k) = c.L[k];
            R.cellc(3 + k) = c.invL[k];
            R.cellc(6 + k) = c.thr[k];
        }
    }
    rep_sync_r<NW>();
    return ok;
}

struct RowArgs {
    double bound2;  // (skin / 2)^2: reach of a row without cell moves
    double skin;    // Angstrom
    double cutb;    // cut_pre + skin: the rows hold every image within this distance (build metric)
};

template <int POT, int NS, int FEAT, int NW>
__global__ void __launch_bounds__(LayoutR<POT, NS, NW>::THREADS, 1) sweep_rows_kernel(const __grid_constant__ SweepArgs a, const RowArgs ra) {
    using LY = LayoutR<POT, NS, NW>;
    constexpr bool UNIFORM = (FEAT & F_CELLVAR) == 0;
    constexpr bool CELLMOVES = (FEAT & F_CELL) != 0;
    constexpr int NSTAT = QMC_MAX_MOVES * 3;  // counts[NSTAT] = status, +1 = pair evaluations, +2 = next move, +3 = next atom
    static_assert((FEAT & (F_EXCHANGE | F_COMPOSITE)) == 0, "the row-based kernel runs displacement and cell moves only");
    extern __shared__ __align__(16) double smem_d[];
    const int slot = threadIdx.x / (32 * NW);  // replica slot inside the CTA
    const int tid = threadIdx.x % (32 * NW);   // thread inside the replica's group of NW warps
    const int warp = tid >> 5, lane = tid & 31;
    const unsigned lt = lanemask_lt();
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    RepR<LY> R;
    R.b = smem_d + slot * (LY::BYTES / 8);
    {  // keep the (32-bit shared-window) base in a register: otherwise it is re-derived from the thread id at every other use
        unsigned sb = (unsigned)__cvta_generic_to_shared(R.b);
        asm volatile("" : "+r"(sb));
        R.b = reinterpret_cast<double *>(__cvta_shared_to_generic((size_t)sb));
    }
    R.wb = reinterpret_cast<double *>(reinterpret_cast<char *>(R.b) + LY::REG0_B + warp * LY::REG_B);
    const CellView<LY, UNIFORM> C{&a.cell, R.b + LY::CELLC};
    const int r = blockIdx.x * LY::REPS + slot;
    if (r >= a.b.n_replicas) return;  // the whole group leaves; barriers are per group

    // ---- load the replica ----
    R.n = __ldcg(a.b.n_atoms + r);
    replica_from_global<POT, NW>(a.b, R, r, tid);
    if (tid < NSTAT + 4) R.counts()[tid] = 0u;
    if (tid < QMC_MAX_MOVES) {
        R.mv_cdf(tid) = tid < a.n_moves ? a.cdf[tid] : 2.0;
        R.mv_step(tid) = a.mv[tid].step;
        R.mv_info(tid) = make_int4(a.mv[tid].type, a.mv[tid].op, a.mv[tid].axes, 0);
    }
    if (tid == 0) {
        R.ps(12) = __ldcg(a.b.energy + r);
        R.ps(13) = 1.0 / (a.b.temperature[r] * KB);
    }
    rep_sync_r<NW>();
    auto flag = [&](const int bits) {
        if (bits != 0 && lane == 0) atomicOr(R.counts() + NSTAT, (uint32_t)bits);
    };
    double cell[3] = {a.cell.L[0], a.cell.L[1], a.cell.L[2]};
    if (!UNIFORM) {
        cell[0] = __ldcg(a.b.cell + (size_t)r * 9 + 0);
        cell[1] = __ldcg(a.b.cell + (size_t)r * 9 + 4);
        cell[2] = __ldcg(a.b.cell + (size_t)r * 9 + 8);
        if (!store_cell_r<NW>(R, cell, a.b.pbc, pot.cut_pre, tid)) flag(QMC_STATUS_CELL_TOO_SMALL);
    }
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    if (FEAT & F_LABELS) {
        if (warp == 0) rebuild_movable(R, a.mv, a.n_moves, r, nmax, LY::LABW);
        rep_sync_r<NW>();
    }

    // ---- neighbour rows: state, metric, (re)build ----
    // ps[0..2] = q_k = L_build,k / L_now,k, ps[3] = bound^2, ps[4] = largest squared committed displacement (build metric),
    // ps[5..7] = build cell.  Without cell moves q = 1, the bound is a kernel argument and every committed position has passed
    // the trial test, so none of them is needed.
    auto row_of = [&](const int atom) { return a.b.nbr_rows + ((size_t)opaque(r) * nmax + atom) * LY::ROWW; };
    auto meta_of = [&]() { return a.b.nbr_ref + (size_t)opaque(r) * 8; };  // build cell [3], largest squared committed displacement
    const bool have_rows = __ldcg(a.b.nbr_state + r) != 0;
    if (CELLMOVES && tid == 0) {
        const double *meta = meta_of();
        R.ps(4) = have_rows ? __ldcg(meta + 3) : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) R.ps(5 + k) = have_rows ? __ldcg(meta + k) : cell[k];
    }
    rep_sync_r<NW>();
    // metric of the build cell seen from the current cell; slack S = skin - (max q - 1) cutoff; bound = S / 2
    auto set_metric = [&](const double Lnow[3]) {
        if (!CELLMOVES) return;
        rep_sync_r<NW>();
        if (tid == 0) {
            double qm = 1.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double qk = Lnow[k] != 0.0 ? R.ps(5 + k) / Lnow[k] : 1.0;  // (positions scale along non-periodic axes too)
                R.ps(k) = qk;
                qm = fmax(qm, qk);
            }
            const double S = ra.skin - (qm - 1.0) * pot.cut_pre;
            R.ps(3) = S > 0.0 ? 0.25 * S * S : -1.0;
        }
        rep_sync_r<NW>();
    };
    auto bound2 = [&]() { return CELLMOVES ? R.ps(3) : ra.bound2; };
    auto in_reach = [&]() { return !CELLMOVES || R.ps(4) <= R.ps(3); };  // every committed atom within the bound
    auto rebuild = [&](const double Lnow[3]) {
        // the rows need cutoff + skin <= L (no self-images) in the cell they are built in
        for (int k = 0; k < 3; ++k)
            if (a.b.pbc[k] && !(Lnow[k] >= ra.cutb)) flag(QMC_STATUS_CELL_TOO_SMALL);
        rep_sync_r<NW>();
        flag(build_rows<LY, NW>(R, C, tid, row_of(0), ra.cutb));
        if (tid == 0) {
            double *meta = meta_of();
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                meta[k] = Lnow[k];
                if (CELLMOVES) R.ps(5 + k) = Lnow[k];
            }
            meta[3] = 0.0;
            if (CELLMOVES) R.ps(4) = 0.0;
            a.b.nbr_state[opaque(r)] += 1;
        }
        rep_sync_r<NW>();
        set_metric(Lnow);
    };
    set_metric(cell);
    if (!have_rows || !in_reach()) rebuild(cell);

    // ---- the attempt one ahead: move and atom, so that its row can be fetched early ----
    auto select_move = [&](const double u_select) {
        int m = 0;
        while (m < a.n_moves - 1 && R.mv_cdf(m) <= u_select) ++m;
        return m;
    };
    auto select_atom = [&](const int m, const double u_label) {
        if (R.mv_info(m).x != QMC_MOVE_DISPLACEMENT) return -1;
        const int nu = (FEAT & F_LABELS) ? count_bits(R.labbits(m), nullptr, LY::LABW) : R.n;
        if (nu <= 0) return -1;
        const int idx = (int)(u_label * (double)nu);
        return (FEAT & F_LABELS) ? select_bit(R.labbits(m), nullptr, LY::LABW, idx) : idx;
    };
    if constexpr (NW == 1) {
        const int k_end = a.n_attempts;
        draw_group(R, key, a.attempt_base, (uint32_t)(a.b.replica_offset + r), lane, 0);
        rep_sync_r<NW>();
        int m_cur, i_cur;
        {
            const double2 d0 = R.draws()[0];
            m_cur = select_move(d0.x);
            i_cur = select_atom(m_cur, d0.y);
        }
        if (i_cur >= 0) row_prefetch<LY::ROWW>(R.rowbuf(0), row_of(i_cur), lane);

        for (int k = 0; k < k_end; ++k) {
            const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
            const int dslot = k & (DRAW_AHEAD - 1), grp = (k / DRAW_AHEAD) & 1;
            if (dslot == 0 && k + DRAW_AHEAD < k_end) {  // the group after this one (the launch's first group exists already)
                draw_group(R, key, attempt + DRAW_AHEAD, (uint32_t)(a.b.replica_offset + opaque(r)), lane, grp ^ 1);
                rep_sync_r<NW>();
            }
            const double2 *dk = R.draws() + (grp * DRAW_AHEAD + dslot) * 4;
            // (u_accept is read where it is needed: nothing but the launch's persistent state should live across the list passes)
            const double o0 = dk[1].x, o1 = dk[1].y, o2 = dk[2].x;
            const int m = m_cur, i = i_cur;
            const int4 mvi = R.mv_info(m);
            const int mv_type = mvi.x, mv_op = mvi.y;
            const double mv_step = R.mv_step(m);
            // next attempt: selection now, prefetch below (after this attempt's row is safe)
            int m_next = 0, i_next = -1;
            if (k + 1 < k_end) {
                const int k1 = k + 1;
                const double2 dn = R.draws()[(((k1 / DRAW_AHEAD) & 1) * DRAW_AHEAD + (k1 & (DRAW_AHEAD - 1))) * 4];
                m_next = select_move(dn.x);
                i_next = select_atom(m_next, dn.y);
            }
            if (tid == 0) {  // parked for the next iteration (read back after the barrier that ends this attempt)
                R.counts()[NSTAT + 2] = (uint32_t)m_next;
                R.counts()[NSTAT + 3] = (uint32_t)i_next;
            }

            int accepted = -1, moved = -1;
            double delta = __longlong_as_double(0x7ff8000000000000LL);  // NaN

            if (mv_type == QMC_MOVE_DISPLACEMENT) {
                if (i >= 0) {
                    row_wait();
                    const uint32_t *row = R.rowbuf(k & 1);
                    double po[3], pn[3], d2;
                    // The trial position must stay within the bound of the atom's build position (build metric).  If it does not, the
                    // rows are rebuilt (then the build position IS the old position) and the proposal is simply evaluated again, so that
                    // nothing but the launch's persistent state is alive across the (rare) rebuild.
                    for (int tries = 0;; ++tries) {
                        double t[3];
                        propose_translation(mv_op, mv_step, o0, o1, o2, t);
                        po[0] = R.x(i), po[1] = R.y(i), po[2] = R.z(i);
                        pn[0] = po[0] + t[0], pn[1] = po[1] + t[1], pn[2] = po[2] + t[2];
                        const double *ref = reinterpret_cast<const double *>(row + 2);
                        const double ux = CELLMOVES ? fma(pn[0], R.ps(0), -ref[0]) : pn[0] - ref[0];
                        const double uy = CELLMOVES ? fma(pn[1], R.ps(1), -ref[1]) : pn[1] - ref[1];
                        const double uz = CELLMOVES ? fma(pn[2], R.ps(2), -ref[2]) : pn[2] - ref[2];
                        d2 = fma(uz, uz, fma(uy, uy, ux * ux));
                        if (d2 <= bound2()) break;
                        if (tries) {  // step > skin / 2: refused by the library beforehand
                            flag(QMC_STATUS_LIST_OVERFLOW);
                            break;
                        }
                        rebuild(cell);
                        row_prefetch<LY::ROWW>(R.rowbuf(k & 1), row_of(i), lane);
                        row_wait();
                    }
                    if (i_next >= 0) row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
                    if (tid == 0) {  // parked until the commit
                        R.ps(8) = pn[0];
                        R.ps(9) = pn[1];
                        R.ps(10) = pn[2];
                        if (CELLMOVES) R.ps(11) = d2;
                    }
                    const Trial T = local_delta_rows<LY, NW>(pot, R, C, warp, lane, lt, row, i, po, pn);
                    const unsigned np = __reduce_add_sync(FULL, (unsigned)T.pairs);
                    if (lane == 0) atomicAdd(R.counts() + NSTAT + 1, np);
                    delta = T.dE;
                    int st = 0;
                    const bool acc = metropolis(dk[2].y, -delta * R.ps(13), st);
                    flag(st);
                    if (acc) {
                        commit_rows<LY>(R, lane, T);
                        if (tid == 0) {
                            R.x(i) = R.ps(8);
                            R.y(i) = R.ps(9);
                            R.z(i) = R.ps(10);
                            if (POT == QMC_POT_EMT) {
                                R.sig(i) = T.sig_i;
                                R.eat(i) = T.eat_i;
                            }
                            if (CELLMOVES) R.ps(4) = fmax(R.ps(4), R.ps(11));
                            R.ps(12) += delta;
                        }
                    }
                    accepted = acc ? 1 : 0;
                    moved = i;
                } else if (i_next >= 0) {
                    row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
                }
            } else if (CELLMOVES && mv_type == QMC_MOVE_CELL) {
                // snapshot the last accepted state in global memory, deform in place, recompute everything through the rows
                replica_to_global<POT, NW>(a.b, R, opaque(r), tid);
                const double s = exp(__dadd_rn(-mv_step, __dmul_rn(mv_step - -mv_step, o0)));
                const int mv_axes = mvi.z;
                double ncell[3], scale[3];
    #pragma unroll
                for (int d = 0; d < 3; ++d) {
                    ncell[d] = (((mv_axes >> d) & 1) ? s : 1.0) * cell[d];
                    scale[d] = ncell[d] / cell[d];
                }
                // the rows must reach every pair of the TRIAL geometry: if the slack left by the committed displacements does not cover
                // the compression, rebuild at the CURRENT geometry first (normal density; still valid if the trial is rejected)
                set_metric(ncell);
                if (!in_reach()) {
                    rebuild(cell);
                    set_metric(ncell);
                }
                for (int j = tid; j < R.n; j += 32 * NW) {
                    R.x(j) = R.x(j) * scale[0];
                    R.y(j) = R.y(j) * scale[1];
                    R.z(j) = R.z(j) * scale[2];
                }
                if (!in_reach()) {  // a single deformation beyond skin / cutoff (12 % for Cu): rows of the trial geometry itself
                    store_cell_r<NW>(R, ncell, a.b.pbc, pot.cut_pre, tid);
                    rebuild(ncell);
                }
                const bool cell_ok = store_cell_r<NW>(R, ncell, a.b.pbc, pot.cut_pre, tid);
                if (!cell_ok) flag(QMC_STATUS_CELL_TOO_SMALL);
                const double E_old = R.ps(12);
                double e_new = E_old;
                if (cell_ok) {
                    const FullEnergy fe = full_energy_rows<LY, NW>(pot, R, C, warp, lane, lt, row_of(0));
                    e_new = fe.e;
                    const unsigned np = __reduce_add_sync(FULL, fe.pairs);
                    if (lane == 0) atomicAdd(R.counts() + NSTAT + 1, np);
                }
                delta = e_new - E_old;
                const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
                const double kT = a.b.temperature[r] * KB;
                const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
                int st = 0;
                const bool acc = cell_ok && metropolis(dk[2].y, x, st);
                flag(st);
                rep_sync_r<NW>();  // everybody has read E_old
                if (acc) {
                    if (tid == 0) R.ps(12) = e_new;
                    cell[0] = ncell[0];
                    cell[1] = ncell[1];
                    cell[2] = ncell[2];
                } else {
                    replica_from_global<POT, NW>(a.b, R, opaque(r), tid);
                    store_cell_r<NW>(R, cell, a.b.pbc, pot.cut_pre, tid);
                }
                set_metric(cell);
                if (!in_reach()) rebuild(cell);
                if (i_next >= 0) row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
                accepted = acc ? 1 : 0;
            }

            if (tid == 0) {
                uint32_t *cnt = R.counts() + m * 3;
                cnt[0] += 1u;
                if (accepted == 1) cnt[1] += 1u;
                if (accepted < 0) cnt[2] += 1u;
            }
            rep_sync_r<NW>();  // B3: the committed state (and the parked selection) is visible to every warp
            if (tid == 0 && a.has_trace) {
                const size_t ti = (size_t)r * a.n_attempts + k;
                if (a.tr.move) a.tr.move[ti] = (int8_t)m;
                if (a.tr.accepted) a.tr.accepted[ti] = (int8_t)accepted;
                if (a.tr.energy) a.tr.energy[ti] = R.ps(12);
                if (a.tr.delta) a.tr.delta[ti] = delta;
                if (a.tr.moved) a.tr.moved[ti] = moved;
            }
            m_cur = (int)R.counts()[NSTAT + 2];
            i_cur = (int)R.counts()[NSTAT + 3];
            if (NW > 1) rep_sync_r<NW>();  // ... and read before the next attempt parks its own
        }
    } else {
        // ---- NW warps per replica: the SCOUT.  The scalar part of an attempt (draws, move selection, atom, translation) depends on
        // nothing but the random stream and the static tables, so the last warp -- which owns the fourth block of a row and is idle
        // whenever a row has at most 96 entries -- works it out TWO ATTEMPTS AHEAD and publishes it in shared memory (prop[k % 3]),
        // together with the row of the atom (cp.async into srow[k % 3]: a whole attempt passes before the row is needed, the rows of
        // C3 live in HBM).  The other warps read the proposal, form the trial position and go straight to their block of the row:
        // the scalar work leaves the critical path, and is done once instead of NW times (profiles/r2_a_rows.md: ~500 of a warp's
        // ~1 080 instructions).  One cp.async group per published attempt, in order: "all but the newest group" = the next row.
        constexpr int SCOUT = NW - 1;
        const int k_end = a.n_attempts;
        auto fetch = [&](const int k1, const int i) {  // SCOUT only: one group per attempt (empty if the attempt needs no row)
            if (i >= 0) row_prefetch<LY::ROWW>(R.srow(k1 % 3), row_of(i), lane);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto publish = [&](const int k1) {  // SCOUT only: proposal of attempt k1 -> prop[k1 % 3], its row -> srow[k1 % 3] (in flight)
            const unsigned long long att = a.attempt_base + (unsigned long long)k1;
            if ((k1 & (DRAW_AHEAD - 1)) == 0) draw_group(R, key, att, (uint32_t)(a.b.replica_offset + opaque(r)), lane, 0);
            const double2 *dk = R.draws() + (k1 & (DRAW_AHEAD - 1)) * 4;
            const int m = select_move(dk[0].x);
            const int i = select_atom(m, dk[0].y);
            const int4 mvi = R.mv_info(m);
            double t[3] = {0.0, 0.0, 0.0};
            const bool disp = mvi.x == QMC_MOVE_DISPLACEMENT && i >= 0;
            fetch(k1, disp ? i : -1);
            if (disp) propose_translation(mvi.y, R.mv_step(m), dk[1].x, dk[1].y, dk[2].x, t);
            if (lane == 0) {
                double *P = R.prop(k1 % 3);
                P[0] = t[0], P[1] = t[1], P[2] = t[2];
                P[3] = dk[2].y;  // u_accept
                P[4] = dk[1].x;  // o0: the cell move's draw
                *reinterpret_cast<int4 *>(P + 6) = make_int4(m, i, mvi.x, mvi.z);
            }
        };
        auto refetch = [&](const int k1) {  // SCOUT only: the rows were rebuilt after the row of attempt k1 had been fetched
            if (k1 >= k_end) return;
            const int4 q = *reinterpret_cast<const int4 *>(R.prop(k1 % 3) + 6);
            fetch(k1, q.z == QMC_MOVE_DISPLACEMENT ? q.y : -1);
        };
        if (warp == SCOUT) {
            publish(0);
            if (k_end > 1) publish(1);
            row_wait();
        }
        rep_sync_r<NW>();
        for (int k = 0; k < k_end; ++k) {
            const double *P = R.prop(k % 3);
            const int4 pi = *reinterpret_cast<const int4 *>(P + 6);
            const int m = pi.x, i = pi.y, mv_type = pi.z;
            int accepted = -1, moved = -1;
            double delta = __longlong_as_double(0x7ff8000000000000LL);  // NaN
            const bool displace = mv_type == QMC_MOVE_DISPLACEMENT && i >= 0;
            const uint32_t *row = R.srow(k % 3);
            double po[3] = {0.0, 0.0, 0.0}, pn[3] = {0.0, 0.0, 0.0}, d2 = 0.0;
            bool rebuilt = false;
            if (displace) {
                // trial position and its distance from the atom's build position (see the one-warp loop above); every warp decides alike
                for (int tries = 0;; ++tries) {
                    po[0] = R.x(i), po[1] = R.y(i), po[2] = R.z(i);
                    pn[0] = po[0] + P[0], pn[1] = po[1] + P[1], pn[2] = po[2] + P[2];
                    const double *ref = reinterpret_cast<const double *>(row + 2);
                    const double ux = CELLMOVES ? fma(pn[0], R.ps(0), -ref[0]) : pn[0] - ref[0];
                    const double uy = CELLMOVES ? fma(pn[1], R.ps(1), -ref[1]) : pn[1] - ref[1];
                    const double uz = CELLMOVES ? fma(pn[2], R.ps(2), -ref[2]) : pn[2] - ref[2];
                    d2 = fma(uz, uz, fma(uy, uy, ux * ux));
                    if (d2 <= bound2()) break;
                    if (tries) {
                        flag(QMC_STATUS_LIST_OVERFLOW);
                        break;
                    }
                    rebuild(cell);
                    rebuilt = true;
                    if (warp == SCOUT) {  // this attempt's row again, and (below) the next one's
                        row_wait();
                        row_prefetch<LY::ROWW>(R.srow(k % 3), row_of(i), lane);
                        row_wait();
                    }
                    rep_sync_r<NW>();
                }
            }
            // the attempt after next (after a possible rebuild, so that its row is read from the rows that are valid now)
            if (warp == SCOUT) {
                if (rebuilt) refetch(k + 1);
                if (k + 2 < k_end) publish(k + 2);
                else asm volatile("cp.async.commit_group;" ::: "memory");  // keeps "all but the newest group" = the next row
            }
            if (displace) {
                if (tid == 0) {  // parked until the commit
                    R.ps(8) = pn[0];
                    R.ps(9) = pn[1];
                    R.ps(10) = pn[2];
                    if (CELLMOVES) R.ps(11) = d2;
                }
                const Trial T = local_delta_rows<LY, NW>(pot, R, C, warp, lane, lt, row, i, po, pn);
                const unsigned np = __reduce_add_sync(FULL, (unsigned)T.pairs);
                if (lane == 0 && np) atomicAdd(R.counts() + NSTAT + 1, np);
                delta = T.dE;
                int st = 0;
                const bool acc = metropolis(P[3], -delta * R.ps(13), st);
                flag(st);
                if (acc) {
                    commit_rows<LY>(R, lane, T);
                    if (tid == 0) {
                        R.x(i) = R.ps(8);
                        R.y(i) = R.ps(9);
                        R.z(i) = R.ps(10);
                        if (POT == QMC_POT_EMT) {
                            R.sig(i) = T.sig_i;
                            R.eat(i) = T.eat_i;
                        }
                        if (CELLMOVES) R.ps(4) = fmax(R.ps(4), R.ps(11));
                        R.ps(12) += delta;
                    }
                }
                accepted = acc ? 1 : 0;
                moved = i;
            } else if (CELLMOVES && mv_type == QMC_MOVE_CELL) {
                replica_to_global<POT, NW>(a.b, R, opaque(r), tid);
                const double mv_step = R.mv_step(m);
                const double s = exp(__dadd_rn(-mv_step, __dmul_rn(mv_step - -mv_step, P[4])));
                double ncell[3], scale[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    ncell[d] = (((pi.w >> d) & 1) ? s : 1.0) * cell[d];
                    scale[d] = ncell[d] / cell[d];
                }
                set_metric(ncell);
                if (!in_reach()) {
                    rebuild(cell);
                    set_metric(ncell);
                }
                for (int j = tid; j < R.n; j += 32 * NW) {
                    R.x(j) = R.x(j) * scale[0];
                    R.y(j) = R.y(j) * scale[1];
                    R.z(j) = R.z(j) * scale[2];
                }
                if (!in_reach()) {
                    store_cell_r<NW>(R, ncell, a.b.pbc, pot.cut_pre, tid);
                    rebuild(ncell);
                }
                const bool cell_ok = store_cell_r<NW>(R, ncell, a.b.pbc, pot.cut_pre, tid);
                if (!cell_ok) flag(QMC_STATUS_CELL_TOO_SMALL);
                const double E_old = R.ps(12);
                double e_new = E_old;
                if (cell_ok) {
                    const FullEnergy fe = full_energy_rows<LY, NW>(pot, R, C, warp, lane, lt, row_of(0));
                    e_new = fe.e;
                    const unsigned np = __reduce_add_sync(FULL, fe.pairs);
                    if (lane == 0) atomicAdd(R.counts() + NSTAT + 1, np);
                }
                delta = e_new - E_old;
                const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
                const double kT = a.b.temperature[r] * KB;
                const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
                int st = 0;
                const bool acc = cell_ok && metropolis(P[3], x, st);
                flag(st);
                rep_sync_r<NW>();  // everybody has read E_old
                if (acc) {
                    if (tid == 0) R.ps(12) = e_new;
                    cell[0] = ncell[0];
                    cell[1] = ncell[1];
                    cell[2] = ncell[2];
                } else {
                    replica_from_global<POT, NW>(a.b, R, opaque(r), tid);
                    store_cell_r<NW>(R, cell, a.b.pbc, pot.cut_pre, tid);
                }
                set_metric(cell);
                if (!in_reach()) rebuild(cell);
                // the rows may have been rebuilt since the scout fetched the rows of the next two attempts: fetch them again
                if (warp == SCOUT) {
                    row_wait();
                    refetch(k + 1);
                    refetch(k + 2);
                    if (k + 2 >= k_end) asm volatile("cp.async.commit_group;" ::: "memory");
                }
                accepted = acc ? 1 : 0;
            }
            if (tid == 0) {
                uint32_t *cnt = R.counts() + m * 3;
                cnt[0] += 1u;
                if (accepted == 1) cnt[1] += 1u;
                if (accepted < 0) cnt[2] += 1u;
            }
            if (warp == SCOUT) {  // the next attempt's row has landed before the barrier publishes it (the one after may be in flight)
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
            }
            rep_sync_r<NW>();               // B3: committed state, next proposal and its row are visible to every warp
            if (tid == 0 && a.has_trace) {
                const size_t ti = (size_t)r * a.n_attempts + k;
                if (a.tr.move) a.tr.move[ti] = (int8_t)m;
                if (a.tr.accepted) a.tr.accepted[ti] = (int8_t)accepted;
                if (a.tr.energy) a.tr.energy[ti] = R.ps(12);
                if (a.tr.delta) a.tr.delta[ti] = delta;
                if (a.tr.moved) a.tr.moved[ti] = moved;
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // ---- store the replica ----
    replica_to_global<POT, NW>(a.b, R, r, tid);
    if (tid < NSTAT) {
        const uint32_t c = R.counts()[tid];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(a.b.counters) + (size_t)r * NSTAT + tid, (unsigned long long)c);
    }
    if (tid == 0) {
        a.b.energy[r] = R.ps(12);
        if (CELLMOVES) {
            a.b.cell[(size_t)r * 9 + 0] = cell[0];
            a.b.cell[(size_t)r * 9 + 4] = cell[1];
            a.b.cell[(size_t)r * 9 + 8] = cell[2];
            meta_of()[3] = R.ps(4);
        }
        const int status = (int)R.counts()[NSTAT];
        if (status) atomicOr(a.b.status + r, status);
        if (a.b.pair_evals) a.b.pair_evals[r] += (long long)R.counts()[NSTAT + 1];
    }
}

}  // namespace qmc
