// Row-based Monte Carlo sweep kernel: the same attempts as sweep_kernel.cuh (one warp per replica, replica state resident
// in shared memory), but the neighbours of the moved atom come from a per-atom NEIGHBOUR ROW instead of a scan over all
// atoms.  The reference side of this is ASE's NeighborList behind `atoms.get_potential_energy()`
// (src/quansino/mc/criteria.py:104-106); the attempt semantics are src/quansino/mc/core.py:353-384 as in sweep_kernel.cuh.
//
//   rows      global memory, [replica][atom][ROWW] u32.  Word 0 = number of entries, words 2..7 = the atom's position when the
//             rows were built (3 doubles), words 8.. = entries (atom j | signed 7-bit image shift per axis << 10 / 17 / 24):
//             every periodic image of another atom within cutoff + skin, the images of one atom next to each other.
//             The row of the atom that the NEXT attempt moves is fetched with cp.async (16 B per lane) into a double buffer
//             in shared memory while the current attempt computes: the draws are produced one group of four attempts ahead,
//             so the next atom is known.
//   validity  distances are judged in the metric of the cell the rows were built in (cell moves scale that metric):
//             with q = max_k L_build,k / L_now,k the slack is S = skin - (q - 1) cutoff.  Every committed atom stays within
//             S / 2 of its build position and a TRIAL position is evaluated only within S / 2 of it, so an image that is
//             inside the cutoff of the old or the trial position is in the row.  Otherwise the replica rebuilds its rows
//             (brute force, all images with shifts k0 - 1, k0, k0 + 1 per axis) and carries on.
//   attempt   row pass: decode (j, shift), image vector = (x_j - x_i) - shift * L for the old and the trial position, entries
//             inside the cutoff on EITHER side compacted into aligned slot pairs (-r2_old, +r2_new) -- no nearest-image
//             arithmetic, no second-image queue.  Pair passes as in replica_warp.cuh (the sigma_1 term overwrites the slot).
//             Embedding: the first entry of an atom sums the terms of its further images (adjacent by construction), so the
//             list of touched atoms needs no bookkeeping; the moved atom is one more item.
//   cell move the full recompute walks the rows (one row per atom, prefetched) instead of N x N distance tests.
#pragma once

#include "sweep_kernel.cuh"

namespace qmc {

constexpr int ROW_HDR = 8;          // header words of a row
constexpr int ROW_SHIFT_MAX = 63;   // |image shift| per axis an entry can hold

template <int POT, int NS>
struct LayoutR {
    static constexpr bool EMT = POT == QMC_POT_EMT;
    // row words (header + entries), a multiple of 4 (16-byte cp.async chunks).  Cu EMT: 78 images inside the cutoff, ~88 with the
    // default skin; Lennard-Jones with a 3 sigma cutoff: ~125, ~160 with the skin
    static constexpr int ROWW = EMT ? (NS <= 128 ? 112 : 128) : 208;
    static constexpr int CAPS = ROWW - ROW_HDR;  // entries per row = slot pairs of one attempt
    static constexpr int X = 0, Y = NS, Z = 2 * NS;
    static constexpr int SIG = 3 * NS, EAT = 4 * NS;
    static constexpr int SL = EMT ? 5 * NS + (NS & 1) : 3 * NS + (NS & 1);  // slot pairs (double2, 16-byte aligned)
    static constexpr int CELLC = SL + 2 * CAPS;                             // 9 doubles
    static constexpr int PS = CELLC + 10;                                   // q[3], bound^2, largest committed displacement^2, build cell[3]
    static constexpr int L2_B = (PS + 8) * 8;                               // u16 atom index per slot pair (+ sentinel)
    static constexpr int ROW_B = round_up(L2_B + (CAPS + 2) * 2, 16);       // 2 x ROWW u32
    static constexpr int DRAW_B = ROW_B + 2 * ROWW * 4;                     // 2 groups x DRAW_AHEAD attempts x 4 blocks x double2
    static constexpr int CNT_B = DRAW_B + 2 * DRAW_AHEAD * 64;
    static constexpr int LABW = round_up(NS, 32) / 32;
    static constexpr int LABBIT_B = CNT_B + QMC_MAX_MOVES * 3 * 4;
    static constexpr int BYTES = round_up(LABBIT_B + QMC_MAX_MOVES * LABW * 4, 16);
    static constexpr int fit = (SMEM_PER_SM - 1024) / BYTES;
    static constexpr int WARPS = fit > 28 ? 28 : (fit < 1 ? 1 : fit);
    static constexpr bool FITS = WARPS * BYTES <= SMEM_PER_CTA_MAX && fit > 0;
};

template <class LY>
struct RepR {
    static constexpr bool EMT = LY::EMT;
    double *b;
    int n;
    __device__ __forceinline__ double &x(int j) const { return b[LY::X + j]; }
    __device__ __forceinline__ double &y(int j) const { return b[LY::Y + j]; }
    __device__ __forceinline__ double &z(int j) const { return b[LY::Z + j]; }
    __device__ __forceinline__ double &sig(int j) const { return b[LY::SIG + j]; }
    __device__ __forceinline__ double &eat(int j) const { return b[LY::EAT + j]; }
    __device__ __forceinline__ double &sl(int e) const { return b[LY::SL + e]; }  // flat view of the slot pairs
    __device__ __forceinline__ double2 *slots() const { return reinterpret_cast<double2 *>(b + LY::SL); }
    __device__ __forceinline__ double &cellc(int k) const { return b[LY::CELLC + k]; }
    __device__ __forceinline__ double &ps(int k) const { return b[LY::PS + k]; }
    __device__ __forceinline__ uint16_t &l2(int e) const { return reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(b) + LY::L2_B)[e]; }
    __device__ __forceinline__ uint32_t *rowbuf(int which) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::ROW_B) + which * LY::ROWW; }
    __device__ __forceinline__ double2 *draws() const { return reinterpret_cast<double2 *>(reinterpret_cast<char *>(b) + LY::DRAW_B); }
    __device__ __forceinline__ uint32_t *counts() const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::CNT_B); }
    __device__ __forceinline__ uint32_t *labbits(int m) const { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(b) + LY::LABBIT_B) + m * LY::LABW; }
};

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

// ---- build -----------------------------------------------------------------------------------------------------------
// Rows of every atom of the replica, from the positions and cell in shared memory: all images within sqrt(cutb2) with image
// shifts k0 + {-1, 0, +1} per axis (k0 = nearest image; complete while cutoff + skin <= 1.5 L, and self-images do not occur
// while cutoff + skin <= L -- checked by the caller).  Entries ordered by atom index, the images of one atom adjacent.
template <class LY, class CELL>
__device__ __noinline__ int build_rows(const RepR<LY> R, const CELL C, const int lane, uint32_t *grows, const double cutb) {
    int status = 0;
    const Axis A0 = C.axis(0), A1 = C.axis(1), A2 = C.axis(2);
    const double cutb2 = cutb * cutb;
    // a second image along an axis can be inside only if |d0| > L - cutb
    const double t0 = C.per(0) ? A0.L - cutb : 1e300, t1 = C.per(1) ? A1.L - cutb : 1e300, t2 = C.per(2) ? A2.L - cutb : 1e300;
    for (int a = 0; a < R.n; ++a) {
        const double xa = R.x(a), ya = R.y(a), za = R.z(a);
        uint32_t *row = grows + (size_t)a * LY::ROWW;
        int cnt = 0;
        for (int base = 0; base < R.n; base += 32) {
            const int j = base + lane;
            const bool valid = j < R.n && j != a;
            const int jc = valid ? j : 0;
            const double dx = R.x(jc) - xa, dy = R.y(jc) - ya, dz = R.z(jc) - za;
            const double kx = fma(dx, A0.invL, RINT_MAGIC) - RINT_MAGIC, ky = fma(dy, A1.invL, RINT_MAGIC) - RINT_MAGIC,
                         kz = fma(dz, A2.invL, RINT_MAGIC) - RINT_MAGIC;
            const double dx0 = fma(-kx, A0.L, dx), dy0 = fma(-ky, A1.L, dy), dz0 = fma(-kz, A2.L, dz);
            // the further image along an axis lies on the other side: shift k0 + sign(d0)
            const double ex = copysign(1.0, dx0), ey = copysign(1.0, dy0), ez = copysign(1.0, dz0);
            const double dx1 = fma(-(kx + ex), A0.L, dx), dy1 = fma(-(ky + ey), A1.L, dy), dz1 = fma(-(kz + ez), A2.L, dz);
            const unsigned fl = (fabs(dx0) > t0 ? 1u : 0u) | (fabs(dy0) > t1 ? 2u : 0u) | (fabs(dz0) > t2 ? 4u : 0u);
            // pass 1: count the images of this atom inside; pass 2: write them at the lane's offset
            int m = 0;
            if (valid) {
                unsigned c = 0;
                for (;;) {
                    const double ux = (c & 1u) ? dx1 : dx0, uy = (c & 2u) ? dy1 : dy0, uz = (c & 4u) ? dz1 : dz0;
                    if (fma(uz, uz, fma(uy, uy, ux * ux)) < cutb2) ++m;
                    if (c == fl) break;
                    c = ((c | (~fl & 7u)) + 1u) & fl;  // next subset of fl
                }
            }
            int incl = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            int at = cnt + incl - m;
            if (m > 0) {
                const int ikx = __double2int_rn(kx), iky = __double2int_rn(ky), ikz = __double2int_rn(kz);
                unsigned c = 0;
                for (;;) {
                    const double ux = (c & 1u) ? dx1 : dx0, uy = (c & 2u) ? dy1 : dy0, uz = (c & 4u) ? dz1 : dz0;
                    if (fma(uz, uz, fma(uy, uy, ux * ux)) < cutb2) {
                        const int sx = ikx + ((c & 1u) ? (int)ex : 0), sy = iky + ((c & 2u) ? (int)ey : 0), sz = ikz + ((c & 4u) ? (int)ez : 0);
                        if (max(max(abs(sx), abs(sy)), abs(sz)) > ROW_SHIFT_MAX) status |= QMC_STATUS_ROW_SHIFT;
                        if (at < LY::CAPS)
                            row[ROW_HDR + at] = (unsigned)j | ((unsigned)(sx & 0x7F) << 10) | ((unsigned)(sy & 0x7F) << 17) | ((unsigned)(sz & 0x7F) << 24);
                        ++at;
                    }
                    if (c == fl) break;
                    c = ((c | (~fl & 7u)) + 1u) & fl;
                }
            }
            cnt += total;
        }
        if (cnt > LY::CAPS) {
            status |= QMC_STATUS_LIST_OVERFLOW;
            cnt = LY::CAPS;
        }
        if (lane == 0) {
            row[0] = (uint32_t)cnt;
            row[1] = 0u;
            double *ref = reinterpret_cast<double *>(row + 2);
            ref[0] = xa;
            ref[1] = ya;
            ref[2] = za;
        }
    }
    __threadfence_block();
    __syncwarp();
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

// Local energy delta of the displacement po -> pn of atom i from its row (in shared memory).
template <class LY, class CELL>
__device__ __forceinline__ Trial local_delta_rows(const PotDev &pot, const RepR<LY> &R, const CELL &C, const int lane, const unsigned lt,
                                                  const uint32_t *row, const int i, const double po[3], const double pn[3]) {
    Trial T;
    const double Lx = C.axis(0).L, Ly = C.axis(1).L, Lz = C.axis(2).L;
    const int len = (int)row[0];
    double2 *slots = R.slots();
    int cnt = 0;
    for (int e0 = 0; e0 < len; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < len;
        const uint32_t w = valid ? row[ROW_HDR + e] : 0u;
        int j;
        double sx, sy, sz;
        row_decode(w, j, sx, sy, sz);
        const double xj = R.x(j), yj = R.y(j), zj = R.z(j);
        const double dxo = fma(-sx, Lx, xj - po[0]), dyo = fma(-sy, Ly, yj - po[1]), dzo = fma(-sz, Lz, zj - po[2]);
        const double dxn = fma(-sx, Lx, xj - pn[0]), dyn = fma(-sy, Ly, yj - pn[1]), dzn = fma(-sz, Lz, zj - pn[2]);
        const double r2o = fma(dzo, dzo, fma(dyo, dyo, dxo * dxo));
        const double r2n = fma(dzn, dzn, fma(dyn, dyn, dxn * dxn));
        const bool any = valid && (r2o < pot.cut_pre2 || r2n < pot.cut_pre2);
        const unsigned b = __ballot_sync(FULL, any);
        const int pos = cnt + __popc(b & lt);  // <= len <= CAPS by construction
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
    const double vs = warp_sum(acc.v);
    T.pairs = acc.pairs;
    T.n2 = cnt;
    T.sig_i = 0.0;
    T.eat_i = 0.0;
    if (!LY::EMT) {
        T.dE = vs;
        return T;
    }
    T.sig_i = warp_sum(acc.snew);
    double part = 0.0, e_i_lane = 0.0;
    for (int e = lane; e <= cnt; e += 32) {
        const bool nb = e < cnt;
        const int j = nb ? (int)R.l2(e) : i;
        const bool head = !nb || e == 0 || (int)R.l2(e - 1) != j;
        if (head) {
            double2 t = nb ? slots[e] : make_double2(0.0, 0.0);
            if (nb)
                for (int q = e + 1; (int)R.l2(q) == j; ++q) {  // further images of the same atom
                    const double2 u = slots[q];
                    t.x += u.x;
                    t.y += u.y;
                }
            const double sn = nb ? R.sig(j) + (t.y - t.x) : T.sig_i;
            const double en = emt_embed(pot, sn);
            part += en - R.eat(j);
            if (nb) slots[e] = make_double2(sn, en);
            else e_i_lane = en;
        }
    }
    T.eat_i = __shfl_sync(FULL, e_i_lane, cnt & 31);
    T.dE = warp_sum(part) + 2.0 * pot.neghalfv0overgamma2 * vs;
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

// Full evaluation of the replica walking the rows (cell moves): fills sig / eat (EMT), returns the total energy on every lane.
struct FullEnergy {
    double e;
    long long pairs;  // per-lane partial
};

template <class LY, class CELL>
__device__ __noinline__ FullEnergy full_energy_rows(const PotDev &pot, const RepR<LY> R, const CELL C, const int lane, const unsigned lt,
                                                    const uint32_t *grows) {
    long long pairs = 0;
    const double Lx = C.axis(0).L, Ly = C.axis(1).L, Lz = C.axis(2).L;
    double epart = 0.0;
    if (R.n > 0) row_prefetch<LY::ROWW>(R.rowbuf(0), grows, lane);
    for (int a = 0; a < R.n; ++a) {
        row_wait();
        if (a + 1 < R.n) row_prefetch<LY::ROWW>(R.rowbuf((a + 1) & 1), grows + (size_t)(a + 1) * LY::ROWW, lane);
        const uint32_t *row = R.rowbuf(a & 1);
        const double xa = R.x(a), ya = R.y(a), za = R.z(a);
        const int len = (int)row[0];
        int cnt = 0;
        for (int e0 = 0; e0 < len; e0 += 32) {
            const int e = e0 + lane;
            const bool valid = e < len;
            const uint32_t w = valid ? row[ROW_HDR + e] : 0u;
            int j;
            double sx, sy, sz;
            row_decode(w, j, sx, sy, sz);
            const double dx = fma(-sx, Lx, R.x(j) - xa), dy = fma(-sy, Ly, R.y(j) - ya), dz = fma(-sz, Lz, R.z(j) - za);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = valid && r2 < pot.cut_pre2;
            const unsigned b = __ballot_sync(FULL, in);
            if (in) R.sl(cnt + __popc(b & lt)) = r2;
            cnt += __popc(b);
        }
        __syncwarp();
        DeltaAcc acc = {0.0, 0.0, 0};
        eval_slots<RepR<LY>, false>(pot, R, lane, cnt, acc);
        pairs += acc.pairs;
        const double vs = warp_sum(acc.v);
        if (LY::EMT) {
            const double s1 = warp_sum(acc.snew);
            if (lane == 0) {
                R.sig(a) = s1;
                R.eat(a) = vs;  // sigma_2 sum parked here until the embedding pass below
            }
        } else if (lane == (a & 31)) {
            epart += 0.5 * vs;
        }
        __syncwarp();
    }
    if (LY::EMT) {
        for (int a = lane; a < R.n; a += 32) {
            const double ea = emt_embed(pot, R.sig(a));
            epart += ea + pot.neghalfv0overgamma2 * R.eat(a);
            R.eat(a) = ea;
        }
        __syncwarp();
    }
    return FullEnergy{warp_sum(epart), pairs};
}

// the draws of DRAW_AHEAD attempts into group `g` of the double-buffered draw area
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
template <int POT, class RV>
__device__ __noinline__ void replica_to_global(const qmc_batch_t &b, const RV R, const int r, const int lane) {
    const int nmax = b.n_max;
    double *gx = b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = lane; j < R.n; j += 32) {
        gx[j] = R.x(j);
        gy[j] = R.y(j);
        gz[j] = R.z(j);
        if (POT == QMC_POT_EMT) {
            b.sigma1[(size_t)r * nmax + j] = R.sig(j);
            b.eatom[(size_t)r * nmax + j] = R.eat(j);
        }
    }
    __syncwarp();
}

template <int POT, class RV>
__device__ __noinline__ void replica_from_global(const qmc_batch_t &b, const RV R, const int r, const int lane) {
    const int nmax = b.n_max;
    const double *gx = b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? __ldcg(gx + j) : 0.0;
        R.y(j) = in ? __ldcg(gy + j) : 0.0;
        R.z(j) = in ? __ldcg(gz + j) : 0.0;
        if (POT == QMC_POT_EMT) {
            R.sig(j) = in ? __ldcg(b.sigma1 + (size_t)r * nmax + j) : 0.0;
            R.eat(j) = in ? __ldcg(b.eatom + (size_t)r * nmax + j) : 0.0;
        }
    }
    __syncwarp();
}

struct RowArgs {
    double bound2;  // (skin / 2)^2: reach of a row without cell moves
    double skin;    // Angstrom
    double cutb;    // cut_pre + skin: the rows hold every image within this distance (build metric)
};

template <int POT, int NS, int FEAT>
__global__ void __launch_bounds__(LayoutR<POT, NS>::WARPS * 32, 1) sweep_rows_kernel(const __grid_constant__ SweepArgs a, const RowArgs ra) {
    using LY = LayoutR<POT, NS>;
    constexpr int WARPS = LY::WARPS;
    constexpr bool UNIFORM = (FEAT & F_CELLVAR) == 0;
    constexpr bool CELLMOVES = (FEAT & F_CELL) != 0;
    static_assert((FEAT & (F_EXCHANGE | F_COMPOSITE)) == 0, "the row-based kernel runs displacement and cell moves only");
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    RepR<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    {
        unsigned sb = (unsigned)__cvta_generic_to_shared(R.b);
        asm volatile("" : "+r"(sb));
        R.b = reinterpret_cast<double *>(__cvta_shared_to_generic((size_t)sb));
    }
    const CellView<LY, UNIFORM> C{&a.cell, R.b + LY::CELLC};
    const int r = blockIdx.x * WARPS + warp;
    if (r >= a.b.n_replicas) return;

    // ---- load the replica ----
    R.n = __ldcg(a.b.n_atoms + r);
    replica_from_global<POT>(a.b, R, r, lane);
    int status = 0;
    double cell[3] = {a.cell.L[0], a.cell.L[1], a.cell.L[2]};
    if (!UNIFORM) {
        cell[0] = __ldcg(a.b.cell + (size_t)r * 9 + 0);
        cell[1] = __ldcg(a.b.cell + (size_t)r * 9 + 4);
        cell[2] = __ldcg(a.b.cell + (size_t)r * 9 + 8);
        if (!store_cell(R, cell, a.b.pbc, pot.cut_pre, lane)) status |= QMC_STATUS_CELL_TOO_SMALL;
    }
    double E = __ldcg(a.b.energy + r);
    const double inv_kT = 1.0 / (a.b.temperature[r] * KB);
    if (lane < QMC_MAX_MOVES * 3) R.counts()[lane] = 0u;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    unsigned n_pairs = 0;  // per-lane partial of this launch
    __syncwarp();
    if (FEAT & F_LABELS) rebuild_movable(R, a.mv, a.n_moves, r, nmax, LY::LABW);

    // ---- neighbour rows: state, metric, (re)build ----
    // Scalars that live as long as the launch but are touched rarely sit in shared memory (ps), not in registers:
    // ps[0..2] = q_k = L_build,k / L_now,k, ps[3] = bound^2, ps[4] = largest squared committed displacement (build metric),
    // ps[5..7] = build cell.  Without cell moves q = 1, the bound is a kernel argument and every committed position has passed
    // the trial test, so none of them is needed.
    auto row_of = [&](const int atom) { return a.b.nbr_rows + ((size_t)opaque(r) * nmax + atom) * LY::ROWW; };
    auto meta_of = [&]() { return a.b.nbr_ref + (size_t)opaque(r) * 8; };  // build cell [3], largest squared committed displacement
    const bool have_rows = __ldcg(a.b.nbr_state + r) != 0;
    if (CELLMOVES && lane == 0) {
        const double *meta = meta_of();
        R.ps(4) = have_rows ? __ldcg(meta + 3) : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) R.ps(5 + k) = have_rows ? __ldcg(meta + k) : cell[k];
    }
    __syncwarp();
    // metric of the build cell seen from the current cell; slack S = skin - (max q - 1) cutoff; bound = S / 2
    auto set_metric = [&](const double Lnow[3]) {
        if (!CELLMOVES) return;
        if (lane == 0) {
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
        __syncwarp();
    };
    auto bound2 = [&]() { return CELLMOVES ? R.ps(3) : ra.bound2; };
    auto in_reach = [&]() { return !CELLMOVES || R.ps(4) <= R.ps(3); };  // every committed atom within the bound
    auto rebuild = [&](const double Lnow[3]) {
        // the rows need cutoff + skin <= L (no self-images) in the cell they are built in
        for (int k = 0; k < 3; ++k)
            if (a.b.pbc[k] && !(Lnow[k] >= ra.cutb)) status |= QMC_STATUS_CELL_TOO_SMALL;
        status |= build_rows<LY>(R, C, lane, row_of(0), ra.cutb);
        if (lane == 0) {
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
        __syncwarp();
        set_metric(Lnow);
    };
    set_metric(cell);
    if (!have_rows || !in_reach()) rebuild(cell);

    // ---- the attempt one ahead: move and atom, so that its row can be fetched early ----
    auto select_move = [&](const double u_select) {
        int m = 0;
        while (m < a.n_moves - 1 && a.cdf[m] <= u_select) ++m;
        return m;
    };
    auto select_atom = [&](const int m, const double u_label) {
        if (a.mv[m].type != QMC_MOVE_DISPLACEMENT) return -1;
        const int nu = (FEAT & F_LABELS) ? count_bits(R.labbits(m), nullptr, LY::LABW) : R.n;
        if (nu <= 0) return -1;
        const int idx = (int)(u_label * (double)nu);
        return (FEAT & F_LABELS) ? select_bit(R.labbits(m), nullptr, LY::LABW, idx) : idx;
    };
    const int k_end = a.n_attempts;
    draw_group(R, key, a.attempt_base, grep, lane, 0);
    int m_cur, i_cur;
    {
        const double2 d0 = R.draws()[0];
        m_cur = select_move(d0.x);
        i_cur = select_atom(m_cur, d0.y);
    }
    if (i_cur >= 0) row_prefetch<LY::ROWW>(R.rowbuf(0), row_of(i_cur), lane);

    for (int k = 0; k < k_end; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        const int slot = k & (DRAW_AHEAD - 1), grp = (k / DRAW_AHEAD) & 1;
        if (slot == 0 && k + DRAW_AHEAD < k_end)  // the group after this one (the launch's first group exists already)
            draw_group(R, key, attempt + DRAW_AHEAD, grep, lane, grp ^ 1);
        const double2 *dk = R.draws() + (grp * DRAW_AHEAD + slot) * 4;
        const double2 b1 = dk[1], b2 = dk[2];
        const double o0 = b1.x, o1 = b1.y, o2 = b2.x, u_accept = b2.y;
        const int m = m_cur, i = i_cur;
        const MoveDev &mv = a.mv[m];
        // next attempt: selection now, prefetch below (after this attempt's row is safe)
        int m_next = 0, i_next = -1;
        if (k + 1 < k_end) {
            const int k1 = k + 1;
            const double2 dn = R.draws()[(((k1 / DRAW_AHEAD) & 1) * DRAW_AHEAD + (k1 & (DRAW_AHEAD - 1))) * 4];
            m_next = select_move(dn.x);
            i_next = select_atom(m_next, dn.y);
        }

        int accepted = -1, moved = -1;
        double delta = __longlong_as_double(0x7ff8000000000000LL);  // NaN

        if (mv.type == QMC_MOVE_DISPLACEMENT) {
            if (i >= 0) {
                row_wait();
                const uint32_t *row = R.rowbuf(k & 1);
                double po[3], pn[3], d2;
                // The trial position must stay within the bound of the atom's build position (build metric).  If it does not, the
                // rows are rebuilt (then the build position IS the old position) and the proposal is simply evaluated again, so that
                // nothing but the launch's persistent state is alive across the (rare) rebuild call.
                for (int tries = 0;; ++tries) {
                    double t[3];
                    propose_translation(mv.op, mv.step, o0, o1, o2, t);
                    po[0] = R.x(i), po[1] = R.y(i), po[2] = R.z(i);
                    pn[0] = po[0] + t[0], pn[1] = po[1] + t[1], pn[2] = po[2] + t[2];
                    const double *ref = reinterpret_cast<const double *>(row + 2);
                    const double ux = CELLMOVES ? fma(pn[0], R.ps(0), -ref[0]) : pn[0] - ref[0];
                    const double uy = CELLMOVES ? fma(pn[1], R.ps(1), -ref[1]) : pn[1] - ref[1];
                    const double uz = CELLMOVES ? fma(pn[2], R.ps(2), -ref[2]) : pn[2] - ref[2];
                    d2 = fma(uz, uz, fma(uy, uy, ux * ux));
                    if (d2 <= bound2()) break;
                    if (tries) {  // step > skin / 2: refused by the library beforehand
                        status |= QMC_STATUS_LIST_OVERFLOW;
                        break;
                    }
                    rebuild(cell);
                    row_prefetch<LY::ROWW>(R.rowbuf(k & 1), row_of(i), lane);
                    row_wait();
                }
                if (i_next >= 0) row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
                const Trial T = local_delta_rows<LY>(pot, R, C, lane, lt, row, i, po, pn);
                n_pairs += T.pairs;
                delta = T.dE;
                const bool acc = metropolis(u_accept, -delta * inv_kT, status);
                if (acc) {
                    commit_rows<LY>(R, lane, T);
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
                    if (CELLMOVES && lane == 0) R.ps(4) = fmax(R.ps(4), d2);
                }
                __syncwarp();
                accepted = acc ? 1 : 0;
                moved = i;
            } else if (i_next >= 0) {
                row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
            }
        } else if (CELLMOVES && mv.type == QMC_MOVE_CELL) {
            // snapshot the last accepted state in global memory, deform in place, recompute everything through the rows
            replica_to_global<POT>(a.b, R, opaque(r), lane);
            const double s = exp(__dadd_rn(-mv.step, __dmul_rn(mv.step - -mv.step, o0)));
            double ncell[3], scale[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                ncell[d] = (((mv.axes >> d) & 1) ? s : 1.0) * cell[d];
                scale[d] = ncell[d] / cell[d];
            }
            for (int j = lane; j < R.n; j += 32) {
                R.x(j) = R.x(j) * scale[0];
                R.y(j) = R.y(j) * scale[1];
                R.z(j) = R.z(j) * scale[2];
            }
            __syncwarp();
            const bool cell_ok = store_cell(R, ncell, a.b.pbc, pot.cut_pre, lane);
            if (!cell_ok) status |= QMC_STATUS_CELL_TOO_SMALL;
            double e_new = E;
            if (cell_ok) {
                set_metric(ncell);
                if (!in_reach()) rebuild(ncell);  // the scaled configuration left the rows' reach
                const FullEnergy fe = full_energy_rows<LY>(pot, R, C, lane, lt, row_of(0));
                e_new = fe.e;
                n_pairs += (unsigned)fe.pairs;
            }
            delta = e_new - E;
            const double v_new = fabs((ncell[0] * ncell[1]) * ncell[2]), v_old = fabs((cell[0] * cell[1]) * cell[2]);
            const double kT = a.b.temperature[r] * KB;
            const double x = -(delta + a.b.pressure * (v_new - v_old)) / kT + (double)(R.n + 1) * log(v_new / v_old);
            const bool acc = cell_ok && metropolis(u_accept, x, status);
            if (acc) {
                E = e_new;
                cell[0] = ncell[0];
                cell[1] = ncell[1];
                cell[2] = ncell[2];
            } else {
                replica_from_global<POT>(a.b, R, opaque(r), lane);
                store_cell(R, cell, a.b.pbc, pot.cut_pre, lane);
            }
            __syncwarp();
            set_metric(cell);
            if (!in_reach()) rebuild(cell);
            if (i_next >= 0) row_prefetch<LY::ROWW>(R.rowbuf((k + 1) & 1), row_of(i_next), lane);
            accepted = acc ? 1 : 0;
        }

        if (lane == 0) {
            uint32_t *cnt = R.counts() + m * 3;
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
        m_cur = m_next;
        i_cur = i_next;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // ---- store the replica ----
    __syncwarp();
    replica_to_global<POT>(a.b, R, r, lane);
    if (lane < QMC_MAX_MOVES * 3) {
        const uint32_t c = R.counts()[lane];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(a.b.counters) + (size_t)r * QMC_MAX_MOVES * 3 + lane, (unsigned long long)c);
    }
    if (lane == 0) {
        a.b.energy[r] = E;
        if (CELLMOVES) {
            a.b.cell[(size_t)r * 9 + 0] = cell[0];
            a.b.cell[(size_t)r * 9 + 4] = cell[1];
            a.b.cell[(size_t)r * 9 + 8] = cell[2];
        }
        if (CELLMOVES) meta_of()[3] = R.ps(4);
        if (status) atomicOr(a.b.status + r, status);
    }
    if (a.b.pair_evals) {
        long long np = n_pairs;
        for (int o = 16; o > 0; o >>= 1) np += __shfl_xor_sync(FULL, np, o);
        if (lane == 0) a.b.pair_evals[r] += np;
    }
}

}  // namespace qmc
