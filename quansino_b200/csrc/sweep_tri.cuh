// Sweeps in GENERAL (triclinic) cells: single-atom displacement moves and cell moves with the three deformation operations of
// the reference (operations/cell.py:56-169: IsotropicDeformation, AnisotropicDeformation, ShapeDeformation), judged by
// IsobaricCriteria or IsotensionCriteria (mc/criteria.py:143-219; driver mc/isotension.py:27-121).
//
// One warp per replica with the replica in shared memory, exactly like csrc/sweep_kernel.cuh, and the SAME local-delta, commit and
// full-energy building blocks (csrc/replica_warp.cuh) -- only the image arithmetic differs: fractional coordinates through H^-1,
// rounding, and up to 2 x 2 x 2 images per pair (TriView).  A cell move recomputes everything (the reference does, too).
//
// Plain single-atom exchange moves (moves/exchange.py:88-176 with Translation placement) run here as well.
//
// Not in this kernel (the host refuses them for triclinic batches): molecular and composite exchange, composite moves, label
// groups, forced moves, neighbour rows, the Hamiltonian / ForceBias force kernels.
#pragma once

#include "sweep_kernel.cuh"

namespace qmc {

constexpr int TRI_T_DOUBLES = 12;  // H^-1[9], thr[3]; H[9] sits where the diagonal kernels keep their cell constants (LY::CELLC)

// The general-cell kernels run no composite, forced or label-bitmap code: those regions of the replica layout (from DISP_B to the
// end) hold H^-1 and the thresholds, so that the kernels keep the scan kernel's geometry (28 replicas per SM for 108 atoms --
// with 27, 4096 replicas need a second wave of 4 CTAs that doubles the launch time).  Layouts without room put them behind the
// replicas and give up warps.
template <class LY>
__host__ __device__ constexpr bool tri_inplace() {
    return LY::BYTES - LY::DISP_B >= TRI_T_DOUBLES * 8 && LY::DISP_B % 8 == 0;
}

template <class LY>
__host__ __device__ constexpr int tri_warps() {
    int w = LY::WARPS;
    if constexpr (!tri_inplace<LY>())
        while (w > 1 && w * (LY::BYTES + TRI_T_DOUBLES * 8) > SMEM_PER_CTA_MAX) --w;
    return w;
}

template <class LY>
__host__ __device__ constexpr size_t tri_smem() {
    return (size_t)tri_warps<LY>() * (LY::BYTES + (tri_inplace<LY>() ? 0 : TRI_T_DOUBLES * 8));
}

// shared-memory homes of the cell constants of warp `warp`
template <class LY>
__device__ __forceinline__ double *tri_t_ptr(double *smem_d, double *rb, int warp) {
    if (tri_inplace<LY>()) return reinterpret_cast<double *>(reinterpret_cast<char *>(rb) + LY::DISP_B);
    return smem_d + tri_warps<LY>() * (LY::BYTES / 8) + warp * TRI_T_DOUBLES;
}

__host__ __device__ inline double det3(const double h[9]) {
    return h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]);
}

__host__ __device__ inline void inv3(const double h[9], double inv[9]) {
    const double d = 1.0 / det3(h);
    inv[0] = (h[4] * h[8] - h[5] * h[7]) * d;
    inv[1] = (h[2] * h[7] - h[1] * h[8]) * d;
    inv[2] = (h[1] * h[5] - h[2] * h[4]) * d;
    inv[3] = (h[5] * h[6] - h[3] * h[8]) * d;
    inv[4] = (h[0] * h[8] - h[2] * h[6]) * d;
    inv[5] = (h[2] * h[3] - h[0] * h[5]) * d;
    inv[6] = (h[3] * h[7] - h[4] * h[6]) * d;
    inv[7] = (h[1] * h[6] - h[0] * h[7]) * d;
    inv[8] = (h[0] * h[4] - h[1] * h[3]) * d;
}

__host__ __device__ inline void matmul3(const double a[9], const double b[9], double c[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}

// heights of the cell: w_k = |det H| / |a_(k+1) x a_(k+2)|
__host__ __device__ inline void cell_heights(const double h[9], double w[3]) {
    const double vol = fabs(det3(h));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double *p = h + 3 * ((k + 1) % 3), *q = h + 3 * ((k + 2) % 3);
        const double cx = p[1] * q[2] - p[2] * q[1], cy = p[2] * q[0] - p[0] * q[2], cz = p[0] * q[1] - p[1] * q[0];
        w[k] = vol / sqrt(cx * cx + cy * cy + cz * cz);
    }
}

// exp of a 3 x 3 matrix (the reference calls scipy.linalg.expm, operations/cell.py:96,141): scaling and squaring around a Taylor
// polynomial of degree 14 -- with the norm scaled below 1/8 the truncation error is 8^-15 / 15! ~ 2e-26
__host__ __device__ inline void expm3(const double A[9], double E[9]) {
    double nrm = 0.0;
    for (int i = 0; i < 3; ++i) nrm = fmax(nrm, fabs(A[3 * i]) + fabs(A[3 * i + 1]) + fabs(A[3 * i + 2]));
    int s = 0;
    double scale = 1.0;
    while (nrm * scale > 0.125 && s < 60) {
        scale *= 0.5;
        ++s;
    }
    double B[9], P[9], T[9];
    for (int q = 0; q < 9; ++q) B[q] = A[q] * scale;
    for (int q = 0; q < 9; ++q) P[q] = (q % 4 == 0) ? 1.0 : 0.0;
    for (int n = 14; n >= 1; --n) {  // Horner: P = I + B P / n
        matmul3(B, P, T);
        for (int q = 0; q < 9; ++q) P[q] = ((q % 4 == 0) ? 1.0 : 0.0) + T[q] / (double)n;
    }
    for (int t = 0; t < s; ++t) {
        matmul3(P, P, T);
        for (int q = 0; q < 9; ++q) P[q] = T[q];
    }
    for (int q = 0; q < 9; ++q) E[q] = P[q];
}

// cell constants of one replica into shared memory (TriView); false if a periodic height is below the cutoff
__device__ inline bool store_cell_tri(const TriView &V, const double H[9], const int pbc[3], double cut_pre, int lane) {
    double *h = const_cast<double *>(V.h), *t = const_cast<double *>(V.t);
    double inv[9], w[3];
    inv3(H, inv);
    cell_heights(H, w);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (pbc[k] && !(w[k] >= cut_pre)) ok = false;
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) h[q] = H[q];
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int k = 0; k < 3; ++k) t[3 * x + k] = pbc[k] ? inv[3 * x + k] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) t[9 + k] = pbc[k] ? 1.0 - cut_pre / w[k] - 1e-9 : 1e300;  // (the margin only adds candidates)
    }
    __syncwarp();
    return ok;
}

// deformation gradient of a cell move from its draws (operations/cell.py:67-169); c[0..5] are the six uniform draws in call order
__device__ inline void deformation_gradient(int op, double step, int mask, const double u[6], double F[9]) {
    if (op == QMC_OP_ISOTROPIC) {  // eye * exp(U(-max, max)) * mask + eye * ~mask; `mask` = the diagonal's three bits
        const double sc = exp(__dadd_rn(-step, __dmul_rn(step - -step, u[0])));
#pragma unroll
        for (int q = 0; q < 9; ++q) F[q] = (q % 4 == 0) ? (((mask >> (q / 4)) & 1) ? sc : 1.0) : 0.0;
        return;
    }
    double c[6], T[9], E[9];
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = __dadd_rn(-step, __dmul_rn(step - -step, u[q]));
    const double mean = op == QMC_OP_SHAPE ? __ddiv_rn(__dadd_rn(__dadd_rn(c[0], c[1]), c[2]), 3.0) : 0.0;  // volume-preserving: traceless
    T[0] = __dsub_rn(c[0], mean);
    T[4] = __dsub_rn(c[1], mean);
    T[8] = __dsub_rn(c[2], mean);
    T[1] = T[3] = c[3];
    T[2] = T[6] = c[4];
    T[5] = T[7] = c[5];
    expm3(T, E);
#pragma unroll
    for (int q = 0; q < 9; ++q) F[q] = ((mask >> q) & 1) ? E[q] : ((q % 4 == 0) ? 1.0 : 0.0);  // expm(T) * mask + eye * ~mask
}

// One cell move of one replica (moves/cell.py:77-91, operations/cell.py:67-169, mc/criteria.py:160-219).  Kept out of line: its
// 3 x 3 temporaries would otherwise weigh on the registers of the displacement path.  The last accepted cell is read from the
// replica's constants in shared memory; an accepted cell is written through to global memory at once.
template <int POT, int NS>
__device__ __noinline__ bool cell_move_tri(const SweepArgs &a, const Rep<Layout<POT, NS>> &R, const TriView &C, const MoveDev &mv, const Draws &D,
                                           unsigned long long attempt, int r, int lane, double &E, double &delta, long long &n_pairs,
                                           int &status) {
    using LY = Layout<POT, NS>;
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    const double kT = a.b.temperature[r] * KB;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    double *gs = POT == QMC_POT_EMT ? a.b.sigma1 + (size_t)r * nmax : nullptr;
    double *ge = POT == QMC_POT_EMT ? a.b.eatom + (size_t)r * nmax : nullptr;
    double H[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) H[q] = C.h[q];
    __syncwarp();
    // moves/cell.py:77-91: cell' = F @ cell, positions' = positions @ solve(cell, cell') (Atoms.set_cell(scale_atoms=True));
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
    double u[6] = {D.o0, D.o1, D.o2, 0.0, 0.0, 0.0};
    if (mv.op != QMC_OP_ISOTROPIC) {  // components 3..5: the attempt's extra draws 0..2
        double e2, e3;
        uniform_pair(key, attempt, grep, 64u, u[3], u[4]);
        uniform_pair(key, attempt, grep, 65u, e2, e3);
        u[5] = e2;
    }
    double F[9], N[9], inv[9], M[9];
    deformation_gradient(mv.op, mv.step, mv.axes, u, F);
    matmul3(F, H, N);
    inv3(H, inv);
    matmul3(inv, N, M);
    for (int j = lane; j < R.n; j += 32) {
        const double x = R.x(j), y = R.y(j), z = R.z(j);
        R.x(j) = (x * M[0] + y * M[3]) + z * M[6];
        R.y(j) = (x * M[1] + y * M[4]) + z * M[7];
        R.z(j) = (x * M[2] + y * M[5]) + z * M[8];
    }
    __syncwarp();
    const bool cell_ok = store_cell_tri(C, N, a.b.pbc, pot.cut_pre, lane);
    if (!cell_ok) status |= QMC_STATUS_CELL_TOO_SMALL;
    long long fp = 0;
    const double e_new = cell_ok ? replica_full_energy<LY>(pot, R, C, lane, fp, status) : E;
    n_pairs += fp;
    delta = e_new - E;
    const double v_new = fabs(det3(N)), v_old = fabs(det3(H));
    double work = a.b.pressure * (v_new - v_old);
    if (a.b.isotension) {
        // IsotensionCriteria (mc/criteria.py:196-213), literally: strain = 0.5 (inv(old^T) @ new^T @ old @ inv(old) - 1),
        // elastic = p dV + V_old trace((stress - p) @ strain) with the SCALAR p subtracted from every component
        double HT[9], NT[9], iHT[9], P[9], Q[9], S[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            HT[q] = H[3 * (q % 3) + q / 3];
            NT[q] = N[3 * (q % 3) + q / 3];
        }
        inv3(HT, iHT);
        matmul3(iHT, NT, P);
        matmul3(P, H, Q);
        matmul3(Q, inv, S);
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double strain_ji = 0.5 * (S[3 * j + i] - (i == j ? 1.0 : 0.0));
                tr += (a.b.external_stress[3 * i + j] - a.b.pressure) * strain_ji;
            }
        work += v_old * tr;
    }
    const double x = -(delta + work) / kT + (double)(R.n + 1) * log(v_new / v_old);
    const bool acc = cell_ok && metropolis(D.u_accept, x, status);
    if (acc) {
        E = e_new;
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < 9; ++q) a.b.cell[(size_t)r * 9 + q] = N[q];
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
        store_cell_tri(C, H, a.b.pbc, pot.cut_pre, lane);
    }
    __syncwarp();
    return acc;
}

// One plain exchange move of one replica in a general cell (moves/exchange.py:134-176, mc/criteria.py:225-277): the branch of
// csrc/sweep_kernel.cuh with `uniform(0, 1, (1, 3)) @ cell` for a full cell matrix.  Out of line like the cell move.
template <int POT, int NS>
__device__ __noinline__ int exchange_move_tri(const SweepArgs &a, const Rep<Layout<POT, NS>> &Rc, const TriView &C, const MoveDev &mv, const Draws &D,
                                              int r, int lane, double &E, double &delta, long long &n_pairs, int &n_ex, int &n_now,
                                              int &moved, int &status) {
    using LY = Layout<POT, NS>;
    Rep<LY> R = Rc;
    R.n = n_now;
    const PotDev &pot = a.pot;
    const unsigned lt = lanemask_lt();
    const int nmax = a.b.n_max;
    const int *lab = mv.labels ? mv.labels + (size_t)r * nmax : nullptr;
    const bool is_add = D.u_coin < mv.bias;
    int pd = 0, i = -1;
    Trial T;
    double pn[3] = {0, 0, 0}, po[3] = {0, 0, 0};
    if (is_add) {
        if (R.n + 1 > nmax) {
            status |= QMC_STATUS_CAPACITY;
        } else {
            // Translation (operations/displacement.py:173-175): uniform(0,1,(1,3)) @ cell - mean(positions[moving])
            const double f[3] = {D.o0, D.o1, D.o2};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double fc = __dadd_rn(__dadd_rn(__dmul_rn(f[0], C.h[d]), __dmul_rn(f[1], C.h[3 + d])), __dmul_rn(f[2], C.h[6 + d]));
                pn[d] = __dadd_rn(a.b.exchange_pos[d], __dsub_rn(fc, a.b.exchange_pos[d]));
            }
            i = R.n;
            T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, false, true, po, pn, status);
            pd = 1;
        }
    } else {
        const int nu = lab ? count_labelled(lab, R.n) : R.n;
        if (nu > 0) {
            const int idx = (int)(D.u_label * (double)nu);
            i = lab ? select_labelled(lab, R.n, idx) : idx;
            po[0] = R.x(i);
            po[1] = R.y(i);
            po[2] = R.z(i);
            T = local_delta<LY>(pot, R, C, lane, lt, R.n, i, true, false, po, pn, status);
            pd = -1;
        }
    }
    if (pd == 0) return -1;
    n_pairs += T.pairs;
    delta = T.dE;
    bool ov;
    const double prob = gc_probability(delta, pd, n_ex, a.b.accessible_volume, a.b.exchange_mass, a.b.temperature[r], a.b.chemical_potential, ov);
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
    }
    __syncwarp();
    n_now = R.n;
    moved = i;
    return acc ? 1 : 0;
}

template <int POT, int NS>
__global__ void __launch_bounds__(tri_warps<Layout<POT, NS>>() * 32) sweep_tri_kernel(const __grid_constant__ SweepArgs a) {
    using LY = Layout<POT, NS>;
    constexpr int WARPS = tri_warps<LY>();
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const PotDev &pot = a.pot;
    const int nmax = a.b.n_max;
    Rep<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    const TriView C{R.b + LY::CELLC, tri_t_ptr<LY>(smem_d, R.b, warp)};
    const int r = blockIdx.x * WARPS + warp;
    if (r >= a.b.n_replicas) return;
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
    {
        double H[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) H[q] = __ldcg(a.b.cell + (size_t)r * 9 + q);
        if (!store_cell_tri(C, H, a.b.pbc, pot.cut_pre, lane)) status |= QMC_STATUS_CELL_TOO_SMALL;
    }
    double E = __ldcg(a.b.energy + r);
    const double kT = a.b.temperature[r] * KB;
    const double inv_kT = 1.0 / kT;
    if (lane < QMC_MAX_MOVES * 3) R.counts()[lane] = 0u;
    const uint32_t grep = (uint32_t)(a.b.replica_offset + r);
    const PhiloxKey key = {(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    long long n_pairs = 0;
    int n_ex = a.need_coin ? __ldcg(a.b.n_exchange + r) : 0;
    __syncwarp();

    for (int k = 0; k < a.n_attempts; ++k) {
        const unsigned long long attempt = a.attempt_base + (unsigned long long)k;
        const int ahead = k & (DRAW_AHEAD - 1);
        if (ahead == 0) draw_ahead(R, key, attempt, grep, lane);
        const Draws D = read_draws(R, ahead, a.need_coin != 0);
        int m = 0;
        while (m < a.n_moves - 1 && a.cdf[m] <= D.u_select) ++m;
        const MoveDev &mv = a.mv[m];
        int accepted = -1, moved = -1;
        double delta = __longlong_as_double(0x7ff8000000000000LL);

        if (mv.type == QMC_MOVE_DISPLACEMENT) {
            const int *lab = mv.labels ? mv.labels + (size_t)r * nmax : nullptr;
            const int nu = lab ? count_labelled(lab, R.n) : R.n;
            if (nu > 0) {
                const int idx = (int)(D.u_label * (double)nu);
                const int i = lab ? select_labelled(lab, R.n, idx) : idx;
                double t[3];
                propose_translation(mv.op, mv.step, D.o0, D.o1, D.o2, t);
                // CompositeOperation (operations/composite.py:58-73), as in csrc/sweep_kernel.cuh: the further operations draw the
                // attempt's extra draws e = 0, 1, ... = block 64 + e / 2, half e & 1; translations added in list order
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
        } else if (mv.type == QMC_MOVE_CELL) {
            accepted = cell_move_tri<POT, NS>(a, R, C, mv, D, attempt, r, lane, E, delta, n_pairs, status) ? 1 : 0;
        } else if (mv.type == QMC_MOVE_EXCHANGE) {
            accepted = exchange_move_tri<POT, NS>(a, R, C, mv, D, r, lane, E, delta, n_pairs, n_ex, R.n, moved, status);
        }  // (the host admits no other move types for general cells)

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
    }

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
        a.b.energy[r] = E;
        a.b.n_atoms[r] = R.n;
        if (a.need_coin) a.b.n_exchange[r] = n_ex;
        if (status) atomicOr(a.b.status + r, status);
    }
    if (a.b.pair_evals) {
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        if (lane == 0) a.b.pair_evals[r] += n_pairs;
    }
}

// validate_simulation in a general cell: full energy of every replica, caches primed
template <int POT, int NS>
__global__ void __launch_bounds__(tri_warps<Layout<POT, NS>>() * 32) energy_tri_kernel(const __grid_constant__ SweepArgs a, long long *pair_count) {
    using LY = Layout<POT, NS>;
    constexpr int WARPS = tri_warps<LY>();
    extern __shared__ __align__(16) double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + warp;
    if (r >= a.b.n_replicas) return;
    const int nmax = a.b.n_max;
    Rep<LY> R;
    R.b = smem_d + warp * (LY::BYTES / 8);
    const TriView C{R.b + LY::CELLC, tri_t_ptr<LY>(smem_d, R.b, warp)};
    R.n = a.b.n_atoms[r];
    const double *gx = a.b.pos + (size_t)r * 3 * nmax, *gy = gx + nmax, *gz = gy + nmax;
    for (int j = lane; j < nmax; j += 32) {
        const bool in = j < R.n;
        R.x(j) = in ? gx[j] : 0.0;
        R.y(j) = in ? gy[j] : 0.0;
        R.z(j) = in ? gz[j] : 0.0;
    }
    double H[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) H[q] = a.b.cell[(size_t)r * 9 + q];
    const bool ok = store_cell_tri(C, H, a.b.pbc, a.pot.cut_pre, lane);
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
