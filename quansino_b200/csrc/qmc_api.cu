// C ABI of quansino_b200 (include/quansino_b200.h): argument checking, launch geometry, kernel launches.
// No CPU fallback anywhere: without a CUDA device every compute entry point returns QMC_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/quansino_b200.h"
#include "hmc_kernel.cuh"
#include "fbmc_kernel.cuh"
#include "launchers.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver      \
                                               ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA,                            \
                                           "%s failed: %s", #expr, cudaGetErrorString(e_));                  \
    } while (0)

int make_potdev(const qmc_potential_t *p, qmc::PotDev &d) {
    std::memset(&d, 0, sizeof(d));
    d.kind = p->kind;
    if (p->kind == QMC_POT_EMT) {
        if (!(p->gamma1 > 0) || !(p->gamma2 > 0) || !(p->rc_list > 0) || !(p->beta > 0) || !(p->eta2 > 0))
            return fail(QMC_ERR_INVALID, "EMT parameters not initialised");
        d.cut = p->rc_list;
        d.acut = p->acut;
        d.rc = p->rc;
        d.rc_list = p->rc_list;
        d.eta2 = p->eta2;
        d.beta_s0 = p->beta * p->s0;
        d.kappa = p->kappa;
        d.inv_beta = 1.0 / p->beta;
        d.s0 = p->s0;
        d.lambda = p->lambda;
        d.E0 = p->E0;
        d.V0x6 = 6.0 * p->V0;
        d.inv12gamma1 = 1.0 / (12.0 * p->gamma1);
        d.neg_inv_beta_eta2 = -1.0 / (p->beta * p->eta2);
        d.neghalfv0overgamma2 = -0.5 * p->V0 / p->gamma2;
        d.kappa_over_beta = p->kappa / p->beta;
        d.theta_c = -p->acut * p->rc;
        d.sig1_a = -p->eta2;
        d.sig1_c = p->eta2 * p->beta * p->s0;
        d.sig2_a = -p->kappa / p->beta;
        d.sig2_c = p->kappa * p->s0;
    } else if (p->kind == QMC_POT_LJ) {
        if (!(p->sigma > 0) || !(p->lj_rc > 0)) return fail(QMC_ERR_INVALID, "LJ parameters not initialised");
        d.cut = p->lj_rc;
        d.sigma2 = p->sigma * p->sigma;
        d.eps4 = 4.0 * p->epsilon;
        d.lj_rc2 = p->lj_rc * p->lj_rc;
        d.lj_e0 = p->lj_e0;
    } else {
        return fail(QMC_ERR_INVALID, "unknown potential kind %d", p->kind);
    }
    d.cut_pre = d.cut * (1.0 + 1e-9);
    d.cut_pre2 = d.cut_pre * d.cut_pre;
    return QMC_OK;
}

int check_batch(const qmc_batch_t *b, int kind) {
    if (!b) return fail(QMC_ERR_INVALID, "batch is NULL");
    if (b->n_replicas < 0 || b->n_max <= 0) return fail(QMC_ERR_INVALID, "bad batch dimensions");
    if (b->n_max > 32000) return fail(QMC_ERR_UNSUPPORTED, "n_max > 32000 atoms per replica");
    if (!b->pos || !b->cell || !b->n_atoms || !b->energy || !b->temperature || !b->counters || !b->status)
        return fail(QMC_ERR_INVALID, "batch has NULL required pointers");
    if (kind == QMC_POT_EMT && (!b->sigma1 || !b->eatom)) return fail(QMC_ERR_INVALID, "EMT batch needs sigma1 / eatom caches");
    return QMC_OK;
}

struct Inst {
    int ns;
    qmc::sweep_launch_fn sweep;
    qmc::energy_launch_fn energy;
    qmc::sweep_launch_fn sweep_block;
    int block_minb;  // CTAs (= replicas) of the block kernel one SM holds
    qmc::rows_launch_fn sweep_rows;
    int row_words;   // u32 words per neighbour row
    int rows_wide_reps;  // replicas one SM holds with four warps per replica (0: not instantiated)
    qmc::tri_launch_fn sweep_tri;      // general (triclinic) cells
    qmc::energy_launch_fn energy_tri;
};

#define QMC_TABLE_ROW(POT, NS) \
    {NS, qmc::sweep_launch_##POT##_##NS, qmc::energy_launch_##POT##_##NS, qmc::sweep_block_launch_##POT##_##NS, qmc::LayoutB<POT, NS>::REPS, \
     qmc::sweep_rows_launch_##POT##_##NS, qmc::LayoutR<POT, NS, 1>::ROWW, NS >= 256 ? qmc::LayoutR<POT, NS, 4>::REPS : 0, \
     qmc::sweep_tri_launch_##POT##_##NS, qmc::energy_tri_launch_##POT##_##NS},
const Inst g_inst[2][7] = {{QMC_FOR_EACH_NS(QMC_TABLE_ROW, 0)}, {QMC_FOR_EACH_NS(QMC_TABLE_ROW, 1)}};

const Inst *find_inst(int kind, int n_max) {
    const int ns = qmc::capacity_class(n_max);
    if (ns == 0 || kind < 0 || kind > 1) return nullptr;
    for (const Inst &i : g_inst[kind])
        if (i.ns == ns) return &i;
    return nullptr;
}

// fills pbc flags and, for batches whose replicas all share one fixed cell, the cell constants in the arguments
int fill_cell_args(const qmc_batch_t *b, const qmc::PotDev &pd, qmc::SweepArgs &a) {
    const double diag[3] = {b->cell_diag[0], b->cell_diag[1], b->cell_diag[2]};
    const double one[3] = {1.0, 1.0, 1.0};
    const bool ok = qmc::cell_constants(b->cells_uniform && !b->triclinic ? diag : one, b->pbc, pd.cut_pre, a.cell);
    if (b->cells_uniform && !b->triclinic && !ok)
        return fail(QMC_ERR_UNSUPPORTED, "a periodic cell edge is shorter than the neighbour cutoff %.3f A", pd.cut);
    return QMC_OK;
}

__global__ void fp64_peak_kernel(double *out, int iters) {
    // 8 independent DFMA chains per thread, register resident
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000000001, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456) out[0] = a0;
}

__global__ void debug_math_kernel(int which, const double *in, double *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = in[i];
    double s, c;
    switch (which) {
        case 0: out[i] = qmc::exp_fast(x); break;
        case 1: out[i] = qmc::log_fast(x); break;
        case 2: out[i] = qmc::sqrt_fast(x); break;
        case 3: out[i] = qmc::rcp_fast(x); break;
        case 4: qmc::sincos_2pi(x, s, c); out[i] = s; break;
        default: qmc::sincos_2pi(x, s, c); out[i] = c; break;
    }
}

}  // namespace

extern "C" {

int qmc_abi_version(void) { return QMC_ABI_VERSION; }

const char *qmc_last_error(void) { return g_err; }

int qmc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(QMC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    return n;
}

int64_t qmc_sweep_smem_per_replica(const qmc_potential_t *pot, int32_t n_max) {
    if (!pot || n_max <= 0) return fail(QMC_ERR_INVALID, "bad arguments");
    const bool emt = pot->kind == QMC_POT_EMT;
    switch (qmc::capacity_class(n_max)) {
#define QMC_BYTES_CASE(POT, NS) case NS: return emt ? qmc::Layout<0, NS>::BYTES : qmc::Layout<1, NS>::BYTES;
        QMC_FOR_EACH_NS(QMC_BYTES_CASE, 0)
        default: return fail(QMC_ERR_UNSUPPORTED, "more than 1024 atoms per replica do not fit the warp-per-replica sweep kernel");
    }
}

int32_t qmc_nbr_row_words(const qmc_potential_t *pot, int32_t n_max) {
    if (!pot || n_max <= 0) return fail(QMC_ERR_INVALID, "bad arguments");
    const Inst *inst = find_inst(pot->kind, n_max);
    if (!inst) return fail(QMC_ERR_UNSUPPORTED, "more than 1024 atoms per replica do not fit the warp-per-replica sweep kernels");
    return inst->row_words;
}

int qmc_energy_full(const qmc_potential_t *pot, const qmc_batch_t *batch, int64_t *pair_count, void *stream) {
    return qmc_atom_energies(pot, batch, nullptr, pair_count, stream);
}

int qmc_atom_energies(const qmc_potential_t *pot, const qmc_batch_t *batch, double *atom_energies, int64_t *pair_count, void *stream) {
    if (!pot) return fail(QMC_ERR_INVALID, "potential is NULL");
    int rc;
    if ((rc = check_batch(batch, pot->kind))) return rc;
    if (batch->n_replicas == 0) return QMC_OK;
    qmc::SweepArgs a;
    std::memset(&a, 0, sizeof(a));
    if ((rc = make_potdev(pot, a.pot))) return rc;
    a.b = *batch;
    a.atom_energies = atom_energies;
    if ((rc = fill_cell_args(batch, a.pot, a))) return rc;
    const Inst *inst = find_inst(pot->kind, batch->n_max);
    if (!inst) return fail(QMC_ERR_UNSUPPORTED, "more than 1024 atoms per replica do not fit the warp-per-replica kernels");
    return (batch->triclinic ? inst->energy_tri : inst->energy)(a, (long long *)pair_count, (cudaStream_t)stream, g_err, sizeof(g_err));
}

int qmc_sweep(const qmc_potential_t *pot, const qmc_batch_t *batch, const qmc_move_t *moves, int32_t n_moves, uint64_t seed,
              uint64_t attempt_base, int32_t n_attempts, const qmc_trace_t *trace, void *stream) {
    if (!pot || !moves) return fail(QMC_ERR_INVALID, "potential / moves is NULL");
    int rc;
    if ((rc = check_batch(batch, pot->kind))) return rc;
    if (n_moves < 1 || n_moves > QMC_MAX_MOVES) return fail(QMC_ERR_INVALID, "n_moves must be in [1, %d]", QMC_MAX_MOVES);
    if (n_attempts < 0) return fail(QMC_ERR_INVALID, "n_attempts < 0");
    if (batch->n_replicas == 0 || n_attempts == 0) return QMC_OK;
    qmc::SweepArgs a;
    std::memset(&a, 0, sizeof(a));
    if ((rc = make_potdev(pot, a.pot))) return rc;
    a.b = *batch;
    bool any_exchange = false, any_composite = false, any_runs = false;
    double tot = 0.0;
    for (int m = 0; m < n_moves; ++m) {
        const qmc_move_t &mv = moves[m];
        const bool head = m == 0 || !(moves[m - 1].chain & 1);  // a composite move carries ONE probability, on its first entry
        if (head && !(mv.probability >= 0.0)) return fail(QMC_ERR_INVALID, "move %d: negative probability", m);
        if (head) tot += mv.probability;
        if (mv.repeat < 0) return fail(QMC_ERR_INVALID, "move %d: negative repeat", m);
        if (mv.chain < 0 || (mv.chain > 4 && mv.chain != 8))
            return fail(QMC_ERR_INVALID, "move %d: chain must be bits 0 / 1 (composite), bit 2 alone (label groups) or bit 3 alone (run-coded groups)", m);
        // an exchange entry with repeat = k > 1 is a CompositeExchangeMove of k single-atom exchange moves (run-coded groups only)
        const bool composite_exchange = mv.type == QMC_MOVE_EXCHANGE && mv.repeat > 1;
        if ((mv.chain & 8) && ((mv.type != QMC_MOVE_DISPLACEMENT && mv.type != QMC_MOVE_EXCHANGE) || !mv.labels || !head ||
                               (mv.repeat > 1 && !composite_exchange) || mv.n_extra_ops))
            return fail(QMC_ERR_UNSUPPORTED, "move %d: run-coded label groups need a single displacement or exchange move with a label array", m);
        if (composite_exchange && (!(mv.chain & 8) || mv.op != QMC_OP_TRANSLATION || batch->n_exchange_atoms > 1 || mv.repeat > QMC_MAX_SUBMOVES))
            return fail(QMC_ERR_UNSUPPORTED, "move %d: a composite exchange move (repeat > 1) needs run-coded label groups (chain bit 3), Translation "
                        "placement, a single-atom exchange species and at most %d sub-moves", m, QMC_MAX_SUBMOVES);
        if ((mv.chain & 4) && (mv.type != QMC_MOVE_DISPLACEMENT || !mv.labels || !head || mv.repeat > 1))
            return fail(QMC_ERR_UNSUPPORTED, "move %d: label groups need a single displacement move with a label array", m);
        if ((mv.chain & 1) && m + 1 == n_moves) return fail(QMC_ERR_INVALID, "move %d: chain bit 0 set on the last table entry", m);
        if (!head && (mv.chain & 2) != (moves[m - 1].chain & 2))
            return fail(QMC_ERR_INVALID, "move %d: chain bit 1 (plain CompositeMove) must be the same on every entry of a chain", m);
        // a CellMove may close a plain CompositeMove chain (the hybrid NPT move); everything else in a chain is a displacement
        const bool closing_cell = !head && mv.type == QMC_MOVE_CELL && (mv.chain & 2) && !(mv.chain & 1) && mv.repeat <= 1;
        if (((mv.chain & 7) || !head || (mv.repeat > 1 && !composite_exchange)) && mv.type != QMC_MOVE_DISPLACEMENT && !closing_cell)
            return fail(QMC_ERR_UNSUPPORTED, "move %d: a composite move is a chain of displacement moves, optionally closed by a cell move", m);
        if (!head && mv.type == QMC_MOVE_DISPLACEMENT && mv.labels != moves[m - 1].labels)
            return fail(QMC_ERR_UNSUPPORTED, "move %d: the sub-moves of a composite move must share one label array", m);
        any_composite |= mv.chain || mv.repeat > 1;  // (label groups, too, live in the general instantiation)
        any_runs |= (mv.chain & 8) != 0;
        if (mv.minimum_count < 0) return fail(QMC_ERR_INVALID, "move %d: negative minimum_count", m);
        if (head)
            for (int q = 0; q < mv.minimum_count; ++q) {
                if (a.n_forced >= QMC_MAX_FORCED) return fail(QMC_ERR_UNSUPPORTED, "at most %d forced moves per step on the GPU path", QMC_MAX_FORCED);
                a.forced_move[a.n_forced++] = m;
            }
        switch (mv.type) {
            case QMC_MOVE_DISPLACEMENT:
                if ((mv.op == QMC_OP_ROTATION || mv.op == QMC_OP_TRANSLATION || mv.op == QMC_OP_TRANSLATION_ROTATION) && !(mv.chain & 12))
                    return fail(QMC_ERR_UNSUPPORTED, "move %d: Rotation / Translation / TranslationRotation need label groups (chain bit 2 and a label array)", m);
                if (mv.op != QMC_OP_BALL && mv.op != QMC_OP_BOX && mv.op != QMC_OP_SPHERE && mv.op != QMC_OP_ROTATION &&
                    mv.op != QMC_OP_TRANSLATION && mv.op != QMC_OP_TRANSLATION_ROTATION)
                    return fail(QMC_ERR_UNSUPPORTED, "move %d: displacement operation %d is not available on the GPU path", m, mv.op);
                break;
            case QMC_MOVE_CELL:
                if (mv.axes < 0) return fail(QMC_ERR_INVALID, "move %d: negative axes mask", m);
                if (mv.op == QMC_OP_ANISOTROPIC || mv.op == QMC_OP_SHAPE) {
                    if (!batch->triclinic)
                        return fail(QMC_ERR_INVALID, "move %d: AnisotropicDeformation / ShapeDeformation leave general cells behind: set batch.triclinic", m);
                    if (mv.axes > 0x1FF) return fail(QMC_ERR_INVALID, "move %d: axes must be a mask of bits 0..8 (the 3 x 3 deformation mask)", m);
                } else if (mv.op != QMC_OP_ISOTROPIC) {
                    return fail(QMC_ERR_UNSUPPORTED, "move %d: cell operation %d is not available on the GPU path", m, mv.op);
                } else if (mv.axes > 7) {
                    return fail(QMC_ERR_INVALID, "move %d: axes must be a mask of bits 0..2", m);
                }
                break;
            case QMC_MOVE_EXCHANGE:
                if (mv.op != QMC_OP_TRANSLATION && !(mv.op == QMC_OP_TRANSLATION_ROTATION && (mv.chain & 8)))
                    return fail(QMC_ERR_UNSUPPORTED, "move %d: exchange moves place with Translation (TranslationRotation with run-coded groups)", m);
                if (batch->n_exchange_atoms < 0 || batch->n_exchange_atoms > 8)
                    return fail(QMC_ERR_INVALID, "batch.n_exchange_atoms must be in [0, 8]");
                if (batch->n_exchange_atoms > 1 && !(mv.chain & 8))
                    return fail(QMC_ERR_UNSUPPORTED, "move %d: a molecular exchange species needs run-coded label groups (chain bit 3) on every move", m);
                if (!batch->n_exchange) return fail(QMC_ERR_INVALID, "exchange move needs batch.n_exchange");
                any_exchange = true;
                break;
            case QMC_MOVE_HMC:
                return fail(QMC_ERR_INVALID, "Hamiltonian moves run through qmc_hmc, not qmc_sweep");
            default:
                return fail(QMC_ERR_INVALID, "move %d: unknown type %d", m, mv.type);
        }
        a.mv[m].type = mv.type;
        a.mv[m].op = mv.op;
        a.mv[m].step = mv.step;
        a.mv[m].bias = mv.bias_insert;
        a.mv[m].default_label = mv.default_label;
        a.mv[m].repeat = mv.repeat > 0 ? mv.repeat : 1;
        a.mv[m].chain = mv.chain;
        a.mv[m].axes = (mv.op == QMC_OP_ANISOTROPIC || mv.op == QMC_OP_SHAPE) ? ((mv.axes & 0x1FF) ? (mv.axes & 0x1FF) : 0x1FF) : ((mv.axes & 7) ? (mv.axes & 7) : 7);
        if (mv.n_extra_ops < 0 || mv.n_extra_ops > 2) return fail(QMC_ERR_UNSUPPORTED, "move %d: at most three operations per move on the GPU path", m);
        if (mv.n_extra_ops > 0 && (mv.type != QMC_MOVE_DISPLACEMENT || mv.chain || mv.repeat > 1 || mv.op == QMC_OP_ROTATION))
            return fail(QMC_ERR_UNSUPPORTED, "move %d: composite operations belong to plain single-atom displacement moves", m);
        a.mv[m].n_extra_ops = mv.n_extra_ops;
        for (int q = 0; q < mv.n_extra_ops; ++q) {
            if (mv.extra_op[q] != QMC_OP_BALL && mv.extra_op[q] != QMC_OP_BOX && mv.extra_op[q] != QMC_OP_SPHERE)
                return fail(QMC_ERR_UNSUPPORTED, "move %d: operation %d cannot be part of a composite operation on the GPU path", m, mv.extra_op[q]);
            a.mv[m].extra_op[q] = mv.extra_op[q];
            a.mv[m].extra_step[q] = mv.extra_step[q];
        }
        any_composite |= mv.n_extra_ops > 0;
        a.mv[m].labels = mv.labels;
    }
    if (!(tot > 0.0)) return fail(QMC_ERR_INVALID, "move probabilities sum to zero");
    // Random stream layout: drawing sub-move c >= 1 of a composite uses Philox blocks 32 + 2 (c - 1), 33 + 2 (c - 1) up to c = 16 and
    // 0x8000 + 2 (c - 17), + 1 beyond (csrc/sweep_kernel.cuh: submove_block); the extra draws (closing CellMove, CompositeOperation,
    // TranslationRotation angles) start at block 64.
    for (int m = 0, drawn = 0; m < n_moves; ++m) {
        const bool head = m == 0 || !(moves[m - 1].chain & 1);
        if (head) drawn = 0;
        if (moves[m].type == QMC_MOVE_DISPLACEMENT) drawn += moves[m].repeat > 0 ? moves[m].repeat : 1;
        if (drawn > QMC_MAX_COMPOSITE_DISPLACEMENTS)
            return fail(QMC_ERR_UNSUPPORTED, "move %d: a composite move may displace at most %d times per attempt on the GPU path", m,
                        QMC_MAX_COMPOSITE_DISPLACEMENTS);
    }
    if (a.n_forced > 0) {  // mc/core.py:161-165: the forced moves must fit into a step; steps are counted from attempt 0
        a.max_cycles = moves[0].max_cycles;
        if (a.max_cycles < a.n_forced) return fail(QMC_ERR_INVALID, "The number of forced moves exceeds the number of cycles.");
        if (attempt_base % (uint64_t)a.max_cycles != 0 || n_attempts % a.max_cycles != 0)
            return fail(QMC_ERR_INVALID, "with forced moves a launch must cover whole steps of max_cycles attempts");
        any_composite = true;  // forced moves live in the general instantiation of the warp-per-replica kernel
    }
    if (any_exchange)
        for (int m = 0; m < n_moves; ++m)
            if (moves[m].chain & 4) return fail(QMC_ERR_UNSUPPORTED, "rank-coded label groups cannot be combined with exchange moves: use run-coded groups (chain bit 3)");
    if (any_runs)
        for (int m = 0; m < n_moves; ++m)
            if (moves[m].type != QMC_MOVE_CELL && !(moves[m].chain & 8))
                return fail(QMC_ERR_INVALID, "run-coded label groups (chain bit 3) must be set on every displacement / exchange move of the table");
    if (any_exchange)
        for (int m = 0; m < n_moves; ++m)
            if (moves[m].type != QMC_MOVE_CELL && !moves[m].labels)
                return fail(QMC_ERR_INVALID, "with an exchange move every displacement / exchange move needs a label array");
    // mc/core.py:344-347 + numpy choice(p=): p /= sum(p); cdf = cumsum(p); cdf /= cdf[-1]
    double acc = 0.0;
    for (int m = 0; m < n_moves; ++m) {
        if (m == 0 || !(moves[m - 1].chain & 1)) acc += moves[m].probability / tot;
        a.cdf[m] = acc;  // the later entries of a chain repeat the head's value: searchsorted never lands on them
    }
    for (int m = 0; m < n_moves; ++m) a.cdf[m] /= acc;
    a.n_moves = n_moves;
    a.need_coin = any_exchange ? 1 : 0;
    a.seed = seed;
    a.attempt_base = attempt_base;
    a.n_attempts = n_attempts;
    if (trace) {
        a.tr = *trace;
        a.has_trace = 1;
    }
    int feat = batch->cells_uniform ? 0 : qmc::F_CELLVAR;
    for (int m = 0; m < n_moves; ++m) {
        if (moves[m].labels) feat |= qmc::F_LABELS;
        if (moves[m].type == QMC_MOVE_CELL) feat |= qmc::F_CELL | qmc::F_CELLVAR;
        if (moves[m].type == QMC_MOVE_EXCHANGE) feat |= qmc::F_EXCHANGE;
    }
    if (any_composite) feat |= qmc::F_COMPOSITE;
    // label tables, composite and forced moves run in instantiations that also contain the exchange code, which reads and writes
    // n_exchange unconditionally
    if ((feat & (qmc::F_LABELS | qmc::F_EXCHANGE | qmc::F_COMPOSITE)) && !batch->n_exchange)
        return fail(QMC_ERR_INVALID, "batch.n_exchange is required with label tables, composite, forced or exchange moves");
    if ((rc = fill_cell_args(batch, a.pot, a))) return rc;
    const Inst *inst = find_inst(pot->kind, batch->n_max);
    if (!inst) return fail(QMC_ERR_UNSUPPORTED, "more than 1024 atoms per replica do not fit the warp-per-replica sweep kernel");
    if (batch->triclinic) {
        // general cells (csrc/sweep_tri.cuh): plain single-atom displacement moves and cell moves
        if (a.n_forced > 0)
            return fail(QMC_ERR_UNSUPPORTED, "general (triclinic) cells: forced moves are not available on the GPU path");
        for (int m = 0; m < n_moves; ++m) {
            if (moves[m].chain || moves[m].repeat > 1)  // (composite OPERATIONS, n_extra_ops, are fine: they only add translations)
                return fail(QMC_ERR_UNSUPPORTED, "move %d: general (triclinic) cells: composite moves and label groups are not available on the GPU path", m);
            if (moves[m].type == QMC_MOVE_DISPLACEMENT && moves[m].op != QMC_OP_BALL && moves[m].op != QMC_OP_BOX && moves[m].op != QMC_OP_SPHERE)
                return fail(QMC_ERR_UNSUPPORTED, "move %d: general (triclinic) cells take Ball / Box / Sphere displacements of single atoms", m);
            if (moves[m].type == QMC_MOVE_EXCHANGE && (moves[m].op != QMC_OP_TRANSLATION || batch->n_exchange_atoms > 1))
                return fail(QMC_ERR_UNSUPPORTED, "move %d: general (triclinic) cells take single-atom exchange moves placed by Translation", m);
        }
        if (batch->nbr_state) CUDA_TRY(cudaMemsetAsync(batch->nbr_state, 0, (size_t)batch->n_replicas * sizeof(int32_t), (cudaStream_t)stream));
        return inst->sweep_tri(a, (cudaStream_t)stream, g_err, sizeof(g_err));
    }
    // one warp per replica fills the GPU only with thousands of replicas; smaller batches run CTA-per-replica
    int sms = 148, dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // (measured on B200: at 108 atoms a warp owns a single 32-atom tile and the block kernel loses, 1.5e8 vs 1.7e8 attempts/s at
    // 1024 replicas; at 500 atoms it wins 1.9x -- profiles/r1_e_kernel_choice.md)
    bool block = inst->ns >= 256 && batch->n_replicas <= 2 * sms * inst->block_minb;
    const char *force = std::getenv("QMC_FORCE_KERNEL");
    // Row-based kernel (csrc/sweep_rows.cuh): single-atom displacement moves and cell moves of a batch that carries neighbour
    // rows.  It needs cutoff + skin <= every periodic edge (no self-images) and step <= skin / 2 (a trial position must stay
    // inside the reach of the moved atom's row right after a rebuild).
    if (batch->nbr_rows && !(force && std::strcmp(force, "rows") != 0)) {
        bool ok = !any_exchange && !any_composite && (feat & qmc::F_EXCHANGE) == 0 && batch->nbr_ref && batch->nbr_state;
        double tmax = 0.0;
        for (int m = 0; m < n_moves && ok; ++m) {
            if (moves[m].type == QMC_MOVE_DISPLACEMENT) {
                ok = moves[m].op == QMC_OP_BALL || moves[m].op == QMC_OP_BOX || moves[m].op == QMC_OP_SPHERE;
                tmax = std::max(tmax, std::fabs(moves[m].step) * (moves[m].op == QMC_OP_BOX ? std::sqrt(3.0) : 1.0));
            } else if (moves[m].type != QMC_MOVE_CELL) {
                ok = false;
            }
        }
        qmc::RowArgs ra;
        ra.skin = batch->nbr_skin > 0.0 ? batch->nbr_skin : QMC_NBR_SKIN_DEFAULT;
        ra.cutb = a.pot.cut_pre + ra.skin;
        ra.bound2 = 0.25 * ra.skin * ra.skin;
        if (ok && !(tmax * (1.0 + 1e-9) <= 0.5 * ra.skin)) ok = false;
        if (ok && batch->cells_uniform)
            for (int k = 0; k < 3; ++k)
                if (batch->pbc[k] && !(batch->cell_diag[k] >= ra.cutb)) ok = false;
        if (ok) {
            if (batch->nbr_row_cap != inst->row_words)
                return fail(QMC_ERR_INVALID, "batch.nbr_row_cap is %d, the row-based kernel of this capacity class uses %d words (qmc_nbr_row_words)",
                            batch->nbr_row_cap, inst->row_words);
            // four warps per replica when one warp per replica would leave most warp slots of the GPU empty (C3: 1024 replicas of
            // 500 atoms = 7 per SM)
            int wide = inst->rows_wide_reps > 0 && batch->n_replicas <= 2 * sms * inst->rows_wide_reps;
            if (const char *w = std::getenv("QMC_ROWS_WIDE")) wide = inst->rows_wide_reps > 0 && std::atoi(w) != 0;
            return inst->sweep_rows(feat, wide, a, ra, (cudaStream_t)stream, g_err, sizeof(g_err));
        }
    }
    // every other kernel moves atoms without looking at the rows: they are stale from here on
    if (batch->nbr_state) CUDA_TRY(cudaMemsetAsync(batch->nbr_state, 0, (size_t)batch->n_replicas * sizeof(int32_t), (cudaStream_t)stream));
    if (force) block = std::strcmp(force, "block") == 0;
    if (any_composite) block = false;  // composite moves exist in the warp-per-replica kernel only
    return (block ? inst->sweep_block : inst->sweep)(feat, a, (cudaStream_t)stream, g_err, sizeof(g_err));
}

int qmc_forces_full(const qmc_potential_t *pot, const qmc_batch_t *batch, double *forces, void *workspace, int64_t workspace_bytes,
                    void *stream) {
    if (!pot || !forces) return fail(QMC_ERR_INVALID, "potential / forces is NULL");
    int rc;
    if ((rc = check_batch(batch, pot->kind))) return rc;
    if (batch->n_replicas == 0) return QMC_OK;
    qmc::PotDev pd;
    if ((rc = make_potdev(pot, pd))) return rc;
    if (!workspace || workspace_bytes < qmc::hmc_workspace_bytes(pd, *batch)) return fail(QMC_ERR_INVALID, "workspace too small");
    return qmc::hmc_forces(pd, *batch, forces, workspace, (cudaStream_t)stream, g_err, sizeof(g_err));
}

int64_t qmc_hmc_workspace_bytes(const qmc_potential_t *pot, const qmc_batch_t *batch) {
    if (!batch || !pot) return fail(QMC_ERR_INVALID, "potential / batch is NULL");
    qmc::PotDev pd;
    int rc;
    if ((rc = make_potdev(pot, pd))) return rc;
    return qmc::hmc_workspace_bytes(pd, *batch);
}

int qmc_hmc(const qmc_potential_t *pot, const qmc_batch_t *batch, const qmc_move_t *move, uint64_t seed, uint64_t attempt_base,
            int32_t n_attempts, const qmc_trace_t *trace, void *workspace, int64_t workspace_bytes, void *stream) {
    if (!pot || !move) return fail(QMC_ERR_INVALID, "potential / move is NULL");
    int rc;
    if ((rc = check_batch(batch, pot->kind))) return rc;
    if (move->type != QMC_MOVE_HMC || move->op != QMC_OP_VERLET)
        return fail(QMC_ERR_UNSUPPORTED, "qmc_hmc runs HamiltonianDisplacementMove + Verlet only");
    if (!batch->momenta || !batch->kinetic) return fail(QMC_ERR_INVALID, "Hamiltonian moves need batch.momenta and batch.kinetic");
    if (move->n_steps < 0) return fail(QMC_ERR_INVALID, "negative number of Verlet steps");
    qmc::PotDev pd;
    if ((rc = make_potdev(pot, pd))) return rc;
    if (workspace_bytes < qmc::hmc_workspace_bytes(pd, *batch) || !workspace) return fail(QMC_ERR_INVALID, "workspace too small");
    if (batch->n_replicas == 0 || n_attempts <= 0) return QMC_OK;
    qmc_trace_t tr;
    std::memset(&tr, 0, sizeof(tr));
    if (trace) tr = *trace;
    if (batch->nbr_state) CUDA_TRY(cudaMemsetAsync(batch->nbr_state, 0, (size_t)batch->n_replicas * sizeof(int32_t), (cudaStream_t)stream));
    return qmc::hmc_run(pd, *batch, move->dt, move->n_steps, seed, attempt_base, n_attempts, tr, workspace, (cudaStream_t)stream,
                        g_err, sizeof(g_err));
}

int qmc_fbmc_step(const qmc_potential_t *pot, const qmc_batch_t *batch, double delta, uint64_t seed, uint64_t step, double *forces,
                  int32_t forces_valid, double *gamma_max, int32_t *n_calls, void *workspace, int64_t workspace_bytes, void *stream) {
    return qmc_fbmc_step_v(pot, batch, delta, nullptr, seed, step, forces, forces_valid, gamma_max, n_calls, workspace, workspace_bytes,
                           stream);
}

int qmc_fbmc_step_v(const qmc_potential_t *pot, const qmc_batch_t *batch, double delta, const double *delta_per_replica, uint64_t seed,
                    uint64_t step, double *forces, int32_t forces_valid, double *gamma_max, int32_t *n_calls, void *workspace,
                    int64_t workspace_bytes, void *stream) {
    if (!pot || !forces) return fail(QMC_ERR_INVALID, "potential / forces is NULL");
    int rc;
    if ((rc = check_batch(batch, pot->kind))) return rc;
    if (!delta_per_replica && !(delta > 0.0)) return fail(QMC_ERR_INVALID, "delta must be positive");
    if (batch->n_replicas == 0) return QMC_OK;
    qmc::PotDev pd;
    if ((rc = make_potdev(pot, pd))) return rc;
    if (!workspace || workspace_bytes < qmc::hmc_workspace_bytes(pd, *batch)) return fail(QMC_ERR_INVALID, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (batch->nbr_state) CUDA_TRY(cudaMemsetAsync(batch->nbr_state, 0, (size_t)batch->n_replicas * sizeof(int32_t), s));
    if (!forces_valid && (rc = qmc::hmc_forces(pd, *batch, forces, workspace, s, g_err, sizeof(g_err)))) return rc;
    qmc::FbmcArgs a;
    std::memset(&a, 0, sizeof(a));
    a.b = *batch;
    a.forces = forces;
    a.delta = delta;
    a.delta_r = delta_per_replica;
    a.seed = seed;
    a.step = step;
    a.gamma_max = gamma_max;
    a.n_calls = n_calls;
    if ((rc = qmc::fbmc_launch(a, s, g_err, sizeof(g_err)))) return rc;
    return qmc::hmc_forces(pd, *batch, forces, workspace, s, g_err, sizeof(g_err));  // energy + forces at the new positions
}

int qmc_pt_swap(double *temperatures_all, const double *energies_all, int32_t n_global, uint64_t seed, uint64_t round,
                int8_t *accepted, void *stream) {
    if (!temperatures_all || !energies_all || n_global < 0) return fail(QMC_ERR_INVALID, "bad arguments");
    if (n_global < 2) return QMC_OK;
    return qmc::pt_swap(temperatures_all, energies_all, n_global, seed, round, accepted, (cudaStream_t)stream, g_err, sizeof(g_err));
}

// ---- multi-GPU part of the path behind the C ABI: NCCL is resolved at run time -- first among the symbols the host
// process has already loaded (the communicator must come from the same NCCL the calls go to), then from libnccl.so.2 -- so the
// library itself has no link-time dependency on it.
namespace {
struct Nccl {
    int (*all_gather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*comm_count)(void *, int *) = nullptr;
    int (*comm_rank)(void *, int *) = nullptr;
    const char *(*error_string)(int) = nullptr;
    bool ok = false;
};
const Nccl &nccl() {
    static Nccl n = [] {
        Nccl x;
        void *h = RTLD_DEFAULT;  // (a null handle: "not found" must be told apart by the lookup, not by the handle)
        if (!dlsym(h, "ncclAllGather")) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) return x;
        }
        x.all_gather = (decltype(x.all_gather))dlsym(h, "ncclAllGather");
        x.all_reduce = (decltype(x.all_reduce))dlsym(h, "ncclAllReduce");
        x.comm_count = (decltype(x.comm_count))dlsym(h, "ncclCommCount");
        x.comm_rank = (decltype(x.comm_rank))dlsym(h, "ncclCommUserRank");
        x.error_string = (decltype(x.error_string))dlsym(h, "ncclGetErrorString");
        x.ok = x.all_gather && x.all_reduce && x.comm_count && x.comm_rank;
        return x;
    }();
    return n;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h
int nccl_fail(int rc, const char *what) {
    return fail(QMC_ERR_CUDA, "%s failed: %s", what, nccl().error_string ? nccl().error_string(rc) : "NCCL error");
}
}  // namespace

int qmc_pt_exchange(void *nccl_comm, double *temperature_local, const double *energy_local, int32_t n_local, uint64_t seed, uint64_t round,
                    double *scratch, int8_t *accepted, void *stream) {
    if (!temperature_local || !energy_local || !scratch || n_local < 0) return fail(QMC_ERR_INVALID, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    int world = 1, rank = 0;
    if (nccl_comm) {
        if (!nccl().ok) return fail(QMC_ERR_UNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 not found)");
        int rc;
        if ((rc = nccl().comm_count(nccl_comm, &world))) return nccl_fail(rc, "ncclCommCount");
        if ((rc = nccl().comm_rank(nccl_comm, &rank))) return nccl_fail(rc, "ncclCommUserRank");
    }
    const int n_global = n_local * world;
    double *t_all = scratch, *e_all = scratch + n_global;
    if (world > 1) {  // the only collective of the path: 2 x n_global doubles, latency bound
        int rc;
        if ((rc = nccl().all_gather(temperature_local, t_all, (size_t)n_local, NCCL_FLOAT64, nccl_comm, s))) return nccl_fail(rc, "ncclAllGather");
        if ((rc = nccl().all_gather(energy_local, e_all, (size_t)n_local, NCCL_FLOAT64, nccl_comm, s))) return nccl_fail(rc, "ncclAllGather");
    } else {
        CUDA_TRY(cudaMemcpyAsync(t_all, temperature_local, (size_t)n_local * 8, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(e_all, energy_local, (size_t)n_local * 8, cudaMemcpyDeviceToDevice, s));
    }
    int rc = qmc_pt_swap(t_all, e_all, n_global, seed, round, accepted, stream);  // the same decisions on every rank
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(temperature_local, t_all + (size_t)rank * n_local, (size_t)n_local * 8, cudaMemcpyDeviceToDevice, s));
    return QMC_OK;
}

int qmc_reduce_observables(void *nccl_comm, double *values, int32_t n, void *stream) {
    if (!values || n < 0) return fail(QMC_ERR_INVALID, "bad arguments");
    if (!nccl_comm || n == 0) return QMC_OK;
    if (!nccl().ok) return fail(QMC_ERR_UNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 not found)");
    const int rc = nccl().all_reduce(values, values, (size_t)n, NCCL_FLOAT64, NCCL_SUM, nccl_comm, (cudaStream_t)stream);
    return rc ? nccl_fail(rc, "ncclAllReduce") : QMC_OK;
}

int qmc_fp64_peak(double *flops, void *stream) {
    if (!flops) return fail(QMC_ERR_INVALID, "flops is NULL");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = (cudaStream_t)stream;
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 16, blocks = sms * 8, threads = 256;
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, 1 << 10);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, s));
        fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, iters);
        CUDA_TRY(cudaEventRecord(e1, s));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads / (ms * 1e-3);
        best = std::max(best, fl);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    CUDA_TRY(cudaGetLastError());
    *flops = best;
    return QMC_OK;
}

int qmc_debug_math(int32_t which, const double *in, double *out, int32_t n, void *stream) {
    if (!in || !out || n < 0 || which < 0 || which > 5) return fail(QMC_ERR_INVALID, "bad arguments");
    if (n == 0) return QMC_OK;
    debug_math_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(which, in, out, n);
    CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

// ---- library-owned device state -------------------------------------------------------------------------------------
struct qmc_batch_handle {
    int32_t R, N, n_tables;
    bool momenta;
    std::vector<void *> ptr;       // one per field (QMC_FIELD_*), nullptr if absent
    std::vector<size_t> bytes;
    void *rows = nullptr, *rows_ref = nullptr, *rows_state = nullptr;  // neighbour rows (qmc_batch_enable_rows)
    int32_t row_words = 0;
};

int qmc_batch_create(int32_t n_replicas, int32_t n_max, int32_t with_momenta, int32_t n_label_tables, qmc_batch_handle_t **out) {
    if (!out) return fail(QMC_ERR_INVALID, "out is NULL");
    if (n_replicas <= 0 || n_max <= 0 || n_label_tables < 0 || n_label_tables > QMC_MAX_MOVES)
        return fail(QMC_ERR_INVALID, "bad batch dimensions");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(QMC_ERR_NO_DEVICE, "no CUDA device: there is no CPU fallback");
    auto *h = new qmc_batch_handle;
    h->R = n_replicas;
    h->N = n_max;
    h->n_tables = n_label_tables;
    h->momenta = with_momenta != 0;
    const size_t R = (size_t)n_replicas, N = (size_t)n_max;
    const size_t sizes[QMC_FIELD_LABELS0] = {R * 3 * N * 8, R * 9 * 8, R * 4, R * 8, R * 8, R * N * 8, R * N * 8,
                                             with_momenta ? R * 3 * N * 8 : 0, with_momenta ? R * 8 : 0, R * 4,
                                             R * QMC_MAX_MOVES * 3 * 8, R * 4, R * 8};
    for (int f = 0; f < QMC_FIELD_LABELS0 + n_label_tables; ++f) {
        const size_t b = f < QMC_FIELD_LABELS0 ? sizes[f] : R * N * 4;
        void *p = nullptr;
        if (b) {
            cudaError_t e = cudaMalloc(&p, b);
            if (e == cudaSuccess) e = cudaMemset(p, f >= QMC_FIELD_LABELS0 ? 0xFF : 0, b);  // label tables start as -1 (not movable)
            if (e != cudaSuccess) {
                for (void *q : h->ptr) cudaFree(q);
                delete h;
                return fail(QMC_ERR_CUDA, "device allocation of %zu bytes failed: %s", b, cudaGetErrorString(e));
            }
        }
        h->ptr.push_back(p);
        h->bytes.push_back(b);
    }
    {  // energies (and kinetic energies) start as NaN = "not primed", like last_potential_energy in the reference
        std::vector<double> nan(R, std::nan(""));
        cudaMemcpy(h->ptr[QMC_FIELD_ENERGY], nan.data(), R * 8, cudaMemcpyHostToDevice);
        if (with_momenta) cudaMemcpy(h->ptr[QMC_FIELD_KINETIC], nan.data(), R * 8, cudaMemcpyHostToDevice);
    }
    *out = h;
    return QMC_OK;
}

int qmc_batch_destroy(qmc_batch_handle_t *h) {
    if (!h) return QMC_OK;
    for (void *p : h->ptr) cudaFree(p);
    cudaFree(h->rows);
    cudaFree(h->rows_ref);
    cudaFree(h->rows_state);
    delete h;
    return QMC_OK;
}

static int batch_field(qmc_batch_handle_t *h, int32_t field, int64_t bytes, void **p) {
    if (!h) return fail(QMC_ERR_INVALID, "handle is NULL");
    if (field < 0 || field >= (int32_t)h->ptr.size() || !h->ptr[field]) return fail(QMC_ERR_INVALID, "field %d is not part of this batch", field);
    if (bytes != (int64_t)h->bytes[field]) return fail(QMC_ERR_INVALID, "field %d holds %zu bytes, got %lld", field, h->bytes[field], (long long)bytes);
    *p = h->ptr[field];
    return QMC_OK;
}

int qmc_batch_upload(qmc_batch_handle_t *h, int32_t field, const void *host, int64_t bytes) {
    void *p;
    int rc;
    if (!host) return fail(QMC_ERR_INVALID, "host pointer is NULL");
    if ((rc = batch_field(h, field, bytes, &p))) return rc;
    CUDA_TRY(cudaMemcpy(p, host, (size_t)bytes, cudaMemcpyHostToDevice));
    if (h->rows_state && (field == QMC_FIELD_POS || field == QMC_FIELD_CELL || field == QMC_FIELD_N_ATOMS))
        CUDA_TRY(cudaMemset(h->rows_state, 0, (size_t)h->R * 4));  // the rows no longer describe these positions
    return QMC_OK;
}

int qmc_batch_download(qmc_batch_handle_t *h, int32_t field, void *host, int64_t bytes) {
    void *p;
    int rc;
    if (!host) return fail(QMC_ERR_INVALID, "host pointer is NULL");
    if ((rc = batch_field(h, field, bytes, &p))) return rc;
    CUDA_TRY(cudaMemcpy(host, p, (size_t)bytes, cudaMemcpyDeviceToHost));
    return QMC_OK;
}

int qmc_batch_view(qmc_batch_handle_t *h, qmc_batch_t *out) {
    if (!h || !out) return fail(QMC_ERR_INVALID, "NULL argument");
    out->n_replicas = h->R;
    out->n_max = h->N;
    out->pos = (double *)h->ptr[QMC_FIELD_POS];
    out->cell = (double *)h->ptr[QMC_FIELD_CELL];
    out->n_atoms = (int32_t *)h->ptr[QMC_FIELD_N_ATOMS];
    out->energy = (double *)h->ptr[QMC_FIELD_ENERGY];
    out->temperature = (double *)h->ptr[QMC_FIELD_TEMPERATURE];
    out->sigma1 = (double *)h->ptr[QMC_FIELD_SIGMA1];
    out->eatom = (double *)h->ptr[QMC_FIELD_EATOM];
    out->momenta = (double *)h->ptr[QMC_FIELD_MOMENTA];
    out->kinetic = (double *)h->ptr[QMC_FIELD_KINETIC];
    out->n_exchange = (int32_t *)h->ptr[QMC_FIELD_N_EXCHANGE];
    out->counters = (int64_t *)h->ptr[QMC_FIELD_COUNTERS];
    out->status = (int32_t *)h->ptr[QMC_FIELD_STATUS];
    out->pair_evals = (int64_t *)h->ptr[QMC_FIELD_PAIR_EVALS];
    out->nbr_rows = (uint32_t *)h->rows;
    out->nbr_ref = (double *)h->rows_ref;
    out->nbr_state = (int32_t *)h->rows_state;
    out->nbr_row_cap = h->row_words;
    return QMC_OK;
}

int qmc_batch_enable_rows(qmc_batch_handle_t *h, const qmc_potential_t *pot) {
    if (!h || !pot) return fail(QMC_ERR_INVALID, "NULL argument");
    if (h->rows) return QMC_OK;
    const int32_t words = qmc_nbr_row_words(pot, h->N);
    if (words < 0) return words;
    const size_t R = (size_t)h->R, N = (size_t)h->N;
    void *rows = nullptr, *ref = nullptr, *state = nullptr;
    cudaError_t e = cudaMalloc(&rows, R * N * words * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ref, R * 8 * 8);
    if (e == cudaSuccess) e = cudaMalloc(&state, R * 4);
    if (e == cudaSuccess) e = cudaMemset(state, 0, R * 4);
    if (e != cudaSuccess) {
        cudaFree(rows);
        cudaFree(ref);
        cudaFree(state);
        return fail(QMC_ERR_CUDA, "device allocation of the neighbour rows failed: %s", cudaGetErrorString(e));
    }
    h->rows = rows;
    h->rows_ref = ref;
    h->rows_state = state;
    h->row_words = words;
    return QMC_OK;
}

int32_t *qmc_batch_labels(qmc_batch_handle_t *h, int32_t table) {
    if (!h || table < 0 || table >= h->n_tables) return nullptr;
    return (int32_t *)h->ptr[QMC_FIELD_LABELS0 + table];
}

int qmc_sweep_host(const qmc_potential_t *pot, const qmc_batch_t *h, const qmc_move_t *moves_host, int32_t n_moves, uint64_t seed,
                   uint64_t attempt_base, int32_t n_attempts, const qmc_trace_t *trace_host) {
    if (!pot || !h || !moves_host) return fail(QMC_ERR_INVALID, "NULL argument");
    if (h->n_replicas <= 0 || h->n_max <= 0) return fail(QMC_ERR_INVALID, "bad batch dimensions");
    if (!h->pos || !h->cell || !h->n_atoms || !h->energy || !h->temperature) return fail(QMC_ERR_INVALID, "host batch has NULL pointers");
    const size_t R = (size_t)h->n_replicas, N = (size_t)h->n_max;
    std::vector<void *> owned;
    auto dev_alloc = [&](size_t bytes, const void *src, void **out) -> int {
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 8));
        owned.push_back(p);
        if (src) CUDA_TRY(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        else CUDA_TRY(cudaMemset(p, 0, bytes ? bytes : 8));
        *out = p;
        return QMC_OK;
    };
    auto cleanup = [&]() { for (void *p : owned) cudaFree(p); };
    qmc_batch_t d = *h;
    d.cells_uniform = 0;
    d.nbr_rows = nullptr;
    d.nbr_ref = nullptr;
    d.nbr_state = nullptr;
    d.nbr_row_cap = 0;
    int rc = QMC_OK;
#define UP(field, bytes, src) if (rc == QMC_OK) rc = dev_alloc(bytes, src, (void **)&d.field)
    UP(pos, R * 3 * N * 8, h->pos);
    UP(cell, R * 9 * 8, h->cell);
    UP(n_atoms, R * 4, h->n_atoms);
    UP(energy, R * 8, h->energy);
    UP(temperature, R * 8, h->temperature);
    UP(sigma1, R * N * 8, nullptr);
    UP(eatom, R * N * 8, nullptr);
    UP(counters, R * QMC_MAX_MOVES * 3 * 8, h->counters);
    UP(status, R * 4, nullptr);
    UP(n_exchange, R * 4, h->n_exchange);  // (zeros if the caller has none: the general kernels read and write it)
    d.momenta = nullptr;
    d.kinetic = nullptr;
    d.pair_evals = nullptr;
    if (h->pair_evals) UP(pair_evals, R * 8, h->pair_evals);
    std::vector<qmc_move_t> mv(moves_host, moves_host + n_moves);
    // one device table per DISTINCT host label array: the sub-moves of a composite must share theirs (qmc_sweep checks the
    // pointers), and moves that share a host array must not overwrite each other on the way back
    std::map<const int32_t *, int32_t *> tables;
    for (int m = 0; m < n_moves && rc == QMC_OK; ++m) {
        if (!moves_host[m].labels) continue;
        auto it = tables.find(moves_host[m].labels);
        if (it == tables.end()) {
            int32_t *p = nullptr;
            rc = dev_alloc(R * N * 4, moves_host[m].labels, (void **)&p);
            it = tables.emplace(moves_host[m].labels, p).first;
        }
        mv[m].labels = it->second;
    }
    qmc_trace_t tr;
    std::memset(&tr, 0, sizeof(tr));
    const size_t T = R * (size_t)n_attempts;
    if (trace_host && rc == QMC_OK) {
        if (trace_host->move) rc = dev_alloc(T, nullptr, (void **)&tr.move);
        if (trace_host->accepted && rc == QMC_OK) rc = dev_alloc(T, nullptr, (void **)&tr.accepted);
        if (trace_host->energy && rc == QMC_OK) rc = dev_alloc(T * 8, nullptr, (void **)&tr.energy);
        if (trace_host->delta && rc == QMC_OK) rc = dev_alloc(T * 8, nullptr, (void **)&tr.delta);
        if (trace_host->moved && rc == QMC_OK) rc = dev_alloc(T * 4, nullptr, (void **)&tr.moved);
    }
#undef UP
    if (rc == QMC_OK) rc = qmc_energy_full(pot, &d, nullptr, nullptr);
    if (rc == QMC_OK) rc = qmc_sweep(pot, &d, mv.data(), n_moves, seed, attempt_base, n_attempts, trace_host ? &tr : nullptr, nullptr);
    auto down = [&](void *dst, const void *src, size_t bytes) {
        if (rc == QMC_OK && dst && src) {
            cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = fail(QMC_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
        }
    };
    down(h->pos, d.pos, R * 3 * N * 8);
    down(h->cell, d.cell, R * 9 * 8);
    down(h->n_atoms, d.n_atoms, R * 4);
    down(h->energy, d.energy, R * 8);
    down(h->counters, d.counters, R * QMC_MAX_MOVES * 3 * 8);
    down(h->status, d.status, R * 4);
    down(h->sigma1, d.sigma1, R * N * 8);
    down(h->eatom, d.eatom, R * N * 8);
    if (h->n_exchange) down(h->n_exchange, d.n_exchange, R * 4);
    if (h->pair_evals) down(h->pair_evals, d.pair_evals, R * 8);
    for (const auto &t : tables) down(const_cast<int32_t *>(t.first), t.second, R * N * 4);
    if (trace_host) {
        down(trace_host->move, tr.move, T);
        down(trace_host->accepted, tr.accepted, T);
        down(trace_host->energy, tr.energy, T * 8);
        down(trace_host->delta, tr.delta, T * 8);
        down(trace_host->moved, tr.moved, T * 4);
    }
    cleanup();
    return rc;
}

}  // extern "C"
