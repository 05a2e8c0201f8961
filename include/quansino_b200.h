/* quansino_b200 -- C ABI of the B200-native batched Monte Carlo hot path.
 *
 * This header is the drop-in boundary for ONE path of quansino (v0.1.3): the inner body of
 * `MonteCarlo.step` (src/quansino/mc/core.py:353-384) -- select move -> propose -> energy -> criteria ->
 * save / revert -- executed for MANY independent replicas and a whole run of attempts per call, on one
 * B200 (sm_100a).  quansino is pure Python and has no FFI of its own; the entry points below are what a
 * binding for this path would call (see INTEGRATION.md for the ctypes stub), and the Python package
 * `quansino_b200` mirrors the reference's Move / Operation / Criteria / Context / driver classes on top of
 * exactly these calls.
 *
 * Conventions
 *  - plain C, no exceptions across the boundary: every function returns 0 on success or a negative
 *    QMC_ERR_* code; `qmc_last_error()` returns a thread-local message.
 *  - every pointer inside `qmc_batch_t`, `qmc_move_t.labels` and `qmc_trace_t` is a DEVICE pointer owned
 *    by the caller (the Python side allocates them as torch CUDA tensors); `stream` is a `cudaStream_t`
 *    passed as `void *` (NULL = legacy default stream).  The `*_host` entry points take HOST pointers
 *    and do their own transfers.
 *  - units are ASE's: eV, Angstrom, amu, Kelvin; time in ASE units (fs * 0.0982269...).
 *  - replicas are independent; replica r of this batch is GLOBAL replica `replica_offset + r` for the
 *    random stream, so results do not depend on how replicas are sharded over GPUs.
 *  - cells are orthorhombic (diagonal), each periodic edge >= the neighbour cutoff; general (triclinic) cells run through the
 *    general-cell kernels (qmc_batch_t.triclinic) with displacement, plain exchange and cell moves only.
 *
 * Random stream (shared with the oracle, DESIGN.md "Random stream"): Philox4x32-10,
 * key = seed, counter = (attempt_lo, attempt_hi, global_replica, block); block 0 = (u_select, u_label),
 * 1 = (op0, op1), 2 = (op2, u_accept), 3 = (u_coin, u_swap), 16 + p = Box-Muller pair p of the momenta.
 */
#ifndef QUANSINO_B200_H
#define QUANSINO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMC_ABI_VERSION 9
#define QMC_MAX_FORCED 16 /* sum of minimum_count over the move table */
#define QMC_MAX_MOVES 4
#define QMC_NBR_SKIN_DEFAULT 0.7 /* Angstrom: Cu EMT rows stay within three 32-entry passes (profiles/r2_a_rows.md) */
#define QMC_MAX_SUBMOVES 16 /* sub-moves of one composite EXCHANGE move */
#define QMC_MAX_COMPOSITE_DISPLACEMENTS 1024 /* displacements one composite move may perform per attempt (sub-move c >= 1 draws from
                                                Philox blocks 32 + 2 (c - 1), 33 + 2 (c - 1) for c <= 16 and 0x8000 + 2 (c - 17), + 1 beyond) */

enum qmc_error {
    QMC_OK = 0,
    QMC_ERR_INVALID = -1,     /* bad argument (ValueError on the Python side) */
    QMC_ERR_UNSUPPORTED = -2, /* feature the GPU path refuses (NotImplementedError), never a silent fallback */
    QMC_ERR_CUDA = -3,        /* CUDA runtime error; message in qmc_last_error() */
    QMC_ERR_NO_DEVICE = -4    /* no CUDA device: there is NO CPU fallback */
};

/* per-replica status bits written by the kernels into qmc_batch_t.status */
enum qmc_status {
    QMC_STATUS_EXP_OVERFLOW = 1, /* criteria exponent > 709.78: the reference raises OverflowError (mc/criteria.py:108) */
    QMC_STATUS_CAPACITY = 2,     /* insertion beyond n_max */
    QMC_STATUS_CELL_TOO_SMALL = 4, /* a periodic edge shrank below the neighbour cutoff */
    QMC_STATUS_LIST_OVERFLOW = 8,
    QMC_STATUS_ROW_SHIFT = 16    /* neighbour rows: interacting atoms whose unwrapped coordinates are more than 63 cells apart */
};

/* replaces ase.calculators.emt.EMT / ase.calculators.lj.LennardJones behind
 * `atoms.get_potential_energy()` (call sites: src/quansino/mc/criteria.py:104-106,131-135,161-163,240-242)
 * and `atoms.get_forces()` (src/quansino/integrators/displacement.py:63,77) */
enum qmc_potential_kind { QMC_POT_EMT = 0, QMC_POT_LJ = 1 };
typedef struct qmc_potential {
    int32_t kind;
    int32_t _pad;
    /* EMT, single species, as EMT.initialize derives them (eV, Angstrom) */
    double E0, s0, V0, eta2, kappa, lambda, beta, rc, rc_list, acut, gamma1, gamma2;
    /* Lennard-Jones: sigma, epsilon, cutoff rc, shift e0 = 4 eps ((sigma/rc)^12 - (sigma/rc)^6) */
    double sigma, epsilon, lj_rc, lj_e0;
} qmc_potential_t;

/* replaces one MoveStorage (src/quansino/utils/moves.py:19-60) with its Move and Operation:
 * DisplacementMove + Ball/Box/Sphere (moves/displacement.py:143-171, operations/displacement.py:67-153),
 * CellMove + IsotropicDeformation (moves/cell.py:57-94, operations/cell.py:153-169),
 * ExchangeMove + Translation (moves/exchange.py:134-176, operations/displacement.py:170-175),
 * HamiltonianDisplacementMove + Verlet (moves/displacement.py:307-345, integrators/displacement.py:52-81).
 *
 * Composite moves (`move * k`, `move_a + move_b`, moves/core.py:185-243, moves/composite.py:62-151): consecutive table
 * entries with `chain` bit 0 set on all but the last form ONE move -- one selection probability (the first entry's; the
 * others are ignored), every entry performed `repeat` times, ONE criteria evaluation on the whole change, everything
 * reverted together.  `chain` bit 1 clear: CompositeDisplacementMove (moves/displacement.py:363-455), no label is
 * displaced twice.  Bit 1 set (on every entry of the chain): plain CompositeMove (moves/composite.py:43-62), the
 * sub-moves choose independently, and a CellMove may be the last entry (the docs' hybrid NPT move
 * `DisplacementMove * N + CellMove`, judged by IsobaricCriteria).  Counters and the trace refer to the first entry.
 *
 * Forced moves (`minimum_count`, mc/core.py:331-342): at the start of every step (every `max_cycles` attempts, counted from
 * `attempt_base`, which must be a multiple of `max_cycles`) K = sum of minimum_count distinct cycle indices are drawn --
 * a partial Fisher-Yates shuffle of arange(max_cycles) with the doubles of Philox blocks 4 + j / 2 of the step's first
 * attempt: t = j + floor(u_j (max_cycles - j)), swap(perm[j], perm[t]), index_j = perm[j] -- and forced move j (the
 * table's heads repeated minimum_count times, in table order) is performed at cycle index_j instead of a selected one.
 *
 * Label groups (moves/displacement.py:164-167: `where(labels == label)` may select several atoms, which then move by ONE
 * translation): `chain` bit 2 on a single displacement entry says that `labels` holds dense group ranks 0..G-1 in the
 * order of the reference's sorted unique labels (-1 = not movable); every atom of the selected rank is displaced.
 *
 * Run-coded groups (`chain` bit 3; tables with a molecular ExchangeMove): an inserted molecule receives ONE new label
 * max + 1 in every registered move (moves/displacement.py:244-250) and deletions preserve the order, so the non-negative labels
 * of a table stay non-decreasing with the atom index and a group is a contiguous RUN of equal labels: the k-th unique label is
 * the k-th run.  `labels` then holds the reference's own label values (the caller checks that they start non-decreasing);
 * a displacement entry moves the whole run rigidly, an exchange entry deletes the whole run or appends n_exchange_atoms
 * atoms with one new label (placed by Translation or TranslationRotation), ONE criteria evaluation either way. */
enum qmc_move_type { QMC_MOVE_DISPLACEMENT = 0, QMC_MOVE_CELL = 1, QMC_MOVE_EXCHANGE = 2, QMC_MOVE_HMC = 3 };
enum qmc_op { QMC_OP_BALL = 0, QMC_OP_BOX = 1, QMC_OP_SPHERE = 2, QMC_OP_TRANSLATION = 3, QMC_OP_ISOTROPIC = 4, QMC_OP_VERLET = 5,
              QMC_OP_ROTATION = 6 /* Rotation (operations/displacement.py:178-200): label groups only, rigid z-x-z Euler rotation about the centre */,
              QMC_OP_TRANSLATION_ROTATION = 7 /* TranslationRotation (:203-227): label groups only; QMC_OP_TRANSLATION (:156-175) also
                                                 serves displacement moves of label groups: centroid to a uniform point of the cell */,
              QMC_OP_ANISOTROPIC = 8 /* AnisotropicDeformation (operations/cell.py:56-96): expm of a symmetric matrix of six U(-max, max)
                                        components (draw order: xx, yy, zz = operation draws 0..2; xy, xz, yz = extra draws 0..2) */,
              QMC_OP_SHAPE = 9 /* ShapeDeformation (:99-141): the same with the mean of the diagonal removed (volume preserving) */ };
#define QMC_LABEL_NONE INT32_MIN
typedef struct qmc_move {
    int32_t type;          /* qmc_move_type */
    int32_t op;            /* qmc_op */
    double step;           /* Ball/Box/Sphere step_size, IsotropicDeformation max_value */
    double probability;    /* MoveStorage.probability (normalised by the library like mc/core.py:344-347) */
    double bias_insert;    /* ExchangeMove.bias_towards_insert */
    double dt;             /* Verlet dt in ASE time units */
    int32_t n_steps;       /* Verlet max_steps */
    int32_t default_label; /* DisplacementMove.default_label; QMC_LABEL_NONE = None */
    int32_t repeat;        /* displacement moves: how many times this entry is performed per attempt (>= 1; 0 reads as 1).
                              exchange moves: repeat = k > 1 is a CompositeExchangeMove of k single-atom exchange moves holding equal
                              labels (moves/exchange.py:237-316): one coin (`bias_insert` = the composite's), k insertions with their own
                              Translation draws (sub-move c >= 1: extra draws 3 (c - 1) + {0, 1, 2}, blocks 64 + e / 2) or k label deletions
                              without replacement (drawing sub-move c >= 1: block 32 + 2 (c - 1), first double), ONE criteria evaluation
                              with particle_delta = +-(sub-moves that acted); needs chain bit 3 on every move of the table */
    int32_t chain;         /* bit 0: the next table entry belongs to the same composite move; bit 1: plain CompositeMove;
                              bit 2: label groups (labels = dense group ranks); bit 3: label groups coded as RUNS (see below) */
    int32_t axes;          /* cell moves: QMC_OP_ISOTROPIC: bit d set = axis d is deformed (the diagonal of IsotropicDeformation.mask,
                              operations/cell.py:167-169); 0 reads as all three axes.  QMC_OP_ANISOTROPIC / QMC_OP_SHAPE: the 3 x 3 mask,
                              bit 3 i + j = mask[i][j] (F = expm(T) * mask + eye * ~mask); 0 reads as all nine */
    int32_t minimum_count; /* MoveStorage.minimum_count: forced occurrences of this move per step (mc/core.py:331-342) */
    int32_t max_cycles;    /* cycles per step; read from entry 0, needed only if some minimum_count > 0 */
    int32_t n_extra_ops;   /* CompositeOperation (operations/composite.py:22-109): 0..2 further displacement operations whose */
    int32_t extra_op[2];   /* translations are ADDED to the first one's (np.sum in list order), each drawing in turn: their */
    int32_t _pad;          /* draws are the attempt's "extra" draws e = 0, 1, ... = Philox block 64 + e / 2, half e & 1 */
    double extra_step[2];
    int32_t *labels;       /* device [n_replicas][n_max]; NULL = arange(n_atoms) and no exchange moves */
} qmc_move_t;

/* replaces Atoms + Context for R replicas (src/quansino/mc/contexts.py:20-385): the device-resident state */
typedef struct qmc_batch {
    int32_t n_replicas;     /* replicas held by THIS device */
    int32_t n_max;          /* per-replica atom capacity (row stride of the per-atom arrays) */
    int32_t replica_offset; /* global id of local replica 0 */
    int32_t pbc[3];
    int32_t cells_uniform;  /* != 0: every replica has the fixed cell diag(cell_diag) (fast path); cell moves clear it */
    int32_t nbr_row_cap;    /* entries per neighbour row (multiple of 32); 0 = no neighbour rows (brute-force scan kernels) */
    double cell_diag[3];
    double *pos;            /* [R][3][n_max]  x-row, y-row, z-row per replica */
    double *cell;           /* [R][9] row-major; diagonal unless `triclinic` is set */
    int32_t *n_atoms;       /* [R] */
    double *energy;         /* [R] last_potential_energy */
    double *temperature;    /* [R] Kelvin (per replica so that parallel tempering only permutes this array) */
    double *sigma1;         /* [R][n_max] EMT cache: sigma_1 per atom (unnormalised) */
    double *eatom;          /* [R][n_max] EMT cache: embedding energy per atom, E_c + E_AS(sigma) - E0 */
    double *momenta;        /* [R][3][n_max] or NULL */
    double *kinetic;        /* [R] last_kinetic_energy or NULL */
    int32_t *n_exchange;    /* [R] number_of_exchange_particles or NULL */
    int64_t *counters;      /* [R][QMC_MAX_MOVES][3] attempted, accepted, failed -- accumulated */
    int32_t *status;        /* [R] qmc_status bits, OR-ed */
    int64_t *pair_evals;    /* [R] pair evaluations performed by the sweeps, accumulated, or NULL */
    /* Neighbour rows of the row-based sweep kernel (csrc/sweep_rows.cuh; the ASE side is NeighborList behind
     * mc/criteria.py:104-106): one row per atom listing every periodic image of another atom within cutoff + skin as
     * (atom index, integer image shift per axis), padded with 0xFFFFFFFF.  The kernel (re)builds the rows of replica r when
     * nbr_state[r] == 0 and whenever a committed atom has moved further than the skin allows; whoever moves atoms by other
     * means (uploads, qmc_hmc, qmc_fbmc_step, the scan-based sweep kernels) must zero nbr_state.  All NULL / 0: not used. */
    uint32_t *nbr_rows;     /* [R][n_max][nbr_row_cap] */
    double *nbr_ref;        /* [R][3][n_max] positions at the time the rows were built (in units of the cell at that time) */
    int32_t *nbr_state;     /* [R] 0 = rows invalid; otherwise number of builds so far */
    double nbr_skin;        /* Angstrom; 0 = library default */
    double mass;            /* amu, single species */
    double pressure;        /* eV / A^3 */
    double chemical_potential;
    double accessible_volume;
    double exchange_mass;   /* amu */
    double exchange_pos[3];
    /* Molecular exchange (moves/exchange.py:88-132 with an `exchange_atoms` of several atoms; the docstring at :194-203 places
     * it with TranslationRotation): atoms per inserted / deleted molecule (0 or 1 = the single atom at exchange_pos) and the
     * molecule's positions [n_exchange_atoms][3]; exchange_mass is the mass of the WHOLE molecule (mc/criteria.py:262). */
    int32_t n_exchange_atoms;
    int32_t _pad2;
    double exchange_mol[24];
    /* General (triclinic) cells.  triclinic != 0: `cell` holds full 3 x 3 matrices (rows = cell vectors) and qmc_sweep /
     * qmc_energy_full / qmc_atom_energies run the general-cell kernels (csrc/sweep_tri.cuh): single-atom displacement moves,
     * single-atom exchange moves placed by Translation, and cell moves (all three deformation operations); every periodic HEIGHT
     * of the cell must be >= the neighbour cutoff.  qmc_forces_full / qmc_hmc / qmc_fbmc_step take general cells, too (the
     * Verlet-list shifts then count cell vectors).
     * isotension != 0: cell moves are judged by IsotensionCriteria (mc/criteria.py:175-219) with `external_stress` (eV / A^3,
     * row-major 3 x 3) and `pressure`; otherwise by IsobaricCriteria.  Batches with QMC_OP_ANISOTROPIC / QMC_OP_SHAPE moves must
     * set triclinic (an accepted move leaves a general cell behind). */
    int32_t triclinic;
    int32_t isotension;
    double external_stress[9];
} qmc_batch_t;

/* optional per-attempt trace, device [R][n_attempts] each, any pointer may be NULL */
typedef struct qmc_trace {
    int8_t *move;      /* selected move index */
    int8_t *accepted;  /* 1 / 0 / -1 (move failed, the reference's None) */
    double *energy;    /* last_potential_energy after the attempt */
    double *delta;     /* energy difference seen by the criteria */
    int32_t *moved;    /* atom index moved / inserted / deleted, -1 otherwise */
} qmc_trace_t;

int qmc_abi_version(void);
const char *qmc_last_error(void);
/* number of CUDA devices, or QMC_ERR_NO_DEVICE */
int qmc_device_count(void);

/* bytes of dynamic shared memory one replica (one warp) of the sweep kernel needs for `n_max` atoms */
int64_t qmc_sweep_smem_per_replica(const qmc_potential_t *pot, int32_t n_max);

/* ---- device-resident state owned by the library (for hosts without a device allocator of their own) -----------------
 * What `Context` owns per simulation in the reference (src/quansino/mc/contexts.py:20-96: atoms, last_* snapshots,
 * calculator caches) for ALL replicas of one GPU.  The handle owns the device arrays of `qmc_batch_t`; `qmc_batch_view`
 * fills a `qmc_batch_t` with those pointers (the scalar members -- pbc, replica_offset, mass, pressure, ... -- are the
 * caller's to set) for qmc_energy_full / qmc_sweep / qmc_hmc / ...; hosts that already hold device memory (the Python
 * package uses torch tensors) skip the handle and fill `qmc_batch_t` themselves. */
typedef struct qmc_batch_handle qmc_batch_handle_t;
enum qmc_field {
    QMC_FIELD_POS = 0,      /* f64 [R][3][n_max] */
    QMC_FIELD_CELL,         /* f64 [R][9] */
    QMC_FIELD_N_ATOMS,      /* i32 [R] */
    QMC_FIELD_ENERGY,       /* f64 [R] */
    QMC_FIELD_TEMPERATURE,  /* f64 [R] */
    QMC_FIELD_SIGMA1,       /* f64 [R][n_max] */
    QMC_FIELD_EATOM,        /* f64 [R][n_max] */
    QMC_FIELD_MOMENTA,      /* f64 [R][3][n_max] (only with_momenta) */
    QMC_FIELD_KINETIC,      /* f64 [R]           (only with_momenta) */
    QMC_FIELD_N_EXCHANGE,   /* i32 [R] */
    QMC_FIELD_COUNTERS,     /* i64 [R][QMC_MAX_MOVES][3] */
    QMC_FIELD_STATUS,       /* i32 [R] */
    QMC_FIELD_PAIR_EVALS,   /* i64 [R] */
    QMC_FIELD_LABELS0       /* i32 [R][n_max]: label table t is field QMC_FIELD_LABELS0 + t, t < n_label_tables */
};
/* all arrays zero-initialised (energy NaN: "not primed") */
int qmc_batch_create(int32_t n_replicas, int32_t n_max, int32_t with_momenta, int32_t n_label_tables, qmc_batch_handle_t **out);
int qmc_batch_destroy(qmc_batch_handle_t *handle);
/* whole-array host <-> device copies; `bytes` must equal the size of the field */
int qmc_batch_upload(qmc_batch_handle_t *handle, int32_t field, const void *host, int64_t bytes);
int qmc_batch_download(qmc_batch_handle_t *handle, int32_t field, void *host, int64_t bytes);
/* n_replicas, n_max, nbr_row_cap and every device pointer of *out; other members are left untouched */
int qmc_batch_view(qmc_batch_handle_t *handle, qmc_batch_t *out);
/* allocates the neighbour rows of the row-based sweep kernel (nbr_rows / nbr_ref / nbr_state, zeroed state = "build at the next
 * launch"); qmc_batch_view then fills the nbr_* members and uploads of positions / cells / atom counts invalidate the rows */
int qmc_batch_enable_rows(qmc_batch_handle_t *handle, const qmc_potential_t *pot);
/* device pointer of label table `table` (for qmc_move_t.labels), NULL if out of range */
int32_t *qmc_batch_labels(qmc_batch_handle_t *handle, int32_t table);

/* words (u32) per neighbour row the row-based sweep kernel expects for this potential and capacity (qmc_batch_t.nbr_row_cap;
 * nbr_rows holds n_replicas * n_max rows), or a negative QMC_ERR_* code */
int32_t qmc_nbr_row_words(const qmc_potential_t *pot, int32_t n_max);

/* validate_simulation (src/quansino/mc/canonical.py:114-122, mc/core.py:221-229): full energy of every
 * replica; fills energy[], sigma1[], eatom[].  `pair_count` (device [R] int64 or NULL) receives the number
 * of (atom, neighbour image) pairs evaluated. */
int qmc_energy_full(const qmc_potential_t *pot, const qmc_batch_t *batch, int64_t *pair_count, void *stream);

/* the same evaluation, also storing the per-atom energies -- `calc.results["energies"]` of ASE's EMT / LennardJones, which
 * AdaptiveForceBias reads under its committee keyword (src/quansino/mc/fbmc.py:329, :431-433) -- into `atom_energies`
 * (device [R][n_max]; NULL = qmc_energy_full).  Replicas of at most 1024 atoms. */
int qmc_atom_energies(const qmc_potential_t *pot, const qmc_batch_t *batch, double *atom_energies, int64_t *pair_count, void *stream);

/* full energy AND forces (device [R][3][n_max]) -- `atoms.get_forces()`; workspace as for qmc_hmc */
int qmc_forces_full(const qmc_potential_t *pot, const qmc_batch_t *batch, double *forces, void *workspace,
                    int64_t workspace_bytes, void *stream);

/* n_attempts cycles of MonteCarlo.step (src/quansino/mc/core.py:353-384) for every replica: move types
 * displacement / cell / exchange; state committed in place. */
int qmc_sweep(const qmc_potential_t *pot, const qmc_batch_t *batch, const qmc_move_t *moves, int32_t n_moves,
              uint64_t seed, uint64_t attempt_base, int32_t n_attempts, const qmc_trace_t *trace, void *stream);

/* n_attempts Hamiltonian attempts (src/quansino/moves/displacement.py:307-345 + mc/criteria.py:117-140) */
int qmc_hmc(const qmc_potential_t *pot, const qmc_batch_t *batch, const qmc_move_t *move, uint64_t seed,
            uint64_t attempt_base, int32_t n_attempts, const qmc_trace_t *trace, void *workspace,
            int64_t workspace_bytes, void *stream);
/* device scratch the Hamiltonian kernels need (double-buffered positions, momenta, neighbour list, partial sums) */
int64_t qmc_hmc_workspace_bytes(const qmc_potential_t *pot, const qmc_batch_t *batch);

/* replaces ForceBias.step (src/quansino/mc/fbmc.py:202-251; force-bias Monte Carlo, doi 10.1063/1.4902136): forces at the
 * current positions -> gamma = clip(F delta / (2 kB T), +-709.782712) -> every coordinate draws zeta in (-1, 1) by rejection
 * against the trial probability (fbmc.py:254-273), redrawing call after call exactly the coordinates still refused ->
 * positions += zeta * delta (momenta round trip m d / m as in the reference) -> energy (and forces) at the new positions.
 * `forces` is a device buffer [R][3][n_max], in / out: if `forces_valid` it holds the forces at the current positions (the
 * previous call left them there), otherwise they are evaluated first; on return it holds the forces at the NEW positions
 * and batch.energy the new energies.  Draws: element p of call c (zeta calls even, uniform calls odd) of step `step` =
 * Philox block ((c + 1) << 16) | (p >> 1) of attempt `step`, half p & 1.  gamma_max, n_calls: device [R], may be NULL.
 * Single species: the mass scaling (min(m) / m)^0.25 is 1. */
int qmc_fbmc_step(const qmc_potential_t *pot, const qmc_batch_t *batch, double delta, uint64_t seed, uint64_t step, double *forces,
                  int32_t forces_valid, double *gamma_max, int32_t *n_calls, void *workspace, int64_t workspace_bytes, void *stream);

/* the same step with one delta per replica (device [R], overrides `delta` when not NULL): AdaptiveForceBias
 * (src/quansino/mc/fbmc.py:285-488, update_delta :382-391) whose "energy" scheme gives every system its own delta. */
int qmc_fbmc_step_v(const qmc_potential_t *pot, const qmc_batch_t *batch, double delta, const double *delta_per_replica, uint64_t seed,
                    uint64_t step, double *forces, int32_t forces_valid, double *gamma_max, int32_t *n_calls, void *workspace,
                    int64_t workspace_bytes, void *stream);

/* parallel tempering between temperature-adjacent replicas (no counterpart in the reference):
 * `energies_all` / `temperatures_all` are device [R_global] arrays every rank holds after the all-gather;
 * the kernel evaluates the swap decisions of round `round` (even / odd pairs) from the shared stream and
 * permutes the TEMPERATURES; `accepted` (device [R_global] int8 or NULL) marks swapped lower partners. */
int qmc_pt_swap(double *temperatures_all, const double *energies_all, int32_t n_global, uint64_t seed,
                uint64_t round, int8_t *accepted, void *stream);

/* The multi-GPU part of the path for hosts without torch.distributed (the reference has none: `MultiDriver` is an
 * empty stub, src/quansino/mc/driver.py:349-371).  `nccl_comm` is an `ncclComm_t` of the caller (one rank per GPU, replicas sharded
 * in equal contiguous blocks, rank-major), passed as `void *`; NULL = a single rank.  NCCL is resolved at run time from the
 * host process (or libnccl.so.2), the library does not link against it.
 *
 * qmc_pt_exchange: one parallel-tempering round -- all-gather of the n_local temperatures and energies of every rank into
 * `scratch` (device, 2 * n_local * ranks doubles), `qmc_pt_swap` on every rank (identical decisions from the shared stream), the
 * rank's slice of the permuted temperatures written back to `temperature_local`.  Configurations never move.
 * qmc_reduce_observables: in-place sum of `n` doubles (device) over the ranks (acceptance counts, energy sums, ...). */
int qmc_pt_exchange(void *nccl_comm, double *temperature_local, const double *energy_local, int32_t n_local, uint64_t seed,
                    uint64_t round, double *scratch, int8_t *accepted, void *stream);
int qmc_reduce_observables(void *nccl_comm, double *values, int32_t n, void *stream);

/* measured-FP64-peak microbenchmark used by bench.py for the roofline denominator: runs a register-resident
 * DFMA loop and returns achieved FLOP/s in *flops (2 flop per DFMA). */
int qmc_fp64_peak(double *flops, void *stream);

/* numerics check of the kernels' branch-free FP64 functions (csrc/fastmath.cuh): out[i] = f(in[i]) on the device,
 * which = 0 exp, 1 log, 2 sqrt, 3 reciprocal, 4 sin(in) for in in [0, 2 pi], 5 cos(in).  Device pointers. */
int qmc_debug_math(int32_t which, const double *in, double *out, int32_t n, void *stream);

/* host-buffer convenience: upload `batch_host` (same struct, HOST pointers), prime, run, download. */
int qmc_sweep_host(const qmc_potential_t *pot, const qmc_batch_t *batch_host, const qmc_move_t *moves_host,
                   int32_t n_moves, uint64_t seed, uint64_t attempt_base, int32_t n_attempts,
                   const qmc_trace_t *trace_host);

#ifdef __cplusplus
}
#endif
#endif
