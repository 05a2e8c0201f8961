/* CPU oracle for the batched Monte Carlo hot path -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * Plain-C restatement of what the reference (quansino v0.1.3 over ASE 3.26.0) computes per Monte Carlo
 * attempt: proposal -> FULL-SYSTEM energy recompute -> acceptance -> save / revert.  See oracle/__init__.py
 * for the parity status (quansino semantics pinned by golden traces of the genuine sources; ASE arithmetic
 * "parity unpinned").  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef QMC_ORACLE_H
#define QMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { QO_POT_EMT = 0, QO_POT_LJ = 1 };
enum { QO_MOVE_DISPLACEMENT = 0, QO_MOVE_CELL = 1, QO_MOVE_EXCHANGE = 2, QO_MOVE_HMC = 3 };
enum { QO_OP_BALL = 0, QO_OP_BOX = 1, QO_OP_SPHERE = 2, QO_OP_TRANSLATION = 3, QO_OP_ISOTROPIC = 4, QO_OP_VERLET = 5, QO_OP_ROTATION = 6,
       QO_OP_TRANSLATION_ROTATION = 7,
       QO_OP_ANISOTROPIC = 8, /* AnisotropicDeformation (operations/cell.py:56-96) */
       QO_OP_SHAPE = 9        /* ShapeDeformation (operations/cell.py:99-141) */ };

/* ase/calculators/emt.py (single species) and ase/calculators/lj.py constants, eV / Angstrom */
typedef struct {
    int32_t kind; /* QO_POT_* */
    int32_t _pad;
    /* EMT */
    double E0, s0, V0, eta2, kappa, lambda, beta, rc, rc_list, acut, gamma1, gamma2;
    /* LJ */
    double sigma, epsilon, lj_rc, lj_e0;
} qo_potential_t;

/* One MoveStorage (src/quansino/utils/moves.py:19-60) with its move and operation lowered to POD. */
typedef struct {
    int32_t type;          /* QO_MOVE_* */
    int32_t op;            /* QO_OP_* */
    double step;           /* step_size (Ball/Box/Sphere) or max_value (IsotropicDeformation) */
    double probability;    /* MoveStorage.probability */
    double bias_insert;    /* ExchangeMove.bias_towards_insert */
    double dt;             /* Verlet: dt in ASE time units (dt_fs * units.fs) */
    int32_t n_steps;       /* Verlet.max_steps */
    int32_t default_label; /* DisplacementMove.default_label, INT32_MIN = None (0 is falsy => None too) */
    int32_t repeat;        /* CompositeDisplacementMove: this sub-move appears `repeat` times in a row (>= 1; 0 reads as 1) */
    int32_t chain;         /* bit 0: the next entry of moves[] belongs to the same composite move (moves/displacement.py:363-455);
                              bit 1: plain CompositeMove (moves/composite.py:43-62): labels may repeat, a CellMove may close it */
    int32_t axes;          /* CellMove: bit d set = axis d is deformed (diagonal of IsotropicDeformation.mask); 0 = all three */
    int32_t minimum_count; /* MoveStorage.minimum_count (mc/core.py:331-342) */
    int32_t max_cycles;    /* cycles per step (entry 0), needed if some minimum_count > 0 */
    int32_t n_extra_ops;   /* CompositeOperation: 0..2 further displacement operations, translations added in list order */
    int32_t extra_op[2];
    int32_t _pad;
    double extra_step[2];
    int32_t *labels;       /* [n_max] this move's labels, updated in place on accepted exchanges */
} qo_move_t;

/* The simulation state of ONE replica (Atoms + Context fields). */
typedef struct {
    int32_t n_atoms;  /* in/out */
    int32_t n_max;    /* capacity of pos / momenta / labels */
    double *pos;      /* [n_max][3] in/out, row-major like Atoms.positions */
    double *momenta;  /* [n_max][3] in/out (HMC) or NULL */
    double cell[9];   /* row-major, in/out */
    int32_t pbc[3];
    double mass;      /* amu, single species */
    double temperature;          /* K */
    double pressure;             /* eV / A^3 */
    double chemical_potential;   /* eV */
    double accessible_volume;    /* A^3 (ExchangeContext.accessible_volume) */
    double exchange_mass;        /* amu: exchange_atoms.get_masses().sum() */
    double exchange_pos[3];      /* position of the single exchange atom before Translation */
    int32_t n_exchange;          /* number_of_exchange_particles, in/out */
    int32_t _pad;
    double last_energy;          /* last_potential_energy; NaN => primed by the first call */
    double last_kinetic;         /* last_kinetic_energy */
    int32_t n_exchange_atoms;    /* atoms of the exchange molecule; 0 / 1: the single atom at exchange_pos */
    int32_t _pad2;
    double exchange_mol[24];     /* [n_exchange_atoms][3] exchange_atoms.positions (at most 8 atoms) */
    int32_t isotension;          /* cell moves judged by IsotensionCriteria (mc/criteria.py:175-219) instead of IsobaricCriteria */
    int32_t _pad3;
    double external_stress[9];   /* eV / A^3, row-major (DeformationContext.external_stress) */
} qo_state_t;

/* Per-attempt trace (all optional, may be NULL). */
typedef struct {
    int8_t *move;       /* index into moves[] of the selected move */
    int8_t *accepted;   /* 1 accepted, 0 rejected, -1 move failed (None) */
    double *energy;     /* last_potential_energy after the attempt */
    double *delta;      /* energy difference the criteria saw (NaN if the move failed) */
    int32_t *n_atoms;   /* atom count after the attempt */
    int32_t *moved;     /* atom index displaced / inserted / deleted, -1 for cell / hmc / failed */
} qo_trace_t;

/* counters[0] pair evaluations the full recomputes performed, [1] P = algorithmic local pair evaluations
 * (n_nb(old) + n_nb(new) of the moved atom; full count for cell / hmc), [2] M = embedding evaluations the
 * local algorithm needs, [3] number of full energy / force evaluations. */

void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void qo_uniform_pair(uint64_t seed, uint64_t attempt, uint32_t replica, uint32_t block, double out[2]);
double qo_normal(uint64_t seed, uint64_t attempt, uint32_t replica, uint32_t m);

/* Full-system energy (and forces / per-atom sigma1 if non-NULL).  Returns the number of
 * (atom, neighbour image) pairs inside the cutoff, or < 0 on error. */
int64_t qo_energy(const qo_potential_t *pot, int32_t n, const double *pos, const double *cell,
                  const int32_t *pbc, double *energy, double *forces, double *sigma1);

/* n_attempts cycles of MonteCarlo.step (mc/core.py:353-384) with every move in moves[] available.
 * Returns 0, or a negative error code (-2 unsupported label groups, -3 capacity, -4 exp overflow). */
/* One ForceBias step (mc/fbmc.py:202-241) of one replica: forces -> gamma -> zeta by rejection -> positions += zeta delta.
 * Draws: element p of call c of step `step` = Philox(seed; step, replica, ((c + 1) << 16) | (p >> 1)), half p & 1.
 * pos [n][3] in/out; returns the energy after the move in *energy and max |gamma| in *gamma_max; iterations in *n_iter. */
int qo_fbmc_step(const qo_potential_t *pot, int32_t n, double *pos, const double *cell, const int32_t *pbc, double mass,
                 double temperature, double delta, uint64_t seed, uint32_t replica, uint64_t step, double *energy,
                 double *gamma_max, int32_t *n_iter);

int qo_mc_run(const qo_potential_t *pot, qo_state_t *st, qo_move_t *moves, int32_t n_moves, uint64_t seed,
              uint32_t replica, uint64_t attempt_base, int32_t n_attempts, qo_trace_t *trace,
              int64_t counters[4]);

/* GrandCanonicalCriteria prefactor pieces (mc/criteria.py:243-271), exposed for the known-answer tests. */
double qo_debroglie_wavelength(double mass_amu, double temperature);
/* exp of a symmetric 3 x 3 matrix (row-major), as used by the Anisotropic / Shape deformations */
void qo_expm_sym3(const double A[9], double E[9]);

double qo_gc_acceptance(double delta_e, int32_t particle_delta, int32_t n_exchange, double accessible_volume,
                        double mass_amu, double temperature, double mu);

/* Parallel-tempering swap decision for the pair (k, k+1) -- no counterpart in the reference. */
int qo_pt_swap_accept(double e_k, double e_k1, double t_k, double t_k1, double u);

#ifdef __cplusplus
}
#endif
#endif
