// One (potential, capacity class) instantiation of the sweep and energy kernels.
// Compile with -DINST_POT=<0|1> -DINST_NS=<32|64|108|128|256|512|1024>.
#include <cstdio>

#include "launchers.h"

#ifndef INST_POT
#error "compile with -DINST_POT and -DINST_NS"
#endif

#define CAT_(a, b, c, d) a##b##_##d
#define CAT(a, b, c, d) CAT_(a, b, c, d)

namespace qmc {

static inline int report(cudaError_t e, const char *what, char *err, size_t errn) {
    if (e == cudaSuccess) return QMC_OK;
    snprintf(err, errn, "%s failed: %s", what, cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? QMC_ERR_NO_DEVICE : QMC_ERR_CUDA;
}

template <int FEAT>
static inline int launch_feat(const SweepArgs &a, cudaStream_t s, char *err, size_t errn) {
    using LY = Layout<INST_POT, INST_NS>;
    static_assert(LY::FITS, "replica does not fit in shared memory");
    constexpr size_t smem = (size_t)LY::WARPS * LY::BYTES;
    auto kernel = sweep_kernel<INST_POT, INST_NS, FEAT>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    const int grid = (a.b.n_replicas + LY::WARPS - 1) / LY::WARPS;
    kernel<<<grid, LY::WARPS * 32, smem, s>>>(a);
    return report(cudaGetLastError(), "sweep kernel launch", err, errn);
}

int CAT(sweep_launch_, INST_POT, _, INST_NS)(int feat, const SweepArgs &a, cudaStream_t s, char *err, size_t errn) {
    // instantiated feature sets: canonical fast path; per-replica cells; + cell moves; everything; everything + composite moves
    if (feat & F_COMPOSITE) return launch_feat<F_COMPOSITE | F_CELLVAR | F_CELL | F_LABELS | F_EXCHANGE>(a, s, err, errn);
    if (feat == 0) return launch_feat<0>(a, s, err, errn);
    if ((feat & ~F_CELLVAR) == 0) return launch_feat<F_CELLVAR>(a, s, err, errn);
    if ((feat & ~(F_CELLVAR | F_CELL)) == 0) return launch_feat<F_CELLVAR | F_CELL>(a, s, err, errn);
    return launch_feat<F_CELLVAR | F_CELL | F_LABELS | F_EXCHANGE>(a, s, err, errn);
}

template <int FEAT>
static inline int launch_block_feat(const SweepArgs &a, cudaStream_t s, char *err, size_t errn) {
    using LY = LayoutB<INST_POT, INST_NS>;
    static_assert(LY::FITS, "replica does not fit in shared memory");
    constexpr size_t smem = (size_t)LY::REPS * LY::BYTES;
    auto kernel = sweep_block_kernel<INST_POT, INST_NS, FEAT>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    kernel<<<(a.b.n_replicas + LY::REPS - 1) / LY::REPS, LY::REPS * BW * 32, smem, s>>>(a);
    return report(cudaGetLastError(), "sweep block kernel launch", err, errn);
}

// CTA-per-replica variant (csrc/sweep_block.cuh); instantiated feature sets: canonical, + per-replica cells + cell moves, everything
int CAT(sweep_block_launch_, INST_POT, _, INST_NS)(int feat, const SweepArgs &a, cudaStream_t s, char *err, size_t errn) {
    if (feat == 0) return launch_block_feat<0>(a, s, err, errn);
    if ((feat & ~(F_CELLVAR | F_CELL)) == 0) return launch_block_feat<F_CELLVAR | F_CELL>(a, s, err, errn);
    return launch_block_feat<F_CELLVAR | F_CELL | F_LABELS | F_EXCHANGE>(a, s, err, errn);
}

template <int FEAT, int NW>
static inline int launch_rows_feat(const SweepArgs &a, const RowArgs &ra, cudaStream_t s, char *err, size_t errn) {
    using LY = LayoutR<INST_POT, INST_NS, NW>;
    static_assert(LY::FITS, "replica does not fit in shared memory");
    constexpr size_t smem = (size_t)LY::REPS * LY::BYTES;
    auto kernel = sweep_rows_kernel<INST_POT, INST_NS, FEAT, NW>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    kernel<<<(a.b.n_replicas + LY::REPS - 1) / LY::REPS, LY::THREADS, smem, s>>>(a, ra);
    return report(cudaGetLastError(), "row-based sweep kernel launch", err, errn);
}

// row-based variant (csrc/sweep_rows.cuh), `wide` = four warps per replica; instantiated feature sets: canonical, per-replica
// cells, + cell moves, + label tables
int CAT(sweep_rows_launch_, INST_POT, _, INST_NS)(int feat, int wide, const SweepArgs &a, const RowArgs &ra, cudaStream_t s, char *err, size_t errn) {
#if INST_NS >= 256
    if (wide) {
        if ((feat & ~(F_CELLVAR | F_CELL)) == 0) return launch_rows_feat<F_CELLVAR | F_CELL, 4>(a, ra, s, err, errn);
        return launch_rows_feat<F_CELLVAR | F_CELL | F_LABELS, 4>(a, ra, s, err, errn);
    }
#endif
    if (feat == 0) return launch_rows_feat<0, 1>(a, ra, s, err, errn);
    if ((feat & ~F_CELLVAR) == 0) return launch_rows_feat<F_CELLVAR, 1>(a, ra, s, err, errn);
    if ((feat & ~(F_CELLVAR | F_CELL)) == 0) return launch_rows_feat<F_CELLVAR | F_CELL, 1>(a, ra, s, err, errn);
    return launch_rows_feat<F_CELLVAR | F_CELL | F_LABELS, 1>(a, ra, s, err, errn);
}

int CAT(energy_launch_, INST_POT, _, INST_NS)(const SweepArgs &a, long long *pair_count, cudaStream_t s, char *err, size_t errn) {
    using LY = Layout<INST_POT, INST_NS>;
    constexpr size_t smem = (size_t)LY::WARPS * LY::BYTES;
    auto kernel = energy_kernel<INST_POT, INST_NS>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    const int grid = (a.b.n_replicas + LY::WARPS - 1) / LY::WARPS;
    kernel<<<grid, LY::WARPS * 32, smem, s>>>(a, pair_count);
    return report(cudaGetLastError(), "energy kernel launch", err, errn);
}

// general (triclinic) cells (csrc/sweep_tri.cuh): the replica layout of the scan kernel plus the cell constants of every warp
int CAT(sweep_tri_launch_, INST_POT, _, INST_NS)(const SweepArgs &a, cudaStream_t s, char *err, size_t errn) {
    using LY = Layout<INST_POT, INST_NS>;
    constexpr int W = tri_warps<LY>();
    constexpr size_t smem = tri_smem<LY>();
    static_assert(smem <= SMEM_PER_CTA_MAX, "replicas and cell constants do not fit in shared memory");
    auto kernel = sweep_tri_kernel<INST_POT, INST_NS>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    const int grid = (a.b.n_replicas + W - 1) / W;
    kernel<<<grid, W * 32, smem, s>>>(a);
    return report(cudaGetLastError(), "general-cell sweep kernel launch", err, errn);
}

int CAT(energy_tri_launch_, INST_POT, _, INST_NS)(const SweepArgs &a, long long *pair_count, cudaStream_t s, char *err, size_t errn) {
    using LY = Layout<INST_POT, INST_NS>;
    constexpr int W = tri_warps<LY>();
    constexpr size_t smem = tri_smem<LY>();
    auto kernel = energy_tri_kernel<INST_POT, INST_NS>;
    int rc = report(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute", err, errn);
    if (rc) return rc;
    const int grid = (a.b.n_replicas + W - 1) / W;
    kernel<<<grid, W * 32, smem, s>>>(a, pair_count);
    return report(cudaGetLastError(), "general-cell energy kernel launch", err, errn);
}

}  // namespace qmc
