// Host-side launch entry points of the per-(potential, capacity class) kernel instantiations.
// Each pair lives in its own object file (sweep_inst.cu compiled with -DINST_POT / -DINST_NS) so that the build
// can compile them in parallel.
#pragma once

#include <cuda_runtime.h>

#include "sweep_block.cuh"
#include "sweep_rows.cuh"
#include "sweep_tri.cuh"

namespace qmc {

typedef int (*sweep_launch_fn)(int feat, const SweepArgs &a, cudaStream_t s, char *err, size_t errn);
typedef int (*rows_launch_fn)(int feat, int wide, const SweepArgs &a, const RowArgs &ra, cudaStream_t s, char *err, size_t errn);
typedef int (*energy_launch_fn)(const SweepArgs &a, long long *pair_count, cudaStream_t s, char *err, size_t errn);
typedef int (*tri_launch_fn)(const SweepArgs &a, cudaStream_t s, char *err, size_t errn);

#define QMC_FOR_EACH_NS(X, POT) X(POT, 32) X(POT, 64) X(POT, 108) X(POT, 128) X(POT, 256) X(POT, 512) X(POT, 1024)

#define QMC_DECLARE_INST(POT, NS)                                                                           \
    int sweep_launch_##POT##_##NS(int feat, const SweepArgs &a, cudaStream_t s, char *err, size_t errn);     \
    int energy_launch_##POT##_##NS(const SweepArgs &a, long long *pair_count, cudaStream_t s, char *err, size_t errn); \
    int sweep_block_launch_##POT##_##NS(int feat, const SweepArgs &a, cudaStream_t s, char *err, size_t errn); \
    int sweep_rows_launch_##POT##_##NS(int feat, int wide, const SweepArgs &a, const RowArgs &ra, cudaStream_t s, char *err, size_t errn); \
    int sweep_tri_launch_##POT##_##NS(const SweepArgs &a, cudaStream_t s, char *err, size_t errn);           \
    int energy_tri_launch_##POT##_##NS(const SweepArgs &a, long long *pair_count, cudaStream_t s, char *err, size_t errn);

QMC_FOR_EACH_NS(QMC_DECLARE_INST, 0)
QMC_FOR_EACH_NS(QMC_DECLARE_INST, 1)

}  // namespace qmc
