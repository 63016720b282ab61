// Row-walk checkerboard sweep, 2D lattices (see sweep_rows.cuh).
#include "sweep_rows_launch.cuh"

namespace ising {

int launch_sweep_rows_2d(const SweepArgs& a, cudaStream_t st) { return launch_sweep_rows_dim<2>(a, st); }

}  // namespace ising
