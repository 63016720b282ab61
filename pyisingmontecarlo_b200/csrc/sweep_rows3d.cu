// Row-walk checkerboard sweep, 3D lattices (see sweep_rows.cuh).
#include "sweep_rows_launch.cuh"

namespace ising {

int launch_sweep_rows_3d(const SweepArgs& a, cudaStream_t st) { return launch_sweep_rows_dim<3>(a, st); }

}  // namespace ising
