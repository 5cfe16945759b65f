// dt_kernels.h -- GPU builders of the 3-D distance transform (GoICP::BuildDT / DT3D::Build).
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace goicp {

// Frame of the grid: 2x-expanded bounding cube of the model (jly_3ddt.cpp:891-929).
// meta4 = {xMin, yMin, zMin, scale}.  Host arithmetic in double, same operation order.
void dt_frame_host(const float* model_xyz, int nm, int S, double expand, double* meta4);

// Builds the S^3 float distance grid ([z][y][x] order) on the device.
//   mode 0 (GOICP_DT_REFERENCE): the reference's sequential 4-pass-per-slice vector propagation,
//          bit-exact incl. its visiting-order effects (jly_3ddt.cpp:710-742) -- inherently serial
//          along every row, run by one persistent CTA.
//   mode 1 (GOICP_DT_EXACT_EDT): exact Euclidean DT, fully parallel.
cudaError_t dt_build_device(const float* model_xyz_host, int nm, int S, double expand, int mode,
                            float* d_grid_out, double* meta4_out, cudaStream_t stream, std::string& msg);

} // namespace goicp
