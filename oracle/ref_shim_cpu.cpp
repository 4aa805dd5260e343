// Shim that exposes the reference's own CPU rasterizer (compiled from the sources where they
// lie under /root/reference, never copied) as a Python module.  TEST INFRASTRUCTURE ONLY.
//
// The stock DSS/csrc/ext.cpp + rasterize_points.h cannot be built CPU-only (they reference the
// CUDA symbols unguarded: rasterize_points.h:138-145,182,276), so we forward-declare the five
// *Cpu entry points with the signatures of DSS/csrc/rasterize_points.h:19-28,130-137,209-218,
// 306-315,317 and bind only those.
#include <torch/extension.h>
#include <tuple>

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizePointsNaiveCpu(
    const torch::Tensor &points, const torch::Tensor &ellipse_params, const torch::Tensor &cutoff_thres,
    const torch::Tensor &radii, const torch::Tensor &cloud_to_packed_first_idx,
    const torch::Tensor &num_points_per_cloud, const float depth_merging_thres, const int image_size,
    const int points_per_pixel);

torch::Tensor RasterizePointsCoarseCpu(
    const torch::Tensor &points, const torch::Tensor &radii, const torch::Tensor &cloud_to_packed_first_idx,
    const torch::Tensor &num_points_per_cloud, const int image_size, const int bin_size,
    const int max_points_per_bin);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizePointsFineCpu(
    const torch::Tensor &points, const torch::Tensor &ellipse_params, const torch::Tensor &cutoff_thres,
    const torch::Tensor &radii, const torch::Tensor &bin_points, const float depth_merging_thres,
    const int image_size, const int bin_size, const int points_per_pixel);

torch::Tensor RasterizePointsOccBackwardCpu(
    const torch::Tensor &points, const torch::Tensor &radii, const torch::Tensor &grad_occ,
    const torch::Tensor &cloud_to_packed_first_idx, const torch::Tensor &num_points_per_cloud,
    const float radii_s, const float depth_merging_thres);

void RasterizeZbufBackwardCpu(const at::Tensor &idx, const at::Tensor &zbuf_grad, at::Tensor &point_z_grad);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "reference DSS/csrc CPU rasterizer (witness for the oracle)";
  m.def("splat_points_naive_cpu", &RasterizePointsNaiveCpu);
  m.def("rasterize_coarse_cpu", &RasterizePointsCoarseCpu);
  m.def("rasterize_fine_cpu", &RasterizePointsFineCpu);
  m.def("splat_points_occ_backward_cpu", &RasterizePointsOccBackwardCpu);
  m.def("backward_zbuf_cpu", &RasterizeZbufBackwardCpu);
}
