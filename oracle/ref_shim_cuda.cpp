// Shim that exposes the reference's own CUDA kernels (DSS/csrc/rasterize_points.cu,
// rasterize_points_backward.cu, FRNN grid.cu / counting_sort.cu), compiled for sm_100a from the
// sources where they lie under /root/reference, as a Python module: the GPU witness that the
// north_star's "match the reference's own DSS/csrc kernels" is checked against.
// TEST INFRASTRUCTURE ONLY.  Signatures: DSS/csrc/rasterize_points.h:32-41,138-145,219-228,
// 318-338; external/FRNN/frnn/csrc/grid/grid.h:43-50, counting_sort.h:4-11.
#include <torch/extension.h>
#include <tuple>

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizePointsNaiveCuda(
    const torch::Tensor &points, const torch::Tensor &ellipse_params, const torch::Tensor &cutoff_thres,
    const torch::Tensor &radii, const torch::Tensor &cloud_to_packed_first_idx,
    const torch::Tensor &num_points_per_cloud, const float depth_merging_thres, const int image_size,
    const int points_per_pixel);

torch::Tensor RasterizePointsCoarseCuda(
    const torch::Tensor &points, const torch::Tensor &radii, const torch::Tensor &cloud_to_packed_first_idx,
    const torch::Tensor &num_points_per_cloud, const int image_size, const int bin_size,
    const int max_points_per_bin);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizePointsFineCuda(
    const torch::Tensor &points, const torch::Tensor &ellipse_params, const torch::Tensor &cutoff_thres,
    const torch::Tensor &radii, const torch::Tensor &bin_points, const float depth_merging_thres,
    const int image_size, const int bin_size, const int points_per_pixel);

torch::Tensor RasterizePointsOccBackwardCuda(
    const torch::Tensor &points, const torch::Tensor &radii, const torch::Tensor &grad_occ,
    const torch::Tensor &cloud_to_packed_first_idx, const torch::Tensor &num_points_per_cloud,
    const float radii_s, const float depth_merging_thres);

at::Tensor RasterizePointsBackwardCudaFast(
    const at::Tensor &points_sorted, const at::Tensor &radii_sorted, const at::Tensor &rs,
    const at::Tensor &grad_occ, const at::Tensor &num_points_per_cloud,
    const at::Tensor &cloud_to_packed_first_idx, const at::Tensor &points_grid_off,
    const at::Tensor &grid_params);

void RasterizeZbufBackwardCuda(const at::Tensor &idx, const at::Tensor &zbuf_grad, at::Tensor &point_z_grad);

void InsertPointsCUDA(const at::Tensor points, const at::Tensor lengths, const at::Tensor params,
                      at::Tensor grid_cnt, at::Tensor grid_cell, at::Tensor grid_idx, int G);

void CountingSortCUDA(const at::Tensor points, const at::Tensor lengths, const at::Tensor grid_cell,
                      const at::Tensor grid_idx, const at::Tensor grid_off, at::Tensor sorted_points,
                      at::Tensor sorted_points_idxs);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "reference DSS/csrc + FRNN grid CUDA kernels built for sm_100a (witness for parity tests)";
  m.def("splat_points_naive_cuda", &RasterizePointsNaiveCuda);
  m.def("rasterize_coarse_cuda", &RasterizePointsCoarseCuda);
  m.def("rasterize_fine_cuda", &RasterizePointsFineCuda);
  m.def("splat_points_occ_backward_cuda", &RasterizePointsOccBackwardCuda);
  m.def("splat_points_occ_fast_cuda_backward", &RasterizePointsBackwardCudaFast);
  m.def("backward_zbuf_cuda", &RasterizeZbufBackwardCuda);
  m.def("insert_points_cuda", &InsertPointsCUDA);
  m.def("counting_sort_cuda", &CountingSortCUDA);
}
