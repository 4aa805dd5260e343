// Shim that exposes the reference's own brute-force fixed-radius K-NN (external/FRNN/frnn/csrc/bruteforce/
// bruteforce_cpu.cpp -- "used as baseline & ground truth" there), compiled from where it lies under /root/reference.
// TEST INFRASTRUCTURE ONLY: the witness the K-NN golden vectors (tests/golden/knn_*.npz) are generated with.
#include <torch/extension.h>
#include <tuple>

std::tuple<at::Tensor, at::Tensor> FRNNBruteForceCPU(const at::Tensor &p1, const at::Tensor &p2,
                                                     const at::Tensor &lengths1, const at::Tensor &lengths2, int K,
                                                     float r);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "reference FRNN brute-force K-NN on the CPU (witness for the oracle)";
  m.def("frnn_bf_cpu", &FRNNBruteForceCPU);
}
