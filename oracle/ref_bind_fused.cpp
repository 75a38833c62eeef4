// TEST INFRASTRUCTURE ONLY (oracle/). Python binding for the *unmodified* reference fused-quant
// CUDA translation unit, compiled where it lies under /root/reference (see build_ref.py).
// Exposes the reference host entry points declared in /root/reference/csrc/fused/fused.h:19-76
// so GPU parity tests can run the real reference kernels next to ours.  No reference source is
// copied into this repository: only the header is #included by absolute path at build time.
#include <torch/extension.h>
#include "/root/reference/csrc/fused/fused.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("quant_per_block_int8_scale_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, float, int, int>(&quant_per_block_int8_cuda));
  m.def("quant_per_block_int8_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, int, int>(&quant_per_block_int8_cuda));
  m.def("quant_per_block_int8_fuse_sub_mean_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int, int>(&quant_per_block_int8_fuse_sub_mean_cuda));
  m.def("quant_per_warp_int8_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, int, int, int>(&quant_per_warp_int8_cuda));
  m.def("transpose_pad_permute_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, int>(&transpose_pad_permute_cuda));
  m.def("scale_fuse_quant_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, int, float, int>(&scale_fuse_quant_cuda));
  m.def("mean_scale_fuse_quant_cuda",
        py::overload_cast<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int, float, int>(&mean_scale_fuse_quant_cuda));
}
