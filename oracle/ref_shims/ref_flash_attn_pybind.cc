// TEST INFRASTRUCTURE — not part of the product.
// Trimmed pybind block for the reference flash-attention kernels we compare against: it registers only the
// split-q + fully-shared-QKV kernel and its f32-accumulating twin (the reference's own
// kernels/flash-attn/pybind/flash_attn.cc registers all 25 and would need all 26 sources, ~2 min each).
// Function definitions come from the UNMODIFIED reference sources compiled next to this file by oracle/build_ref.py.
#include <torch/extension.h>
#include <torch/types.h>

void flash_attn_mma_stages_split_q_shared_qkv(torch::Tensor Q, torch::Tensor K, torch::Tensor V, torch::Tensor O,
                                              int stages);
void flash_attn_mma_stages_split_q_shared_qkv_acc_f32(torch::Tensor Q, torch::Tensor K, torch::Tensor V,
                                                      torch::Tensor O, int stages);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("flash_attn_mma_stages_split_q_shared_qkv", &flash_attn_mma_stages_split_q_shared_qkv,
        "flash_attn_mma_stages_split_q_shared_qkv");
  m.def("flash_attn_mma_stages_split_q_shared_qkv_acc_f32", &flash_attn_mma_stages_split_q_shared_qkv_acc_f32,
        "flash_attn_mma_stages_split_q_shared_qkv_acc_f32");
}
