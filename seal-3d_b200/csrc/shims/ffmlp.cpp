// _ffmlp: ffmlp/src/bindings.cpp:5-10 of the reference.
#include "shim_common.h"
using at::Tensor;

static void chk(const Tensor &t, const char *n) {
    TORCH_CHECK(t.device().is_cuda(), n, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), n, " must be a contiguous tensor");
    TORCH_CHECK(t.scalar_type() == at::ScalarType::Half, n, " must be a half tensor");
}

void ffmlp_forward(const Tensor inputs, const Tensor weights, const uint32_t B, const uint32_t input_dim, const uint32_t output_dim, const uint32_t hidden_dim,
                   const uint32_t num_layers, const uint32_t activation_, const uint32_t output_activation_, Tensor forward_buffer, Tensor outputs) {
    chk(inputs, "inputs"); chk(weights, "weights"); chk(forward_buffer, "forward_buffer"); chk(outputs, "outputs");
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_ffmlp_forward(inputs.data_ptr(), weights.data_ptr(), B, input_dim, output_dim, hidden_dim, num_layers, activation_, output_activation_,
                                forward_buffer.data_ptr(), outputs.data_ptr(), cur_stream(inputs)), "ffmlp_forward");
}
void ffmlp_inference(const Tensor inputs, const Tensor weights, const uint32_t B, const uint32_t input_dim, const uint32_t output_dim, const uint32_t hidden_dim,
                     const uint32_t num_layers, const uint32_t activation_, const uint32_t output_activation_, Tensor inference_buffer, Tensor outputs) {
    chk(inputs, "inputs"); chk(weights, "weights"); chk(outputs, "outputs");
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_ffmlp_inference(inputs.data_ptr(), weights.data_ptr(), B, input_dim, output_dim, hidden_dim, num_layers, activation_, output_activation_,
                                  inference_buffer.data_ptr(), outputs.data_ptr(), cur_stream(inputs)), "ffmlp_inference");
}
void ffmlp_backward(const Tensor grad, const Tensor inputs, const Tensor weights, const Tensor forward_buffer, const uint32_t B, const uint32_t input_dim,
                    const uint32_t output_dim, const uint32_t hidden_dim, const uint32_t num_layers, const uint32_t activation_,
                    const uint32_t output_activation_, const bool calc_grad_inputs, Tensor backward_buffer, Tensor grad_inputs, Tensor grad_weights) {
    chk(grad, "grad"); chk(inputs, "inputs"); chk(weights, "weights"); chk(forward_buffer, "forward_buffer"); chk(backward_buffer, "backward_buffer");
    chk(grad_weights, "grad_weights"); chk(grad_inputs, "grad_inputs");
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_ffmlp_backward(grad.data_ptr(), inputs.data_ptr(), weights.data_ptr(), forward_buffer.data_ptr(), B, input_dim, output_dim, hidden_dim,
                                 num_layers, activation_, output_activation_, calc_grad_inputs ? 1 : 0, backward_buffer.data_ptr(), grad_inputs.data_ptr(),
                                 grad_weights.data_ptr(), cur_stream(inputs)), "ffmlp_backward");
}
void allocate_splitk(size_t size) { s3d_throw(s3d_allocate_splitk(size), "allocate_splitk"); }
void free_splitk() { s3d_throw(s3d_free_splitk(), "free_splitk"); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("ffmlp_forward", &ffmlp_forward, "ffmlp_forward (CUDA)");
    m.def("ffmlp_inference", &ffmlp_inference, "ffmlp_inference (CUDA)");
    m.def("ffmlp_backward", &ffmlp_backward, "ffmlp_backward (CUDA)");
    m.def("allocate_splitk", &allocate_splitk, "allocate_splitk (CUDA)");
    m.def("free_splitk", &free_splitk, "free_splitk (CUDA)");
}
