// _freqencoder: freqencoder/src/bindings.cpp of the reference.
#include "shim_common.h"
using at::Tensor;

void freq_encode_forward(Tensor inputs, const uint32_t B, const uint32_t D, const uint32_t deg, const uint32_t C, Tensor outputs) {
    S3D_CHECK_CUDA(inputs); S3D_CHECK_CUDA(outputs); S3D_CHECK_CONTIGUOUS(inputs); S3D_CHECK_CONTIGUOUS(outputs);
    S3D_CHECK_FLOAT(inputs); S3D_CHECK_FLOAT(outputs);
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_freq_encode_forward(inputs.data_ptr<float>(), B, D, deg, C, outputs.data_ptr<float>(), cur_stream(inputs)), "freq_encode_forward");
}
void freq_encode_backward(Tensor grad, Tensor outputs, const uint32_t B, const uint32_t D, const uint32_t deg, const uint32_t C, Tensor grad_inputs) {
    S3D_CHECK_CUDA(grad); S3D_CHECK_CUDA(outputs); S3D_CHECK_CUDA(grad_inputs);
    S3D_CHECK_CONTIGUOUS(grad); S3D_CHECK_CONTIGUOUS(outputs); S3D_CHECK_CONTIGUOUS(grad_inputs);
    S3D_CHECK_FLOAT(grad); S3D_CHECK_FLOAT(outputs); S3D_CHECK_FLOAT(grad_inputs);
    c10::cuda::CUDAGuard g(grad.device());
    s3d_throw(s3d_freq_encode_backward(grad.data_ptr<float>(), outputs.data_ptr<float>(), B, D, deg, C, grad_inputs.data_ptr<float>(), cur_stream(grad)),
              "freq_encode_backward");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("freq_encode_forward", &freq_encode_forward, "freq encode forward (CUDA)");
    m.def("freq_encode_backward", &freq_encode_backward, "freq encode backward (CUDA)");
}
