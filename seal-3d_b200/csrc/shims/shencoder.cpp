// _shencoder: shencoder/src/bindings.cpp of the reference.
#include "shim_common.h"
using at::Tensor;

void sh_encode_forward(Tensor inputs, Tensor outputs, const uint32_t B, const uint32_t D, const uint32_t C, at::optional<Tensor> dy_dx) {
    S3D_CHECK_CUDA(inputs); S3D_CHECK_CUDA(outputs); S3D_CHECK_CONTIGUOUS(inputs); S3D_CHECK_CONTIGUOUS(outputs);
    S3D_CHECK_FLOAT(inputs); S3D_CHECK_FLOAT(outputs);
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_sh_encode_forward(inputs.data_ptr<float>(), outputs.data_ptr<float>(), B, D, C, opt_ptr<float>(dy_dx), cur_stream(inputs)), "sh_encode_forward");
}
void sh_encode_backward(Tensor grad, Tensor inputs, const uint32_t B, const uint32_t D, const uint32_t C, Tensor dy_dx, Tensor grad_inputs) {
    S3D_CHECK_CUDA(grad); S3D_CHECK_CUDA(inputs); S3D_CHECK_CUDA(dy_dx); S3D_CHECK_CUDA(grad_inputs);
    S3D_CHECK_CONTIGUOUS(grad); S3D_CHECK_CONTIGUOUS(inputs); S3D_CHECK_CONTIGUOUS(dy_dx); S3D_CHECK_CONTIGUOUS(grad_inputs);
    S3D_CHECK_FLOAT(grad); S3D_CHECK_FLOAT(dy_dx); S3D_CHECK_FLOAT(grad_inputs);
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_sh_encode_backward(grad.data_ptr<float>(), inputs.data_ptr<float>(), B, D, C, dy_dx.data_ptr<float>(), grad_inputs.data_ptr<float>(),
                                     cur_stream(inputs)), "sh_encode_backward");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("sh_encode_forward", &sh_encode_forward, "SH encode forward (CUDA)");
    m.def("sh_encode_backward", &sh_encode_backward, "SH encode backward (CUDA)");
}
