// _gridencoder: gridencoder/src/bindings.cpp:5-8 of the reference.
#include "shim_common.h"
using at::Tensor;

static int dtype_of(const Tensor &t, const char *name) {
    if (t.scalar_type() == at::ScalarType::Float) return 0;
    if (t.scalar_type() == at::ScalarType::Half) return 1;
    throw std::runtime_error(std::string(name) + " must be a float32 or float16 tensor");
}

void grid_encode_forward(const Tensor inputs, const Tensor embeddings, const Tensor offsets, Tensor outputs, const uint32_t B, const uint32_t D,
                         const uint32_t C, const uint32_t L, const float S, const uint32_t H, at::optional<Tensor> dy_dx, const uint32_t gridtype,
                         const bool align_corners, const uint32_t interp) {
    S3D_CHECK_CUDA(inputs); S3D_CHECK_CUDA(embeddings); S3D_CHECK_CUDA(offsets); S3D_CHECK_CUDA(outputs);
    S3D_CHECK_CONTIGUOUS(inputs); S3D_CHECK_CONTIGUOUS(embeddings); S3D_CHECK_CONTIGUOUS(offsets); S3D_CHECK_CONTIGUOUS(outputs);
    S3D_CHECK_FLOAT(inputs); S3D_CHECK_INT(offsets);
    c10::cuda::CUDAGuard g(inputs.device());
    const int dt = dtype_of(embeddings, "embeddings");
    TORCH_CHECK(outputs.scalar_type() == embeddings.scalar_type(), "outputs must have the dtype of embeddings");
    s3d_throw(s3d_grid_encode_forward(inputs.data_ptr<float>(), embeddings.data_ptr(), offsets.data_ptr<int>(), outputs.data_ptr(), B, D, C, L, S, H,
                                      opt_ptr<void>(dy_dx), gridtype, align_corners ? 1 : 0, interp, dt, cur_stream(inputs)), "grid_encode_forward");
}
void grid_encode_backward(const Tensor grad, const Tensor inputs, const Tensor embeddings, const Tensor offsets, Tensor grad_embeddings, const uint32_t B,
                          const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H, const at::optional<Tensor> dy_dx,
                          at::optional<Tensor> grad_inputs, const uint32_t gridtype, const bool align_corners, const uint32_t interp) {
    S3D_CHECK_CUDA(grad); S3D_CHECK_CUDA(inputs); S3D_CHECK_CUDA(embeddings); S3D_CHECK_CUDA(offsets); S3D_CHECK_CUDA(grad_embeddings);
    S3D_CHECK_CONTIGUOUS(grad); S3D_CHECK_CONTIGUOUS(inputs); S3D_CHECK_CONTIGUOUS(embeddings); S3D_CHECK_CONTIGUOUS(offsets); S3D_CHECK_CONTIGUOUS(grad_embeddings);
    S3D_CHECK_FLOAT(inputs); S3D_CHECK_INT(offsets);
    c10::cuda::CUDAGuard g(inputs.device());
    const int dt = dtype_of(grad, "grad");
    s3d_throw(s3d_grid_encode_backward(grad.data_ptr(), inputs.data_ptr<float>(), embeddings.data_ptr(), offsets.data_ptr<int>(), grad_embeddings.data_ptr(),
                                       B, D, C, L, S, H, opt_ptr<void>(dy_dx), opt_ptr<void>(grad_inputs), gridtype, align_corners ? 1 : 0, interp, dt,
                                       cur_stream(inputs)), "grid_encode_backward");
}
void grad_total_variation(const Tensor inputs, const Tensor embeddings, Tensor grad, const Tensor offsets, const float weight, const uint32_t B,
                          const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H, const uint32_t gridtype,
                          const bool align_corners) {
    c10::cuda::CUDAGuard g(inputs.device());
    s3d_throw(s3d_grad_total_variation(inputs.data_ptr(), embeddings.data_ptr(), grad.data_ptr(), offsets.data_ptr<int>(), weight, B, D, C, L, S, H,
                                       gridtype, align_corners ? 1 : 0, dtype_of(embeddings, "embeddings"), cur_stream(inputs)), "grad_total_variation");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("grid_encode_forward", &grid_encode_forward, "grid_encode_forward (CUDA)");
    m.def("grid_encode_backward", &grid_encode_backward, "grid_encode_backward (CUDA)");
    m.def("grad_total_variation", &grad_total_variation, "grad_total_variation (CUDA)");
}
