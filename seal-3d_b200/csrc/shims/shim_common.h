// Thin pybind11/torch shims: the reference's extension-module surface on top of the C-ABI.
// Built as _raymarching / _gridencoder / _shencoder / _freqencoder / _ffmlp -- the setup.py names the
// reference wrappers import first (`import _gridencoder as _backend`, gridencoder/grid.py:9-12), so the
// unmodified reference Python picks these modules up when their directory is on sys.path.
#pragma once
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>
#include <stdexcept>
#include <string>
#include "seal3d_b200.h"

#define S3D_CHECK_CUDA(x) TORCH_CHECK(x.device().is_cuda(), #x " must be a CUDA tensor")
#define S3D_CHECK_CONTIGUOUS(x) TORCH_CHECK(x.is_contiguous(), #x " must be a contiguous tensor")
#define S3D_CHECK_FLOAT(x) TORCH_CHECK(x.scalar_type() == at::ScalarType::Float, #x " must be a float32 tensor")
#define S3D_CHECK_INT(x) TORCH_CHECK(x.scalar_type() == at::ScalarType::Int, #x " must be an int tensor")
#define S3D_CHECK_HALF(x) TORCH_CHECK(x.scalar_type() == at::ScalarType::Half, #x " must be a half tensor")

static inline void *cur_stream(const at::Tensor &t) {
    return (void *)c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
}
static inline void s3d_throw(int rc, const char *what) {
    if (rc == 0) return;
    if (rc == -95) throw std::runtime_error(std::string(what) + ": configuration not supported by this build");
    if (rc < 0) throw std::runtime_error(std::string(what) + ": invalid argument");
    throw std::runtime_error(std::string(what) + ": CUDA error " + std::to_string(rc));
}
template <typename T> static inline T *opt_ptr(const at::optional<at::Tensor> &t) {
    return t.has_value() ? (T *)t.value().data_ptr() : nullptr;
}
