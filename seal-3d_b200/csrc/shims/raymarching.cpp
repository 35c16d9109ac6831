// _raymarching: raymarching/src/bindings.cpp:7-18 of the reference, same names and argument lists.
#include "shim_common.h"
using at::Tensor;

static const float *f(const Tensor &t) { S3D_CHECK_CUDA(t); S3D_CHECK_CONTIGUOUS(t); S3D_CHECK_FLOAT(t); return t.data_ptr<float>(); }
static float *fm(Tensor &t) { S3D_CHECK_CUDA(t); S3D_CHECK_CONTIGUOUS(t); S3D_CHECK_FLOAT(t); return t.data_ptr<float>(); }
static int *im(Tensor &t) { S3D_CHECK_CUDA(t); S3D_CHECK_CONTIGUOUS(t); S3D_CHECK_INT(t); return t.data_ptr<int>(); }

void near_far_from_aabb(const Tensor rays_o, const Tensor rays_d, const Tensor aabb, const uint32_t N, const float min_near, Tensor nears, Tensor fars) {
    c10::cuda::CUDAGuard g(rays_o.device());
    s3d_throw(s3d_near_far_from_aabb(f(rays_o), f(rays_d), f(aabb), N, min_near, fm(nears), fm(fars), cur_stream(rays_o)), "near_far_from_aabb");
}
void sph_from_ray(const Tensor rays_o, const Tensor rays_d, const float radius, const uint32_t N, Tensor coords) {
    c10::cuda::CUDAGuard g(rays_o.device());
    s3d_throw(s3d_sph_from_ray(f(rays_o), f(rays_d), radius, N, fm(coords), cur_stream(rays_o)), "sph_from_ray");
}
void morton3D(const Tensor coords, const uint32_t N, Tensor indices) {
    c10::cuda::CUDAGuard g(coords.device());
    Tensor c = coords;
    s3d_throw(s3d_morton3D(im(c), N, im(indices), cur_stream(coords)), "morton3D");
}
void morton3D_invert(const Tensor indices, const uint32_t N, Tensor coords) {
    c10::cuda::CUDAGuard g(indices.device());
    Tensor i = indices;
    s3d_throw(s3d_morton3D_invert(im(i), N, im(coords), cur_stream(indices)), "morton3D_invert");
}
void packbits(const Tensor grid, const uint32_t N, const float density_thresh, Tensor bitfield) {
    c10::cuda::CUDAGuard g(grid.device());
    S3D_CHECK_CUDA(bitfield); S3D_CHECK_CONTIGUOUS(bitfield);
    s3d_throw(s3d_packbits(f(grid), N, density_thresh, bitfield.data_ptr<uint8_t>(), cur_stream(grid)), "packbits");
}
void march_rays_train(const Tensor rays_o, const Tensor rays_d, const Tensor grid, const float bound, const float dt_gamma, const uint32_t max_steps,
                      const uint32_t N, const uint32_t C, const uint32_t H, const uint32_t M, const Tensor nears, const Tensor fars, Tensor xyzs,
                      Tensor dirs, Tensor deltas, Tensor rays, Tensor counter, Tensor noises) {
    c10::cuda::CUDAGuard g(rays_o.device());
    S3D_CHECK_CUDA(grid); S3D_CHECK_CONTIGUOUS(grid);
    s3d_throw(s3d_march_rays_train(f(rays_o), f(rays_d), grid.data_ptr<uint8_t>(), bound, dt_gamma, max_steps, N, C, H, M, f(nears), f(fars),
                                   fm(xyzs), fm(dirs), fm(deltas), im(rays), im(counter), f(noises), cur_stream(rays_o)), "march_rays_train");
}
void composite_rays_train_forward(const Tensor sigmas, const Tensor rgbs, const Tensor deltas, const Tensor rays, const uint32_t M, const uint32_t N,
                                  const float T_thresh, Tensor weights_sum, Tensor depth, Tensor image) {
    c10::cuda::CUDAGuard g(sigmas.device());
    Tensor r = rays;
    s3d_throw(s3d_composite_rays_train_forward(f(sigmas), f(rgbs), f(deltas), im(r), M, N, T_thresh, fm(weights_sum), fm(depth), fm(image),
                                               cur_stream(sigmas)), "composite_rays_train_forward");
}
void composite_rays_train_backward(const Tensor grad_weights_sum, const Tensor grad_image, const Tensor sigmas, const Tensor rgbs, const Tensor deltas,
                                   const Tensor rays, const Tensor weights_sum, const Tensor image, const uint32_t M, const uint32_t N,
                                   const float T_thresh, Tensor grad_sigmas, Tensor grad_rgbs) {
    c10::cuda::CUDAGuard g(sigmas.device());
    Tensor r = rays;
    s3d_throw(s3d_composite_rays_train_backward(f(grad_weights_sum), f(grad_image), f(sigmas), f(rgbs), f(deltas), im(r), f(weights_sum), f(image), M,
                                                N, T_thresh, fm(grad_sigmas), fm(grad_rgbs), cur_stream(sigmas)), "composite_rays_train_backward");
}
void march_rays(const uint32_t n_alive, const uint32_t n_step, const Tensor rays_alive, const Tensor rays_t, const Tensor rays_o, const Tensor rays_d,
                const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H, const Tensor grid,
                const Tensor nears, const Tensor fars, Tensor xyzs, Tensor dirs, Tensor deltas, Tensor noises) {
    c10::cuda::CUDAGuard g(rays_o.device());
    Tensor ra = rays_alive;
    S3D_CHECK_CUDA(grid); S3D_CHECK_CONTIGUOUS(grid);
    s3d_throw(s3d_march_rays(n_alive, n_step, im(ra), f(rays_t), f(rays_o), f(rays_d), bound, dt_gamma, max_steps, C, H, grid.data_ptr<uint8_t>(),
                             f(nears), f(fars), fm(xyzs), fm(dirs), fm(deltas), f(noises), cur_stream(rays_o)), "march_rays");
}
void composite_rays(const uint32_t n_alive, const uint32_t n_step, const float T_thresh, Tensor rays_alive, Tensor rays_t, Tensor sigmas, Tensor rgbs,
                    Tensor deltas, Tensor weights, Tensor depth, Tensor image) {
    c10::cuda::CUDAGuard g(image.device());
    s3d_throw(s3d_composite_rays(n_alive, n_step, T_thresh, im(rays_alive), fm(rays_t), f(sigmas), f(rgbs), f(deltas), fm(weights), fm(depth),
                                 fm(image), cur_stream(image)), "composite_rays");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("near_far_from_aabb", &near_far_from_aabb, "near_far_from_aabb (CUDA)");
    m.def("sph_from_ray", &sph_from_ray, "sph_from_ray (CUDA)");
    m.def("morton3D", &morton3D, "morton3D (CUDA)");
    m.def("morton3D_invert", &morton3D_invert, "morton3D_invert (CUDA)");
    m.def("packbits", &packbits, "packbits (CUDA)");
    m.def("march_rays_train", &march_rays_train, "march_rays_train (CUDA)");
    m.def("composite_rays_train_forward", &composite_rays_train_forward, "composite_rays_train_forward (CUDA)");
    m.def("composite_rays_train_backward", &composite_rays_train_backward, "composite_rays_train_backward (CUDA)");
    m.def("march_rays", &march_rays, "march rays (CUDA)");
    m.def("composite_rays", &composite_rays, "composite rays (CUDA)");
}
