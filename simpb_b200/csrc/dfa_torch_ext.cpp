// dfa_torch_ext.cpp — torch extension `deformable_aggregation_ext` over libdfa_b200's C ABI.
//
// Exposes exactly the two functions the reference's Python binds
// (/root/reference/projects/mmdet3d_plugin/ops/src/deformable_aggregation.cpp:127-138):
//   deformable_aggregation_forward(mc_ms_feat, spatial_shape, scale_start_index,
//                                  sampling_location, weights) -> Tensor[bs, A, C]
//   deformable_aggregation_backward(..., grad_output, grad_mc_ms_feat,
//                                   grad_sampling_location, grad_weights) -> None
// so the built module can replace the reference's .so under its unmodified Python files.
// Unlike the reference it validates its arguments, launches on PyTorch's CURRENT stream (the
// reference uses the legacy default stream, …_cuda.cu:282-283) and checks the launch.  It holds
// no device code; bfloat16 features are accepted as an extension.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "dfa_b200.h"

namespace {

void check_cuda_contig(const at::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

int feat_dtype_of(const at::Tensor &feat) {
  if (feat.scalar_type() == at::kFloat) return DFA_F32;
  if (feat.scalar_type() == at::kBFloat16) return DFA_BF16;
  TORCH_CHECK(false, "mc_ms_feat must be float32 or bfloat16, got ", feat.scalar_type());
}

dfa_dims dims_of(const at::Tensor &feat, const at::Tensor &shape, const at::Tensor &start,
                 const at::Tensor &loc, const at::Tensor &w) {
  check_cuda_contig(feat, "mc_ms_feat");
  check_cuda_contig(shape, "spatial_shape");
  check_cuda_contig(start, "scale_start_index");
  check_cuda_contig(loc, "sampling_location");
  check_cuda_contig(w, "weights");
  TORCH_CHECK(shape.scalar_type() == at::kInt && start.scalar_type() == at::kInt,
              "spatial_shape / scale_start_index must be int32");
  TORCH_CHECK(loc.scalar_type() == at::kFloat && w.scalar_type() == at::kFloat,
              "sampling_location / weights must be float32");
  TORCH_CHECK(feat.dim() == 3 && shape.dim() == 3 && shape.size(2) == 2 && start.dim() == 2 &&
                  loc.dim() == 5 && loc.size(4) == 2 && w.dim() == 6,
              "deformable_aggregation: wrong tensor ranks");
  dfa_dims d;
  d.batch_size = feat.size(0), d.num_feat = feat.size(1), d.num_embeds = feat.size(2);
  d.num_cams = shape.size(0), d.num_scale = shape.size(1);
  d.num_anchors = loc.size(1), d.num_pts = loc.size(2), d.num_groups = w.size(5);
  TORCH_CHECK(start.size(0) == d.num_cams && start.size(1) == d.num_scale,
              "scale_start_index must be [num_cams, num_scale]");
  TORCH_CHECK(loc.size(0) == d.batch_size && loc.size(3) == d.num_cams,
              "sampling_location must be [bs, anchors, pts, cams, 2]");
  TORCH_CHECK(w.size(0) == d.batch_size && w.size(1) == d.num_anchors && w.size(2) == d.num_pts &&
                  w.size(3) == d.num_cams && w.size(4) == d.num_scale,
              "weights must be [bs, anchors, pts, cams, scales, groups]");
  for (const at::Tensor *t : {&shape, &start, &loc, &w})
    TORCH_CHECK(t->device() == feat.device(), "all tensors must live on the same device");
  return d;
}

void raise_on(int rc, const char *what) {
  TORCH_CHECK(rc == 0, what, " failed: ", dfa_error_string(rc), " (code ", rc, ")");
}

at::Tensor deformable_aggregation_forward(const at::Tensor &feat, const at::Tensor &shape,
                                          const at::Tensor &start, const at::Tensor &loc,
                                          const at::Tensor &w) {
  const dfa_dims d = dims_of(feat, shape, start, loc, w);
  const c10::cuda::CUDAGuard guard(feat.device());
  // the kernel writes every element: no zero-fill (reference: at::zeros, .cpp:55)
  at::Tensor out = at::empty({d.batch_size, d.num_anchors, d.num_embeds},
                             feat.options().dtype(at::kFloat));
  raise_on(dfa_forward(feat.data_ptr(), feat_dtype_of(feat), shape.data_ptr<int>(),
                       start.data_ptr<int>(), loc.data_ptr<float>(), w.data_ptr<float>(),
                       out.data_ptr<float>(), &d, at::cuda::getCurrentCUDAStream().stream()),
           "deformable_aggregation_forward");
  return out;
}

void deformable_aggregation_backward(const at::Tensor &feat, const at::Tensor &shape,
                                     const at::Tensor &start, const at::Tensor &loc,
                                     const at::Tensor &w, const at::Tensor &grad_output,
                                     at::Tensor &grad_feat, at::Tensor &grad_loc,
                                     at::Tensor &grad_w) {
  const dfa_dims d = dims_of(feat, shape, start, loc, w);
  check_cuda_contig(grad_output, "grad_output");
  check_cuda_contig(grad_feat, "grad_mc_ms_feat");
  check_cuda_contig(grad_loc, "grad_sampling_location");
  check_cuda_contig(grad_w, "grad_weights");
  TORCH_CHECK(grad_output.scalar_type() == at::kFloat && grad_feat.scalar_type() == at::kFloat &&
                  grad_loc.scalar_type() == at::kFloat && grad_w.scalar_type() == at::kFloat,
              "gradient tensors must be float32");
  TORCH_CHECK(grad_output.numel() == (int64_t)d.batch_size * d.num_anchors * d.num_embeds &&
                  grad_feat.numel() == feat.numel() && grad_loc.numel() == loc.numel() &&
                  grad_w.numel() == w.numel(),
              "gradient tensor sizes do not match their inputs");
  const c10::cuda::CUDAGuard guard(feat.device());
  // reference contract: accumulate into the caller's pre-zeroed buffers
  raise_on(dfa_backward(feat.data_ptr(), feat_dtype_of(feat), shape.data_ptr<int>(),
                        start.data_ptr<int>(), loc.data_ptr<float>(), w.data_ptr<float>(),
                        grad_output.data_ptr<float>(), grad_feat.data_ptr<float>(),
                        grad_loc.data_ptr<float>(), grad_w.data_ptr<float>(), &d,
                        DFA_BWD_ACCUMULATE, at::cuda::getCurrentCUDAStream().stream()),
           "deformable_aggregation_backward");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("deformable_aggregation_forward", &deformable_aggregation_forward,
        "deformable_aggregation_forward");
  m.def("deformable_aggregation_backward", &deformable_aggregation_backward,
        "deformable_aggregation_backward");
}
