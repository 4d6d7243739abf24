// ref_binding.cu — pybind wrappers that LAUNCH THE REFERENCE'S OWN CUDA KERNELS (compiled for
// sm_100a from temp copies of /root/reference/src/lib/*.cu, never vendored) so the GPU parity
// tests can pin our kernels against them.  TEST INFRASTRUCTURE: built by oracle/build_ref.py
// into oracle/_ref/, imported only by tests/.
//
// REF_SLICE is the Eigen-free part of src/lib/droid_kernels.cu (every __global__ kernel,
// accum_cuda and the frame_distance / projmap / depth_filter / iproj host wrappers); the
// Eigen-dependent host code (SparseBlock / schur_block / ba_cuda, :1117-1434) is restated in
// oracle/ref_ba_driver.py on top of the kernels exposed here.
#include REF_SLICE

std::vector<torch::Tensor> corr_index_cuda_forward(torch::Tensor volume, torch::Tensor coords, int radius);
std::vector<torch::Tensor> altcorr_cuda_forward(torch::Tensor fmap1, torch::Tensor fmap2,
                                                torch::Tensor coords, int radius);

#define ACC(t, T, N) t.packed_accessor32<T, N, torch::RestrictPtrTraits>()

std::vector<torch::Tensor> ref_linearize(torch::Tensor poses, torch::Tensor disps,
                                         torch::Tensor intrinsics, torch::Tensor targets,
                                         torch::Tensor weights, torch::Tensor ii, torch::Tensor jj) {
  auto opts = poses.options();
  const int num = ii.size(0), ht = disps.size(1), wd = disps.size(2);
  auto Hs = torch::zeros({4, num, 6, 6}, opts);
  auto vs = torch::zeros({2, num, 6}, opts);
  auto Eii = torch::zeros({num, 6, ht * wd}, opts);
  auto Eij = torch::zeros({num, 6, ht * wd}, opts);
  auto Cii = torch::zeros({num, ht * wd}, opts);
  auto wi = torch::zeros({num, ht * wd}, opts);
  projective_transform_kernel<<<num, THREADS>>>(
      ACC(targets, float, 4), ACC(weights, float, 4), ACC(poses, float, 2), ACC(disps, float, 3),
      ACC(intrinsics, float, 1), ACC(ii, long, 1), ACC(jj, long, 1), ACC(Hs, float, 4),
      ACC(vs, float, 3), ACC(Eii, float, 3), ACC(Eij, float, 3), ACC(Cii, float, 2), ACC(wi, float, 2));
  return {Hs, vs, Eii, Eij, Cii, wi};
}

torch::Tensor ref_accum(torch::Tensor data, torch::Tensor ix, torch::Tensor jx) {
  return accum_cuda(data, ix, jx);
}

torch::Tensor ref_EEt6x6(torch::Tensor E, torch::Tensor Q, torch::Tensor idx) {
  auto S = torch::zeros({idx.size(0), 6, 6}, E.options());
  if (idx.size(0) > 0)
    EEt6x6_kernel<<<idx.size(0), THREADS>>>(ACC(E, float, 3), ACC(Q, float, 2), ACC(idx, long, 2),
                                            ACC(S, float, 3));
  return S;
}

torch::Tensor ref_Ev6x1(torch::Tensor E, torch::Tensor Q, torch::Tensor w, torch::Tensor idx) {
  auto v = torch::zeros({idx.size(0), 6}, E.options());
  if (idx.size(0) > 0)
    Ev6x1_kernel<<<idx.size(0), THREADS>>>(ACC(E, float, 3), ACC(Q, float, 2), ACC(w, float, 2),
                                           ACC(idx, long, 2), ACC(v, float, 2));
  return v;
}

torch::Tensor ref_EvT6x1(torch::Tensor E, torch::Tensor x, torch::Tensor idx) {
  auto w = torch::zeros({idx.size(0), E.size(2)}, E.options());
  if (idx.size(0) > 0)
    EvT6x1_kernel<<<idx.size(0), THREADS>>>(ACC(E, float, 3), ACC(x, float, 2), ACC(idx, long, 1),
                                            ACC(w, float, 2));
  return w;
}

void ref_pose_retr(torch::Tensor poses, torch::Tensor dx, int t0, int t1) {
  pose_retr_kernel<<<1, THREADS>>>(ACC(poses, float, 2), ACC(dx, float, 2), t0, t1);
}

void ref_disp_retr(torch::Tensor disps, torch::Tensor dz, torch::Tensor inds) {
  disp_retr_kernel<<<inds.size(0), THREADS>>>(ACC(disps, float, 3), ACC(dz, float, 2),
                                              ACC(inds, long, 1));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("corr_index_forward", &corr_index_cuda_forward);
  m.def("altcorr_forward", &altcorr_cuda_forward);
  m.def("frame_distance", &frame_distance_cuda);
  m.def("projmap", &projmap_cuda);
  m.def("iproj", &iproj_cuda);
  m.def("depth_filter", &depth_filter_cuda);
  m.def("linearize", &ref_linearize);
  m.def("accum", &ref_accum);
  m.def("EEt6x6", &ref_EEt6x6);
  m.def("Ev6x1", &ref_Ev6x1);
  m.def("EvT6x1", &ref_EvT6x1);
  m.def("pose_retr", &ref_pose_retr);
  m.def("disp_retr", &ref_disp_retr);
}
