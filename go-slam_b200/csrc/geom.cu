// geom.cu — per-pixel reprojection kernels of the droid_backends boundary:
// frame_distance, projmap, iproj, depth_filter and the DepthVideo.reproject fusion.
// All are a few flops per 4-byte pixel => HBM/latency bound; one thread per pixel,
// coalesced along x, relative pose computed once per block into shared memory.
#include "common.cuh"
#include "se3.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------
// frame_distance  (reference: src/lib/droid_kernels.cu:518-657)
// One block per (i,j) pair.  The float summation order is part of the contract: the
// distances are sorted / thresholded into factor-graph edges, so we keep the reference's
// association exactly — thread t sums pixels t, t+256, ... serially, then a fixed
// 128 / 64 / 32 / 16 / 8 / 4 / 2 / 1 tree (src/lib/droid_kernels.cu:36-55).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float tree256(float v, float* s) {
  const int tid = threadIdx.x;
  s[tid] = v;
  __syncthreads();
  if (tid < 128) s[tid] += s[tid + 128];
  __syncthreads();
  if (tid < 64) s[tid] += s[tid + 64];
  __syncthreads();
  float r = 0.f;
  if (tid < 32) {
    r = s[tid] + s[tid + 32];
    r += __shfl_down_sync(0xffffffffu, r, 16);
    r += __shfl_down_sync(0xffffffffu, r, 8);
    r += __shfl_down_sync(0xffffffffu, r, 4);
    r += __shfl_down_sync(0xffffffffu, r, 2);
    r += __shfl_down_sync(0xffffffffu, r, 1);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// distance of the ordered pair (ix -> jx); the value is valid in thread 0.  NOT inlined: the one-way and the
// bidirectional kernel must run the very same instruction sequence (FMA contraction included) so that
// 0.5 * (d_ij + d_ji) is bit-identical in both forms.
__device__ __noinline__ float pair_distance(const float* __restrict__ poses, const float* __restrict__ disps,
                                               const float* __restrict__ intr, int ix, int jx, int ht, int wd,
                                               float beta, float* red) {
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];

  // NB: no stereo special case here — the reference calls relSE3 unconditionally.
  GsSE3 G;
  {
    const float* pi = poses + 7 * (size_t)ix;
    const float* pj = poses + 7 * (size_t)jx;
    gs_rel(pi, pi + 3, pj, pj + 3, G);
  }

  float accum = 0.f, valid = 0.f, total = 0.f;
  const float* dsp = disps + (size_t)ix * ht * wd;
  const float omb = 1 - beta;
  for (int k = threadIdx.x; k < ht * wd; k += kThreads) {
    const float u = (float)(k % wd);
    const float v = (float)(k / wd);
    float Xi[4], Xj[4];
    Xi[0] = (u - cx) / fx;
    Xi[1] = (v - cy) / fy;
    Xi[2] = 1.f;
    Xi[3] = dsp[k];
    gs_act4(G, Xi, Xj);

    float du = fx * (Xj[0] / Xj[2]) + cx - u;
    float dv = fy * (Xj[1] / Xj[2]) + cy - v;
    float d = sqrtf(du * du + dv * dv);
    total += beta;
    if (Xj[2] > GS_MIN_DEPTH) {
      accum += beta * d;
      valid += beta;
    }

    // translation-only term
    Xj[0] = Xi[0] + Xi[3] * G.t[0];
    Xj[1] = Xi[1] + Xi[3] * G.t[1];
    Xj[2] = Xi[2] + Xi[3] * G.t[2];
    du = fx * (Xj[0] / Xj[2]) + cx - u;
    dv = fy * (Xj[1] / Xj[2]) + cy - v;
    d = sqrtf(du * du + dv * dv);
    total += omb;
    if (Xj[2] > GS_MIN_DEPTH) {
      accum += omb * d;
      valid += omb;
    }
  }
  const float a = tree256(accum, red);
  const float t = tree256(total, red);
  const float w = tree256(valid, red);
  return (w / (t + 1e-8f) < 0.75f) ? 1000.0f : a / w;
}

__global__ void __launch_bounds__(kThreads)
frame_distance_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                      const float* __restrict__ intr, const int64_t* __restrict__ ii,
                      const int64_t* __restrict__ jj, float* __restrict__ dist,
                      int ht, int wd, float beta) {
  __shared__ float red[kThreads];
  const float d = pair_distance(poses, disps, intr, (int)ii[blockIdx.x], (int)jj[blockIdx.x], ht, wd, beta, red);
  if (threadIdx.x == 0) dist[blockIdx.x] = d;
}

// DepthVideo.distance(bidirectional=True) (src/depth_video.py:233-245): 0.5 * (d(i->j) + d(j->i)) in one
// launch; each direction keeps the reduction tree above, so the result equals the two-launch form bit for bit.
__global__ void __launch_bounds__(kThreads)
frame_distance_bidir_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                            const float* __restrict__ intr, const int64_t* __restrict__ ii,
                            const int64_t* __restrict__ jj, float* __restrict__ dist,
                            int ht, int wd, float beta) {
  __shared__ float red[kThreads];
  const int ix = (int)ii[blockIdx.x], jx = (int)jj[blockIdx.x];
  const float d1 = pair_distance(poses, disps, intr, ix, jx, ht, wd, beta, red);
  const float d2 = pair_distance(poses, disps, intr, jx, ix, ht, wd, beta, red);
  if (threadIdx.x == 0) dist[blockIdx.x] = __fmul_rn(0.5f, __fadd_rn(d1, d2));
}

// ---------------------------------------------------------------------------------
// projmap  (reference: src/lib/droid_kernels.cu:427-516)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
projmap_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
               const float* __restrict__ intr, const int64_t* __restrict__ ii,
               const int64_t* __restrict__ jj, float* __restrict__ coords,
               float* __restrict__ valid, int ht, int wd) {
  const int e = blockIdx.y;
  const int k = blockIdx.x * kThreads + threadIdx.x;
  const int ix = (int)ii[e], jx = (int)jj[e];
  __shared__ GsSE3 G;
  if (threadIdx.x == 0) {
    const float* pi = poses + 7 * (size_t)ix;
    const float* pj = poses + 7 * (size_t)jx;
    gs_rel(pi, pi + 3, pj, pj + 3, G);
  }
  __syncthreads();
  if (k >= ht * wd) return;
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float u = (float)(k % wd), v = (float)(k / wd);
  float Xi[4] = {(u - cx) / fx, (v - cy) / fy, 1.f, disps[(size_t)ix * ht * wd + k]};
  float Xj[4];
  gs_act4(G, Xi, Xj);
  float cu = u, cv = v;
  if (Xj[2] > 0.01f) {
    cu = fx * (Xj[0] / Xj[2]) + cx;
    cv = fy * (Xj[1] / Xj[2]) + cy;
  }
  float* c = coords + ((size_t)e * ht * wd + k) * 3;
  c[0] = cu; c[1] = cv; c[2] = 0.f;
  valid[(size_t)e * ht * wd + k] = (Xj[2] > GS_MIN_DEPTH) ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------
// reproject: DepthVideo.reproject -> pops.projective_transform(jacobian=False)
// (src/depth_video.py:207-217, src/geom/projective_ops.py:26-44,54-57,88-99,114-144).
// Python-side constants: MIN_DEPTH = 0.2, Z < 0.1 replaced by 1 before dividing.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
reproject_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                 const float* __restrict__ intr_all, const int64_t* __restrict__ ii,
                 const int64_t* __restrict__ jj, float* __restrict__ coords,
                 float* __restrict__ valid, const float* __restrict__ target,
                 float* __restrict__ motion, int ht, int wd) {
  const int e = blockIdx.y;
  const int k = blockIdx.x * kThreads + threadIdx.x;
  const int ix = (int)ii[e], jx = (int)jj[e];
  __shared__ GsSE3 G;
  if (threadIdx.x == 0) gs_edge_pose(poses, ix, jx, G);
  __syncthreads();
  if (k >= ht * wd) return;
  const float* Ki = intr_all + 4 * (size_t)ix;
  const float* Kj = intr_all + 4 * (size_t)jx;
  const float u = (float)(k % wd), v = (float)(k / wd);
  float X0[4] = {(u - Ki[2]) / Ki[0], (v - Ki[3]) / Ki[1], 1.f,
                 disps[(size_t)ix * ht * wd + k]};
  float X1[4];
  gs_act4(G, X0, X1);
  const float Z = (X1[2] < 0.5f * 0.2f) ? 1.0f : X1[2];
  float2 c;
  c.x = Kj[0] * (X1[0] / Z) + Kj[2];
  c.y = Kj[1] * (X1[1] / Z) + Kj[3];
  reinterpret_cast<float2*>(coords)[(size_t)e * ht * wd + k] = c;
  if (valid) valid[(size_t)e * ht * wd + k] = (X1[2] > 0.2f) ? 1.0f : 0.0f;
  if (motion) {
    // FactorGraph.update's motion features (src/factor_graph.py:204-206):
    // cat([coords1 - coords0, target - coords1], -1).permute(0,1,4,2,3).clamp(-64, 64)
    const float2 t = reinterpret_cast<const float2*>(target)[(size_t)e * ht * wd + k];
    float* m = motion + (size_t)e * 4 * ht * wd + k;
    const size_t hw = (size_t)ht * wd;
    m[0] = fminf(fmaxf(c.x - u, -64.f), 64.f);
    m[hw] = fminf(fmaxf(c.y - v, -64.f), 64.f);
    m[2 * hw] = fminf(fmaxf(t.x - c.x, -64.f), 64.f);
    m[3 * hw] = fminf(fmaxf(t.y - c.y, -64.f), 64.f);
  }
}

// ---------------------------------------------------------------------------------
// iproj  (reference: src/lib/droid_kernels.cu:779-850)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
iproj_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
             const float* __restrict__ intr, float* __restrict__ points, int ht, int wd) {
  const int f = blockIdx.y;
  const int k = blockIdx.x * kThreads + threadIdx.x;
  if (k >= ht * wd) return;
  GsSE3 G;
  const float* p = poses + 7 * (size_t)f;
  G.t[0] = p[0]; G.t[1] = p[1]; G.t[2] = p[2];
  G.q[0] = p[3]; G.q[1] = p[4]; G.q[2] = p[5]; G.q[3] = p[6];
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float u = (float)(k % wd), v = (float)(k / wd);
  float Xi[4] = {(u - cx) / fx, (v - cy) / fy, 1.f, disps[(size_t)f * ht * wd + k]};
  float Xj[4];
  gs_act4(G, Xi, Xj);
  float* o = points + ((size_t)f * ht * wd + k) * 3;
  o[0] = Xj[0] / Xj[3];
  o[1] = Xj[1] / Xj[3];
  o[2] = Xj[2] / Xj[3];
}

// ---------------------------------------------------------------------------------
// depth_filter  (reference: src/lib/droid_kernels.cu:661-775).  The reference scatters
// atomicAdd over a (frame, neighbour, tile) grid; each output pixel only ever receives
// its own 6 neighbour votes, so we loop the 6 neighbours inside one thread instead:
// same counts, no atomics, no memset.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
depth_filter_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                    const float* __restrict__ intr, const int64_t* __restrict__ inds,
                    const float* __restrict__ thresh, float* __restrict__ counter,
                    int num, int ht, int wd) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * kThreads + threadIdx.x;
  const int ix = (int)inds[b];
  __shared__ GsSE3 G[6];
  __shared__ int jxs[6];
  if (threadIdx.x < 6) {
    const int nb = threadIdx.x;
    const int jx = (nb < 3) ? ix - nb - 1 : ix + nb;
    jxs[nb] = jx;
    if (jx >= 0 && jx < num) {
      const float* pi = poses + 7 * (size_t)ix;
      const float* pj = poses + 7 * (size_t)jx;
      gs_rel(pi, pi + 3, pj, pj + 3, G[nb]);
    }
  }
  __syncthreads();
  if (k >= ht * wd) return;
  const float t = thresh[b];
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const int i = k / wd, j = k % wd;
  const float ui = (float)j, vi = (float)i;
  float Xi[4] = {(ui - cx) / fx, (vi - cy) / fy, 1.f, disps[(size_t)ix * ht * wd + k]};
  float count = 0.f;
#pragma unroll
  for (int nb = 0; nb < 6; ++nb) {
    const int jx = jxs[nb];
    if (jx < 0 || jx >= num) continue;
    float Xj[4];
    gs_act4(G[nb], Xi, Xj);
    const float uj = fx * (Xj[0] / Xj[2]) + cx;
    const float vj = fy * (Xj[1] / Xj[2]) + cy;
    const float dj = Xj[3] / Xj[2];
    const int u0 = (int)floorf(uj);
    const int v0 = (int)floorf(vj);
    if (u0 >= 0 && v0 >= 0 && u0 < wd - 1 && v0 < ht - 1) {
      const float* dj_map = disps + (size_t)jx * ht * wd;
      const float d00 = dj_map[(v0 + 0) * wd + u0 + 0];
      const float d01 = dj_map[(v0 + 0) * wd + u0 + 1];
      const float d10 = dj_map[(v0 + 1) * wd + u0 + 0];
      const float d11 = dj_map[(v0 + 1) * wd + u0 + 1];
      // the reference evaluates these in double (1.0/dj literals), then compares to float t
      const double inv = 1.0 / dj;
      if (fabs(inv - 1.0 / d00) < t) count += 1.0f;
      else if (fabs(inv - 1.0 / d01) < t) count += 1.0f;
      else if (fabs(inv - 1.0 / d10) < t) count += 1.0f;
      else if (fabs(inv - 1.0 / d11) < t) count += 1.0f;
    }
  }
  counter[(size_t)b * ht * wd + k] = count;
}

}  // namespace

extern "C" {

int goslam_frame_distance(const float* poses, const float* disps, const float* intrinsics,
                          const int64_t* ii, const int64_t* jj, float* dist, int K, int ht,
                          int wd, float beta, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  frame_distance_kernel<<<K, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ii,
                                                                  jj, dist, ht, wd, beta);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_frame_distance_bidir(const float* poses, const float* disps, const float* intrinsics,
                                const int64_t* ii, const int64_t* jj, float* dist, int K, int ht, int wd,
                                float beta, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  frame_distance_bidir_kernel<<<K, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ii, jj, dist, ht,
                                                                        wd, beta);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_projmap(const float* poses, const float* disps, const float* intrinsics,
                   const int64_t* ii, const int64_t* jj, float* coords, float* valid, int K,
                   int ht, int wd, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(ht * wd, kThreads), K);
  projmap_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ii, jj,
                                                              coords, valid, ht, wd);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_reproject(const float* poses, const float* disps, const float* intrinsics_all,
                     const int64_t* ii, const int64_t* jj, float* coords, float* valid, int K,
                     int ht, int wd, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(ht * wd, kThreads), K);
  reproject_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics_all, ii,
                                                                jj, coords, valid, nullptr, nullptr, ht, wd);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_reproject_motion(const float* poses, const float* disps, const float* intrinsics_all,
                            const int64_t* ii, const int64_t* jj, const float* target, float* coords,
                            float* valid, float* motion, int K, int ht, int wd, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0 || target == nullptr || motion == nullptr) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(ht * wd, kThreads), K);
  reproject_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics_all, ii,
                                                                jj, coords, valid, target, motion, ht, wd);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_iproj(const float* poses, const float* disps, const float* intrinsics, float* points,
                 int num, int ht, int wd, void* stream) {
  if (num < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (num == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(ht * wd, kThreads), num);
  iproj_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, points, ht,
                                                            wd);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_depth_filter(const float* poses, const float* disps, const float* intrinsics,
                        const int64_t* ix, const float* thresh, float* counter, int K, int num,
                        int ht, int wd, void* stream) {
  if (K < 0 || ht <= 0 || wd <= 0) return GOSLAM_EINVAL;
  if (K == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(ht * wd, kThreads), K);
  depth_filter_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ix,
                                                                   thresh, counter, num, ht, wd);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

}  // extern "C"
