// upsample.cu — convex 8x upsampling of a per-pixel field (the step right after BA on every
// update when `upsample: True`: FactorGraph.update -> DepthVideo.upsample -> cvx_upsample,
// src/factor_graph.py:249-250, src/depth_video.py:194-196, src/droid_net.py:9-23).
//
//   out[b, 8y+sy, 8x+sx, :] = sum_k softmax_k(mask[b, k, sy, sx, y, x]) * data[b, y+ky-1, x+kx-1, :]
//   k = 3*ky + kx over the zero-padded 3x3 neighbourhood (F.unfold order).
//
// HBM-bound: reads the 576-channel mask once (x is its fastest axis), writes the 64x larger
// field once.  One block per (b, y, 32-wide x tile): each warp walks (sy, sx) pairs with its
// lanes along x (coalesced 64/128-byte mask reads), the 8 x 256 output tile is staged in shared
// memory and written as whole contiguous rows (a thread-per-output-pixel mapping would write
// 4-byte pieces of 32-byte sectors, 8x write amplification).
#include "common.cuh"

namespace {

constexpr int kUpTX = 32;          // x positions per block
constexpr int kUpThreads = 256;    // 8 warps
constexpr int kUpMaxDim = 4;

template <typename MT>
__device__ __forceinline__ float mask_load(const MT* p);
template <>
__device__ __forceinline__ float mask_load<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float mask_load<__half>(const __half* p) { return __half2float(__ldg(p)); }

template <typename MT, int DIM>
__global__ void __launch_bounds__(kUpThreads)
cvx_upsample_kernel(const float* __restrict__ data, const MT* __restrict__ mask, float* __restrict__ out,
                    int ht, int wd) {
  __shared__ float nb[3][kUpTX + 2][DIM];                 // data rows y-1..y+1, x0-1..x0+32
  __shared__ float tile[8][kUpTX * 8 * DIM + 4];          // output rows 8y..8y+7
  const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * kUpTX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t hw = (size_t)ht * wd;
  for (int i = tid; i < 3 * (kUpTX + 2) * DIM; i += kUpThreads) {
    const int dch = i % DIM, xi = (i / DIM) % (kUpTX + 2), r = i / (DIM * (kUpTX + 2));
    const int yy = y + r - 1, xx = x0 + xi - 1;
    nb[r][xi][dch] = (yy >= 0 && yy < ht && xx >= 0 && xx < wd)
                         ? data[(((size_t)b * ht + yy) * wd + xx) * DIM + dch] : 0.f;
  }
  __syncthreads();
  const int x = x0 + lane;
  const bool ok = x < wd;
  const MT* mb = mask + (size_t)b * 576 * hw + (size_t)y * wd + (ok ? x : 0);
  for (int s = warp; s < 64; s += kUpThreads / 32) {      // s = sy*8 + sx
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = ok ? mask_load<MT>(mb + (size_t)(k * 64 + s) * hw) : 0.f;
    float mx = m[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) mx = fmaxf(mx, m[k]);
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { m[k] = expf(m[k] - mx); den += m[k]; }
    float acc[DIM];
#pragma unroll
    for (int dch = 0; dch < DIM; ++dch) acc[dch] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float p = m[k] / den;
      if (sizeof(MT) == 2) p = __half2float(__float2half_rn(p));     // torch.softmax(half) returns half
#pragma unroll
      for (int dch = 0; dch < DIM; ++dch) acc[dch] += p * nb[k / 3][lane + (k % 3)][dch];
    }
    const int sy = s >> 3, sx = s & 7;
#pragma unroll
    for (int dch = 0; dch < DIM; ++dch) tile[sy][(lane * 8 + sx) * DIM + dch] = acc[dch];
  }
  __syncthreads();
  const int ncol = min(kUpTX, wd - x0) * 8 * DIM;           // valid floats per output row of the tile
  const size_t row_stride = (size_t)wd * 8 * DIM;
  float* ob = out + ((size_t)b * ht * 8 + (size_t)y * 8) * row_stride + (size_t)x0 * 8 * DIM;
  for (int i = tid; i < 8 * kUpTX * 8 * DIM; i += kUpThreads) {
    const int r = i / (kUpTX * 8 * DIM), c = i % (kUpTX * 8 * DIM);
    if (c < ncol) ob[(size_t)r * row_stride + c] = tile[r][c];
  }
}

template <typename MT>
int launch_up(const float* data, const MT* mask, float* out, int B, int ht, int wd, int dim, cudaStream_t st) {
  dim3 grid(gs_cdiv(wd, kUpTX), ht, B);
  switch (dim) {
    case 1: cvx_upsample_kernel<MT, 1><<<grid, kUpThreads, 0, st>>>(data, mask, out, ht, wd); break;
    case 2: cvx_upsample_kernel<MT, 2><<<grid, kUpThreads, 0, st>>>(data, mask, out, ht, wd); break;
    case 3: cvx_upsample_kernel<MT, 3><<<grid, kUpThreads, 0, st>>>(data, mask, out, ht, wd); break;
    case 4: cvx_upsample_kernel<MT, 4><<<grid, kUpThreads, 0, st>>>(data, mask, out, ht, wd); break;
    default: return GOSLAM_EINVAL;
  }
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

}  // namespace

extern "C" {

int goslam_cvx_upsample(const float* data, const void* mask, int mask_dtype, float* out, int B, int ht,
                        int wd, int dim, void* stream) {
  if (B < 0 || ht <= 0 || wd <= 0 || dim < 1 || dim > kUpMaxDim || B > 65535 || ht > 65535) return GOSLAM_EINVAL;
  if (B == 0) return GOSLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mask_dtype == GOSLAM_F32) return launch_up<float>(data, reinterpret_cast<const float*>(mask), out, B, ht, wd, dim, st);
  if (mask_dtype == GOSLAM_F16) return launch_up<__half>(data, reinterpret_cast<const __half*>(mask), out, B, ht, wd, dim, st);
  return GOSLAM_EINVAL;
}

}  // extern "C"
