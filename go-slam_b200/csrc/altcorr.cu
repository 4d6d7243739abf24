// altcorr.cu — on-the-fly windowed correlation (no 4D volume).
//
// Reference: altcorr_forward_kernel (src/lib/altcorr_kernel.cu:27-149), called per pyramid
// level by AltCorrBlock.corr_fn (src/modules/corr.py:112-131) with fp32 NHWC feature maps.
//   s(iy,ix) = < fmap1[b,h,w,:], fmap2[b, floor(y)-r+iy, floor(x)-r+ix, :] >   (0 outside)
//   corr[b,n,ox*(2r+1)+oy,h,w] = bilinear blend of s(oy..oy+1, ox..ox+1) with (dy,dx).
//
// v1 mapping (CUDA cores, fp32 like the reference): one warp per source pixel, lane t owns
// taps t and t+32 of the 8x8 window and walks the C channels with 128-bit loads; the source
// feature vector is staged once per pixel in shared memory; a block covers 32 consecutive
// pixels and writes the 49 channels with coalesced 128-byte rows.
#include "common.cuh"

namespace {

constexpr int kPixPerBlock = 32;
constexpr int kWarps = 8;

template <int R>
__global__ void __launch_bounds__(kWarps * 32)
altcorr_kernel(const float* __restrict__ fmap1, const float* __restrict__ fmap2,
               const float* __restrict__ coords, float* __restrict__ corr, int S, int H, int W,
               int H2, int W2, int C) {
  constexpr int RD = 2 * R + 1;
  constexpr int NT = (RD + 1) * (RD + 1);          // 64 taps
  static_assert(NT == 64, "lane mapping assumes r = 3");
  extern __shared__ float smem[];
  float* f1s = smem;                               // [kWarps][C]
  float* taps = f1s + kWarps * C;                  // [kWarps][NT]
  float* stage = taps + kWarps * NT;               // [RD*RD][kPixPerBlock+1]

  const int b = blockIdx.z, s = blockIdx.y;
  const int k0 = blockIdx.x * kPixPerBlock;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = H * W;
  const float* f2b = fmap2 + (size_t)b * H2 * W2 * C;

  for (int pp = warp; pp < kPixPerBlock; pp += kWarps) {
    const int k = k0 + pp;
    if (k >= HW) break;                            // warp-uniform
    const float* f1 = fmap1 + ((size_t)b * HW + k) * C;
    for (int c = lane; c < C; c += 32) f1s[warp * C + c] = f1[c];
    const float2 xy = *reinterpret_cast<const float2*>(coords + (((size_t)b * S + s) * HW + k) * 2);
    const float fx0 = floorf(xy.x), fy0 = floorf(xy.y);
    const float dx = xy.x - fx0, dy = xy.y - fy0;
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int t = lane + 32 * half;
      const int iy = t / (RD + 1), ix = t % (RD + 1);
      const int h2 = (int)fy0 - R + iy, w2 = (int)fx0 - R + ix;
      float acc = 0.f;
      if (h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2) {
        const float* f2 = f2b + ((size_t)h2 * W2 + w2) * C;
        if ((C & 3) == 0) {
          const float4* f2v = reinterpret_cast<const float4*>(f2);
          const float4* f1v = reinterpret_cast<const float4*>(f1s + warp * C);
          for (int c = 0; c < C / 4; ++c) {
            const float4 a = f1v[c], q = __ldg(f2v + c);
            acc = fmaf(a.x, q.x, acc); acc = fmaf(a.y, q.y, acc);
            acc = fmaf(a.z, q.z, acc); acc = fmaf(a.w, q.w, acc);
          }
        } else {
          for (int c = 0; c < C; ++c) acc = fmaf(f1s[warp * C + c], __ldg(f2 + c), acc);
        }
      }
      taps[warp * NT + t] = acc;
    }
    __syncwarp();
    const float w_se = (1 - dy) * (1 - dx), w_sw = (1 - dy) * dx;
    const float w_ne = dy * (1 - dx), w_nw = dy * dx;
    for (int o = lane; o < RD * RD; o += 32) {
      const int ox = o / RD, oy = o % RD;          // channel = ox*RD + oy
      const float* tw = taps + warp * NT;
      float v = tw[oy * (RD + 1) + ox] * w_se;
      v += tw[oy * (RD + 1) + ox + 1] * w_sw;
      v += tw[(oy + 1) * (RD + 1) + ox] * w_ne;
      v += tw[(oy + 1) * (RD + 1) + ox + 1] * w_nw;
      stage[o * (kPixPerBlock + 1) + pp] = v;
    }
    __syncwarp();
  }
  __syncthreads();
  const int npx = min(kPixPerBlock, HW - k0);
  float* outp = corr + (((size_t)b * S + s) * RD * RD) * HW + k0;
  for (int idx = threadIdx.x; idx < RD * RD * kPixPerBlock; idx += kWarps * 32) {
    const int c = idx / kPixPerBlock, p = idx % kPixPerBlock;
    if (p < npx) outp[(size_t)c * HW + p] = stage[c * (kPixPerBlock + 1) + p];
  }
}


// ---------------------------------------------------------------------------------------
// AltCorrBlock.__call__ in one launch (src/modules/corr.py:112-145, S == 1): half-precision
// NHWC feature pyramids indexed per edge ON THE DEVICE (no gathered / float-converted copies:
// the reference materialises pyramid[i][:, jj].float() per level per call), all levels in one
// grid, fp32 accumulation of exact fp16 products — the same numbers the reference's fp32 kernel
// produces from the same half-valued inputs, up to summation order.
// ---------------------------------------------------------------------------------------
struct AltPyrArgs {
  const __half* pyr[4];
  int Hl[4], Wl[4];
  float inv_scale[4];
  const float* coords;      // [N, H, W, 2]
  const int64_t* ii; const int64_t* jj;
  float* out;               // [N, L*49, H*W]
  int N, H, W, C, L;
};

template <int R>
__global__ void __launch_bounds__(kWarps * 32)
altcorr_pyramid_kernel(const AltPyrArgs a) {
  constexpr int RD = 2 * R + 1;
  constexpr int NT = (RD + 1) * (RD + 1);
  extern __shared__ float smem[];
  float* f1s = smem;                               // [kWarps][C]
  float* taps = f1s + kWarps * a.C;                // [kWarps][NT]
  float* stage = taps + kWarps * NT;               // [RD*RD][kPixPerBlock+1]
  const int e = blockIdx.y, lvl = blockIdx.z;
  const int k0 = blockIdx.x * kPixPerBlock;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = a.H * a.W, C = a.C;
  const int H2 = a.Hl[lvl], W2 = a.Wl[lvl];
  const __half* f1b = a.pyr[0] + (size_t)a.ii[e] * HW * C;
  const __half* f2b = a.pyr[lvl] + (size_t)a.jj[e] * H2 * W2 * C;
  const float sc = a.inv_scale[lvl];

  for (int pp = warp; pp < kPixPerBlock; pp += kWarps) {
    const int k = k0 + pp;
    if (k >= HW) break;                            // warp-uniform
    const __half2* f1 = reinterpret_cast<const __half2*>(f1b + (size_t)k * C);
    for (int c = lane; c < C / 2; c += 32) {
      const float2 v = __half22float2(f1[c]);
      f1s[warp * C + 2 * c] = v.x; f1s[warp * C + 2 * c + 1] = v.y;
    }
    const float2 xy = *reinterpret_cast<const float2*>(a.coords + ((size_t)e * HW + k) * 2);
    const float cx = xy.x * sc, cy = xy.y * sc;
    const float fx0 = floorf(cx), fy0 = floorf(cy);
    const float dx = cx - fx0, dy = cy - fy0;
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int t = lane + 32 * hh;
      const int iy = t / (RD + 1), ix = t % (RD + 1);
      const int h2 = (int)fy0 - R + iy, w2 = (int)fx0 - R + ix;
      float acc = 0.f;
      if (h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2) {
        const uint4* f2v = reinterpret_cast<const uint4*>(f2b + ((size_t)h2 * W2 + w2) * C);
        const float4* f1v = reinterpret_cast<const float4*>(f1s + warp * C);
        for (int c = 0; c < C / 8; ++c) {
          const uint4 q = __ldg(f2v + c);
          const float4 a0 = f1v[2 * c], a1 = f1v[2 * c + 1];
          const float2 q0 = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
          const float2 q1 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
          const float2 q2 = __half22float2(*reinterpret_cast<const __half2*>(&q.z));
          const float2 q3 = __half22float2(*reinterpret_cast<const __half2*>(&q.w));
          acc = fmaf(a0.x, q0.x, acc); acc = fmaf(a0.y, q0.y, acc);
          acc = fmaf(a0.z, q1.x, acc); acc = fmaf(a0.w, q1.y, acc);
          acc = fmaf(a1.x, q2.x, acc); acc = fmaf(a1.y, q2.y, acc);
          acc = fmaf(a1.z, q3.x, acc); acc = fmaf(a1.w, q3.y, acc);
        }
      }
      taps[warp * NT + t] = acc;
    }
    __syncwarp();
    const float w_se = (1 - dy) * (1 - dx), w_sw = (1 - dy) * dx;
    const float w_ne = dy * (1 - dx), w_nw = dy * dx;
    for (int o = lane; o < RD * RD; o += 32) {
      const int ox = o / RD, oy = o % RD;
      const float* tw = taps + warp * NT;
      float v = tw[oy * (RD + 1) + ox] * w_se;
      v += tw[oy * (RD + 1) + ox + 1] * w_sw;
      v += tw[(oy + 1) * (RD + 1) + ox] * w_ne;
      v += tw[(oy + 1) * (RD + 1) + ox + 1] * w_nw;
      stage[o * (kPixPerBlock + 1) + pp] = v;
    }
    __syncwarp();
  }
  __syncthreads();
  const int npx = min(kPixPerBlock, HW - k0);
  float* outp = a.out + (((size_t)e * a.L + lvl) * RD * RD) * HW + k0;
  for (int idx = threadIdx.x; idx < RD * RD * kPixPerBlock; idx += kWarps * 32) {
    const int c = idx / kPixPerBlock, p = idx % kPixPerBlock;
    if (p < npx) outp[(size_t)c * HW + p] = stage[c * (kPixPerBlock + 1) + p];
  }
}

}  // namespace

extern "C" {

int goslam_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords,
                           float* corr, int B, int S, int H, int W, int H2, int W2, int C,
                           int radius, void* stream) {
  if (B < 0 || S <= 0 || H <= 0 || W <= 0 || H2 <= 0 || W2 <= 0 || C <= 0) return GOSLAM_EINVAL;
  if (radius != 3) return GOSLAM_EINVAL;
  if (B == 0) return GOSLAM_OK;
  if (S > 65535 || B > 65535) return GOSLAM_EINVAL;
  const size_t smem = (size_t)(kWarps * C + kWarps * 64 + 49 * (kPixPerBlock + 1)) * sizeof(float);
  if (smem > 200 * 1024) return GOSLAM_EINVAL;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(altcorr_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         200 * 1024);
    attr = true;
  }
  dim3 grid(gs_cdiv(H * W, kPixPerBlock), S, B);
  altcorr_kernel<3><<<grid, kWarps * 32, smem, (cudaStream_t)stream>>>(fmap1, fmap2, coords, corr,
                                                                       S, H, W, H2, W2, C);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_altcorr_pyramid(const void* const* pyramid, int num_levels, const float* coords,
                           const int64_t* ii, const int64_t* jj, float* out, int N, int H, int W,
                           int C, int radius, void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8) != 0 || num_levels < 1 || num_levels > 4)
    return GOSLAM_EINVAL;
  if (radius != 3 || N > 65535) return GOSLAM_EINVAL;
  if (N == 0) return GOSLAM_OK;
  AltPyrArgs a{};
  for (int l = 0; l < num_levels; ++l) {
    a.pyr[l] = reinterpret_cast<const __half*>(pyramid[l]);
    a.Hl[l] = H >> l; a.Wl[l] = W >> l;
    a.inv_scale[l] = 1.0f / (float)(1 << l);
    if (a.Hl[l] <= 0 || a.Wl[l] <= 0) return GOSLAM_EINVAL;
  }
  a.coords = coords; a.ii = ii; a.jj = jj; a.out = out;
  a.N = N; a.H = H; a.W = W; a.C = C; a.L = num_levels;
  const size_t smem = (size_t)(kWarps * C + kWarps * 64 + 49 * (kPixPerBlock + 1)) * sizeof(float);
  if (smem > 200 * 1024) return GOSLAM_EINVAL;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(altcorr_pyramid_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  dim3 grid(gs_cdiv(H * W, kPixPerBlock), N, num_levels);
  altcorr_pyramid_kernel<3><<<grid, kWarps * 32, smem, (cudaStream_t)stream>>>(a);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_altcorr_backward(void) { return GOSLAM_EUNSUPPORTED; }

}  // extern "C"
