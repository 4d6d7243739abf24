// conv_tc.cu — the update operator's ConvGRU on tcgen05: 3x3 / 1x1 convolutions as an
// implicit GEMM with the GRU gates fused into the epilogues.
//
// Replaces ConvGRU.forward (src/modules/gru.py:21-39) as called by UpdateModule.forward
// (src/droid_net.py:125, `self.gru(net, inp, corr, flow)`):
//     glo = mean_hw(sigmoid(w(net)) * net)                                   1x1 conv + pooling
//     z = sigmoid(convz([net|inp|corr|flow]) + convz_glo(glo))               3x3, 448 -> 128
//     r = sigmoid(convr([net|inp|corr|flow]) + convr_glo(glo))               3x3, 448 -> 128
//     q = tanh(convq([r*net|inp|corr|flow]) + convq_glo(glo))                3x3, 448 -> 128
//     net' = (1 - z) * net + z * q
// The reference runs 7 cuDNN convolutions, 2 torch.cat of the 448-channel input and ~12 elementwise kernels.
// Here: three launches of ONE kernel (conv_tc_kernel) with different epilogues, plus two tiny ones:
//   pass G  1x1 conv of net, epilogue: sigmoid(.)*net, per-image channel sums       -> glo_sum [B,128]
//   (gru_glo_fc_kernel: the three 128x128 matvecs on the pooled vector              -> glo     [B,384])
//   pass ZR 3x3 conv with z and r stacked to N = 256 (the activation tile is loaded once for both
//           gates), epilogue: bias + glo, sigmoid, z and r*net written                -> z, rnet
//   pass Q  3x3 conv over [rnet|inp|corr|flow], epilogue: tanh, (1-z)*net + z*q       -> net'
//
// Implicit GEMM: M = 128 output pixels (an 8x16 image patch), N = output channels (128 or 256),
// K = taps x input channels in chunks of 64.  Activations are NHWC fp16, so the patch shifted by a tap is ONE
// TMA box (64 ch, 16, 8, 1) of a 4-D tensor map (ch, x, y, image) — image borders are the TMA's zero fill, the
// concatenated input never exists (one tensor map per source tensor) — and lands K-major / SWIZZLE_128B, exactly
// the A operand.  Weights are [tap][cout][cin] fp16: box (64, N, 1) = the B operand.  tcgen05.mma M128 N{128,256}
// K16 accumulates in TMEM (two accumulator stages, so the epilogue of a tile overlaps the MMAs of the next);
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (one per TMEM lane quadrant).
// fp16 operands, fp32 accumulation, fp32 gate arithmetic, fp16 state — what the reference's autocast region
// computes (src/factor_graph.py:198, torch.cuda.amp.autocast) with one rounding less per gate.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <mutex>

using namespace gs_tc;

namespace {

constexpr int kPY = 8, kPX = 16, kBM = kPY * kPX;   // output patch = MMA M
constexpr int kKC = 64;                             // channels per K chunk (128-byte rows)
constexpr int kABytes = kBM * kKC * 2;              // 16 KB
constexpr int kMaxN = 256;
constexpr int kBBytes = kMaxN * kKC * 2;            // 32 KB (a 128-channel pass uses half)
constexpr int kStages = 4;
constexpr int kThreadsC = 6 * 32;
constexpr int kSmemC = 1024 + kStages * (kABytes + kBBytes) + 256;
constexpr int kMaxIn = 4;

enum { EPI_GLO = 1, EPI_ZR = 2, EPI_Q = 3 };

struct ConvMaps {
  CUtensorMap in[kMaxIn];
  CUtensorMap w;
};

struct ConvParams {
  int B, h, w, n_yb, n_xb, n_tiles;
  int taps;                 // 1 or 9
  int n_in;
  int chunks[kMaxIn];       // K chunks (64 channels) of each input tensor
  int N;                    // output channels of the pass (128 or 256)
  int epi;
  const float* bias;        // [N]
  const float* glo;         // [B, 384] (z | r | q) or nullptr
  const __half* net;        // [B, h, w, 128] NHWC
  const __half* z_in;       // EPI_Q
  __half* z_out;            // EPI_ZR
  __half* rnet_out;         // EPI_ZR
  __half* net_out;          // EPI_Q
  float* glo_sum;           // EPI_GLO: [B, 128] += sum over pixels of sigmoid(.) * net
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(kThreadsC, 1)
conv_tc_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smA = base;
  unsigned char* smB = base + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * (kABytes + kBBytes));
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = full + kStages;         // [kStages]
  uint64_t* tm_full = empty + kStages;      // [2]
  uint64_t* tm_empty = tm_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * kMaxN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  int kiters = 0;
  for (int i = 0; i < p.n_in; ++i) kiters += p.chunks[i];
  kiters *= p.taps;
  const uint32_t stage_tx = kABytes + (uint32_t)p.N * kKC * 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int xb = tile % p.n_xb, yb = (tile / p.n_xb) % p.n_yb, b = tile / (p.n_xb * p.n_yb);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          int gchunk = 0;
          for (int ci = 0; ci < p.n_in; ++ci)
            for (int kc = 0; kc < p.chunks[ci]; ++kc, ++gchunk) {
              mbar_wait(&empty[s], ph ^ 1);
              mbar_expect_tx(&full[s], stage_tx);
              tma_load_4d(&maps.in[ci], &full[s], smA + s * kABytes, kc * kKC, xb * kPX + dx, yb * kPY + dy, b);
              tma_load_3d(&maps.w, &full[s], smB + s * kBBytes, gchunk * kKC, 0, tap);
              if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, p.N);
      int s = 0, ph = 0, ts = 0, tph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(&tm_empty[ts], tph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ts * kMaxN;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + s * kABytes), b_addr = smem_u32(smB + s * kBBytes);
#pragma unroll
          for (int k = 0; k < kKC / 16; ++k)
            umma_f16(d_tmem, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32), idesc,
                     (it | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tm_full[ts]);
        if (++ts == 2) { ts = 0; tph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                // pixel of the patch
    const int py = row / kPX, px = row % kPX;
    int ts = 0, tph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int xb = tile % p.n_xb, yb = (tile / p.n_xb) % p.n_yb, b = tile / (p.n_xb * p.n_yb);
      const int y = yb * kPY + py, x = xb * kPX + px;
      const bool ok = y < p.h && x < p.w;
      const size_t pix = ((size_t)b * p.h + (ok ? y : 0)) * p.w + (ok ? x : 0);
      mbar_wait(&tm_full[ts], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ts * kMaxN + ((uint32_t)(quad * 32) << 16);
      const float* glo = p.glo ? p.glo + (size_t)b * 384 : nullptr;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        const int ch = c0 & 127;                     // channel of the 128-channel state this chunk maps to
        uint4 netv[4];
        if (ok) {
          const uint4* np = reinterpret_cast<const uint4*>(p.net + pix * 128 + ch);
#pragma unroll
          for (int i = 0; i < 4; ++i) netv[i] = __ldg(np + i);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) netv[i] = make_uint4(0, 0, 0, 0);
        }
        const __half2* nh = reinterpret_cast<const __half2*>(netv);
        if (p.epi == EPI_GLO) {
          // g = sigmoid(conv + b) * net; per-image channel sums (the mean's divisor is applied by the fc kernel)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float nv = (i & 1) ? __high2float(nh[i >> 1]) : __low2float(nh[i >> 1]);
            float g = ok ? sigmoidf_(__uint_as_float(v[i]) + __ldg(p.bias + c0 + i)) * nv : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
            if (lane == 0) atomicAdd(p.glo_sum + (size_t)b * 128 + c0 + i, g);
          }
        } else if (p.epi == EPI_ZR) {
          const bool is_r = c0 >= 128;
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float g0 = sigmoidf_(__uint_as_float(v[i]) + __ldg(p.bias + c0 + i) + glo[c0 + i]);
            const float g1 = sigmoidf_(__uint_as_float(v[i + 1]) + __ldg(p.bias + c0 + i + 1) + glo[c0 + i + 1]);
            const float2 nv = __half22float2(nh[i >> 1]);
            const __half2 h = is_r ? __floats2half2_rn(g0 * nv.x, g1 * nv.y) : __floats2half2_rn(g0, g1);
            o[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (ok) {
            uint4* dst = reinterpret_cast<uint4*>((is_r ? p.rnet_out : p.z_out) + pix * 128 + ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        } else {   // EPI_Q
          uint4 zv[4];
          if (ok) {
            const uint4* zp = reinterpret_cast<const uint4*>(p.z_in + pix * 128 + ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) zv[i] = __ldg(zp + i);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) zv[i] = make_uint4(0, 0, 0, 0);
          }
          const __half2* zh = reinterpret_cast<const __half2*>(zv);
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float q0 = tanhf(__uint_as_float(v[i]) + __ldg(p.bias + c0 + i) + glo[256 + c0 + i]);
            const float q1 = tanhf(__uint_as_float(v[i + 1]) + __ldg(p.bias + c0 + i + 1) + glo[256 + c0 + i + 1]);
            const float2 nv = __half22float2(nh[i >> 1]), zz = __half22float2(zh[i >> 1]);
            const __half2 h = __floats2half2_rn((1.0f - zz.x) * nv.x + zz.x * q0, (1.0f - zz.y) * nv.y + zz.y * q1);
            o[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.net_out + pix * 128 + ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
      }
      // all TMEM reads of this warp for this stage are complete (tmem_ld32 waits): hand the stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tm_empty[ts]);
      if (++ts == 2) { ts = 0; tph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kMaxN);
}

// glo[b] = W_glo (glo_sum[b] / hw) + b_glo : the three 1x1 "global" convolutions on the pooled vector
__global__ void gru_glo_fc_kernel(const float* __restrict__ glo_sum, const float* __restrict__ w_glo,
                                  const float* __restrict__ b_glo, float* __restrict__ glo, float inv_hw) {
  __shared__ float v[128];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) v[i] = glo_sum[(size_t)b * 128 + i] * inv_hw;
  __syncthreads();
  for (int o = threadIdx.x; o < 384; o += blockDim.x) {
    float s = b_glo[o];
    const float* wr = w_glo + (size_t)o * 128;
#pragma unroll 8
    for (int k = 0; k < 128; ++k) s += wr[k] * v[k];
    glo[(size_t)b * 384 + o] = s;
  }
}

// [B, C, hw] <-> [B, hw, C] fp16 through a 32x32 shared-memory tile
__global__ void transpose_f16_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
  __shared__ __half t[32][34];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[i][threadIdx.x] = src[boff + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[boff + (size_t)c * rows + r] = t[threadIdx.x][i];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn conv_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// activation map: NHWC [B, h, w, C] as (ch, x, y, image), box 64 ch x 16 x 8
bool act_map(EncodeTiledFn enc, const void* base, int B, int h, int w, int C, CUtensorMap* out) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)w * C * 2, (cuuint64_t)h * w * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kKC, (cuuint32_t)kPX, (cuuint32_t)kPY, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weight map: [taps, N, Cin] as (cin, cout, tap), box 64 x N
bool weight_map(EncodeTiledFn enc, const void* base, int taps, int N, int Cin, CUtensorMap* out) {
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)N * Cin * 2};
  cuuint32_t box[3] = {(cuuint32_t)kKC, (cuuint32_t)N, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int conv_launch(const ConvMaps& maps, ConvParams p, cudaStream_t st) {
  static int sm_count[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int sms;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (sm_count[dev] == 0) {
      if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemC) != cudaSuccess)
        return GOSLAM_ELAUNCH;
      int n = 148;
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      sm_count[dev] = n > 0 ? n : 148;
    }
    sms = sm_count[dev];
  }
  p.n_yb = gs_cdiv(p.h, kPY); p.n_xb = gs_cdiv(p.w, kPX);
  p.n_tiles = p.B * p.n_yb * p.n_xb;
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  conv_tc_kernel<<<grid, kThreadsC, kSmemC, st>>>(maps, p);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

struct GruWs {
  float* glo_sum;   // [B,128]
  float* glo;       // [B,384]
  __half* z;        // [B,h,w,128]
  __half* rnet;     // [B,h,w,128]
};
size_t gru_layout(int B, int h, int w, void* base, size_t cap, GruWs* ws) {
  GsArena a(base, cap);
  GruWs g{};
  g.glo_sum = a.take<float>((size_t)B * 128);
  g.glo = a.take<float>((size_t)B * 384);
  g.z = a.take<__half>((size_t)B * h * w * 128);
  g.rnet = a.take<__half>((size_t)B * h * w * 128);
  if (ws) *ws = g;
  return a.off;
}

}  // namespace

extern "C" {

int goslam_nchw_to_nhwc_f16(const void* src, void* dst, int B, int C, int hw, void* stream) {
  if (B < 0 || C <= 0 || hw <= 0) return GOSLAM_EINVAL;
  if (B == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(hw, 32), gs_cdiv(C, 32), B), block(32, 8);
  transpose_f16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(src),
                                                                  reinterpret_cast<__half*>(dst), C, hw);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_nhwc_to_nchw_f16(const void* src, void* dst, int B, int C, int hw, void* stream) {
  if (B < 0 || C <= 0 || hw <= 0) return GOSLAM_EINVAL;
  if (B == 0) return GOSLAM_OK;
  dim3 grid(gs_cdiv(C, 32), gs_cdiv(hw, 32), B), block(32, 8);
  transpose_f16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(src),
                                                                  reinterpret_cast<__half*>(dst), hw, C);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

size_t goslam_conv_gru_workspace_bytes(int B, int h, int w) {
  if (B <= 0 || h <= 0 || w <= 0) return 256;
  return gru_layout(B, h, w, nullptr, 0, nullptr) + 256;
}

int goslam_conv_gru(const goslam_gru_weights* wts, const void* net, const void* inp, const void* corr,
                    const void* flow, void* net_out, int B, int h, int w, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (!wts || B < 0 || h <= 0 || w <= 0) return GOSLAM_EINVAL;
  if (B == 0) return GOSLAM_OK;
  GruWs ws;
  const size_t need = gru_layout(B, h, w, workspace, workspace_bytes, &ws);
  if (workspace == nullptr || need > workspace_bytes) return GOSLAM_EWORKSPACE;
  EncodeTiledFn enc = conv_encode_fn();
  if (!enc) return GOSLAM_ELAUNCH;
  cudaStream_t st = (cudaStream_t)stream;
  ConvMaps m{};
  ConvParams p{};
  p.B = B; p.h = h; p.w = w;
  p.net = reinterpret_cast<const __half*>(net);
  // ---- pass G: glo_sum = sum_px sigmoid(w(net)) * net
  cudaMemsetAsync(ws.glo_sum, 0, (size_t)B * 128 * sizeof(float), st);
  if (!act_map(enc, net, B, h, w, 128, &m.in[0]) || !weight_map(enc, wts->w_w, 1, 128, 128, &m.w)) return GOSLAM_ELAUNCH;
  p.taps = 1; p.n_in = 1; p.chunks[0] = 2; p.N = 128; p.epi = EPI_GLO; p.bias = wts->b_w; p.glo = nullptr;
  p.glo_sum = ws.glo_sum;
  int rc = conv_launch(m, p, st);
  if (rc) return rc;
  gru_glo_fc_kernel<<<B, 128, 0, st>>>(ws.glo_sum, wts->w_glo, wts->b_glo, ws.glo, 1.0f / (float)(h * w));
  GS_CHECK_LAUNCH();
  // ---- pass ZR: z, r*net
  if (!act_map(enc, inp, B, h, w, 128, &m.in[1]) || !act_map(enc, corr, B, h, w, 128, &m.in[2]) ||
      !act_map(enc, flow, B, h, w, 64, &m.in[3]) || !weight_map(enc, wts->w_zr, 9, 256, 448, &m.w))
    return GOSLAM_ELAUNCH;
  p.taps = 9; p.n_in = 4; p.chunks[0] = 2; p.chunks[1] = 2; p.chunks[2] = 2; p.chunks[3] = 1;
  p.N = 256; p.epi = EPI_ZR; p.bias = wts->b_zr; p.glo = ws.glo; p.z_out = ws.z; p.rnet_out = ws.rnet;
  rc = conv_launch(m, p, st);
  if (rc) return rc;
  // ---- pass Q: net' = (1 - z) net + z tanh(convq([r*net | inp | corr | flow]) + glo_q)
  if (!act_map(enc, ws.rnet, B, h, w, 128, &m.in[0]) || !weight_map(enc, wts->w_q, 9, 128, 448, &m.w)) return GOSLAM_ELAUNCH;
  p.N = 128; p.epi = EPI_Q; p.bias = wts->b_q; p.z_in = ws.z; p.net_out = reinterpret_cast<__half*>(net_out);
  return conv_launch(m, p, st);
}

}  // extern "C"
