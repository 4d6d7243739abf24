// neus.cu — fused hash-grid neural-surface ray marcher (InstantNeuS.forward).
//
// Replaces, in ONE persistent kernel: tiny-cuda-nn HashGrid forward + its input-gradient
// backward (SDF normal), nn.Linear(35,32), NeuS alpha, sin-embedding, tiny-cuda-nn
// FullyFusedMLP 67(->80)->64->64->3(->16), sigmoid, and front-to-back compositing
// (src/InstantNeuS.py:295-370 with :12-32, :35-94, :97-160, :162-205, :258-293), i.e. what
// the reference runs as ~60 eager kernels + two tcnn launches + an autograd pass.
//
// tiny-cuda-nn is an un-vendored, un-pinned dependency of the reference (README.md:95); its
// arithmetic is restated from its published algorithm (see oracle/neus_oracle.py header):
//   * level l: scale = exp2(l*log2(b))*16 - 1, res = ceil(scale)+1, pos = x*scale + 0.5,
//     8-corner trilinear; index = x + y*res + z*res^2 while the stride fits the table,
//     else (x*1) ^ (y*2654435761) ^ (z*805459861); mod table size; table entries half2;
//     features accumulated in half:  r += (half)(w * (float)v);
//   * d(enc)/dx in fp32 from half table values, dL/dy rounded to half (tcnn backward);
//   * MLP: half inputs padded with 1.0 to 80, half activations, ReLU, no bias; we accumulate
//     in fp32 on mma.sync (tcnn: half accumulators) — documented tolerance in the tests.
//
// Work decomposition: a block owns kRaysPerGroup rays at a time (persistent loop).  A warp
// takes 32 consecutive samples: phase 1 is one thread per sample (gather + SDF head +
// alpha + embedding), phase 2 is the warp-wide MLP on m16n8k16 tensor-core tiles with the
// weights resident in shared memory, phase 3 composites each ray with a warp scan.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int kLevels = 16;
constexpr int kThreadsN = 256;
constexpr int kWarpsN = kThreadsN / 32;
constexpr int kRaysPerGroup = 8;
constexpr int kIn = 80, kInPad = 88;     // MLP input width / padded smem row (halves)
constexpr int kHid = 64, kHidPad = 72;
constexpr int kOutW = 16;

struct GridMeta {
  float scale[kLevels];
  int res[kLevels];
  unsigned offset[kLevels];   // in entries (half2)
  unsigned size[kLevels];     // entries in level
};

GridMeta make_grid_meta(int64_t* total_entries) {
  GridMeta g{};
  const float log2_b = log2f(1.447269237440378f);
  unsigned off = 0;
  for (int l = 0; l < kLevels; ++l) {
    const float scale = exp2f((float)l * log2_b) * 16.0f - 1.0f;
    const unsigned res = (unsigned)ceilf(scale) + 1u;
    unsigned long long dense = (unsigned long long)res * res * res;
    const unsigned long long maxp = 0xFFFFFFFFull / 2;
    unsigned long long p = dense > maxp ? maxp : dense;
    p = (p + 7) / 8 * 8;
    if (p > (1ull << 19)) p = 1ull << 19;
    g.scale[l] = scale; g.res[l] = (int)res; g.offset[l] = off; g.size[l] = (unsigned)p;
    off += (unsigned)p;
  }
  if (total_entries) *total_entries = off;
  return g;
}

__constant__ GridMeta c_grid;

struct NeusArgs {
  goslam_neus_params p;
  goslam_neus_out o;
  const float* rays_o; const float* rays_d; const float* z_vals; const float* dists;
  int R, S;
  float* blk_gerr;        // [grid] partial sums of the eikonal term
  unsigned* blk_count;    // [grid] in-bound sample counts
  int* flag;              // [1] written by the finalize kernel: 1 = nothing in bound
  int mode;               // 0 main pass, 1 fix-up pass (mask[:100] = True)
};

__device__ __forceinline__ unsigned grid_index(int l, unsigned x, unsigned y, unsigned z) {
  const unsigned res = (unsigned)c_grid.res[l];
  const unsigned size = c_grid.size[l];
  unsigned stride = 1, index = 0;
  // mirrors tcnn grid_index: dense strides while they fit the table, else the hash
  index += x * stride; stride *= res;
  if (stride <= size) { index += y * stride; stride *= res; }
  if (stride <= size) { index += z * stride; stride *= res; }
  if (size < stride) index = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
  return index % size;
}

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2(unsigned (&r)[2], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4],
                                         const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One dense layer for a 32-row warp tile: out[32][N] = in[32][K] * W[N][K]^T (fp32 accum).
// NT = N/8 n-tiles, KT = K/16 k-steps.  acc[mt][nt][4].
template <int NT, int KT, int IN_LD, int W_LD>
__device__ __forceinline__ void warp_layer(const __half* in, const __half* W,
                                           float (&acc)[2][NT][4], int lane) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
  const int lr = lane & 15, lc = (lane >> 4) * 8;       // ldmatrix.x4 A addressing
  const int br = lane & 7, bc = ((lane >> 3) & 1) * 8;  // ldmatrix.x2 B addressing
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    unsigned a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
      ldmatrix_x4(a[mt], in + (mt * 16 + lr) * IN_LD + kt * 16 + lc);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      unsigned b[2];
      ldmatrix_x2(b, W + (nt * 8 + br) * W_LD + kt * 16 + bc);
      mma16816(acc[0][nt], a[0], b);
      mma16816(acc[1][nt], a[1], b);
    }
  }
}

template <int NT, int OUT_LD>
__device__ __forceinline__ void store_relu_half(const float (&acc)[2][NT][4], __half* out,
                                                int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = nt * 8 + 2 * t;
      const __half2 lo = __floats2half2_rn(fmaxf(acc[mt][nt][0], 0.f), fmaxf(acc[mt][nt][1], 0.f));
      const __half2 hi = __floats2half2_rn(fmaxf(acc[mt][nt][2], 0.f), fmaxf(acc[mt][nt][3], 0.f));
      *reinterpret_cast<__half2*>(out + (mt * 16 + g) * OUT_LD + col) = lo;
      *reinterpret_cast<__half2*>(out + (mt * 16 + g + 8) * OUT_LD + col) = hi;
    }
}

struct Smem {
  alignas(16) __half W1[kHid * kInPad];
  alignas(16) __half W2[kHid * kHidPad];
  alignas(16) __half W3[kOutW * kHidPad];
  alignas(16) float sdfWT[35 * 32];          // transposed Linear weight: [k][out]
  float sdfB[32];
  float colB[3 * 33];
  alignas(16) __half actA[kWarpsN][32 * kInPad];
  alignas(16) __half actB[kWarpsN][32 * kHidPad];
  // per-sample results of the current ray group
  float alpha[kRaysPerGroup * 128];
  float zmid[kRaysPerGroup * 128];
  float grad[3][kRaysPerGroup * 128];
  float rgb[3][kRaysPerGroup * 128];
  float maskf[kRaysPerGroup * 128];
  float red_g[kWarpsN];
  unsigned red_c[kWarpsN];
};

__global__ void __launch_bounds__(kThreadsN, 1)
neus_forward_kernel(const NeusArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S;

  int num_groups = gs_cdiv_dev(a.R, kRaysPerGroup);
  if (a.mode == 1) {
    if (*a.flag == 0) return;                    // something was in bound: no fix-up needed
    num_groups = gs_cdiv_dev(gs_cdiv_dev(100, S), kRaysPerGroup);
    const int rg = gs_cdiv_dev(a.R, kRaysPerGroup);
    if (num_groups > rg) num_groups = rg;
  }

  // ---- stage the network weights once per block (persistent) ----
  {
    const __half* w = reinterpret_cast<const __half*>(a.p.mlp_w);
    for (int i = tid; i < kHid * kIn; i += kThreadsN) sm.W1[(i / kIn) * kInPad + i % kIn] = w[i];
    for (int i = tid; i < kHid * kHid; i += kThreadsN)
      sm.W2[(i / kHid) * kHidPad + i % kHid] = w[kHid * kIn + i];
    for (int i = tid; i < kOutW * kHid; i += kThreadsN)
      sm.W3[(i / kHid) * kHidPad + i % kHid] = w[kHid * kIn + kHid * kHid + i];
    for (int i = tid; i < 32 * 35; i += kThreadsN) sm.sdfWT[(i % 35) * 32 + i / 35] = a.p.sdf_w[i];
    for (int i = tid; i < 32; i += kThreadsN) sm.sdfB[i] = a.p.sdf_b[i];
    for (int i = tid; i < 99; i += kThreadsN) sm.colB[i] = a.p.color_B[i];
  }
  __syncthreads();

  const __half2* table = reinterpret_cast<const __half2*>(a.p.grid);
  float gerr_local = 0.f;
  unsigned count_local = 0;

  for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
    const int ray0 = group * kRaysPerGroup;
    const int nrays = min(kRaysPerGroup, a.R - ray0);
    const int nsamp = nrays * S;
    const int ntiles = gs_cdiv_dev(nsamp, 32);

    for (int tile = warp; tile < ntiles; tile += kWarpsN) {
      const int ls = tile * 32 + lane;           // local sample index in the group
      const bool valid = ls < nsamp;
      const int lr = valid ? ls / S : 0;
      const int sidx = valid ? ls % S : 0;
      const int ray = ray0 + lr;
      const size_t gidx = (size_t)ray * S + sidx;
      __half* rowA = sm.actA[warp] + lane * kInPad;

      float zm = 0.f, dist = 0.f, alpha = 0.f, sdf = 100.f;
      float g3[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f}, pt[3] = {0.f, 0.f, 0.f};
      bool inb = false;
      if (valid) {
        dist = a.dists[gidx];
        zm = __fadd_rn(a.z_vals[gidx], dist / 2.0f);
        // op-by-op fp32 like the reference's separate torch kernels (no FMA contraction): the
        // hash grid turns a 1-ulp difference in the position into a different fine-level cell
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          dir[c] = a.rays_d[(size_t)ray * 3 + c];
          pt[c] = __fadd_rn(a.rays_o[(size_t)ray * 3 + c], __fmul_rn(dir[c], zm));
        }
        inb = pt[0] < a.p.rt_bound[1] && pt[0] > a.p.rt_bound[0] &&
              pt[1] < a.p.rt_bound[3] && pt[1] > a.p.rt_bound[2] &&
              pt[2] < a.p.rt_bound[5] && pt[2] > a.p.rt_bound[4];
        if (a.mode == 1 && gidx < 100) inb = true;
      }

      float feat[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) feat[i] = 0.f;

      if (inb) {
        // normalised coordinate, clamp, unit cube
        float xn[3], x01[3], dscale[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float b0 = a.p.bound[2 * c], b1 = a.p.bound[2 * c + 1];
          const float raw = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(pt[c], b0), __fsub_rn(b1, b0)), 2.0f), 1.0f);
          xn[c] = fminf(fmaxf(raw, -1.0f), 1.0f);
          dscale[c] = (raw >= -1.0f && raw <= 1.0f) ? 2.0f / (b1 - b0) : 0.0f;
          x01[c] = __fdiv_rn(__fadd_rn(xn[c], 1.0f), 2.0f);
        }
        // SDF head accumulators start from bias + xyz part
        float out[32];
#pragma unroll
        for (int o = 0; o < 32; ++o)
          out[o] = sm.sdfB[o] + sm.sdfWT[0 * 32 + o] * xn[0] + sm.sdfWT[1 * 32 + o] * xn[1] +
                   sm.sdfWT[2 * 32 + o] * xn[2];
        float genc[3] = {0.f, 0.f, 0.f};
#pragma unroll 2
        for (int l = 0; l < kLevels; ++l) {
          const float scale = c_grid.scale[l];
          float fr[3]; unsigned pg[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float pos = fmaf(scale, x01[c], 0.5f);
            const float fl = floorf(pos);
            pg[c] = (unsigned)(int)fl;
            fr[c] = pos - fl;
          }
          const __half2* lvl = table + c_grid.offset[l];
          __half2 v[8];
#pragma unroll
          for (int idx = 0; idx < 8; ++idx) {
            const unsigned cx = pg[0] + (idx & 1), cy = pg[1] + ((idx >> 1) & 1),
                           cz = pg[2] + ((idx >> 2) & 1);
            v[idx] = __ldg(lvl + grid_index(l, cx, cy, cz));
          }
          __half r0 = __float2half_rn(0.f), r1 = r0;
#pragma unroll
          for (int idx = 0; idx < 8; ++idx) {
            float w = 1.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) w *= ((idx >> c) & 1) ? fr[c] : 1.f - fr[c];
            r0 = __hadd_rn(r0, __float2half_rn(w * __low2float(v[idx])));
            r1 = __hadd_rn(r1, __float2half_rn(w * __high2float(v[idx])));
          }
          const float e0 = __half2float(r0), e1 = __half2float(r1);
          // dL/dy for the normal: sdf row of the Linear weight, rounded to half (tcnn bwd)
          const float gy0 = __half2float(__float2half_rn(sm.sdfWT[(3 + 2 * l) * 32]));
          const float gy1 = __half2float(__float2half_rn(sm.sdfWT[(4 + 2 * l) * 32]));
#pragma unroll
          for (int gd = 0; gd < 3; ++gd) {
            const int d1 = (gd + 1) % 3, d2 = (gd + 2) % 3;
            float gsum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int b1 = q & 1, b2 = (q >> 1) & 1;
              const float w = scale * (b1 ? fr[d1] : 1.f - fr[d1]) * (b2 ? fr[d2] : 1.f - fr[d2]);
              const int il = (b1 << d1) | (b2 << d2);
              const int ir = il | (1 << gd);
              gsum += w * ((__low2float(v[ir]) - __low2float(v[il])) * gy0 +
                           (__high2float(v[ir]) - __high2float(v[il])) * gy1);
            }
            genc[gd] += gsum;
          }
#pragma unroll
          for (int o = 0; o < 32; ++o)
            out[o] += sm.sdfWT[(3 + 2 * l) * 32 + o] * e0 + sm.sdfWT[(4 + 2 * l) * 32 + o] * e1;
        }
        sdf = out[0];
#pragma unroll
        for (int i = 0; i < 31; ++i) feat[i] = out[1 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          g3[c] = (sm.sdfWT[c * 32] + 0.5f * genc[c]) * dscale[c];
      }

      if (valid) {
        // NeuS alpha (get_alpha, src/InstantNeuS.py:276-293)
        const float true_cos = dir[0] * g3[0] + dir[1] * g3[1] + dir[2] * g3[2];
        const float car = a.p.cos_anneal_ratio;
        const float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - car) +
                                 fmaxf(-true_cos, 0.f) * car);
        const float half_step = iter_cos * dist / 2.0f;
        const float prev_cdf = 1.0f / (1.0f + expf(-(sdf - half_step) * a.p.inv_s));
        const float next_cdf = 1.0f / (1.0f + expf(-(sdf + half_step) * a.p.inv_s));
        alpha = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
        alpha = fminf(fmaxf(alpha, 0.f), 1.f);
        alpha = inb ? alpha : 0.f;
        a.o.sdf[gidx] = sdf;
        a.o.z_mid[gidx] = zm;
        if (inb) {
          const float gn = sqrtf(g3[0] * g3[0] + g3[1] * g3[1] + g3[2] * g3[2]) - 1.0f;
          gerr_local += gn * gn;
          ++count_local;
        }
        sm.alpha[ls] = alpha;
        sm.zmid[ls] = zm;
        sm.grad[0][ls] = g3[0]; sm.grad[1][ls] = g3[1]; sm.grad[2][ls] = g3[2];
        sm.maskf[ls] = inb ? 1.f : 0.f;
      }

      // ---- MLP input row: [sin(p B)(33) | normal(3) | feat(31) | 1-padding(13)] ----
      {
#pragma unroll
        for (int j = 0; j < 33; ++j) {
          const float arg = pt[0] * sm.colB[j] + pt[1] * sm.colB[33 + j] + pt[2] * sm.colB[66 + j];
          rowA[j] = __float2half_rn(sinf(arg));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) rowA[33 + c] = __float2half_rn(g3[c]);
#pragma unroll
        for (int i = 0; i < 31; ++i) rowA[36 + i] = __float2half_rn(feat[i]);
#pragma unroll
        for (int i = 67; i < kIn; ++i) rowA[i] = __float2half_rn(1.0f);
      }
      __syncwarp();

      // ---- warp-wide MLP on tensor cores ----
      float rgbv[3] = {0.f, 0.f, 0.f};
      {
        float acc[2][8][4];
        warp_layer<8, kIn / 16, kInPad, kInPad>(sm.actA[warp], sm.W1, acc, lane);
        store_relu_half<8, kHidPad>(acc, sm.actB[warp], lane);
        __syncwarp();
        warp_layer<8, kHid / 16, kHidPad, kHidPad>(sm.actB[warp], sm.W2, acc, lane);
        __syncwarp();
        // hidden 2 goes back into actA (row stride kInPad; only the first 64 columns used)
        store_relu_half<8, kInPad>(acc, sm.actA[warp], lane);
        __syncwarp();
        float acc3[2][2][4];
        warp_layer<2, kHid / 16, kInPad, kHidPad>(sm.actA[warp], sm.W3, acc3, lane);
        __syncwarp();
        // scatter the 32x16 result so each lane can pick up its own sample's rgb
        float* scratch = reinterpret_cast<float*>(sm.actB[warp]);   // 32 x 4 floats
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (t < 2) {
            scratch[(mt * 16 + g) * 4 + 2 * t] = acc3[mt][0][0];
            scratch[(mt * 16 + g) * 4 + 2 * t + 1] = acc3[mt][0][1];
            scratch[(mt * 16 + g + 8) * 4 + 2 * t] = acc3[mt][0][2];
            scratch[(mt * 16 + g + 8) * 4 + 2 * t + 1] = acc3[mt][0][3];
          }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float x = __half2float(__float2half_rn(scratch[lane * 4 + c]));  // tcnn output is half
          const float sg = 1.0f / (1.0f + expf(-x));
          rgbv[c] = __half2float(__float2half_rn(sg));                           // torch.sigmoid(half)
        }
        __syncwarp();
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.rgb[c][ls] = inb ? rgbv[c] : 0.f;
      }
    }
    __syncthreads();

    // ---- compositing: one warp per ray, exclusive product scan of (1 - alpha + 1e-7) ----
    for (int lr = warp; lr < nrays; lr += kWarpsN) {
      const int ray = ray0 + lr;
      float carry = 1.0f;
      float wsum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
      // pass 1: weights; keep them in the alpha slot
      for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        const int ls = lr * S + s;
        const float al = (s < S) ? sm.alpha[ls] : 0.f;
        float fct = (s < S) ? (1.0f - al + 1e-7f) : 1.0f;
        // inclusive product scan
        float inc = fct;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float nb = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) inc *= nb;
        }
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.0f;
        const float T = carry * exc;
        const float wgt = al * T;
        carry *= __shfl_sync(0xffffffffu, inc, 31);
        if (s < S) {
          sm.alpha[ls] = wgt;
          const float z = sm.zmid[ls];
          const float m = sm.maskf[ls];
          wsum += wgt; dep += z * wgt;
          cr += sm.rgb[0][ls] * wgt; cg += sm.rgb[1][ls] * wgt; cb += sm.rgb[2][ls] * wgt;
          nx += sm.grad[0][ls] * wgt * m; ny += sm.grad[1][ls] * wgt * m; nz += sm.grad[2][ls] * wgt * m;
        }
      }
      wsum = gs_warp_sum(wsum); dep = gs_warp_sum(dep);
      cr = gs_warp_sum(cr); cg = gs_warp_sum(cg); cb = gs_warp_sum(cb);
      nx = gs_warp_sum(nx); ny = gs_warp_sum(ny); nz = gs_warp_sum(nz);
      float var = 0.f;
      for (int s = lane; s < S; s += 32) {
        const int ls = lr * S + s;
        const float dz = sm.zmid[ls] - dep;
        var += dz * dz * sm.alpha[ls];
      }
      var = gs_warp_sum(var);
      if (lane == 0) {
        a.o.color[(size_t)ray * 3 + 0] = cr; a.o.color[(size_t)ray * 3 + 1] = cg;
        a.o.color[(size_t)ray * 3 + 2] = cb;
        a.o.depth[ray] = dep; a.o.depth_variance[ray] = var; a.o.weight_sum[ray] = wsum;
        a.o.normal[(size_t)ray * 3 + 0] = nx; a.o.normal[(size_t)ray * 3 + 1] = ny;
        a.o.normal[(size_t)ray * 3 + 2] = nz;
      }
    }
    __syncthreads();
  }

  // ---- per-block partials of the eikonal term / in-bound count (deterministic order) ----
  gerr_local = gs_warp_sum(gerr_local);
  unsigned cnt = count_local;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) { sm.red_g[warp] = gerr_local; sm.red_c[warp] = cnt; }
  __syncthreads();
  if (tid == 0) {
    float g = 0.f; unsigned c = 0;
    for (int w = 0; w < kWarpsN; ++w) { g += sm.red_g[w]; c += sm.red_c[w]; }
    a.blk_gerr[blockIdx.x] = g;
    a.blk_count[blockIdx.x] = c;
  }
}

__global__ void neus_finalize_kernel(const float* blk_gerr, const unsigned* blk_count, int nblk,
                                     long long total_samples, float* gradient_error, int* flag,
                                     int mode) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (mode == 1 && *flag == 0) return;
  double g = 0.0; unsigned long long c = 0;
  for (int i = 0; i < nblk; ++i) { g += (double)blk_gerr[i]; c += blk_count[i]; }
  *gradient_error = (float)(g / (double)total_samples);
  if (mode == 0) *flag = (c == 0) ? 1 : 0;
}

}  // namespace

extern "C" {

int64_t goslam_hashgrid_layout(int64_t* offsets, int* resolutions, float* scales) {
  int64_t total = 0;
  GridMeta g = make_grid_meta(&total);
  for (int l = 0; l < kLevels; ++l) {
    if (offsets) offsets[l] = (int64_t)g.offset[l] * 2;
    if (resolutions) resolutions[l] = g.res[l];
    if (scales) scales[l] = g.scale[l];
  }
  if (offsets) offsets[kLevels] = total * 2;
  return total * 2;
}

size_t goslam_neus_workspace_bytes(int R, int S) {
  (void)R; (void)S;
  return gs_align(148 * 4 * sizeof(float)) + gs_align(148 * 4 * sizeof(unsigned)) + 256;
}

int goslam_neus_forward(const goslam_neus_params* params, const float* rays_o, const float* rays_d,
                        const float* z_vals, const float* dists, int R, int S,
                        const goslam_neus_out* out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (!params || !out || R < 0 || S <= 0 || S > 128) return GOSLAM_EINVAL;
  if (R == 0) return GOSLAM_OK;
  if (workspace == nullptr || workspace_bytes < goslam_neus_workspace_bytes(R, S))
    return GOSLAM_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  static bool init = false;
  if (!init) {
    GridMeta g = make_grid_meta(nullptr);
    if (cudaMemcpyToSymbol(c_grid, &g, sizeof(g)) != cudaSuccess) return GOSLAM_ELAUNCH;
    if (cudaFuncSetAttribute(neus_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(Smem)) != cudaSuccess) return GOSLAM_ELAUNCH;
    init = true;
  }
  GsArena ar(workspace, workspace_bytes);
  NeusArgs a{};
  a.p = *params; a.o = *out;
  a.rays_o = rays_o; a.rays_d = rays_d; a.z_vals = z_vals; a.dists = dists;
  a.R = R; a.S = S;
  const int grid = 148;
  a.blk_gerr = ar.take<float>(148 * 4);
  a.blk_count = ar.take<unsigned>(148 * 4);
  a.flag = reinterpret_cast<int*>(ar.take<int>(1));
  const int groups = gs_cdiv(R, kRaysPerGroup);
  const int nblk = groups < grid ? groups : grid;
  for (int mode = 0; mode < 2; ++mode) {
    a.mode = mode;
    neus_forward_kernel<<<nblk, kThreadsN, sizeof(Smem), st>>>(a);
    GS_CHECK_LAUNCH();
    neus_finalize_kernel<<<1, 32, 0, st>>>(a.blk_gerr, a.blk_count, nblk, (long long)R * S,
                                           out->gradient_error, a.flag, mode);
    GS_CHECK_LAUNCH();
  }
  return GOSLAM_OK;
}

}  // extern "C"
