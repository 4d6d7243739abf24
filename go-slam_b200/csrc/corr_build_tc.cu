// corr_build_tc.cu — all-pairs correlation volume on Blackwell tensor cores (tcgen05 + TMA)
// with the 4-level average-pool pyramid produced in the epilogue.
//
// Reference: CorrBlock.__init__ / CorrBlock.corr (src/modules/corr.py:25-41,67-76) —
// torch.matmul (cuBLAS) writes level 0, then three F.avg_pool2d passes re-read the volume.
//
// Design (one CTA per SM, persistent, warp-specialised, 320 threads):
//   warp 0      TMA producer: A tile = 128 source pixels x 128 channels (2 boxes of 64 ch,
//               128B-swizzled, K-major) once per work item; B tile = an 8x16 patch of TARGET
//               pixels x 128 channels (4-D tensor map (ch, x, y, edge) -> rows ordered
//               y*16+x), 3-stage ring.  Out-of-image rows/cols are zero-filled by TMA.
//   warp 1      MMA issuer: 8 x tcgen05.mma (M128,N128,K16, fp16 in / fp32 accumulate in
//               TMEM) per tile, 2 accumulator stages (256 TMEM columns), tcgen05.commit
//               releases smem stages / publishes accumulators through mbarriers.
//   warps 2-9   epilogue (two groups of 4, one per TMEM stage): tcgen05.ld (32 lanes x 32 columns) -> x 1/16 -> fp16 -> level 0
//               rows (32 B per (source pixel, target row)); because one thread holds the whole
//               8x16 target patch of its source pixel, levels 1..3 (4x8, 2x4, 1x2) are pooled
//               in registers from the ROUNDED finer level, exactly like avg_pool2d on fp16,
//               and written directly — the volume is never re-read.
// The kernel is output-write bound by construction (K = 128 => 128 FLOP per output byte at
// level 0): DESIGN.md gives the roofline.  Feature maps arrive channel-major ([N,128,h,w],
// as DepthVideo stores them); a small prepass re-lays them K-major ([N,hw,128]) so that both
// operands use the canonical K-major SWIZZLE_128B UMMA layout.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>

int gs_corr_build_simt_f16(const __half* f1, const __half* f2, __half* const* levels,
                           int num_levels, int N, int D, int h, int w, cudaStream_t st);

namespace {

constexpr int kEpiGroupsC = 2;
constexpr int kD = 128;                 // channels (K)
constexpr int kBM = 128;                // source pixels per tile
constexpr int kPY = 8, kPX = 16;        // target patch
constexpr int kBN = kPY * kPX;          // 128
constexpr int kKBox = 64;               // channels per TMA box (128 B)
constexpr int kTileBytes = kBM * kD * 2;          // 32 KB (A or B tile)
constexpr int kBoxBytes = kBM * kKBox * 2;        // 16 KB
constexpr int kAStages = 2, kBStages = 2, kTStages = 2;
constexpr int kEpiGroups = 2;               // one epilogue warp-group (4 warps) per TMEM stage
constexpr int kThreadsTC = 64 + kEpiGroups * 128;
constexpr int kStageL0 = kPY * kBM * kPX * 2;              // [8][128][16] halves = 32 KB
constexpr int kStageL1 = (kPY / 2) * kBM * (kPX / 2) * 2;  // [4][128][8]  halves =  8 KB
constexpr int kStageBytes = kStageL0 + kStageL1;           // per epilogue group
// pooled-level staging of one work item (all x-tiles of an 8-row band), so that levels 1-3 leave
// the SM as long contiguous runs of complete 32-byte sectors (partial-sector writes cost an ECC
// read-modify-write in L2 and were measured to cost more than all of level 0):
constexpr int kMaxXB = 5;                                  // x-tiles per band supported by the stage (w <= 80)
constexpr int kP1Row = kMaxXB * 8 * 2;                     // 80 B  : one level-1 row of the band
constexpr int kP1Src = 4 * kP1Row + 16;                    // 336 B : 4 rows + pad (bank-conflict-free)
constexpr int kP2Row = kMaxXB * 4 * 2;                     // 40 B
constexpr int kP2Src = 2 * kP2Row + 8;                     // 88 B
constexpr int kP3Src = kMaxXB * 2 * 2 + 4;                 // 24 B
constexpr int kPoolBytes = kBM * (kP1Src + kP2Src + kP3Src);   // 57,344 B
static_assert(kPoolBytes <= kEpiGroupsC * kStageBytes, "pooled staging aliases the TMA-store staging");
constexpr int kSmemTC = 1024 + (kAStages + kBStages) * kTileBytes + kEpiGroupsC * kStageBytes + 256;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)map),
      "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)map),
      "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          (uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void group_bar(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(=1)<<16 | SBO(=1024B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=f16 (0), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

struct TcParams {
  __half* lvl[4];
  int num_levels, N, h, w, hw;
  int n_mt, n_yb, n_xb;       // m-tiles, y-blocks, x-blocks
  int n_items;                // N * n_mt * n_yb
  // optional edge -> feature-map-slot indirection (video-level K-major feature maps):
  // slot1 = rig*ii[e], slot2 = rig*jj[e] + (ii[e]==jj[e])   (src/factor_graph.py:108-113,290)
  const int64_t* ii; const int64_t* jj; int rig;
  int tma_l0, tma_l1;         // level 0 / 1 leave through TMA tensor stores (else direct STG)
  int experiment;             // profiling only: 1 = no output writes, 2 = no TMEM reads
  int pool_stage;             // levels 1-3 staged per band in smem and written as contiguous runs
};

// ---- packed fp16 rows live in registers as uint32 pairs (lo = even column) ----
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float lo_f(uint32_t u) {
  return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu)));
}
__device__ __forceinline__ float hi_f(uint32_t u) {
  return __half2float(__ushort_as_half((unsigned short)(u >> 16)));
}
// 2x2 mean of ROUNDED halves, fp32 sum in row-major window order, one rounding (avg_pool2d)
__device__ __forceinline__ float pool_pair(uint32_t top, uint32_t bot) {
  float s = lo_f(top);
  s += hi_f(top);
  s += lo_f(bot);
  s += hi_f(bot);
  return s * 0.25f;
}
// store NW packed words (2*NW halves) to dst, honouring alignment and the valid count
template <int NW>
__device__ __forceinline__ void store_row(__half* dst, const uint32_t (&r)[NW], int nvalid) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
  if (nvalid >= 2 * NW) {
    if (NW == 8 && (a & 31) == 0) {     // one full 32-byte sector per lane (STG.256)
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(r[0]),
                   "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                   : "memory");
      return;
    }
    if (NW >= 4 && (a & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NW; i += 4)
        *reinterpret_cast<uint4*>(dst + 2 * i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      return;
    }
    if (NW == 2 && (a & 7) == 0) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(r[0], r[1]);
      return;
    }
    if ((a & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NW; ++i) *reinterpret_cast<uint32_t*>(dst + 2 * i) = r[i];
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    if (2 * i < nvalid) dst[2 * i] = __ushort_as_half((unsigned short)(r[i] & 0xffffu));
    if (2 * i + 1 < nvalid) dst[2 * i + 1] = __ushort_as_half((unsigned short)(r[i] >> 16));
  }
}

__global__ void __launch_bounds__(kThreadsTC, 1)
corr_build_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                     const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapL0,
                     const __grid_constant__ CUtensorMap mapL1, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* base =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smA = base;
  unsigned char* smB = base + kAStages * kTileBytes;
  unsigned char* smStage = base + (kAStages + kBStages) * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smStage + kEpiGroupsC * kStageBytes);
  uint64_t* full_a = bars;                       // [kAStages]
  uint64_t* empty_a = full_a + kAStages;
  uint64_t* full_b = empty_a + kAStages;         // [kBStages]
  uint64_t* empty_b = full_b + kBStages;
  uint64_t* tm_full = empty_b + kBStages;        // [kTStages]
  uint64_t* tm_empty = tm_full + kTStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tm_empty + kTStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAStages; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < kTStages; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTStages * kBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int as = 0, aph = 0, bs = 0, bph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int yb = item % p.n_yb;
        const int mt = (item / p.n_yb) % p.n_mt;
        const int n = item / (p.n_yb * p.n_mt);
        int n1 = n, n2 = n;
        if (p.ii != nullptr) {
          const int fi = (int)p.ii[n], fj = (int)p.jj[n];
          n1 = p.rig * fi;
          n2 = p.rig * fj + (fi == fj ? 1 : 0) * (p.rig > 1 ? 1 : 0);
        }
        mbar_wait(&empty_a[as], aph ^ 1);
        mbar_expect_tx(&full_a[as], kTileBytes);
        tma_load_3d(&mapA, &full_a[as], smA + as * kTileBytes, 0, mt * kBM, n1);
        tma_load_3d(&mapA, &full_a[as], smA + as * kTileBytes + kBoxBytes, kKBox, mt * kBM, n1);
        if (++as == kAStages) { as = 0; aph ^= 1; }
        for (int xb = 0; xb < p.n_xb; ++xb) {
          mbar_wait(&empty_b[bs], bph ^ 1);
          mbar_expect_tx(&full_b[bs], kTileBytes);
          tma_load_4d(&mapB, &full_b[bs], smB + bs * kTileBytes, 0, xb * kPX, yb * kPY, n2);
          tma_load_4d(&mapB, &full_b[bs], smB + bs * kTileBytes + kBoxBytes, kKBox, xb * kPX,
                      yb * kPY, n2);
          if (++bs == kBStages) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int as = 0, aph = 0, bs = 0, bph = 0, ts = 0, tph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        mbar_wait(&full_a[as], aph);
        const uint32_t a_addr = smem_u32(smA + as * kTileBytes);
        for (int xb = 0; xb < p.n_xb; ++xb) {
          mbar_wait(&tm_empty[ts], tph ^ 1);
          mbar_wait(&full_b[bs], bph);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smB + bs * kTileBytes);
          const uint32_t d_tmem = tmem_base + ts * kBN;
#pragma unroll
          for (int kb = 0; kb < kD / kKBox; ++kb) {
#pragma unroll
            for (int k = 0; k < kKBox / 16; ++k) {
              const uint64_t da = make_desc_sw128(a_addr + kb * kBoxBytes + k * 32);
              const uint64_t db = make_desc_sw128(b_addr + kb * kBoxBytes + k * 32);
              umma_f16(d_tmem, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_b[bs]);      // smem B stage reusable when these MMAs retire
          umma_commit(&tm_full[ts]);      // accumulator ready for the epilogue
          if (++bs == kBStages) { bs = 0; bph ^= 1; }
          if (++ts == kTStages) { ts = 0; tph ^= 1; }
        }
        umma_commit(&empty_a[as]);
        if (++as == kAStages) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9: two groups of 4) =====================
    // group g drains TMEM stage g only, so two tiles are in flight in the store path.
    const int group = (warp - 2) >> 2;
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;             // row of the 128-row tile
    int ts = group;
    const bool single = p.experiment == 5;        // profiling: one epilogue group drains both stages
    const bool use_tma = (p.tma_l0 | p.tma_l1) != 0 && p.experiment != 1;
    const bool elected = (warp == 2 + 4 * group) && lane == 0;
    unsigned char* stg0 = smStage + group * kStageBytes;      // [8][128][16] halves
    unsigned char* stg1 = stg0 + kStageL0;                    // [4][128][8]  halves
    unsigned char* pool1 = smStage;                            // [128][kP1Src]
    unsigned char* pool2 = pool1 + kBM * kP1Src;               // [128][kP2Src]
    unsigned char* pool3 = pool2 + kBM * kP2Src;               // [128][kP3Src]
    const int etid = threadIdx.x - 64;                         // 0..255 within the epilogue warps
    int tph = 0, tile = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int yb = item % p.n_yb;
      const int mt = (item / p.n_yb) % p.n_mt;
      const int n = item / (p.n_yb * p.n_mt);
      const int src = mt * kBM + row;
      const bool src_ok = src < p.hw && p.experiment != 1;
      const long long plane_id = (long long)n * p.hw + src;
      const int y0 = yb * kPY;
      for (int xb = 0; xb < p.n_xb; ++xb, ++tile) {
        if (single) {
          if (group != 0) continue;
          ts = tile & (kTStages - 1);
          tph = (tile >> 1) & 1;
        } else if ((tile & (kTStages - 1)) != ts) continue;
        const int x0 = xb * kPX;
        mbar_wait(&tm_full[ts], tph);
        tc_fence_after();
        if (use_tma) {
          // the previous tensor store issued from this group's staging buffer must have been read
          if (elected) bulk_wait_read0();
          group_bar(1 + group);
        }
        const uint32_t taddr = tmem_base + ts * kBN + ((uint32_t)(quad * 32) << 16);
        uint32_t l1[4][4];     // level-1 rows (8 halves each) of this patch
        uint32_t l2[2][2];     // level-2 rows (4 halves each)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          if (p.experiment != 2) tmem_ld32(taddr + c * 32, v);
          else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0x3f800000u + q;
          }
          uint32_t h0[2][8];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 8; ++x)
              h0[r][x] = pack2(__uint_as_float(v[r * 16 + 2 * x]) * 0.0625f,
                               __uint_as_float(v[r * 16 + 2 * x + 1]) * 0.0625f);
          if (src_ok) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int y = y0 + 2 * c + r;
              if (p.tma_l0) continue;
              if (y < p.h) store_row<8>(p.lvl[0] + (plane_id * p.h + y) * p.w + x0, h0[r], p.w - x0);
            }
          }
          if (p.tma_l0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              uint4* dst = reinterpret_cast<uint4*>(stg0 + ((2 * c + r) * kBM + row) * (kPX * 2));
              dst[0] = make_uint4(h0[r][0], h0[r][1], h0[r][2], h0[r][3]);
              dst[1] = make_uint4(h0[r][4], h0[r][5], h0[r][6], h0[r][7]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            l1[c][j] = pack2(pool_pair(h0[0][2 * j], h0[1][2 * j]),
                             pool_pair(h0[0][2 * j + 1], h0[1][2 * j + 1]));
          if (p.pool_stage) {
            *reinterpret_cast<uint4*>(pool1 + row * kP1Src + c * kP1Row + xb * 16) =
                make_uint4(l1[c][0], l1[c][1], l1[c][2], l1[c][3]);
          } else if (p.tma_l1) {
            *reinterpret_cast<uint4*>(stg1 + (c * kBM + row) * (kPX)) =
                make_uint4(l1[c][0], l1[c][1], l1[c][2], l1[c][3]);
          } else if (src_ok && p.num_levels > 1) {
            const int h1 = p.h >> 1, w1 = p.w >> 1;
            const int y = (y0 >> 1) + c, x = x0 >> 1;
            if (y < h1 && x < w1) store_row<4>(p.lvl[1] + (plane_id * h1 + y) * w1 + x, l1[c], w1 - x);
          }
          if (c & 1) {
            const int q = c >> 1;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              l2[q][j] = pack2(pool_pair(l1[c - 1][2 * j], l1[c][2 * j]),
                               pool_pair(l1[c - 1][2 * j + 1], l1[c][2 * j + 1]));
            if (p.pool_stage) {
              *reinterpret_cast<uint2*>(pool2 + row * kP2Src + q * kP2Row + xb * 8) = make_uint2(l2[q][0], l2[q][1]);
            } else if (src_ok && p.num_levels > 2) {
              const int h2 = p.h >> 2, w2 = p.w >> 2;
              const int y = (y0 >> 2) + q, x = x0 >> 2;
              if (y < h2 && x < w2) store_row<2>(p.lvl[2] + (plane_id * h2 + y) * w2 + x, l2[q], w2 - x);
            }
          }
        }
        // all TMEM reads of this stage are complete (tmem_ld32 waits): hand the stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tm_empty[ts]);
        if (use_tma) {
          fence_async_smem();                      // generic-proxy smem writes -> async proxy
          group_bar(1 + group);
          if (elected) {
            if (p.tma_l0) tma_store_4d(&mapL0, stg0, x0, mt * kBM, y0, n);
            if (p.tma_l1) tma_store_4d(&mapL1, stg1, x0 >> 1, mt * kBM, y0 >> 1, n);
            bulk_commit();
          }
        }
        if (p.pool_stage) {
          *reinterpret_cast<uint32_t*>(pool3 + row * kP3Src + xb * 4) =
              pack2(pool_pair(l2[0][0], l2[1][0]), pool_pair(l2[0][1], l2[1][1]));
        } else if (src_ok && p.num_levels > 3) {
          uint32_t l3[1];
          l3[0] = pack2(pool_pair(l2[0][0], l2[1][0]), pool_pair(l2[0][1], l2[1][1]));
          const int h3 = p.h >> 3, w3 = p.w >> 3;
          const int y = y0 >> 3, x = x0 >> 3;
          if (y < h3 && x < w3) store_row<1>(p.lvl[3] + (plane_id * h3 + y) * w3 + x, l3, w3 - x);
        }
        tph ^= 1;
      }
      if (p.pool_stage) {
        // ---- band write-out: both epilogue groups have staged all x-tiles of this 8-row band ----
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const int s_loc = etid >> 1, part = etid & 1;          // two threads per source pixel
        const int s_glb = mt * kBM + s_loc;
        if (s_glb < p.hw) {
          const long long pl = (long long)n * p.hw + s_glb;
          // level 1: 4 full rows = 4*w1 halves contiguous (w1 == n_xb*8), 32-byte aligned
          {
            const int w1b = p.n_xb * 16;                        // bytes per level-1 row
            unsigned char* g = reinterpret_cast<unsigned char*>(p.lvl[1]) + (pl * (p.h >> 1) + (y0 >> 1)) * w1b;
            const unsigned char* sp = pool1 + s_loc * kP1Src;
            const int total = 4 * w1b;                          // bytes, multiple of 64
            for (int off = part * 32; off < total; off += 64) {
              // staging rows are kP1Row apart, global rows w1b apart
              uint32_t rr[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int o = off + 4 * k;
                rr[k] = *reinterpret_cast<const uint32_t*>(sp + (o / w1b) * kP1Row + (o % w1b));
              }
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(g + off), "r"(rr[0]),
                           "r"(rr[1]), "r"(rr[2]), "r"(rr[3]), "r"(rr[4]), "r"(rr[5]), "r"(rr[6]), "r"(rr[7])
                           : "memory");
            }
          }
          if (part == 0 && p.num_levels > 2) {                  // level 2: 2 rows of w2 halves
            const int w2b = p.n_xb * 8;
            unsigned char* g = reinterpret_cast<unsigned char*>(p.lvl[2]) + (pl * (p.h >> 2) + (y0 >> 2)) * w2b;
            const unsigned char* sp = pool2 + s_loc * kP2Src;
            for (int off = 0; off < 2 * w2b; off += 16) {       // 2*w2b = n_xb*16: whole 16-byte chunks
              uint32_t rr[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int o = off + 4 * k;
                rr[k] = *reinterpret_cast<const uint32_t*>(sp + (o / w2b) * kP2Row + (o % w2b));
              }
              *reinterpret_cast<uint4*>(g + off) = make_uint4(rr[0], rr[1], rr[2], rr[3]);
            }
          }
          if (part == 1 && p.num_levels > 3) {                  // level 3: 1 row of w3 halves
            const int w3b = p.n_xb * 4;
            unsigned char* g = reinterpret_cast<unsigned char*>(p.lvl[3]) + (pl * (p.h >> 3) + (y0 >> 3)) * w3b;
            const unsigned char* sp = pool3 + s_loc * kP3Src;
            for (int off = 0; off < w3b; off += 4)
              *reinterpret_cast<uint32_t*>(g + off) = *reinterpret_cast<const uint32_t*>(sp + off);
          }
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
    }
    if (use_tma && elected) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTStages * kBN);
}

// [N, D, hw] (channel-major) -> [N, hw, D] (K-major), D = 128
__global__ void __launch_bounds__(256)
to_kmajor_kernel(const __half* __restrict__ in, __half* __restrict__ out, int hw) {
  __shared__ __half tile[kD][64 + 2];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 64;
  const __half* src = in + (size_t)n * kD * hw;
  for (int idx = threadIdx.x; idx < kD * 64; idx += 256) {
    const int k = idx / 64, pp = idx % 64;
    tile[k][pp] = (p0 + pp < hw) ? src[(size_t)k * hw + p0 + pp] : __half(0.f);
  }
  __syncthreads();
  __half* dst = out + ((size_t)n * hw + p0) * kD;
  for (int idx = threadIdx.x; idx < 64 * (kD / 2); idx += 256) {
    const int pp = idx / (kD / 2), k2 = idx % (kD / 2);
    if (p0 + pp < hw) {
      __half2 v = __halves2half2(tile[2 * k2][pp], tile[2 * k2 + 1][pp]);
      reinterpret_cast<__half2*>(dst + (size_t)pp * kD)[k2] = v;
    }
  }
}

// GOSLAM_TC_DIRECT_STORE=1 forces the direct-STG epilogue (A/B switch for profiling)
static const bool g_tma_store = [] { const char* e = getenv("GOSLAM_TC_DIRECT_STORE"); return !(e && e[0] == '1'); }();

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// launch the tensor-core kernel on K-major operands: f1t/f2t = [F1|F2, hw, 128]
int launch_tc(const __half* f1t, int F1, const __half* f2t, int F2, const int64_t* ii,
              const int64_t* jj, int rig, __half* const* levels, int num_levels, int N, int h, int w,
              cudaStream_t st) {
  const int hw = h * w;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return GOSLAM_ELAUNCH;
  CUtensorMap mapA, mapB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)hw, (cuuint64_t)F1};
    cuuint64_t strides[2] = {(cuuint64_t)kD * 2, (cuuint64_t)hw * kD * 2};
    cuuint32_t box[3] = {(cuuint32_t)kKBox, (cuuint32_t)kBM, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(f1t), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GOSLAM_ELAUNCH;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)F2};
    cuuint64_t strides[3] = {(cuuint64_t)kD * 2, (cuuint64_t)w * kD * 2, (cuuint64_t)hw * kD * 2};
    cuuint32_t box[4] = {(cuuint32_t)kKBox, (cuuint32_t)kPX, (cuuint32_t)kPY, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(f2t), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GOSLAM_ELAUNCH;
  }
  TcParams p{};
  for (int i = 0; i < 4; ++i) p.lvl[i] = i < num_levels ? levels[i] : nullptr;
  const char* exp_env = getenv("GOSLAM_TC_EXPERIMENT");
  const int exp_id = exp_env ? atoi(exp_env) : 0;
  // output tensor maps (x, source pixel, y, edge): the epilogue stages [y][src][x] tiles in
  // shared memory and the TMA engine streams the 32-byte rows out, clipping ragged edges
  CUtensorMap mapL0, mapL1;
  memset(&mapL0, 0, sizeof(mapL0)); memset(&mapL1, 0, sizeof(mapL1));
  auto make_out_map = [&](CUtensorMap* m, __half* ptr, int hl, int wl, int bx, int by) -> bool {
    if (ptr == nullptr || (wl * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
    cuuint64_t dims[4] = {(cuuint64_t)wl, (cuuint64_t)hw, (cuuint64_t)hl, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)hl * wl * 2, (cuuint64_t)wl * 2, (cuuint64_t)hw * hl * wl * 2};
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)kBM, (cuuint32_t)by, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ptr, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  p.tma_l0 = (g_tma_store && make_out_map(&mapL0, levels[0], h, w, kPX, kPY)) ? 1 : 0;
  p.tma_l1 = (g_tma_store && num_levels > 1 && make_out_map(&mapL1, levels[1], h >> 1, w >> 1, kPX / 2, kPY / 2)) ? 1 : 0;
  p.num_levels = num_levels; p.N = N; p.h = h; p.w = w; p.hw = hw;
  p.n_mt = gs_cdiv(hw, kBM); p.n_yb = gs_cdiv(h, kPY); p.n_xb = gs_cdiv(w, kPX);
  p.n_items = N * p.n_mt * p.n_yb;
  p.ii = ii; p.jj = jj; p.rig = rig;
  p.experiment = exp_id;
  // band staging of the pooled levels needs whole tiles and sector-aligned level-1 bands
  static const bool no_pool_stage = [] { const char* e = getenv("GOSLAM_TC_NO_POOL_STAGE"); return e && e[0] == '1'; }();
  p.pool_stage = (!no_pool_stage && num_levels == 4 && w % 16 == 0 && h % 8 == 0 && p.n_xb <= kMaxXB &&
                  exp_id != 1 && exp_id != 3) ? 1 : 0;
  if (p.pool_stage) { p.tma_l0 = 0; p.tma_l1 = 0; }       // level 0 leaves as full-sector STG.256
  if (exp_id == 1) { p.tma_l0 = 0; p.tma_l1 = 0; }
  if (exp_id == 3) { p.num_levels = 1; p.tma_l1 = 0; }   // level 0 only
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(corr_build_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSmemTC) != cudaSuccess)
      return GOSLAM_ELAUNCH;
    attr = true;
  }
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = p.n_items < sms ? p.n_items : sms;
  corr_build_tc_kernel<<<grid, kThreadsTC, kSmemTC, st>>>(mapA, mapB, mapL0, mapL1, p);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int build_tc(const __half* f1, const __half* f2, __half* const* levels, int num_levels, int N,
             int h, int w, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int hw = h * w;
  const size_t per = (size_t)N * hw * kD * sizeof(__half);
  if (workspace == nullptr || workspace_bytes < 2 * gs_align(per)) return GOSLAM_EWORKSPACE;
  __half* f1t = reinterpret_cast<__half*>(workspace);
  __half* f2t = reinterpret_cast<__half*>(reinterpret_cast<char*>(workspace) + gs_align(per));
  dim3 tg(gs_cdiv(hw, 64), N);
  to_kmajor_kernel<<<tg, 256, 0, st>>>(f1, f1t, hw);
  to_kmajor_kernel<<<tg, 256, 0, st>>>(f2, f2t, hw);
  GS_CHECK_LAUNCH();
  return launch_tc(f1t, N, f2t, N, nullptr, nullptr, 1, levels, num_levels, N, h, w, st);
}

}  // namespace

extern "C" {

size_t goslam_corr_build_workspace_bytes(int N, int D, int h, int w) {
  if (N <= 0 || D != kD) return 256;
  return 2 * gs_align((size_t)N * h * w * kD * sizeof(__half)) + 256;
}

int goslam_fmaps_to_kmajor(const void* fmaps, void* out, int F, int D, int h, int w, void* stream) {
  if (F < 0 || D != kD || h <= 0 || w <= 0) return GOSLAM_EINVAL;
  if (F == 0) return GOSLAM_OK;
  dim3 tg(gs_cdiv(h * w, 64), F);
  to_kmajor_kernel<<<tg, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(fmaps),
                                                         reinterpret_cast<__half*>(out), h * w);
  GS_CHECK_LAUNCH();
  return GOSLAM_OK;
}

int goslam_corr_build_indexed(const void* fmaps_kmajor, int F, int rig, const int64_t* ii,
                              const int64_t* jj, void* const* levels, int num_levels, int N, int D,
                              int h, int w, void* stream) {
  if (N < 0 || F <= 0 || rig < 1 || D != kD || h <= 0 || w <= 0 || num_levels < 1 || num_levels > 4)
    return GOSLAM_EINVAL;
  if ((h >> (num_levels - 1)) <= 0 || (w >> (num_levels - 1)) <= 0) return GOSLAM_EINVAL;
  if (N == 0) return GOSLAM_OK;
  const __half* f = reinterpret_cast<const __half*>(fmaps_kmajor);
  return launch_tc(f, F, f, F, ii, jj, rig, reinterpret_cast<__half* const*>(levels), num_levels, N, h,
                   w, (cudaStream_t)stream);
}

int goslam_corr_build(const void* fmap1, const void* fmap2, void* const* levels, int num_levels,
                      int N, int D, int h, int w, int impl, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (N < 0 || D <= 0 || h <= 0 || w <= 0 || num_levels < 1 || num_levels > 4) return GOSLAM_EINVAL;
  if ((h >> (num_levels - 1)) <= 0 || (w >> (num_levels - 1)) <= 0) return GOSLAM_EINVAL;
  if (N == 0) return GOSLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const __half* f1 = reinterpret_cast<const __half*>(fmap1);
  const __half* f2 = reinterpret_cast<const __half*>(fmap2);
  __half* const* lv = reinterpret_cast<__half* const*>(levels);
  if (impl == 0) impl = (D == kD) ? 1 : 2;
  if (impl == 1) {
    if (D != kD) return GOSLAM_EINVAL;
    return build_tc(f1, f2, lv, num_levels, N, h, w, workspace, workspace_bytes, st);
  }
  if (impl == 2) return gs_corr_build_simt_f16(f1, f2, lv, num_levels, N, D, h, w, st);
  return GOSLAM_EINVAL;
}

}  // extern "C"
