// write-pattern microbenchmark for the correlation-volume epilogue (tools only, not shipped)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// volume [N][hw][h][w] halves; h=40,w=80,hw=3200. tile: 128 src x (8 rows x 16 cols)
constexpr int H = 40, W = 80, HW = 3200;
__global__ void p1(uint4* out, int N) {   // thread = src pixel, 8 rows x 32 B, loop x tiles (current pattern)
  const int items = N * 25 * 5;            // (n, mt, yb)
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int yb = item % 5, mt = (item / 5) % 25, n = item / 125;
    const int src = mt * 128 + threadIdx.x;
    char* plane = (char*)out + ((size_t)n * HW + src) * (H * W * 2);
    for (int xb = 0; xb < 5; ++xb)
      for (int r = 0; r < 8; ++r) {
        uint4* p = (uint4*)(plane + ((yb * 8 + r) * W + xb * 16) * 2);
        uint4 v = make_uint4(item, xb, r, src);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
  }
}
__global__ void p3(uint4* out, int N) {   // warp = one src pixel at a time, 2 full rows (320 B contiguous) per instr
  const int items = N * 25;                // (n, mt): loop all 20 row-pairs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int mt = item % 25, n = item / 25;
    for (int rp = 0; rp < 20; ++rp)
      for (int s = warp; s < 128; s += 4) {
        const int src = mt * 128 + s;
        char* plane = (char*)out + ((size_t)n * HW + src) * (H * W * 2);
        if (lane < 20) {
          uint4* p = (uint4*)(plane + rp * 320) + lane;
          *p = make_uint4(item, rp, s, lane);
        }
      }
  }
}
__global__ void p5(uint4* out, int N) {   // warp = one src pixel, 8 full rows = 1280 B contiguous (80 x 16 B)
  const int items = N * 25;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int mt = item % 25, n = item / 25;
    for (int yb = 0; yb < 5; ++yb)
      for (int s = warp; s < 128; s += 4) {
        const int src = mt * 128 + s;
        uint4* p = (uint4*)((char*)out + ((size_t)n * HW + src) * (H * W * 2) + yb * 1280);
        for (int k = lane; k < 80; k += 32) p[k] = make_uint4(item, yb, s, k);
      }
  }
}
// ---- tiled slot-pool layout (levels 0: 4x4-element tiles, tile-row-major: a band of 8 rows = 1280 B contiguous per plane)
// p6: the build kernel's epilogue pattern: thread = src pixel; per MMA tile (8x16 patch) two 128-byte runs (one per
//     tile-row, 640 B apart) as 4 + 4 STG.256; 5 x-tiles per band, 5 bands
__global__ void p6(uint4* out, int N) {
  const int items = N * 25 * 5;
  const int group = threadIdx.x >> 7, lane_src = threadIdx.x & 127;      // blockDim/128 groups alternate x-tiles
  const int ngroups = blockDim.x >> 7;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int yb = item % 5, mt = (item / 5) % 25, n = item / 125;
    const int src = mt * 128 + lane_src;
    char* band = (char*)out + ((size_t)n * HW + src) * (H * W * 2) + yb * 1280;
    for (int xb = group; xb < 5; xb += ngroups)
      for (int tr = 0; tr < 2; ++tr)
        for (int t = 0; t < 4; ++t) {
          char* p = band + tr * 640 + (xb * 4 + t) * 32;
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p), "r"(item), "r"(xb), "r"(tr), "r"(src) : "memory");
        }
  }
}
// p7: same bytes, but one store instruction = 8 planes x 128 contiguous bytes (lane -> plane lane/4, 32-byte piece lane%4)
__global__ void p7(uint4* out, int N) {
  const int items = N * 25 * 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int yb = item % 5, mt = (item / 5) % 25, n = item / 125;
    // work units: (x-tile, tile-row, group of 8 planes): 5 * 2 * 16 = 160 per item
    for (int u = warp; u < 160; u += nwarps) {
      const int pg = u % 16, tr = (u / 16) % 2, xb = u / 32;
      const int src = mt * 128 + pg * 8 + (lane >> 2);
      char* p = (char*)out + ((size_t)n * HW + src) * (H * W * 2) + yb * 1280 + tr * 640 + (xb * 4 + (lane & 3)) * 32;
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p), "r"(item), "r"(xb), "r"(tr), "r"(src) : "memory");
    }
  }
}
// p8: staged band: one warp writes one plane's whole 1280-byte band (40 x 32 B: 32 lanes + 8 lanes)
__global__ void p8(uint4* out, int N) {
  const int items = N * 25 * 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int yb = item % 5, mt = (item / 5) % 25, n = item / 125;
    for (int s = warp; s < 128; s += nwarps) {
      char* band = (char*)out + ((size_t)n * HW + mt * 128 + s) * (H * W * 2) + yb * 1280;
      for (int k = lane; k < 40; k += 32) {
        char* p = band + k * 32;
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p), "r"(item), "r"(k), "r"(s), "r"(lane) : "memory");
      }
    }
  }
}
__global__ void p4(uint4* out, size_t n16) {   // fully coalesced stream
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_uint4(i, 1, 2, 3);
}
int main() {
  const int N = 36;
  const size_t bytes = (size_t)N * HW * H * W * 2;
  uint4* buf; cudaMalloc(&buf, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-48s %8.1f us  %7.1f GB/s  (%s)\n", name, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("p1 thread=src, 8x32B rows, planes apart (x128thr)", [&] { p1<<<148, 128>>>(buf, N); });
  run("p1 same, 2 CTAs/SM", [&] { p1<<<296, 128>>>(buf, N); });
  run("p1 same, 4 CTAs/SM", [&] { p1<<<592, 128>>>(buf, N); });
  run("p3 warp=src, 320 B contiguous", [&] { p3<<<148, 128>>>(buf, N); });
  run("p3 same, 4 CTAs/SM", [&] { p3<<<592, 128>>>(buf, N); });
  run("p5 warp=src, 1280 B contiguous", [&] { p5<<<148, 128>>>(buf, N); });
  run("p5 same, 4 CTAs/SM", [&] { p5<<<592, 128>>>(buf, N); });
  run("p4 coalesced stream", [&] { p4<<<148 * 8, 256>>>(buf, bytes / 16); });
  run("p6 tiled, thread=src, 2x128 B runs, 128 thr/SM", [&] { p6<<<148, 128>>>(buf, N); });
  run("p6 same, 256 thr/SM", [&] { p6<<<148, 256>>>(buf, N); });
  run("p6 same, 512 thr/SM (the kernel's 16 epilogue warps)", [&] { p6<<<148, 512>>>(buf, N); });
  run("p7 tiled, 8 planes x 128 B per instruction, 128 thr/SM", [&] { p7<<<148, 128>>>(buf, N); });
  run("p7 same, 512 thr/SM", [&] { p7<<<148, 512>>>(buf, N); });
  run("p8 tiled, warp = plane, 1280 B band, 128 thr/SM", [&] { p8<<<148, 128>>>(buf, N); });
  run("p8 same, 512 thr/SM", [&] { p8<<<148, 512>>>(buf, N); });
  return 0;
}
