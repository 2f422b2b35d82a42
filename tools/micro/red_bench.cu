// Microbenchmark: throughput of global reductions (red.global.add) with the access patterns the attention backward could
// use for its dQ accumulation.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

// Layout mimics dq_acc: rows x 768 fp32; a CTA (128 threads) adds a [128 rows x 64 cols] tile, 13 CTAs hit each tile.
template <int MODE>
__global__ void red_kernel(float* acc, int rows, int D, int passes) {
  const int tiles_per_head = rows / 128;
  const int tile = blockIdx.x % tiles_per_head;
  const int h = (blockIdx.x / tiles_per_head) % (D / 64);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < passes; ++p) {
    const int t = (tile + p) % tiles_per_head;
    if (MODE == 0) {   // lane pairs: adjacent 16-byte chunks of one row (32-byte sector per pair), as the kernel does today
      const int r = t * 128 + warp * 32 + (lane & ~1);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float* a0 = acc + (long long)r * D + h * 64 + 8 * q + 4 * (lane & 1);
        float* a1 = a0 + D;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a0), "f"(1.0f) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a1), "f"(1.0f) : "memory");
      }
    } else if (MODE == 1) {   // 8 lanes cover one 128-byte line (4 rows per warp instruction)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int r = t * 128 + warp * 32 + (q >> 1) * 4 + (lane >> 3);
        float* a0 = acc + (long long)r * D + h * 64 + (q & 1) * 32 + 4 * (lane & 7);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a0), "f"(1.0f) : "memory");
      }
    } else if (MODE == 2) {   // scalar f32, a warp covers one 128-byte line
      for (int q = 0; q < 64; ++q) {
        const int r = t * 128 + warp * 32 + (q >> 1);
        float* a0 = acc + (long long)r * D + h * 64 + (q & 1) * 32 + lane;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a0), "f"(1.0f) : "memory");
      }
    } else if (MODE == 3) {   // bf16x2 packed, v4 (16 bytes = 8 bf16): rows are 64 bf16 = 128 bytes; 8 lanes per row
      __nv_bfloat16* accb = reinterpret_cast<__nv_bfloat16*>(acc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = t * 128 + warp * 32 + q * 4 + (lane >> 3);
        __nv_bfloat16* a0 = accb + (long long)r * D + h * 64 + 8 * (lane & 7);
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" ::"l"(a0), "r"(0x3f803f80u) : "memory");
      }
    } else if (MODE == 4) {   // plain 16-byte stores with the MODE 1 pattern (upper bound: no atomic ALU)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int r = t * 128 + warp * 32 + (q >> 1) * 4 + (lane >> 3);
        float4* a0 = reinterpret_cast<float4*>(acc + (long long)r * D + h * 64 + (q & 1) * 32 + 4 * (lane & 7));
        *a0 = make_float4(1.f, 1.f, 1.f, 1.f);
      }
    }
  }
}

template <int MODE>
void run(const char* name, float* acc, int rows, int D, int passes) {
  const int grid = (rows / 128) * (D / 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) red_kernel<MODE><<<grid, 128>>>(acc, rows, D, passes);
  cudaEventRecord(e0);
  for (int it = 0; it < 5; ++it) red_kernel<MODE><<<grid, 128>>>(acc, rows, D, passes);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double elems = (double)grid * passes * 128 * 64;
  printf("%-44s %.3f ms  %.1f G elem/s  (%.2f TB/s as fp32)  err=%s\n", name, ms, elems / ms / 1e6, elems * 4 / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows = 102400, D = 768, passes = 13;   // one encoder layer's dq accumulator at cfg2
  float* acc;
  cudaMalloc(&acc, (size_t)rows * D * 4);
  cudaMemset(acc, 0, (size_t)rows * D * 4);
  run<0>("red.v4.f32, lane pairs share a sector", acc, rows, D, passes);
  run<1>("red.v4.f32, 8 lanes per 128B line", acc, rows, D, passes);
  run<2>("red.f32 scalar, warp per 128B line", acc, rows, D, passes);
  run<3>("red.v4.bf16x2, 8 lanes per 128B row", acc, rows, D, passes);
  run<4>("st.v4.f32 (no atomic), 8 lanes per line", acc, rows, D, passes);
  return 0;
}
