// Microbenchmark: issue rate of tcgen05.mma kind::f16 (M = 128, K = 16) on one SM as a function of N, operand source
// (A from TMEM / shared memory), accumulator reuse and the B shared-memory layout.  Data is zeros: only timing matters.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../pharmacoforge_b200/csrc mma_bench.cu -o mma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "pf_tc.cuh"
using namespace pf;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {  // K-major SWIZZLE_128B: SBO = 1024, LBO = 1 (unused)
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Case {
  int n, a_smem, n_acc, sw128, reps, bg, style;  // bg: other warps hammer TMEM ld/st (1) or shared memory (2)
};

__global__ void __launch_bounds__(288, 1) k(const Case* cases, int ncase, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 8) {
    tc::tmem_alloc(&s_tmem, 512);
    if (lane == 0) {
      tc::mbar_init(&bar, 1);
      tc::fence_mbar_init();
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  uint32_t par = 0;
  for (int ci = 0; ci < ncase; ++ci) {
    const Case c = cases[ci];
    if (threadIdx.x == 0) s_stop = 0;
    __syncthreads();
    if (warp == 8) {
      const uint32_t idesc = tc::make_idesc_f16(128, c.n);
      const uint32_t sb = tc::smem_u32(smem);
      const uint32_t lbo = (uint32_t)(c.n / 8) * 128;
      const uint32_t nmask = (uint32_t)c.n_acc - 1;  // n_acc is a power of two here
      if (c.style == 0) {        // one diverged lane, descriptors rebuilt per MMA (what the fused kernels did)
        if (lane == 0) {
          const long long t0 = clock64();
#pragma unroll 4
          for (int r = 0; r < c.reps; ++r) {
            const uint32_t d = tmem + (r & nmask) * c.n;
            const uint64_t bd = c.sw128 ? desc_sw128(sb + (uint32_t)((r >> 2) & 3) * 32768 + (r & 3) * 32)
                                        : tc::make_smem_desc(sb + (uint32_t)(r & 7) * 8192, lbo, 128);
            if (c.a_smem)
              tc::mma_ss(d, tc::make_smem_desc(sb + 131072 + (r & 3) * 8192, 2048, 128), bd, idesc, r > (int)nmask);
            else
              tc::mma_ts(d, tmem + 384 + (r & 7) * 8, bd, idesc, r > (int)nmask);
          }
          const long long t1 = clock64();
          tc::mma_commit(&bar);
          tc::mbar_wait(&bar, par);
          const long long t2 = clock64();
          out[ci * 2] = t1 - t0;
          out[ci * 2 + 1] = t2 - t0;
          s_stop = 1;
        }
      } else {                   // converged warp, elect.sync around the issue, descriptors advanced by adds
        const long long t0 = clock64();
        const uint64_t bd0 = c.sw128 ? desc_sw128(sb) : tc::make_smem_desc(sb, lbo, 128);
        const uint64_t ad0 = tc::make_smem_desc(sb + 131072, 2048, 128);
        for (int r0 = 0; r0 < c.reps; r0 += 8) {
          if (tc::elect_one()) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t d = tmem + ((uint32_t)j & nmask) * c.n;
              const uint64_t bd = bd0 + (c.sw128 ? (uint64_t)(((j >> 2) * 32768 + (j & 3) * 32) >> 4) : (uint64_t)(j * 8192 >> 4));
              if (c.a_smem)
                tc::mma_ss(d, ad0 + (uint64_t)((j & 3) * 8192 >> 4), bd, idesc, (r0 | j) > (int)nmask);
              else
                tc::mma_ts(d, tmem + 384 + j * 8, bd, idesc, (r0 | j) > (int)nmask);
            }
          }
          __syncwarp();
        }
        const long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(&bar);
        __syncwarp();
        tc::mbar_wait(&bar, par);
        const long long t2 = clock64();
        if (lane == 0) {
          out[ci * 2] = t1 - t0;
          out[ci * 2 + 1] = t2 - t0;
          s_stop = 1;
        }
      }
      par ^= 1;
      __syncwarp();
    } else if (c.bg == 1) {  // TMEM traffic like an epilogue: ld 16 cols, st 16 cols (columns 256..383: not the accumulators)
      const uint32_t base = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
      uint32_t r[16];
      while (!s_stop) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tc::tmem_ld16(base + 16 * j, r);
          tc::wait_ld();
          tc::tmem_st16(base + 16 * j, r);
        }
        tc::wait_st();
      }
    } else if (c.bg == 2) {  // shared-memory traffic
      float4* p = reinterpret_cast<float4*>(smem + 163840) + threadIdx.x;
      float4 acc = make_float4(0, 0, 0, 0);
      while (!s_stop) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = p[j * 256];
          acc.x += v.x;
          p[j * 256] = acc;
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

int main() {
  Case h[64];
  int n = 0;
  for (int style = 0; style < 2; ++style) {
    const int base[][5] = {{128, 0, 1, 0, 0}, {128, 0, 2, 0, 0}, {128, 1, 1, 0, 0}, {128, 0, 1, 1, 0}, {128, 1, 1, 1, 0},
                           {256, 0, 1, 0, 0}, {256, 0, 1, 1, 0}, {64, 0, 1, 0, 0},  {64, 0, 2, 0, 0},  {32, 0, 1, 0, 0},
                           {32, 0, 4, 0, 0},  {16, 0, 1, 0, 0},  {16, 0, 4, 0, 0},  {16, 1, 1, 0, 0},  {128, 0, 1, 0, 1},
                           {128, 0, 1, 0, 2}, {128, 0, 1, 1, 1}, {128, 0, 1, 1, 2}};
    for (auto& b : base) h[n++] = Case{b[0], b[1], b[2], b[3], 96, b[4], style};
  }
  Case* d;
  long long* o;
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&o, n * 16);
  cudaMemcpy(d, h, sizeof(Case) * n, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 288, 196608>>>(d, n, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("error: %s\n", cudaGetErrorString(e));
      return 1;
    }
  }
  long long r[2 * 64];
  cudaMemcpy(r, o, n * 16, cudaMemcpyDeviceToHost);
  printf("%5s %4s %5s %5s %5s %5s %3s | %8s %8s %8s\n", "style", "N", "Asmem", "nacc", "sw128", "reps", "bg", "issue", "total", "cyc/mma");
  for (int i = 0; i < n; ++i)
    printf("%5d %4d %5d %5d %5d %5d %3d | %8lld %8lld %8.1f\n", h[i].style, h[i].n, h[i].a_smem, h[i].n_acc, h[i].sw128, h[i].reps, h[i].bg,
           r[2 * i], r[2 * i + 1], (double)r[2 * i + 1] / h[i].reps);
  return 0;
}
