// Self-test of the tcgen05 building blocks: one 128 x N x K GEMM with A split into bf16 (hi, lo) in TMEM and B as
// pre-packed bf16 (hi, lo) shared-memory images.  Used by tests/test_gpu_tcgen05.py to pin descriptor encodings,
// TMEM operand layout and the mbarrier / commit protocol before the fused kernels rely on them.
#include "pf_common.cuh"
#include "pf_tc.cuh"

namespace pf {

__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(const float* __restrict__ A, const uint8_t* __restrict__ b_hi,
                                                             const uint8_t* __restrict__ b_lo, float* __restrict__ D,
                                                             int K, int N, int npass) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar_w, mbar_d;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t img_bytes = (uint32_t)K * N * 2;
  uint8_t* s_hi = smem;
  uint8_t* s_lo = smem + img_bytes;
  if (warp == 4) {
    tc::tmem_alloc(&s_tmem, 512);
    if (lane == 0) {
      tc::mbar_init(&mbar_w, 1);
      tc::mbar_init(&mbar_d, 1);
      tc::fence_mbar_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t colA_hi = 256, colA_lo = 384;
  if (warp == 4) {
    if (lane == 0) {
      tc::mbar_expect_tx(&mbar_w, npass == 3 ? 2 * img_bytes : img_bytes);
      for (uint32_t o = 0; o < img_bytes; o += 16384) {
        const uint32_t n = img_bytes - o < 16384 ? img_bytes - o : 16384;
        tc::bulk_g2s(s_hi + o, b_hi + o, n, &mbar_w);
        if (npass == 3) tc::bulk_g2s(s_lo + o, b_lo + o, n, &mbar_w);
      }
    }
    tc::named_bar_sync(1, 160);
    tc::fence_after_sync();
    if (lane == 0) {
      tc::mbar_wait(&mbar_w, 0);
      const uint32_t idesc = tc::make_idesc_bf16(128, N);
      const uint32_t lbo = (uint32_t)(N / 8) * 128;
      uint32_t acc = 0;
      for (int s = 0; s < K / 16; ++s) {
        for (int p = 0; p < npass; ++p) {
          const uint32_t a = tmem + (p == 2 ? colA_lo : colA_hi) + s * 8;
          const uint8_t* b = (p == 1 ? s_lo : s_hi) + (size_t)s * 2 * lbo;
          tc::mma_ts(tmem, a, tc::make_smem_desc(tc::smem_u32(b), lbo, 128), idesc, acc);
          acc = 1;
        }
      }
      tc::mma_commit(&mbar_d);
    }
  } else {
    const int r = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_pack(A[(size_t)r * K + k0 + 2 * j], A[(size_t)r * K + k0 + 2 * j + 1], hi[j], lo[j]);
      tc::tmem_st8(tmem + lane_base + colA_hi + k0 / 2, hi);
      tc::tmem_st8(tmem + lane_base + colA_lo + k0 / 2, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    tc::named_bar_sync(1, 160);
    tc::mbar_wait(&mbar_d, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(tmem + lane_base + c0, v);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)r * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

}  // namespace pf

extern "C" int pf_tc_selftest(const float* A, const void* b_hi, const void* b_lo, float* D, int32_t K, int32_t N,
                              int32_t npass, void* stream) {
  PF_CHECK_ARG(A && b_hi && D && (npass == 1 || (npass == 3 && b_lo)), "pf_tc_selftest: null pointer / npass");
  PF_CHECK_ARG(K % 16 == 0 && K >= 16 && K <= 176 && N % 16 == 0 && N >= 16 && N <= 128, "pf_tc_selftest: K, N");
  const size_t smem = (size_t)K * N * 2 * 2;
  cudaError_t e = cudaFuncSetAttribute(pf::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    pf::set_error("pf_tc_selftest: %s", cudaGetErrorString(e));
    return PF_ERR_LAUNCH;
  }
  pf::tc_selftest_kernel<<<1, 160, smem, pf::as_stream(stream)>>>(A, (const uint8_t*)b_hi, (const uint8_t*)b_lo, D, K, N,
                                                                npass);
  PF_CHECK_LAUNCH("pf_tc_selftest");
  return PF_OK;
}
