// fp32-accurate GEMM on the 5th-generation tensor cores for the TRAINING path: the 128-wide Linear layers of every GVP
// (forward, input gradient, weight gradient; pharmacodiff.py:162-297 -> gvp.py:89-116 and its backward).
//
//   C[M][N] (+)= sum_k A(m, k) B(k, n) (+ bias[n]),   A(m, k) = A[m a_rs + k a_cs],  B(k, n) = B[k b_rs + n b_cs]
//
// Numerics: both operands are split on the fly into bf16 (hi, lo) and contracted in three tcgen05.mma kind::f16 passes
// hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM: 16 significand bits per operand with the exponent range of fp32.
// The fp16 split of the sampling kernels (22 bits) is not usable here: gradients reach 1e-8, far below fp16's subnormal
// range (a 2 % error on the smallest weight gradients when tried), and the training parity bar is 1e-3, not 1e-4.
// The tensor core accumulates with truncation, so the error of one accumulator grows linearly with the number of MMAs
// (measured: 1.6e-3 absolute on sums of magnitude 180 after 384 MMAs); a split-K CTA therefore flushes its accumulator
// after at most ~2048 contraction elements per CTA when the problem is large enough to fill the grid (k_per_cta below).
//
// One persistent CTA per SM, warp-specialised (416 threads):
//   warps 0-3   A producers: thread r owns tile row r = TMEM lane r; per K-chunk of 16 it loads its 16 values, splits them
//               and stores the chunk as the MMA's A operand in TMEM (tcgen05.st; 8 chunks in flight)
//   warps 4-7   epilogue: accumulator (TMEM, double buffered) -> registers -> C (+ bias / accumulate)
//   warps 8-11  B producers: B as bf16 (hi, lo) UMMA K-major SWIZZLE_NONE images in shared memory
//   warp 12     MMA issue (one elected lane), tcgen05.commit hands buffers back through mbarriers
// MODE 0 (forward, dgrad): N, K <= 176; the whole B image is built once per CTA and stays resident; tiles run over M.
// MODE 1 (wgrad): M <= 128 output rows, the contraction index (edges / nodes) is split across CTAs; A and B chunks are both
//   streamed; every CTA writes its partial [128 x n_pad] to a workspace and reduce_partials_kernel sums the partials in
//   CTA order -- a deterministic two-stage reduction (the FFMA path used split-K atomics).
#include "pf_common.cuh"
#include "pf_tc.cuh"

namespace pf {
namespace tcg {

constexpr int kRows = 128;
constexpr int kMaxN = 176;
constexpr int kAStages = 8;    // A chunks in flight in TMEM: 16 columns each (hi 8 | lo 8)
constexpr int kBStages = 4;    // streamed B chunks in flight (MODE 1)
constexpr int kThreads = 32 * 13;
constexpr uint32_t kColD = 0;              // two accumulators of up to 176 columns
constexpr uint32_t kColA = 2 * kMaxN;      // 352 .. 480

struct Params {
  const float* A;
  long long a_rs, a_cs;
  const float* B;
  long long b_rs, b_cs;
  const float* bias;
  float* C;
  int ldc, M, N, K, n_pad, k_pad, accumulate;
  float* partial;     // MODE 1: [gridDim.x][128][n_pad]
  int k_per_cta;      // MODE 1: contraction elements per CTA (multiple of 16)
};

struct Bars {
  uint64_t a_full[kAStages], a_empty[kAStages], b_full[kBStages], b_empty[kBStages], d_full[2], d_empty[2], b_res;
};

// 8 consecutive k of column n, split, into the K-major SWIZZLE_NONE image of one 16-wide chunk:
// byte (k8)*lbo + (n/8)*128 + (n%8)*16, hi image at +0, lo image at +lo_off
__device__ __forceinline__ void store_b8(uint8_t* chunk, uint32_t lbo, uint32_t lo_off, int n, int k8, const float (&x)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::split_pack(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
  uint8_t* a = chunk + (size_t)k8 * lbo + (n >> 3) * 128 + (n & 7) * 16;
  *reinterpret_cast<uint4*>(a) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(a + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pad = p.n_pad;
  const uint32_t lbo = (uint32_t)(n_pad / 8) * 128;       // bytes between the two 8-k halves of a chunk image
  const uint32_t chunk_bytes = (uint32_t)n_pad * 32;      // one 16-k chunk, one of (hi, lo)
  // MODE 0: contraction = p.K (<= 176), tiles over M.  MODE 1: contraction = this CTA's slice of p.K, one tile.
  const int k_begin = MODE == 1 ? blockIdx.x * p.k_per_cta : 0;
  const int k_end = MODE == 1 ? min(p.K, k_begin + p.k_per_cta) : p.K;
  const int n_chunks = k_end > k_begin ? (k_end - k_begin + 15) / 16 : 0;
  const int n_tiles = MODE == 1 ? 1 : (p.M + kRows - 1) / kRows;
  const int my_tiles = MODE == 1 ? 1 : (n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
  const uint32_t b_lo_off = MODE == 0 ? (uint32_t)(p.k_pad / 16) * chunk_bytes : chunk_bytes;   // hi block | lo block
  const uint32_t b_stage_bytes = 2 * chunk_bytes;                                                // MODE 1 stage = hi | lo

  if (warp == 12) {
    tc::tmem_alloc(&s_tmem, 512);
    if (lane == 0) {
      for (int s = 0; s < kAStages; ++s) {
        tc::mbar_init(&bars.a_full[s], 128);
        tc::mbar_init(&bars.a_empty[s], 1);
      }
      for (int s = 0; s < kBStages; ++s) {
        tc::mbar_init(&bars.b_full[s], 128);
        tc::mbar_init(&bars.b_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        tc::mbar_init(&bars.d_full[s], 1);
        tc::mbar_init(&bars.d_empty[s], 128);
      }
      tc::mbar_init(&bars.b_res, 128);
      tc::fence_mbar_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers
    const int r = threadIdx.x;                                    // tile row == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t it = 0;
    // the loads of chunk i + 1 are issued before chunk i is split and stored: one global-memory latency per tile row is
    // paid once, not once per chunk (the chunk sequence runs across tile boundaries)
    auto load_chunk = [&](float (&x)[16], int t, int c) {
      const long long m = MODE == 1 ? r : ((long long)(blockIdx.x + (long long)t * gridDim.x) * kRows + r);
      const bool row_ok = t < my_tiles && m < p.M;
      const float* arow = p.A + (row_ok ? m : 0) * p.a_rs;
      const int k0 = k_begin + 16 * c;
      if (row_ok && p.a_cs == 1 && k0 + 16 <= k_end && ((reinterpret_cast<uintptr_t>(arow + k0) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(arow + k0) + j);
          x[4 * j] = v.x, x[4 * j + 1] = v.y, x[4 * j + 2] = v.z, x[4 * j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (row_ok && k0 + j < k_end) ? __ldg(arow + (long long)(k0 + j) * p.a_cs) : 0.f;
      }
    };
    float xn[16];
    if (my_tiles > 0 && n_chunks > 0) load_chunk(xn, 0, 0);
    for (int t = 0; t < my_tiles; ++t) {
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const int s = it % kAStages;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = xn[j];
        if (c + 1 < n_chunks)
          load_chunk(xn, t, c + 1);
        else if (t + 1 < my_tiles)
          load_chunk(xn, t + 1, 0);
        if (it >= kAStages) tc::mbar_wait(&bars.a_empty[s], ((it / kAStages) - 1) & 1);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::split_pack(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
        tc::tmem_st8(tmem + lane_base + kColA + 16 * s, hi);
        tc::tmem_st8(tmem + lane_base + kColA + 16 * s + 8, lo);
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&bars.a_full[s]);
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp - 4, r = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      tc::mbar_wait(&bars.d_full[buf], (t >> 1) & 1);
      tc::fence_after_sync();
      const long long m = MODE == 1 ? r : ((long long)(blockIdx.x + (long long)t * gridDim.x) * kRows + r);
      for (int c0 = 0; c0 < n_pad; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + lane_base + kColD + kMaxN * buf + c0, v);
        tc::wait_ld();
        if (MODE == 1) {
          float* out = p.partial + ((size_t)blockIdx.x * kRows + r) * n_pad + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else if (m < p.M) {
          float* out = p.C + m * p.ldc + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < p.N) {
              float y = __uint_as_float(v[j]);
              if (p.bias) y += __ldg(p.bias + c0 + j);
              if (p.accumulate) y += out[j];
              out[j] = y;
            }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&bars.d_empty[buf]);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ B producers
    const int tid = threadIdx.x - 256;   // 0 .. 127
    if (MODE == 0) {
      if (my_tiles > 0) {
        const int groups = p.k_pad / 8;
        for (int item = tid; item < n_pad * groups; item += 128) {
          const int n = item % n_pad, g = item / n_pad;           // adjacent threads: adjacent n
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = 8 * g + i;
            x[i] = (n < p.N && k < p.K) ? __ldg(p.B + (long long)k * p.b_rs + (long long)n * p.b_cs) : 0.f;
          }
          store_b8(smem + (size_t)(g >> 1) * chunk_bytes, lbo, b_lo_off, n, g & 1, x);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&bars.b_res);
      }
    } else {
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % kBStages;
        if (c >= kBStages) tc::mbar_wait(&bars.b_empty[s], ((c / kBStages) - 1) & 1);
        const int k0 = k_begin + 16 * c;
        uint8_t* stage = smem + (size_t)s * b_stage_bytes;
        for (int item = tid; item < n_pad * 2; item += 128) {
          const int n = item % n_pad, g = item / n_pad;
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = k0 + 8 * g + i;
            x[i] = (n < p.N && k < k_end) ? __ldg(p.B + (long long)k * p.b_rs + (long long)n * p.b_cs) : 0.f;
          }
          store_b8(stage, lbo, b_lo_off, n, g, x);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&bars.b_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issue
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = tc::make_idesc_bf16(kRows, n_pad);
      const uint32_t smem_a = tc::smem_u32(smem);
      if (MODE == 0) tc::mbar_wait(&bars.b_res, 0);
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t & 1;
        if (t >= 2) tc::mbar_wait(&bars.d_empty[buf], ((t >> 1) - 1) & 1);
        const uint32_t dcol = tmem + kColD + kMaxN * buf;
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const int s = it % kAStages;
          tc::mbar_wait(&bars.a_full[s], (it / kAStages) & 1);
          uint32_t b_base;
          if (MODE == 0) {
            b_base = smem_a + (uint32_t)c * chunk_bytes;
          } else {
            const int bs = c % kBStages;
            tc::mbar_wait(&bars.b_full[bs], (c / kBStages) & 1);
            b_base = smem_a + (uint32_t)bs * b_stage_bytes;
          }
          tc::fence_after_sync();
          const uint32_t a_hi = tmem + kColA + 16 * s, a_lo = a_hi + 8;
          const uint64_t b_hi = tc::make_smem_desc(b_base, lbo, 128), b_lo = tc::make_smem_desc(b_base + b_lo_off, lbo, 128);
          tc::mma_ts(dcol, a_hi, b_hi, idesc, c > 0);
          tc::mma_ts(dcol, a_hi, b_lo, idesc, 1);
          tc::mma_ts(dcol, a_lo, b_hi, idesc, 1);
          tc::mma_commit(&bars.a_empty[s]);
          if (MODE == 1) tc::mma_commit(&bars.b_empty[c % kBStages]);
        }
        if (n_chunks == 0) {   // empty contraction slice: the partial is zero
          // (only MODE 1 can get here; accumulator left undefined, handled by the reducer through k_per_cta)
        }
        tc::mma_commit(&bars.d_full[buf]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 12) tc::tmem_dealloc(tmem, 512);
}

// C[r][c] (+)= sum over CTAs, in CTA order, of partial[cta][r][c] (+ nothing else): the deterministic second stage of MODE 1
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int rows, int n_pad, int N,
                                       float* __restrict__ C, int ldc, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * N) return;
  const int r = idx / N, c = idx - r * N;
  float acc = 0.f;
  for (int q = 0; q < n_parts; ++q) acc += partial[((size_t)q * kRows + r) * n_pad + c];
  C[(size_t)r * ldc + c] = accumulate ? C[(size_t)r * ldc + c] + acc : acc;
}

}  // namespace tcg
}  // namespace pf

using namespace pf;

extern "C" size_t pf_tc_gemm_workspace_bytes(int32_t M, int32_t N, int64_t K) {
  // MODE 1 partials: one [128][n_pad] block per CTA, at most 2 x 148 CTAs
  (void)M;
  (void)K;
  const int n_pad = (N + 15) / 16 * 16;
  return (size_t)2 * kNumSms * tcg::kRows * n_pad * sizeof(float);
}

extern "C" int pf_tc_gemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int64_t K,
                          int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  PF_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0 && ldc >= N, "pf_tc_gemm: arguments");
  if (M == 0 || N == 0) return PF_OK;
  PF_CHECK_ARG(N <= tcg::kMaxN, "pf_tc_gemm: N > 176");
  const int n_pad = (N + 15) / 16 * 16;
  tcg::Params p{};
  p.A = A, p.a_rs = a_rs, p.a_cs = a_cs, p.B = B, p.b_rs = b_rs, p.b_cs = b_cs, p.bias = bias, p.C = C, p.ldc = ldc;
  p.M = M, p.N = N, p.n_pad = n_pad, p.accumulate = accumulate;
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  const int kMaxSmem = 2 * tcg::kMaxN * tcg::kMaxN * 2;   // resident B: hi + lo images of a 176 x 176 operand = 123,904 B
  if (!configured.done[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(tcg::tc_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tcg::tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) {
      set_error("pf_tc_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PF_ERR_LAUNCH;
    }
    configured.done[dev_] = true;
  }
  if (K <= tcg::kMaxN) {
    // MODE 0: B resident, persistent over the row tiles
    p.K = (int)K;
    p.k_pad = ((int)K + 15) / 16 * 16;
    if (p.k_pad == 0) p.k_pad = 16;
    const size_t smem = (size_t)2 * n_pad * p.k_pad * 2;
    const long long tiles = ((long long)M + tcg::kRows - 1) / tcg::kRows;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    tcg::tc_gemm_kernel<0><<<grid, tcg::kThreads, smem, as_stream(stream)>>>(p);
    PF_CHECK_LAUNCH("pf_tc_gemm");
    return PF_OK;
  }
  // MODE 1: M <= 128 output rows, the long contraction split over CTAs, deterministic two-stage reduction
  PF_CHECK_ARG(M <= tcg::kRows, "pf_tc_gemm: K > 176 needs M <= 128 (weight-gradient shape)");
  PF_CHECK_ARG(K < (1LL << 31), "pf_tc_gemm: K too large");
  PF_CHECK_ARG(workspace && workspace_bytes >= pf_tc_gemm_workspace_bytes(M, N, K), "pf_tc_gemm: workspace too small");
  long long want = (K + 2047) / 2048;                       // >= 2048 contraction elements per CTA
  const long long cap = 2LL * num_sms();
  int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  long long per = ((K + grid - 1) / grid + 15) / 16 * 16;
  grid = (int)((K + per - 1) / per);
  p.K = (int)K;
  p.k_pad = 16;
  p.k_per_cta = (int)per;
  p.partial = static_cast<float*>(workspace);
  const size_t smem = (size_t)tcg::kBStages * 2 * n_pad * 32;
  tcg::tc_gemm_kernel<1><<<grid, tcg::kThreads, smem, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_tc_gemm (split-K)");
  const int total = M * N;
  tcg::reduce_partials_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(p.partial, grid, M, n_pad, N, C, ldc,
                                                                                 accumulate);
  PF_CHECK_LAUNCH("pf_tc_gemm (reduce)");
  return PF_OK;
}
