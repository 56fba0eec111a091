// Differentiable primitives of the TRAINING path (PharmacophoreDiff.forward + backward, pharmacodiff.py:162-243):
// forward and backward kernels of the operations GVP / GVPLayerNorm / GVPMultiEdgeConv are composed of (gvp.py:89-116,
// 159-166, 459-551), fp32 throughout.  The Python host (pharmacoforge_b200/train_ops.py) registers them as
// torch.library custom ops with register_autograd and composes the reference's graph out of them; this first training
// path is UNFUSED (every per-edge tensor is materialised, like the reference) -- the fused tcgen05 kernels serve
// sampling and evaluation.  Vector features are component-major: [rows][3][channels].
#include "pf_common.cuh"

namespace pf {
namespace train {

// ------------------------------------------------------------------------------------------------ deterministic reductions
// Reductions that are split over CTAs park one partial per CTA in the caller's workspace and add the partials up in a FIXED
// order, so the result does not depend on the order the CTAs ran in (run-to-run bit-identical gradients): the split-K weight
// gradients through a second small launch (splitk_reduce_kernel, in place of the memset the atomic form needs), the column
// sums and LayerNorm affine gradients inside the same launch by the LAST CTA to arrive (a ticket counter per output block in
// the zeroed tail of the workspace, reset by that CTA).  Without a workspace the kernels fall back to atomicAdd into a
// zeroed output (pf_train_sgemm's legacy behaviour).
constexpr int kWsTailBytes = PF_TRAIN_WS_TAIL;                 // zeroed ticket counters at the END of the workspace
constexpr int kWsCounters = kWsTailBytes / (int)sizeof(unsigned);
struct SplitWs {
  float* part;         // partials (nullptr: atomic fallback)
  unsigned* counters;  // [kWsCounters], all zero between calls
};
// sum of q[z * stride] over z = first, first + step, ... < n in a FIXED association: four running sums over every fourth term,
// combined as (a0 + a1) + (a2 + a3) -- independent loads in flight, the same result whatever order the CTAs ran in
__device__ __forceinline__ float ordered_sum(const float* q, size_t stride, int first, int step, int n) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int z = first;
  for (; z + 3 * step < n; z += 4 * step) {
    const float x0 = __ldcg(q + (size_t)z * stride), x1 = __ldcg(q + (size_t)(z + step) * stride);
    const float x2 = __ldcg(q + (size_t)(z + 2 * step) * stride), x3 = __ldcg(q + (size_t)(z + 3 * step) * stride);
    a0 += x0;
    a1 += x1;
    a2 += x2;
    a3 += x3;
  }
  for (; z < n; z += step) a0 += __ldcg(q + (size_t)z * stride);
  return (a0 + a1) + (a2 + a3);
}
// Second stage of a split-K GEMM: C (+)= sum over the splits of the row-major partial tiles part[split][tile][TM][TN] (+ bias).
// One warp per output element: lane l adds splits l, l + 32, ... in order, then a fixed xor tree over the lanes -- the result
// does not depend on the order the first stage's CTAs ran in.  (A last-CTA-per-tile reduction inside the GEMM kernel was
// measured first: one CTA adding 592 x 289 partials is latency-bound at 40-60 us; this launch replaces the memset of C that
// the atomic form needs, so the launch count is the same.)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int n_splits, int TM, int TN,
                                                            int tiles_x, int n_tiles, const float* __restrict__ bias,
                                                            float* __restrict__ C, int M, int N, int ldc, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= (long long)M * N) return;
  const int gm = (int)(o / N), gn = (int)(o - (long long)gm * N);
  const int tile = (gm / TM) * tiles_x + gn / TN;
  const float* q = part + (size_t)tile * (TM * TN) + (gm % TM) * TN + gn % TN;
  float v = ordered_sum(q, (size_t)n_tiles * (TM * TN), lane, 32, n_splits);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) {
    if (bias != nullptr) v += bias[gn];
    float* out = C + (size_t)gm * ldc + gn;
    *out = accumulate ? *out + v : v;
  }
}
// returns true in every thread of the LAST CTA of `tile` (of `n_splits`) once all partials are visible
__device__ __forceinline__ bool last_cta_of_tile(unsigned* counters, int tile, int n_splits) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counters + tile, 1u);
    s_last = prev == (unsigned)(n_splits - 1);
    if (s_last) counters[tile] = 0;   // nobody else touches this counter before the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ------------------------------------------------------------------------------------------------ strided SGEMM
// C[M][N] (row-major, ldc) (+)= A(m, k) * B(k, n) (+ bias[n]); A(m, k) = A[m * a_rs + k * a_cs], B(k, n) = B[k * b_rs +
// n * b_cs].  One kernel covers y = x W^T + b, dx = dy W and dW = dy^T x.  K step 16, 256 threads, 128 x 128 tiles with
// 8 x 8 outputs per thread (64 x 64 / 4 x 4 for narrow outputs); gridDim.z > 1 splits K and accumulates with atomicAdd
// into a zeroed / accumulating C.
constexpr int kTK = 16;
template <int TM, int TN, int R>  // TM x TN tile, R x R outputs per thread, (TM / R) * (TN / R) = 256 threads
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    const float* __restrict__ bias, float* __restrict__ C, int M, int N,
                                                    int K, long long a_rs, long long a_cs, long long b_rs,
                                                    long long b_cs, int ldc, int accumulate, int k_chunk, SplitWs ws) {
  static_assert((TM / R) * (TN / R) == 256 && R % 4 == 0, "tile shape");
  __shared__ __align__(16) float sA[kTK][TM + 4];
  __shared__ __align__(16) float sB[kTK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int tm = (tid / (TN / R)) * R, tn = (tid % (TN / R)) * R;
  float acc[R][R] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += kTK) {
#pragma unroll
    for (int i = 0; i < TM * kTK / 256; ++i) {
      const int e = tid + 256 * i;
      // pick the index order that makes consecutive threads walk the unit-stride dimension of the operand
      const int m = a_cs == 1 ? e / kTK : e % TM, k = a_cs == 1 ? e % kTK : e / TM;
      const int gm = m0 + m, gk = k0 + k;
      sA[k][m] = (gm < M && gk < k_end) ? A[gm * a_rs + gk * a_cs] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TN * kTK / 256; ++i) {
      const int e = tid + 256 * i;
      const int n = b_cs == 1 ? e % TN : e / kTK, k = b_cs == 1 ? e / TN : e % kTK;
      const int gn = n0 + n, gk = k0 + k;
      sB[k][n] = (gn < N && gk < k_end) ? B[gk * b_rs + gn * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      float av[R], bv[R];
#pragma unroll
      for (int i = 0; i < R; i += 4) {
        *reinterpret_cast<float4*>(&av[i]) = *reinterpret_cast<const float4*>(&sA[k][tm + i]);
        *reinterpret_cast<float4*>(&bv[i]) = *reinterpret_cast<const float4*>(&sB[k][tn + i]);
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (gridDim.z > 1 && ws.part != nullptr) {
    // deterministic split-K, first stage: this split's partial tile, row-major; splitk_reduce_kernel adds the splits up
    const int tile = blockIdx.y * gridDim.x + blockIdx.x, n_tiles = gridDim.x * gridDim.y;
    float* mine = ws.part + ((size_t)blockIdx.z * n_tiles + tile) * (TM * TN);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; j += 4)
        if (m0 + tm + i < M && n0 + tn + j < N)   // only the part of the tile that exists is ever read back
          *reinterpret_cast<float4*>(mine + (tm + i) * TN + tn + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
    return;
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      float* c = C + (size_t)gm * ldc + gn;
      if (gridDim.z > 1)
        atomicAdd(c, v);
      else
        *c = accumulate ? *c + v : v;
    }
  }
}

// The same contract with a software pipeline, for the GEMMs that carry the step (the Linear(161 / 144 -> 128) of every GVP
// over the pp edges: forward, dgrad and wgrad are ~7.4 GFLOP each at 200 k edges): TM x TN tile with (TM / 16) x (TN / 16)
// outputs per thread (8 x 8 at 128 x 128: one LDS.128 per 16 FFMA instead of one per 8), K step 8, two shared-memory
// stages, and the next stage's global loads issued into registers before the current stage's FFMAs.  The single-stage
// 64 x 64 kernel above stays for the skinny vector-channel contractions (N = 16 / 17 / 32), which are memory-bound.
template <int TM, int TN>
__global__ void __launch_bounds__(256, 2) sgemm_pipe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         const float* __restrict__ bias, float* __restrict__ C, int M,
                                                         int N, int K, long long a_rs, long long a_cs, long long b_rs,
                                                         long long b_cs, int ldc, int accumulate, int k_chunk, SplitWs ws) {
  constexpr int TK = 8, RM = TM / 16, RN = TN / 16, LA = TM * TK / 256, LB = TN * TK / 256;
  static_assert((RM == 4 || RM == 8) && (RN == 4 || RN == 8), "tile shape");
  __shared__ __align__(16) float sA[2][TK][TM + 4];
  __shared__ __align__(16) float sB[2][TK][TN + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const bool a_kmajor = a_cs == 1, b_nmajor = b_cs == 1;
  float ra[LA], rb[LB];
  auto gload = [&](const int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int e = tid + 256 * i;
      const int m = a_kmajor ? e / TK : e % TM, k = a_kmajor ? e % TK : e / TM;
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < M && gk < k_end) ? A[gm * a_rs + gk * a_cs] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + 256 * i;
      const int n = b_nmajor ? e % TN : e / TK, k = b_nmajor ? e / TN : e % TK;
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < N && gk < k_end) ? B[gk * b_rs + gn * b_cs] : 0.f;
    }
  };
  auto sstore = [&](const int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int e = tid + 256 * i;
      const int m = a_kmajor ? e / TK : e % TM, k = a_kmajor ? e % TK : e / TM;
      sA[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + 256 * i;
      const int n = b_nmajor ? e % TN : e / TK, k = b_nmajor ? e / TN : e % TK;
      sB[buf][k][n] = rb[i];
    }
  };
  float acc[RM][RN] = {};
  const int nk = (k_end - k_begin + TK - 1) / TK;
  if (nk > 0) {
    gload(k_begin);
    sstore(0);
  }
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) gload(k_begin + (t + 1) * TK);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[RM], bv[RN];
#pragma unroll
      for (int i = 0; i < RM; i += 4)  // rows ty * 4 + (0..3), then TM / 2 + ty * 4 + (0..3)
        *reinterpret_cast<float4*>(&av[i]) = *reinterpret_cast<const float4*>(&sA[buf][k][(i / 4) * (TM / 2) + ty * 4]);
#pragma unroll
      for (int j = 0; j < RN; j += 4)
        *reinterpret_cast<float4*>(&bv[j]) = *reinterpret_cast<const float4*>(&sB[buf][k][(j / 4) * (TN / 2) + tx * 4]);
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
  if (gridDim.z > 1 && ws.part != nullptr) {   // deterministic split-K, first stage (see sgemm_kernel)
    const int tile = blockIdx.y * gridDim.x + blockIdx.x, n_tiles = gridDim.x * gridDim.y;
    float* mine = ws.part + ((size_t)blockIdx.z * n_tiles + tile) * (TM * TN);
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int r = (i / 4) * (TM / 2) + ty * 4 + (i & 3);
#pragma unroll
      for (int j4 = 0; j4 < RN; j4 += 4) {
        const int c = (j4 / 4) * (TN / 2) + tx * 4;
        if (m0 + r < M && n0 + c < N)
          *reinterpret_cast<float4*>(mine + r * TN + c) = make_float4(acc[i][j4], acc[i][j4 + 1], acc[i][j4 + 2], acc[i][j4 + 3]);
      }
    }
    return;
  }
  const bool vec = gridDim.z == 1 && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int gm = m0 + (i / 4) * (TM / 2) + ty * 4 + (i & 3);
    if (gm >= M) continue;
#pragma unroll
    for (int j4 = 0; j4 < RN; j4 += 4) {
      const int gn = n0 + (j4 / 4) * (TN / 2) + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][j4 + j];
        if (bias != nullptr && blockIdx.z == 0 && gn + j < N) v[j] += bias[gn + j];
      }
      float* c = C + (size_t)gm * ldc + gn;
      if (vec && gn + 3 < N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(c);
          o.x += old.x;
          o.y += old.y;
          o.z += old.z;
          o.w += old.w;
        }
        *reinterpret_cast<float4*>(c) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (gn + j >= N) continue;
          if (gridDim.z > 1)
            atomicAdd(c + j, v[j]);
          else
            c[j] = accumulate ? c[j] + v[j] : v[j];
        }
      }
    }
  }
}

// column sums of a row-major [M][N] matrix (bias gradients): out[n] += sum_m x[m][n].  256 threads = 32 columns x 8 row lanes,
// gridDim.y row chunks: lane j adds rows r0 + j, r0 + j + 8, ... in order, the lanes are added in lane order.  With a workspace
// the chunks' partial sums go to part[chunk][column] and the last CTA of each column block adds them the same way (8 chunk
// lanes, lane order) -- deterministic; without one, atomicAdd.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int N,
                                                     SplitWs ws) {
  __shared__ float s_p[8][32];
  const int cx = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + cx;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (long long r = r0 + j; r < r1; r += 8) s += x[r * N + n];
  s_p[j][cx] = s;
  __syncthreads();
  if (j == 0) {
#pragma unroll
    for (int q = 1; q < 8; ++q) s += s_p[q][cx];
    if (ws.part == nullptr) {
      if (n < N) atomicAdd(out + n, s);
    } else {
      ws.part[(size_t)blockIdx.y * (gridDim.x * 32) + n] = s;
    }
  }
  if (ws.part == nullptr) return;
  if (!last_cta_of_tile(ws.counters, blockIdx.x, gridDim.y)) return;
  s_p[j][cx] = ordered_sum(ws.part + n, (size_t)gridDim.x * 32, j, 8, gridDim.y);
  __syncthreads();
  if (j == 0 && n < N) {
    float v = s_p[0][cx];
#pragma unroll
    for (int q = 1; q < 8; ++q) v += s_p[q][cx];
    out[n] += v;
  }
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void silu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = silu_f(x[i]);
}
__global__ void silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = sigmoid_f(x[i]);
    dx[i] = dy[i] * s * (1.0f + x[i] * (1.0f - s));
  }
}

// Vout[m][c][u] = act(gate[m][u]) * Vu[m][c][u], act = sigmoid (GVP default) or identity (last noise GVP)
__global__ void gate_fwd_kernel(const float* __restrict__ gate, const float* __restrict__ vu, float* __restrict__ out,
                                long long rows, int U, int act_sigmoid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 3 * U) return;
  const long long m = i / (3 * U);
  const int u = (int)(i % U);
  const float g = gate[m * U + u];
  out[i] = (act_sigmoid ? sigmoid_f(g) : g) * vu[i];
}
__global__ void gate_bwd_kernel(const float* __restrict__ gate, const float* __restrict__ vu,
                                const float* __restrict__ dout, float* __restrict__ dgate, float* __restrict__ dvu,
                                long long rows, int U, int act_sigmoid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * U) return;
  const long long m = i / U;
  const int u = (int)(i % U);
  const float g = gate[i];
  const float a = act_sigmoid ? sigmoid_f(g) : g;
  const float da = act_sigmoid ? a * (1.0f - a) : 1.0f;
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long long j = (m * 3 + c) * U + u;
    dvu[j] = a * dout[j];
    dot = fmaf(dout[j], vu[j], dot);
  }
  dgate[i] = da * dot;
}

// sh[m][h] = sqrt(max(sum_c Vh[m][c][h]^2, 1e-8))   (_norm_no_nan, gvp.py:12-19)
__global__ void vecnorm_fwd_kernel(const float* __restrict__ vh, float* __restrict__ sh, long long rows, int H,
                                   int ld_sh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const long long m = i / H;
  const int h = (int)(i % H);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = vh[(m * 3 + c) * H + h];
    s = fmaf(v, v, s);
  }
  sh[m * ld_sh + h] = sqrtf(fmaxf(s, 1e-8f));
}
__global__ void vecnorm_bwd_kernel(const float* __restrict__ vh, const float* __restrict__ dsh, float* __restrict__ dvh,
                                   long long rows, int H, int ld_sh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const long long m = i / H;
  const int h = (int)(i % H);
  float v[3], s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    v[c] = vh[(m * 3 + c) * H + h];
    s = fmaf(v[c], v[c], s);
  }
  const float k = s > 1e-8f ? dsh[m * ld_sh + h] / sqrtf(s) : 0.f;  // the clamp has zero gradient below eps
#pragma unroll
  for (int c = 0; c < 3; ++c) dvh[(m * 3 + c) * H + h] = k * v[c];
}

// ------------------------------------------------------------------------------------------------ layer norms
// nn.LayerNorm(D) with affine parameters, eps 1e-5; one warp per row.  Saves mean / rstd for the backward.
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ b, float* __restrict__ y, float* __restrict__ stats,
                                     long long rows, int D) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + r * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += xr[i];
  const float mean = warp_sum(s) / D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float d = xr[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
  for (int i = lane; i < D; i += 32) y[r * D + i] = (xr[i] - mean) * rstd * w[i] + b[i];
  if (lane == 0) {
    stats[2 * r] = mean;
    stats[2 * r + 1] = rstd;
  }
}
// dx per row; dw / db accumulated over rows with atomicAdd (one partial per warp-row): the fallback without a workspace
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ stats, const float* __restrict__ dy,
                                     float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                     long long rows, int D) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float mean = stats[2 * r], rstd = stats[2 * r + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float xh = (x[r * D + i] - mean) * rstd;
    const float g = dy[r * D + i] * w[i];
    s1 += g;
    s2 = fmaf(g, xh, s2);
    atomicAdd(dw + i, dy[r * D + i] * xh);
    atomicAdd(db + i, dy[r * D + i]);
  }
  s1 = warp_sum(s1) / D;
  s2 = warp_sum(s2) / D;
  for (int i = lane; i < D; i += 32) {
    const float xh = (x[r * D + i] - mean) * rstd;
    dx[r * D + i] = rstd * (dy[r * D + i] * w[i] - s1 - xh * s2);
  }
}
// The same with deterministic dw / db (D <= 128): a CTA of 8 warps owns kLnRows consecutive rows (an argument), each warp sums its rows'
// contributions in row order (4 columns per lane), the warps' sums are added in warp order, the CTA's partial goes to the
// workspace and the last CTA adds the partials in CTA order: dw[i] += ..., db[i] += ... exactly once.
__global__ void __launch_bounds__(256) layernorm_bwd_det_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ stats, const float* __restrict__ dy,
                                                                float* __restrict__ dx, float* __restrict__ dw,
                                                                float* __restrict__ db, long long rows, int D,
                                                                int kLnRows, SplitWs ws) {
  __shared__ float s_w[8][2][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  const long long rbase = (long long)blockIdx.x * kLnRows;
  for (int rr = warp; rr < kLnRows; rr += 8) {
    const long long r = rbase + rr;
    if (r >= rows) break;
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float xh[4], dyv[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = lane + 32 * j;
      xh[j] = dyv[j] = 0.f;
      if (i < D) {
        xh[j] = (x[r * D + i] - mean) * rstd;
        dyv[j] = dy[r * D + i];
        const float g = dyv[j] * w[i];
        s1 += g;
        s2 = fmaf(g, xh[j], s2);
        aw[j] = fmaf(dyv[j], xh[j], aw[j]);
        ab[j] += dyv[j];
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = lane + 32 * j;
      if (i < D) dx[r * D + i] = rstd * (dyv[j] * w[i] - s1 - xh[j] * s2);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_w[warp][0][lane + 32 * j] = aw[j];
    s_w[warp][1][lane + 32 * j] = ab[j];
  }
  __syncthreads();
  const int which = threadIdx.x >> 7, col = threadIdx.x & 127;   // 256 threads = (dw | db) x 128 columns
  float v = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) v += s_w[q][which][col];
  ws.part[(size_t)blockIdx.x * 256 + threadIdx.x] = v;
  if (!last_cta_of_tile(ws.counters, 0, gridDim.x)) return;
  if (col >= D) return;
  float* o = which ? db : dw;
  o[col] += ordered_sum(ws.part + threadIdx.x, 256, 0, 1, gridDim.x);
}


// GVPLayerNorm on vectors (gvp.py:163-165): out = v / vn, vn = sqrt(mean_u(max(|v_u|^2, 1e-8)) + 1e-5) + 1e-5.
// One thread per row ([3][U], U <= 32).
__global__ void vecln_fwd_kernel(const float* __restrict__ v, float* __restrict__ out, long long rows, int U) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const float* p = v + m * 3 * U;
  float acc = 0.f;
  for (int u = 0; u < U; ++u) {
    const float n = p[u] * p[u] + p[U + u] * p[U + u] + p[2 * U + u] * p[2 * U + u];
    acc += fmaxf(n, 1e-8f);
  }
  const float vn = sqrtf(acc / U + 1e-5f) + 1e-5f;
  for (int i = 0; i < 3 * U; ++i) out[m * 3 * U + i] = p[i] / vn;
}
__global__ void vecln_bwd_kernel(const float* __restrict__ v, const float* __restrict__ dout, float* __restrict__ dv,
                                 long long rows, int U) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const float* p = v + m * 3 * U;
  const float* g = dout + m * 3 * U;
  float acc = 0.f, dot = 0.f;
  for (int u = 0; u < U; ++u) {
    const float n = p[u] * p[u] + p[U + u] * p[U + u] + p[2 * U + u] * p[2 * U + u];
    acc += fmaxf(n, 1e-8f);
  }
  for (int i = 0; i < 3 * U; ++i) dot = fmaf(g[i], p[i], dot);
  const float root = sqrtf(acc / U + 1e-5f);
  const float vn = root + 1e-5f;
  // d vn / d v[c][u] = [|v_u|^2 > 1e-8] * v[c][u] / (U * root)
  const float k = dot / (vn * vn) / (U * root);
  for (int u = 0; u < U; ++u) {
    const float n = p[u] * p[u] + p[U + u] * p[U + u] + p[2 * U + u] * p[2 * U + u];
    const float live = n > 1e-8f ? k : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) dv[m * 3 * U + c * U + u] = g[c * U + u] / vn - live * p[c * U + u];
  }
}

// ------------------------------------------------------------------------------------------------ graph data movement
// out[e][:] = x[idx[e]][:]  /  dx[idx[e]][:] += dout[e][:]   (edges.src[...], gvp.py:543-545)
__global__ void gather_fwd_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out,
                                  long long E, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * D) return;
  const long long e = i / D;
  out[i] = x[(long long)idx[e] * D + (i % D)];
}
__global__ void gather_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx, float* __restrict__ dx,
                                  long long E, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * D) return;
  const long long e = i / D;
  atomicAdd(dx + (long long)idx[e] * D + (i % D), dout[i]);
}
// deterministic form of gather_bwd: perm lists the edges grouped by source row (a stable sort of idx), row n owns
// perm[ptr[n] .. ptr[n + 1]); one thread per (row, column) adds its edges' gradients in that order and WRITES dx
__global__ void gather_bwd_sorted_kernel(const float* __restrict__ dout, const int* __restrict__ perm,
                                         const int* __restrict__ ptr, float* __restrict__ dx, long long n_rows, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * D) return;
  const long long n = i / D;
  const int d = (int)(i % D);
  float acc = 0.f;
  for (int e = ptr[n]; e < ptr[n + 1]; ++e) acc += dout[(long long)perm[e] * D + d];
  dx[i] = acc;
}
// mean of the message rows of every destination (fn.mean + cross_reducer sum, gvp.py:488-497): edges are sorted by
// destination, segment s = rows [ptr[s], ptr[s+1]) aggregates onto node seg_dst[s] (or s); out is ACCUMULATED into.
__global__ void segmean_fwd_kernel(const float* __restrict__ msg, const int* __restrict__ ptr,
                                   const int* __restrict__ seg_dst, float* __restrict__ out, int n_seg, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_seg * D) return;
  const int s = (int)(i / D), d = (int)(i % D);
  const int r0 = ptr[s], r1 = ptr[s + 1];
  if (r1 == r0) return;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += msg[(long long)r * D + d];
  const int node = seg_dst ? seg_dst[s] : s;
  out[(long long)node * D + d] += acc / (float)(r1 - r0);
}
__global__ void segmean_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ ptr,
                                   const int* __restrict__ seg_dst, float* __restrict__ dmsg, int n_seg, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_seg * D) return;
  const int s = (int)(i / D), d = (int)(i % D);
  const int r0 = ptr[s], r1 = ptr[s + 1];
  if (r1 == r0) return;
  const int node = seg_dst ? seg_dst[s] : s;
  const float g = dout[(long long)node * D + d] / (float)(r1 - r0);
  for (int r = r0; r < r1; ++r) dmsg[(long long)r * D + d] = g;
}
// x_diff / rbf of every edge (gvp.py:472-480); inputs are data, there is no backward
__global__ void edge_geom_kernel(const float* __restrict__ src_x, const float* __restrict__ dst_x,
                                 const int* __restrict__ src, const int* __restrict__ dst, float* __restrict__ xdiff,
                                 float* __restrict__ rbf, long long E) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int s = src[e], d = dst[e];
  const float dx = src_x[3 * s] - dst_x[3 * d], dy = src_x[3 * s + 1] - dst_x[3 * d + 1],
              dz = src_x[3 * s + 2] - dst_x[3 * d + 2];
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float dist = sqrtf(fmaxf(d2, 1e-8f)) + 1e-8f;
  xdiff[3 * e] = dx / dist;
  xdiff[3 * e + 1] = dy / dist;
  xdiff[3 * e + 2] = dz / dist;
#pragma unroll
  for (int k = 0; k < kRbf; ++k) {
    const float z = (dist - (float)k) / 0.9375f;
    rbf[e * kRbf + k] = expf(-(z * z));
  }
}

}  // namespace train
}  // namespace pf

using namespace pf;
namespace T = pf::train;

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// workspace -> (partials, ticket counters) when it holds `need_floats` partials in front of its zeroed tail and the launch
// needs at most kWsCounters tickets; {nullptr, nullptr} (atomic fallback) otherwise
extern "C" size_t pf_train_workspace_bytes(void) { return pf_tc_gemm_workspace_bytes(128, 176, 0) + PF_TRAIN_WS_TAIL; }

static T::SplitWs split_ws(void* workspace, size_t workspace_bytes, size_t need_floats, long long tickets) {
  if (workspace == nullptr || workspace_bytes < (size_t)T::kWsTailBytes + need_floats * sizeof(float) || tickets > T::kWsCounters)
    return T::SplitWs{nullptr, nullptr};
  return T::SplitWs{static_cast<float*>(workspace),
                    reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + workspace_bytes - T::kWsTailBytes)};
}

static int sgemm_impl(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                      int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate,
                      int32_t split_k, void* workspace, size_t workspace_bytes, void* stream) {
  PF_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0 && ldc >= N, "pf_train_sgemm: arguments");
  if (M == 0 || N == 0) return PF_OK;
  int splits = split_k < 1 ? 1 : split_k;
  int chunk = ((K + splits - 1) / splits + T::kTK - 1) / T::kTK * T::kTK;
  if (chunk < T::kTK) chunk = T::kTK;
  splits = K > 0 ? (K + chunk - 1) / chunk : 1;
  // Shape dispatch: the pipelined 8 x 8 (or 8 x 4 / 4 x 8) kernel for the GEMMs with at least ~100 columns and rows (the
  // scalar Linear of every GVP: forward, dgrad, wgrad), the single-stage 64 x 64 kernel for the skinny vector-channel
  // contractions (N = 16 / 17 / 32 over 3E rows), which are memory-bound and measured slower on large tiles.
  const long long work = (long long)M * N;
  const bool pipe = N >= 96 && M >= 96 && work >= (1 << 16);
  const bool wide_n = N > 64 && (N % 128 == 0 || N % 128 > 64);   // 161 -> 3 x 64, 128 / 144 -> 128-wide tiles
  const bool wide_m = M > 64 && (M % 128 == 0 || M % 128 > 64 || M >= 1024);
  const int TM = pipe ? (wide_m ? 128 : 64) : 64, TN = pipe ? ((wide_m && !wide_n) ? 64 : 128) : 64;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, splits);
  // split-K: deterministic in-kernel reduction through the workspace when it fits, else atomicAdd into a zeroed C
  T::SplitWs ws{nullptr, nullptr};
  if (splits > 1) {
    ws = split_ws(workspace, workspace_bytes, (size_t)splits * grid.x * grid.y * TM * TN, 0);   // two launches, no tickets
    if (ws.part == nullptr && !accumulate) {
      cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, as_stream(stream));
      if (e != cudaSuccess) {
        set_error("pf_train_sgemm: memset: %s", cudaGetErrorString(e));
        return PF_ERR_LAUNCH;
      }
    }
  }
  if (pipe) {
    if (wide_m && wide_n)
      T::sgemm_pipe_kernel<128, 128><<<grid, 256, 0, as_stream(stream)>>>(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs,
                                                                         ldc, accumulate, chunk, ws);
    else if (wide_m)
      T::sgemm_pipe_kernel<128, 64><<<grid, 256, 0, as_stream(stream)>>>(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs,
                                                                        ldc, accumulate, chunk, ws);
    else
      T::sgemm_pipe_kernel<64, 128><<<grid, 256, 0, as_stream(stream)>>>(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs,
                                                                        ldc, accumulate, chunk, ws);
  } else {
    T::sgemm_kernel<64, 64, 4><<<grid, 256, 0, as_stream(stream)>>>(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc,
                                                                   accumulate, chunk, ws);
  }
  PF_CHECK_LAUNCH("pf_train_sgemm");
  if (ws.part != nullptr) {
    const long long n_out = (long long)M * N;
    T::splitk_reduce_kernel<<<(unsigned)((n_out + 7) / 8), 256, 0, as_stream(stream)>>>(
        ws.part, splits, TM, TN, (int)grid.x, (int)(grid.x * grid.y), bias, C, M, N, ldc, accumulate);
    PF_CHECK_LAUNCH("pf_train_sgemm (split-K reduce)");
  }
  return PF_OK;
}

extern "C" int pf_train_sgemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                              int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate,
                              int32_t split_k, void* stream) {
  return sgemm_impl(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, accumulate, split_k, nullptr, 0, stream);
}

extern "C" int pf_train_colsum(const float* x, float* out, int64_t M, int32_t N, void* workspace, size_t workspace_bytes,
                               void* stream) {
  PF_CHECK_ARG(x && out && M >= 0 && N > 0, "pf_train_colsum: arguments");
  if (M == 0) return PF_OK;
  // rows are split over many CTAs (about 1024 rows = 128 per row lane each): bandwidth-trivial, latency-bound per thread
  const long long want = (M + 1023) / 1024;
  const int ysplit = (int)(want > 2048 ? 2048 : (want > 0 ? want : 1));
  dim3 grid((N + 31) / 32, ysplit);
  const T::SplitWs ws = split_ws(workspace, workspace_bytes, (size_t)ysplit * grid.x * 32, grid.x);
  T::colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, out, M, N, ws);
  PF_CHECK_LAUNCH("pf_train_colsum");
  return PF_OK;
}

extern "C" int pf_train_silu(const float* x, const float* dy, float* out, int64_t n, void* stream) {
  PF_CHECK_ARG(x && out && n >= 0, "pf_train_silu: arguments");
  if (n == 0) return PF_OK;
  if (dy == nullptr)
    T::silu_fwd_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, out, n);
  else
    T::silu_bwd_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, dy, out, n);
  PF_CHECK_LAUNCH("pf_train_silu");
  return PF_OK;
}

extern "C" int pf_train_gate(const float* gate, const float* vu, const float* dout, float* out_or_dgate, float* dvu,
                             int64_t rows, int32_t U, int32_t act_sigmoid, void* stream) {
  PF_CHECK_ARG(gate && vu && out_or_dgate && rows >= 0 && U > 0, "pf_train_gate: arguments");
  if (rows == 0) return PF_OK;
  if (dout == nullptr) {
    T::gate_fwd_kernel<<<blocks_for(rows * 3 * U, 256), 256, 0, as_stream(stream)>>>(gate, vu, out_or_dgate, rows, U,
                                                                                   act_sigmoid);
  } else {
    PF_CHECK_ARG(dvu != nullptr, "pf_train_gate: dvu");
    T::gate_bwd_kernel<<<blocks_for(rows * U, 256), 256, 0, as_stream(stream)>>>(gate, vu, dout, out_or_dgate, dvu, rows,
                                                                               U, act_sigmoid);
  }
  PF_CHECK_LAUNCH("pf_train_gate");
  return PF_OK;
}

extern "C" int pf_train_vecnorm(const float* vh, const float* dsh, float* out, int64_t rows, int32_t H, void* stream) {
  PF_CHECK_ARG(vh && out && rows >= 0 && H > 0, "pf_train_vecnorm: arguments");
  if (rows == 0) return PF_OK;
  if (dsh == nullptr)
    T::vecnorm_fwd_kernel<<<blocks_for(rows * H, 256), 256, 0, as_stream(stream)>>>(vh, out, rows, H, H);
  else
    T::vecnorm_bwd_kernel<<<blocks_for(rows * H, 256), 256, 0, as_stream(stream)>>>(vh, dsh, out, rows, H, H);
  PF_CHECK_LAUNCH("pf_train_vecnorm");
  return PF_OK;
}

extern "C" int pf_train_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* stats,
                                      int64_t rows, int32_t D, void* stream) {
  PF_CHECK_ARG(x && w && b && y && stats && rows >= 0 && D > 0, "pf_train_layernorm_fwd: arguments");
  if (rows == 0) return PF_OK;
  T::layernorm_fwd_kernel<<<blocks_for(rows, 8), 256, 0, as_stream(stream)>>>(x, w, b, y, stats, rows, D);
  PF_CHECK_LAUNCH("pf_train_layernorm_fwd");
  return PF_OK;
}
extern "C" int pf_train_layernorm_bwd(const float* x, const float* w, const float* stats, const float* dy, float* dx,
                                      float* dw, float* db, int64_t rows, int32_t D, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  PF_CHECK_ARG(x && w && stats && dy && dx && dw && db && rows >= 0 && D > 0, "pf_train_layernorm_bwd: arguments");
  if (rows == 0) return PF_OK;
  // deterministic dw / db: at most ~300 CTAs of at least 128 rows, one 256-float partial each
  long long per = (rows + 295) / 296;
  per = per < 128 ? 128 : (per + 7) / 8 * 8;
  const long long ctas = (rows + per - 1) / per;
  const T::SplitWs ws = D <= 128 ? split_ws(workspace, workspace_bytes, (size_t)ctas * 256, 1) : T::SplitWs{nullptr, nullptr};
  if (ws.part != nullptr)
    T::layernorm_bwd_det_kernel<<<(unsigned)ctas, 256, 0, as_stream(stream)>>>(x, w, stats, dy, dx, dw, db, rows, D, (int)per, ws);
  else
    T::layernorm_bwd_kernel<<<blocks_for(rows, 8), 256, 0, as_stream(stream)>>>(x, w, stats, dy, dx, dw, db, rows, D);
  PF_CHECK_LAUNCH("pf_train_layernorm_bwd");
  return PF_OK;
}

extern "C" int pf_train_vecln(const float* v, const float* dout, float* out, int64_t rows, int32_t U, void* stream) {
  PF_CHECK_ARG(v && out && rows >= 0 && U > 0, "pf_train_vecln: arguments");
  if (rows == 0) return PF_OK;
  if (dout == nullptr)
    T::vecln_fwd_kernel<<<blocks_for(rows, 128), 128, 0, as_stream(stream)>>>(v, out, rows, U);
  else
    T::vecln_bwd_kernel<<<blocks_for(rows, 128), 128, 0, as_stream(stream)>>>(v, dout, out, rows, U);
  PF_CHECK_LAUNCH("pf_train_vecln");
  return PF_OK;
}

extern "C" int pf_train_gather(const float* x_or_dout, const int32_t* idx, float* out_or_dx, int64_t E, int32_t D,
                               int32_t backward, void* stream) {
  PF_CHECK_ARG(x_or_dout && idx && out_or_dx && E >= 0 && D > 0, "pf_train_gather: arguments");
  if (E == 0) return PF_OK;
  if (!backward)
    T::gather_fwd_kernel<<<blocks_for(E * D, 256), 256, 0, as_stream(stream)>>>(x_or_dout, idx, out_or_dx, E, D);
  else
    T::gather_bwd_kernel<<<blocks_for(E * D, 256), 256, 0, as_stream(stream)>>>(x_or_dout, idx, out_or_dx, E, D);
  PF_CHECK_LAUNCH("pf_train_gather");
  return PF_OK;
}

extern "C" int pf_train_gather_bwd_sorted(const float* dout, const int32_t* perm, const int32_t* ptr, float* dx,
                                          int64_t n_rows, int32_t D, void* stream) {
  PF_CHECK_ARG(dout && perm && ptr && dx && n_rows >= 0 && D > 0, "pf_train_gather_bwd_sorted: arguments");
  if (n_rows == 0) return PF_OK;
  T::gather_bwd_sorted_kernel<<<blocks_for(n_rows * D, 256), 256, 0, as_stream(stream)>>>(dout, perm, ptr, dx, n_rows, D);
  PF_CHECK_LAUNCH("pf_train_gather_bwd_sorted");
  return PF_OK;
}

extern "C" int pf_train_segmean(const float* msg_or_dout, const int32_t* ptr, const int32_t* seg_dst, float* out_or_dmsg,
                                int32_t n_seg, int32_t D, int32_t backward, void* stream) {
  PF_CHECK_ARG(msg_or_dout && ptr && out_or_dmsg && n_seg >= 0 && D > 0, "pf_train_segmean: arguments");
  if (n_seg == 0) return PF_OK;
  if (!backward)
    T::segmean_fwd_kernel<<<blocks_for((long long)n_seg * D, 256), 256, 0, as_stream(stream)>>>(msg_or_dout, ptr, seg_dst,
                                                                                              out_or_dmsg, n_seg, D);
  else
    T::segmean_bwd_kernel<<<blocks_for((long long)n_seg * D, 256), 256, 0, as_stream(stream)>>>(msg_or_dout, ptr, seg_dst,
                                                                                              out_or_dmsg, n_seg, D);
  PF_CHECK_LAUNCH("pf_train_segmean");
  return PF_OK;
}

extern "C" int pf_train_edge_geom(const float* src_x, const float* dst_x, const int32_t* src, const int32_t* dst,
                                  float* xdiff, float* rbf, int64_t E, void* stream) {
  PF_CHECK_ARG(src_x && dst_x && src && dst && xdiff && rbf && E >= 0, "pf_train_edge_geom: arguments");
  if (E == 0) return PF_OK;
  T::edge_geom_kernel<<<blocks_for(E, 256), 256, 0, as_stream(stream)>>>(src_x, dst_x, src, dst, xdiff, rbf, E);
  PF_CHECK_LAUNCH("pf_train_edge_geom");
  return PF_OK;
}

// ------------------------------------------------------------------------------------------------ one GVP per call
// GVP.forward (gvp.py:89-116) and its backward as ONE host call each: the same kernels as above, enqueued back to back
// (the Python-side cost of ~16 custom-op round trips per GVP dominated the first training step).  All buffers are the
// caller's.  Shapes: feats [M][n], vec [M][3][vi], Wh [vi][h], Wu [h][vo], Wf [no][n + h], Wg [vo][no];
// saved for the backward: Vh [3M][h], Vu [3M][vo], s = [feats | sh] [M][n + h], z [M][no], f [M][no], gates [M][vo].
// GEMM dispatch of the training path: the tensor-core kernel (pf_tc_gemm.cu: fp16 hi/lo split, 3 tcgen05 passes, fp32
// accumulation; weight gradients through its deterministic two-stage split-K) wherever the shape fits it -- N <= 176 and
// either a short contraction with many rows (forward, input gradients) or a long contraction with <= 128 output rows
// (weight gradients, needs the workspace) -- and the fp32 FFMA kernels otherwise (small problems, no workspace).
struct GemmWs {
  void* ptr;
  size_t bytes;
};
static int sgemm_ld(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, long long a_rs,
                    long long a_cs, long long b_rs, long long b_cs, int ldc, int accumulate, int split_k, void* stream,
                    GemmWs ws = GemmWs{nullptr, 0}) {
  const bool tc_ok = ws.ptr != nullptr && N <= 176 && N >= 1 && M >= 1;
  // row threshold: a persistent launch pays TMEM allocation, the B image build and a pipeline fill per CTA (~5 us); the
  // small edge types (ff / pf / fp: a few thousand rows) stay on the FFMA kernels
  // and so do the skinny vector-channel contractions (N, K = 16 / 17 / 32): one or two K-steps per 128-row tile make the
  // tile hand-offs, not the arithmetic, the cost (measured 80 us per launch against 25-37 us on the FFMA kernel)
  if (tc_ok && K <= 176 && M >= 16384 && N >= 64 && K >= 64)
    return pf_tc_gemm(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, accumulate, ws.ptr, ws.bytes, stream);
  if (tc_ok && K >= 16384 && M <= 128 && M >= 64 && N >= 64 && bias == nullptr && ws.bytes >= pf_tc_gemm_workspace_bytes(M, N, K))
    return pf_tc_gemm(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, accumulate, ws.ptr, ws.bytes, stream);
  return sgemm_impl(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, accumulate, split_k, ws.ptr, ws.bytes, stream);
}

extern "C" int pf_train_gemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                             int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate,
                             int32_t split_k, void* workspace, size_t workspace_bytes, void* stream) {
  return sgemm_ld(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, accumulate, split_k, stream,
                  GemmWs{workspace, workspace_bytes});
}
// split-K factor of a weight gradient dW[m][n] = sum over `rows`: enough CTAs to fill the 148 SMs four times over even when
// dW is a single 17 x 17 tile (the vector-channel weights: with a fixed cap of 64 splits those reductions over 3E = 600 k
// rows ran on 64 CTAs), at least 512 rows per split.
static int wgrad_splits(long long rows, int m, int n) {
  const long long tiles = (long long)((m + 63) / 64) * ((n + 63) / 64);
  long long cap = (4 * num_sms() + tiles - 1) / tiles;
  if (cap < 1) cap = 1;
  const long long by_rows = rows / 512;
  const long long sp = by_rows < cap ? by_rows : cap;
  return sp > 1 ? (int)sp : 1;
}

extern "C" int pf_train_gvp_fwd(const float* feats, const float* vec, const float* Wh, const float* Wu, const float* Wf,
                                const float* bf, const float* Wg, const float* bg, int64_t M, int32_t n, int32_t vi,
                                int32_t h, int32_t vo, int32_t no, int32_t act_sigmoid, float* Vh, float* Vu, float* s,
                                float* z, float* f, float* gates, float* vout, void* workspace, size_t workspace_bytes,
                                void* stream) {
  const GemmWs ws{workspace, workspace_bytes};
  PF_CHECK_ARG(feats && vec && Wh && Wu && Wf && bf && Wg && bg && Vh && Vu && s && z && f && gates && vout,
               "pf_train_gvp_fwd: null pointer");
  if (M == 0) return PF_OK;
  const int M3 = (int)(3 * M), K = n + h;
  int rc;
  if ((rc = sgemm_ld(vec, Wh, nullptr, Vh, M3, h, vi, vi, 1, h, 1, h, 0, 1, stream, ws)) != PF_OK) return rc;   // Vh = V Wh
  if ((rc = sgemm_ld(Vh, Wu, nullptr, Vu, M3, vo, h, h, 1, vo, 1, vo, 0, 1, stream, ws)) != PF_OK) return rc;   // Vu = Vh Wu
  cudaError_t e = cudaMemcpy2DAsync(s, (size_t)K * 4, feats, (size_t)n * 4, (size_t)n * 4, (size_t)M,
                                    cudaMemcpyDeviceToDevice, as_stream(stream));
  if (e != cudaSuccess) {
    set_error("pf_train_gvp_fwd: copy: %s", cudaGetErrorString(e));
    return PF_ERR_LAUNCH;
  }
  T::vecnorm_fwd_kernel<<<blocks_for(M * h, 256), 256, 0, as_stream(stream)>>>(Vh, s + n, M, h, K);      // s = [feats|sh]
  PF_CHECK_LAUNCH("pf_train_gvp_fwd(vecnorm)");
  if ((rc = sgemm_ld(s, Wf, bf, z, (int)M, no, K, K, 1, 1, K, no, 0, 1, stream, ws)) != PF_OK) return rc;       // z = s Wf^T + bf
  T::silu_fwd_kernel<<<blocks_for(M * no, 256), 256, 0, as_stream(stream)>>>(z, f, M * no);
  PF_CHECK_LAUNCH("pf_train_gvp_fwd(silu)");
  if ((rc = sgemm_ld(f, Wg, bg, gates, (int)M, vo, no, no, 1, 1, no, vo, 0, 1, stream, ws)) != PF_OK) return rc;
  T::gate_fwd_kernel<<<blocks_for(M * 3 * vo, 256), 256, 0, as_stream(stream)>>>(gates, Vu, vout, M, vo, act_sigmoid);
  PF_CHECK_LAUNCH("pf_train_gvp_fwd(gate)");
  return PF_OK;
}

// grads: dfeats [M][n], dvec [M][3][vi], dWh, dWu, dWf, dbf, dWg, dbg (all overwritten; dbf / dbg must come ZEROED).
// scratch: dgates [M][vo], dVu [3M][vo], dfz [M][no] (df then dz), ds [M][n + h], dVh [3M][h].
extern "C" int pf_train_gvp_bwd(const float* vec, const float* Wh, const float* Wu, const float* Wf, const float* Wg,
                                const float* Vh, const float* Vu, const float* s, const float* z, const float* f,
                                const float* gates, const float* df_out, const float* dvout, int64_t M, int32_t n,
                                int32_t vi, int32_t h, int32_t vo, int32_t no, int32_t act_sigmoid, float* dgates,
                                float* dVu, float* dfz, float* ds, float* dVh, float* dfeats, float* dvec, float* dWh,
                                float* dWu, float* dWf, float* dbf, float* dWg, float* dbg, void* workspace,
                                size_t workspace_bytes, void* stream) {
  const GemmWs ws{workspace, workspace_bytes};
  PF_CHECK_ARG(vec && Wh && Wu && Wf && Wg && Vh && Vu && s && z && f && gates && df_out && dvout && dgates && dVu && dfz &&
                   ds && dVh && dfeats && dvec && dWh && dWu && dWf && dbf && dWg && dbg,
               "pf_train_gvp_bwd: null pointer");
  if (M == 0) return PF_OK;
  const int M3 = (int)(3 * M), K = n + h, Mi = (int)M;

  cudaStream_t st = as_stream(stream);
  int rc;
  T::gate_bwd_kernel<<<blocks_for(M * vo, 256), 256, 0, st>>>(gates, Vu, dvout, dgates, dVu, M, vo, act_sigmoid);
  PF_CHECK_LAUNCH("pf_train_gvp_bwd(gate)");
  // gates = f Wg^T + bg
  if ((rc = sgemm_ld(dgates, f, nullptr, dWg, vo, no, Mi, 1, vo, no, 1, no, 0, wgrad_splits(M, vo, no), stream, ws)) != PF_OK) return rc;  // dWg = dgates^T f
  if ((rc = pf_train_colsum(dgates, dbg, M, vo, ws.ptr, ws.bytes, stream)) != PF_OK) return rc;
  cudaError_t e = cudaMemcpyAsync(dfz, df_out, (size_t)M * no * 4, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) {
    set_error("pf_train_gvp_bwd: copy: %s", cudaGetErrorString(e));
    return PF_ERR_LAUNCH;
  }
  if ((rc = sgemm_ld(dgates, Wg, nullptr, dfz, Mi, no, vo, vo, 1, no, 1, no, 1, 1, stream, ws)) != PF_OK) return rc;  // df += dgates Wg
  T::silu_bwd_kernel<<<blocks_for(M * no, 256), 256, 0, st>>>(z, dfz, dfz, M * no);                              // dz in place
  PF_CHECK_LAUNCH("pf_train_gvp_bwd(silu)");
  // z = s Wf^T + bf
  if ((rc = sgemm_ld(dfz, s, nullptr, dWf, no, K, Mi, 1, no, K, 1, K, 0, wgrad_splits(M, no, K), stream, ws)) != PF_OK) return rc;       // dWf = dz^T s
  if ((rc = pf_train_colsum(dfz, dbf, M, no, ws.ptr, ws.bytes, stream)) != PF_OK) return rc;
  if ((rc = sgemm_ld(dfz, Wf, nullptr, ds, Mi, K, no, no, 1, K, 1, K, 0, 1, stream, ws)) != PF_OK) return rc;        // ds = dz Wf
  e = cudaMemcpy2DAsync(dfeats, (size_t)n * 4, ds, (size_t)K * 4, (size_t)n * 4, (size_t)M, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) {
    set_error("pf_train_gvp_bwd: copy: %s", cudaGetErrorString(e));
    return PF_ERR_LAUNCH;
  }
  T::vecnorm_bwd_kernel<<<blocks_for(M * h, 256), 256, 0, st>>>(Vh, ds + n, dVh, M, h, K);                        // via sh
  PF_CHECK_LAUNCH("pf_train_gvp_bwd(vecnorm)");
  // Vu = Vh Wu
  if ((rc = sgemm_ld(dVu, Wu, nullptr, dVh, M3, h, vo, vo, 1, 1, vo, h, 1, 1, stream, ws)) != PF_OK) return rc;      // dVh += dVu Wu^T
  if ((rc = sgemm_ld(Vh, dVu, nullptr, dWu, h, vo, M3, 1, h, vo, 1, vo, 0, wgrad_splits(3 * M, h, vo), stream, ws)) != PF_OK) return rc;    // dWu = Vh^T dVu
  // Vh = V Wh
  if ((rc = sgemm_ld(dVh, Wh, nullptr, dvec, M3, vi, h, h, 1, 1, h, vi, 0, 1, stream, ws)) != PF_OK) return rc;      // dV = dVh Wh^T
  if ((rc = sgemm_ld(vec, dVh, nullptr, dWh, vi, h, M3, 1, vi, h, 1, h, 0, wgrad_splits(3 * M, vi, h), stream, ws)) != PF_OK) return rc;    // dWh = V^T dVh
  return PF_OK;
}
