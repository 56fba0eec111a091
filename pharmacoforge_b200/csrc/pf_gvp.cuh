// One GVP (reference gvp.py:89-116) applied to a 64-row tile held in shared memory, fp32 FFMA path.
//
// Layout of a tile: scalars row-major with stride kLds floats (features in columns [0,SI), the vector
// norms `sh` are appended in [SI, SI+VH), zero padding up to a multiple of 4); vectors row-major with an
// odd stride kLdv, element (c, u) of a row at c*V + u (component-major).  256 threads.
//
//   Vh[c][h] = sum_v V[c][v] Wh[v][h]          sh[h] = sqrt(max(sum_c Vh[c][h]^2, 1e-8))
//   Vu[c][u] = sum_h Vh[c][h] Wu[h][u]         f = SiLU(WfT^T [s, sh] + bf)
//   gate[u]  = WgT^T f + bg                    Vout[c][u] = act(gate[u]) * Vu[c][u]
#pragma once
#include "pf_common.cuh"

namespace pf {

constexpr int kTileRows = PF_TILE_ROWS;  // 64
constexpr int kThreads = 256;
constexpr int kLds = 164;  // >= 128 + 16 + 17, multiple of 4; 164 % 32 == 4 keeps 8-row float4 reads conflict-free
constexpr int kLdv = 51;   // 3 * 17, odd

__device__ __forceinline__ float f4c(const float4& a, int i) {
  return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.w));
}

template <int VI, int VO, int SI, int SO, bool SIGMOID>
__device__ __forceinline__ void gvp_tile(const float* __restrict__ w, float* __restrict__ sIn,
                                         float* __restrict__ sOut, const float* __restrict__ vIn,
                                         float* __restrict__ vH, float* __restrict__ vOut) {
  constexpr int VH = VI > VO ? VI : VO;
  constexpr int K = SI + VH;
  constexpr int K4 = (K + 3) & ~3;
  const GvpLayout L = gvp_layout(VI, VO, SI, SO);
  const int tid = threadIdx.x;

  // ---- vector hidden channels and their norms
  {
    const int r = tid & (kTileRows - 1);
    const int q = tid >> 6;
    const float* vr = vIn + r * kLdv;
    for (int h = q; h < VH; h += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int v = 0; v < VI; ++v) {
        const float wv = __ldg(w + L.wh + v * VH + h);
        a0 = fmaf(vr[v], wv, a0);
        a1 = fmaf(vr[VI + v], wv, a1);
        a2 = fmaf(vr[2 * VI + v], wv, a2);
      }
      vH[r * kLdv + h] = a0;
      vH[r * kLdv + VH + h] = a1;
      vH[r * kLdv + 2 * VH + h] = a2;
      sIn[r * kLds + SI + h] = sqrtf(fmaxf(a0 * a0 + a1 * a1 + a2 * a2, 1e-8f));
    }
    if (q == 0) {
#pragma unroll
      for (int k = K; k < K4; ++k) sIn[r * kLds + k] = 0.f;
    }
  }
  __syncthreads();

  // ---- vector outputs before gating
  {
    const int r = tid & (kTileRows - 1);
    const int q = tid >> 6;
    const float* hr = vH + r * kLdv;
    for (int u = q; u < VO; u += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int h = 0; h < VH; ++h) {
        const float wv = __ldg(w + L.wu + h * VO + u);
        a0 = fmaf(hr[h], wv, a0);
        a1 = fmaf(hr[VH + h], wv, a1);
        a2 = fmaf(hr[2 * VH + h], wv, a2);
      }
      vOut[r * kLdv + u] = a0;
      vOut[r * kLdv + VO + u] = a1;
      vOut[r * kLdv + 2 * VO + u] = a2;
    }
  }

  // ---- scalar GEMM [64 x K4] x [K4 x SO]: each thread 4 rows x 8 columns
  {
    const int cg = tid & 15;
    const int rg = tid >> 4;
    if (8 * cg < SO) {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const float* a_base = sIn + (4 * rg) * kLds;
      const float* w_base = w + L.wf + 8 * cg;
#pragma unroll 2
      for (int k = 0; k < K4; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a_base + i * kLds + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_base + (k + kk) * SO));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_base + (k + kk) * SO + 4));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = f4c(a[i], kk);
            acc[i][0] = fmaf(av, w0.x, acc[i][0]);
            acc[i][1] = fmaf(av, w0.y, acc[i][1]);
            acc[i][2] = fmaf(av, w0.z, acc[i][2]);
            acc[i][3] = fmaf(av, w0.w, acc[i][3]);
            acc[i][4] = fmaf(av, w1.x, acc[i][4]);
            acc[i][5] = fmaf(av, w1.y, acc[i][5]);
            acc[i][6] = fmaf(av, w1.z, acc[i][6]);
            acc[i][7] = fmaf(av, w1.w, acc[i][7]);
          }
        }
      }
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(w + L.bf + 8 * cg));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(w + L.bf + 8 * cg + 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 o0, o1;
        o0.x = silu_f(acc[i][0] + b0.x);
        o0.y = silu_f(acc[i][1] + b0.y);
        o0.z = silu_f(acc[i][2] + b0.z);
        o0.w = silu_f(acc[i][3] + b0.w);
        o1.x = silu_f(acc[i][4] + b1.x);
        o1.y = silu_f(acc[i][5] + b1.y);
        o1.z = silu_f(acc[i][6] + b1.z);
        o1.w = silu_f(acc[i][7] + b1.w);
        float* o = sOut + (4 * rg + i) * kLds + 8 * cg;
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + 4) = o1;
      }
    }
  }
  __syncthreads();

  // ---- gates from the new scalars, applied to the vector outputs
  {
    const int r = tid >> 2;
    const int ug = tid & 3;
    constexpr int NU = VO >= 4 ? 4 : VO;
    if (4 * ug < VO) {
      float g[NU];
#pragma unroll
      for (int j = 0; j < NU; ++j) g[j] = __ldg(w + L.bg + 4 * ug + j);
      const float* fr = sOut + r * kLds;
      const float* wg = w + L.wg + 4 * ug;
#pragma unroll 4
      for (int n = 0; n < SO; n += 4) {
        const float4 f = *reinterpret_cast<const float4*>(fr + n);
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          const float fv = f4c(f, nn);
          if constexpr (NU == 4) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wg + (n + nn) * VO));
            g[0] = fmaf(fv, wv.x, g[0]);
            g[1] = fmaf(fv, wv.y, g[1]);
            g[2] = fmaf(fv, wv.z, g[2]);
            g[3] = fmaf(fv, wv.w, g[3]);
          } else {
#pragma unroll
            for (int j = 0; j < NU; ++j) g[j] = fmaf(fv, __ldg(wg + (n + nn) * VO + j), g[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const float gate = SIGMOID ? sigmoid_f(g[j]) : g[j];
        const int u = 4 * ug + j;
        vOut[r * kLdv + u] *= gate;
        vOut[r * kLdv + VO + u] *= gate;
        vOut[r * kLdv + 2 * VO + u] *= gate;
      }
    }
  }
  __syncthreads();
}

}  // namespace pf
