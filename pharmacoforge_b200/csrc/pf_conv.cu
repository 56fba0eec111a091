// K0 encoders, K3 fused edge conv, K4 node update, K5a noise head -- fp32 FFMA path.
#include "pf_gvp.cuh"

namespace pf {

// ------------------------------------------------------------------------------------------------
// K0: h = LayerNorm(SiLU(W [feats, t] + b)).  One CTA walks whole graphs (t is per graph), one warp per
// node, each lane owns 4 of the 128 outputs.
// ------------------------------------------------------------------------------------------------
constexpr int kEncMaxIn = 32;

// One encoder row: LN(SiLU(W [x, t] + b)) of one node by one warp (4 columns per lane); x[i] is read through `xin`.
template <typename XIn>
__device__ __forceinline__ float4 encode_row(const XIn& xin, int nf, float tg, const float* s_w, const float* s_p, int lane) {
  float4 z = *reinterpret_cast<const float4*>(s_p + 4 * lane);
  for (int i = 0; i <= nf; ++i) {
    const float xi = i < nf ? xin(i) : tg;
    const float4 wv = *reinterpret_cast<const float4*>(s_w + i * kHidden + 4 * lane);
    z.x = fmaf(xi, wv.x, z.x);
    z.y = fmaf(xi, wv.y, z.y);
    z.z = fmaf(xi, wv.z, z.z);
    z.w = fmaf(xi, wv.w, z.w);
  }
  z.x = silu_f(z.x);
  z.y = silu_f(z.y);
  z.z = silu_f(z.z);
  z.w = silu_f(z.w);
  const float mean = warp_sum(z.x + z.y + z.z + z.w) * (1.0f / kHidden);
  const float dx = z.x - mean, dy = z.y - mean, dz = z.z - mean, dw = z.w - mean;
  const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / kHidden);
  const float rstd = rsqrtf(var + 1e-5f);
  const float4 lw = *reinterpret_cast<const float4*>(s_p + kHidden + 4 * lane);
  const float4 lb = *reinterpret_cast<const float4*>(s_p + 2 * kHidden + 4 * lane);
  float4 o;
  o.x = dx * rstd * lw.x + lb.x;
  o.y = dy * rstd * lw.y + lb.y;
  o.z = dz * rstd * lw.z + lb.z;
  o.w = dw * rstd * lw.w + lb.w;
  return o;
}

// A one-hot node (the reference's element encoding, dev.yml:57) has only nf distinct encoder rows per graph and timestep:
// the CTA computes them once per graph with the SAME instruction sequence as the general path (fmaf(0, w, z) == z, so the
// rows are bit-identical to what the per-node loop produces) and a one-hot node then costs one 44-byte read and one
// 512-byte write instead of ~150 warp instructions; any other feature row takes the general path.
__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ feats, int nf,
                                                     const int* __restrict__ node_ptr, int n_graphs,
                                                     const float* __restrict__ t, const float* __restrict__ w,
                                                     float* __restrict__ h_out) {
  __shared__ __align__(16) float s_w[(kEncMaxIn + 1) * kHidden];
  __shared__ __align__(16) float s_p[3 * kHidden];
  __shared__ __align__(16) float s_tab[kEncMaxIn * kHidden];
  const int nin = nf + 1;
  for (int i = threadIdx.x; i < nin * kHidden; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) s_p[i] = w[nin * kHidden + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int n0 = node_ptr[g], n1 = node_ptr[g + 1];
    const float tg = t[g];
    const bool use_table = n1 - n0 > 2 * nf;  // small graphs (pharmacophores): not worth nf extra rows
    if (use_table) {
      for (int k = warp; k < nf; k += nwarps) {
        const float4 o = encode_row([&](int i) { return i == k ? 1.0f : 0.0f; }, nf, tg, s_w, s_p, lane);
        *reinterpret_cast<float4*>(s_tab + k * kHidden + 4 * lane) = o;
      }
      __syncthreads();
    }
    for (int n = n0 + warp; n < n1; n += nwarps) {
      const float* fr = feats + (size_t)n * nf;
      int hot = -1;
      if (use_table) {
        const float x = lane < nf ? __ldg(fr + lane) : 0.0f;
        const unsigned ones = __ballot_sync(0xffffffffu, x == 1.0f), nonzero = __ballot_sync(0xffffffffu, x != 0.0f);
        if (ones == nonzero && __popc(ones) == 1) hot = __ffs(ones) - 1;
      }
      float4 o;
      if (hot >= 0)
        o = *reinterpret_cast<const float4*>(s_tab + hot * kHidden + 4 * lane);
      else
        o = encode_row([&](int i) { return __ldg(fr + i); }, nf, tg, s_w, s_p, lane);
      *reinterpret_cast<float4*>(h_out + (size_t)n * kHidden + 4 * lane) = o;
    }
    if (use_table) __syncthreads();  // the table is rewritten for the next graph
  }
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up of a compute tile
// ------------------------------------------------------------------------------------------------
struct TileBufs {
  float *sA, *sB, *vA, *vB, *vH;
};
constexpr int kScalarBufFloats = kTileRows * kLds;
constexpr int kVectorBufFloats = kTileRows * kLdv;
constexpr size_t kTileSmemBytes = (2 * kScalarBufFloats + 3 * kVectorBufFloats) * sizeof(float);

__device__ __forceinline__ TileBufs carve(float* base) {
  TileBufs b;
  b.sA = base;
  b.sB = b.sA + kScalarBufFloats;
  b.vA = b.sB + kScalarBufFloats;
  b.vB = b.vA + kVectorBufFloats;
  b.vH = b.vB + kVectorBufFloats;
  return b;
}

// ------------------------------------------------------------------------------------------------
// K3: gather -> edge features -> GVP chain -> segmented mean.  Persistent CTAs stride over the tile list.
// ------------------------------------------------------------------------------------------------
struct EdgeConvParams {
  const float *src_h, *src_v, *src_x, *dst_x;
  const int *seg_start, *seg_cnt, *seg_dst, *col, *tiles, *n_tiles;
  const float* w;
  int n_gvps;
  float *agg_h, *agg_v;
  int accumulate;
};

__global__ void __launch_bounds__(kThreads, 1) edge_conv_kernel(const EdgeConvParams p) {
  extern __shared__ __align__(16) float smem[];
  TileBufs b = carve(smem);
  int* s_off = reinterpret_cast<int*>(b.vH + kVectorBufFloats);  // [kTileRows + 1]
  int* s_start = s_off + kTileRows + 1;                          // [kTileRows]
  int* s_dst = s_start + kTileRows;                              // [kTileRows]
  int* s_rowseg = s_dst + kTileRows;                             // [kTileRows]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = *p.n_tiles;
  const GvpLayout L0 = gvp_layout(kVec + 1, kVec, kHidden + kRbf, kHidden);
  const GvpLayout L1 = gvp_layout(kVec, kVec, kHidden, kHidden);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int s0 = p.tiles[2 * tile], s1 = p.tiles[2 * tile + 1];
    const int nseg = s1 - s0;
    if (tid < nseg) {
      s_start[tid] = p.seg_start[s0 + tid];
      s_dst[tid] = p.seg_dst ? p.seg_dst[s0 + tid] : s0 + tid;
    }
    if (warp == 0) {  // exclusive scan of the segment sizes (nseg <= 64)
      const int c0 = lane < nseg ? p.seg_cnt[s0 + lane] : 0;
      const int c1 = lane + 32 < nseg ? p.seg_cnt[s0 + lane + 32] : 0;
      int i0 = c0, i1 = c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, i0, o);
        const int t1 = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= o) {
          i0 += t0;
          i1 += t1;
        }
      }
      const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
      s_off[lane] = i0 - c0;
      s_off[lane + 32] = tot0 + i1 - c1;
      if (lane == 31) s_off[64] = tot0 + i1;
    }
    __syncthreads();
    const int nrows = s_off[nseg];
    if (tid < nseg) {
      for (int r = s_off[tid]; r < s_off[tid + 1]; ++r) s_rowseg[r] = tid;
    }
    __syncthreads();

    // ---- gather + edge features, one warp per row
    for (int r = warp; r < kTileRows; r += kThreads / 32) {
      float* sr = b.sA + r * kLds;
      float* vr = b.vA + r * kLdv;
      if (r < nrows) {
        const int j = s_rowseg[r];
        const int e = s_start[j] + (r - s_off[j]);
        const int src = __ldg(p.col + e);
        const int dst = s_dst[j];
        const float4 hv = __ldg(reinterpret_cast<const float4*>(p.src_h + (size_t)src * kHidden) + lane);
        *reinterpret_cast<float4*>(sr + 4 * lane) = hv;
        if (p.src_v != nullptr) {
          if (lane < 12) {
            const float4 vv = __ldg(reinterpret_cast<const float4*>(p.src_v + (size_t)src * kVRow) + lane);
            const int c = lane >> 2, u = (lane & 3) * 4;
            float* o = vr + c * (kVec + 1) + 1 + u;
            o[0] = vv.x;
            o[1] = vv.y;
            o[2] = vv.z;
            o[3] = vv.w;
          }
        } else {
          for (int k = lane; k < kVRow; k += 32) vr[(k >> 4) * (kVec + 1) + 1 + (k & 15)] = 0.f;
        }
        const float dx = __ldg(p.src_x + (size_t)src * 3 + 0) - __ldg(p.dst_x + (size_t)dst * 3 + 0);
        const float dy = __ldg(p.src_x + (size_t)src * 3 + 1) - __ldg(p.dst_x + (size_t)dst * 3 + 1);
        const float dz = __ldg(p.src_x + (size_t)src * 3 + 2) - __ldg(p.dst_x + (size_t)dst * 3 + 2);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float d = sqrtf(fmaxf(d2, 1e-8f)) + 1e-8f;
        if (lane < 3) vr[lane * (kVec + 1)] = (lane == 0 ? dx : (lane == 1 ? dy : dz)) / d;
        if (lane < kRbf) {
          const float z = (d - (float)lane) / 0.9375f;
          sr[kHidden + lane] = expf(-(z * z));
        }
      } else {
        for (int k = lane; k < kHidden + kRbf; k += 32) sr[k] = 0.f;
        for (int k = lane; k < kLdv; k += 32) vr[k] = 0.f;
      }
    }
    __syncthreads();

    // ---- message GVP chain (gvp.py:392-415): first GVP sees 17 vector / 144 scalar inputs
    gvp_tile<kVec + 1, kVec, kHidden + kRbf, kHidden, true>(p.w, b.sA, b.sB, b.vA, b.vH, b.vB);
    float *sCur = b.sB, *sNext = b.sA, *vCur = b.vB, *vNext = b.vA;
    for (int gi = 1; gi < p.n_gvps; ++gi) {
      gvp_tile<kVec, kVec, kHidden, kHidden, true>(p.w + L0.total + (gi - 1) * L1.total, sCur, sNext, vCur, b.vH,
                                                   vNext);
      float* ts = sCur;
      sCur = sNext;
      sNext = ts;
      float* tv = vCur;
      vCur = vNext;
      vNext = tv;
    }

    // ---- segmented mean over each destination's rows, in edge order (deterministic, no atomics)
    for (int j = warp; j < nseg; j += kThreads / 32) {
      const int r0 = s_off[j], r1 = s_off[j + 1];
      const int cnt = r1 - r0;
      if (p.accumulate && cnt == 0) continue;
      const int dst = s_dst[j];
      const float denom = (float)(cnt > 0 ? cnt : 1);
      float* oh = p.agg_h + (size_t)dst * kHidden;
#pragma unroll
      for (int k = lane; k < kHidden; k += 32) {
        float acc = 0.f;
        for (int r = r0; r < r1; ++r) acc += sCur[r * kLds + k];
        acc = acc / denom;
        oh[k] = p.accumulate ? oh[k] + acc : acc;
      }
      float* ov = p.agg_v + (size_t)dst * kVRow;
      for (int k = lane; k < kVRow; k += 32) {
        float acc = 0.f;
        for (int r = r0; r < r1; ++r) acc += vCur[r * kLdv + k];
        acc = acc / denom;
        ov[k] = p.accumulate ? ov[k] + acc : acc;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// GVPLayerNorm (gvp.py:159-166) of one row held by a warp: lane owns scalars 4*lane..4*lane+3 and vector
// entries lane and lane+32 (<48) of the [c][u] row.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void row_layernorm(float4& h, float& e0, float& e1, const float* __restrict__ lw,
                                              const float* __restrict__ lb, int lane) {
  const float mean = warp_sum(h.x + h.y + h.z + h.w) * (1.0f / kHidden);
  const float dx = h.x - mean, dy = h.y - mean, dz = h.z - mean, dw = h.w - mean;
  const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / kHidden);
  const float rstd = rsqrtf(var + 1e-5f);
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(lw) + lane);
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(lb) + lane);
  h.x = dx * rstd * w4.x + b4.x;
  h.y = dy * rstd * w4.y + b4.y;
  h.z = dz * rstd * w4.z + b4.z;
  h.w = dw * rstd * w4.w + b4.w;
  // vector channels: lane u<16 holds (c0,u) in e0 and (c2,u) in e1; lane 16+u holds (c1,u) in e0
  const float sq0 = e0 * e0;
  const float up = __shfl_down_sync(0xffffffffu, sq0, 16);
  float nrm = lane < 16 ? fmaxf(sq0 + up + e1 * e1, 1e-8f) : 0.f;
  nrm = warp_sum(nrm) * (1.0f / kVec);
  const float vn = sqrtf(nrm + 1e-5f) + 1e-5f;
  e0 = e0 / vn;
  e1 = e1 / vn;
}

// ------------------------------------------------------------------------------------------------
// K4: node update
// ------------------------------------------------------------------------------------------------
struct NodeUpdateParams {
  const float *h_in, *v_in, *agg_h, *agg_v;
  long long n_nodes;
  const float* w;
  int n_gvps;
  float *h_out, *v_out;
};
constexpr size_t kNodeSmemBytes = kTileSmemBytes + (kTileRows * kHidden + kTileRows * kVRow) * sizeof(float);

__global__ void __launch_bounds__(kThreads, 1) node_update_kernel(const NodeUpdateParams p) {
  extern __shared__ __align__(16) float smem[];
  TileBufs b = carve(smem);
  float* sRes = b.vH + kVectorBufFloats;        // [64][128] normalised scalars kept for the residual
  float* vRes = sRes + kTileRows * kHidden;     // [64][48]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GvpLayout L1 = gvp_layout(kVec, kVec, kHidden, kHidden);
  const float* ln = p.w;
  const float* wg = p.w + 4 * kHidden;
  const long long n_tiles = (p.n_nodes + kTileRows - 1) / kTileRows;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n0 = tile * kTileRows;
    for (int r = warp; r < kTileRows; r += kThreads / 32) {
      const long long n = n0 + r;
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
      float e0 = 0.f, e1 = 0.f;
      if (n < p.n_nodes) {
        h = *(reinterpret_cast<const float4*>(p.h_in + n * kHidden) + lane);  // plain load: h_out may alias h_in
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.agg_h + n * kHidden) + lane);
        h.x += a.x;
        h.y += a.y;
        h.z += a.z;
        h.w += a.w;
        e0 = __ldg(p.agg_v + n * kVRow + lane);
        if (lane < 16) e1 = __ldg(p.agg_v + n * kVRow + 32 + lane);
        if (p.v_in != nullptr) {
          e0 += p.v_in[n * kVRow + lane];
          if (lane < 16) e1 += p.v_in[n * kVRow + 32 + lane];
        }
        row_layernorm(h, e0, e1, ln, ln + kHidden, lane);
      }
      *reinterpret_cast<float4*>(b.sA + r * kLds + 4 * lane) = h;
      *reinterpret_cast<float4*>(sRes + r * kHidden + 4 * lane) = h;
      b.vA[r * kLdv + lane] = e0;
      vRes[r * kVRow + lane] = e0;
      if (lane < 16) {
        b.vA[r * kLdv + 32 + lane] = e1;
        vRes[r * kVRow + 32 + lane] = e1;
      }
    }
    __syncthreads();

    float *sCur = b.sA, *sNext = b.sB, *vCur = b.vA, *vNext = b.vB;
    for (int gi = 0; gi < p.n_gvps; ++gi) {
      gvp_tile<kVec, kVec, kHidden, kHidden, true>(wg + gi * L1.total, sCur, sNext, vCur, b.vH, vNext);
      float* ts = sCur;
      sCur = sNext;
      sNext = ts;
      float* tv = vCur;
      vCur = vNext;
      vNext = tv;
    }

    for (int r = warp; r < kTileRows; r += kThreads / 32) {
      const long long n = n0 + r;
      if (n >= p.n_nodes) continue;
      float4 h = *reinterpret_cast<const float4*>(sRes + r * kHidden + 4 * lane);
      const float4 d = *reinterpret_cast<const float4*>(sCur + r * kLds + 4 * lane);
      h.x += d.x;
      h.y += d.y;
      h.z += d.z;
      h.w += d.w;
      float e0 = vRes[r * kVRow + lane] + vCur[r * kLdv + lane];
      float e1 = lane < 16 ? vRes[r * kVRow + 32 + lane] + vCur[r * kLdv + 32 + lane] : 0.f;
      row_layernorm(h, e0, e1, ln + 2 * kHidden, ln + 3 * kHidden, lane);
      *reinterpret_cast<float4*>(p.h_out + n * kHidden + 4 * lane) = h;
      p.v_out[n * kVRow + lane] = e0;
      if (lane < 16) p.v_out[n * kVRow + 32 + lane] = e1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K5a: noise head
// ------------------------------------------------------------------------------------------------
struct NoiseHeadParams {
  const float *h, *v;
  long long n_nodes;
  const float* w;
  int n_gvps, n_out;
  float *eps_h, *eps_x;
};

__global__ void __launch_bounds__(kThreads, 1) noise_head_kernel(const NoiseHeadParams p) {
  extern __shared__ __align__(16) float smem[];
  TileBufs b = carve(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GvpLayout L1 = gvp_layout(kVec, kVec, kHidden, kHidden);
  const GvpLayout LL = gvp_layout(kVec, 1, kHidden, 64);
  const long long n_tiles = (p.n_nodes + kTileRows - 1) / kTileRows;
  const int n_out4 = round_up4(p.n_out);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n0 = tile * kTileRows;
    for (int r = warp; r < kTileRows; r += kThreads / 32) {
      const long long n = n0 + r;
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
      float e0 = 0.f, e1 = 0.f;
      if (n < p.n_nodes) {
        h = __ldg(reinterpret_cast<const float4*>(p.h + n * kHidden) + lane);
        e0 = __ldg(p.v + n * kVRow + lane);
        if (lane < 16) e1 = __ldg(p.v + n * kVRow + 32 + lane);
      }
      *reinterpret_cast<float4*>(b.sA + r * kLds + 4 * lane) = h;
      b.vA[r * kLdv + lane] = e0;
      if (lane < 16) b.vA[r * kLdv + 32 + lane] = e1;
    }
    __syncthreads();
    float *sCur = b.sA, *sNext = b.sB, *vCur = b.vA, *vNext = b.vB;
    for (int gi = 0; gi + 1 < p.n_gvps; ++gi) {
      gvp_tile<kVec, kVec, kHidden, kHidden, true>(p.w + gi * L1.total, sCur, sNext, vCur, b.vH, vNext);
      float* ts = sCur;
      sCur = sNext;
      sNext = ts;
      float* tv = vCur;
      vCur = vNext;
      vNext = tv;
    }
    const float* wl = p.w + (p.n_gvps - 1) * L1.total;
    gvp_tile<kVec, 1, kHidden, 64, false>(wl, sCur, sNext, vCur, b.vH, vNext);
    // Linear(64 -> n_out) on the scalars, single output vector channel is eps_x
    const float* wo = wl + LL.total;         // Wt[64][n_out4]
    const float* bo = wo + 64 * n_out4;      // b[n_out4]
    for (int idx = tid; idx < kTileRows * p.n_out; idx += kThreads) {
      const int r = idx / p.n_out, o = idx - r * p.n_out;
      const long long n = n0 + r;
      if (n >= p.n_nodes) continue;
      float acc = __ldg(bo + o);
      const float* fr = sNext + r * kLds;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) acc = fmaf(fr[k], __ldg(wo + k * n_out4 + o), acc);
      p.eps_h[n * p.n_out + o] = acc;
    }
    for (int idx = tid; idx < kTileRows * 3; idx += kThreads) {
      const int r = idx / 3, c = idx - r * 3;
      const long long n = n0 + r;
      if (n < p.n_nodes) p.eps_x[n * 3 + c] = vNext[r * kLdv + c];
    }
    __syncthreads();
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
    return PF_ERR_LAUNCH;
  }
  return PF_OK;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_encode(const float* feats, int32_t n_feats, const int32_t* node_ptr, int32_t n_graphs,
                         const float* t, const float* w, float* h_out, void* stream) {
  PF_CHECK_ARG(feats && node_ptr && t && w && h_out, "pf_encode: null pointer");
  PF_CHECK_ARG(n_feats >= 1 && n_feats <= kEncMaxIn, "pf_encode: n_feats out of range");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 8 * num_sms() ? n_graphs : 8 * num_sms();
  encode_kernel<<<grid, 256, 0, as_stream(stream)>>>(feats, n_feats, node_ptr, n_graphs, t, w, h_out);
  PF_CHECK_LAUNCH("pf_encode");
  return PF_OK;
}

extern "C" int pf_edge_conv(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                            const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                            const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                            const float* w, int32_t n_gvps, float* agg_h, float* agg_v, int32_t accumulate,
                            void* stream) {
  PF_CHECK_ARG(src_h && src_x && dst_x && seg_start && seg_cnt && col && tiles && n_tiles && w && agg_h && agg_v,
               "pf_edge_conv: null pointer");
  PF_CHECK_ARG(n_gvps >= 1, "pf_edge_conv: n_gvps < 1");
  if (max_tiles <= 0) return PF_OK;
  const size_t smem = kTileSmemBytes + (4 * kTileRows + 1) * sizeof(int);
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    int rc = set_smem(edge_conv_kernel, smem);
    if (rc) return rc;
    configured.done[dev_] = true;
  }
  EdgeConvParams p{src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles,
                   w,     n_gvps, agg_h, agg_v, accumulate};
  const int grid = max_tiles < num_sms() ? max_tiles : num_sms();
  edge_conv_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_edge_conv");
  return PF_OK;
}

extern "C" int pf_node_update(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v,
                              int64_t n_nodes, const float* w, int32_t n_gvps, float* h_out, float* v_out,
                              void* stream) {
  PF_CHECK_ARG(h_in && agg_h && agg_v && w && h_out && v_out, "pf_node_update: null pointer");
  PF_CHECK_ARG(n_gvps >= 1, "pf_node_update: n_gvps < 1");
  if (n_nodes <= 0) return PF_OK;
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    int rc = set_smem(node_update_kernel, kNodeSmemBytes);
    if (rc) return rc;
    configured.done[dev_] = true;
  }
  NodeUpdateParams p{h_in, v_in, agg_h, agg_v, (long long)n_nodes, w, n_gvps, h_out, v_out};
  const long long tiles = (n_nodes + kTileRows - 1) / kTileRows;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  node_update_kernel<<<grid, kThreads, kNodeSmemBytes, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_node_update");
  return PF_OK;
}

extern "C" int pf_noise_head(const float* h, const float* v, int64_t n_nodes, const float* w, int32_t n_gvps,
                             int32_t n_out, float* eps_h, float* eps_x, void* stream) {
  PF_CHECK_ARG(h && v && w && eps_h && eps_x, "pf_noise_head: null pointer");
  PF_CHECK_ARG(n_gvps >= 1 && n_out >= 1 && n_out <= 64, "pf_noise_head: bad n_gvps / n_out");
  if (n_nodes <= 0) return PF_OK;
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    int rc = set_smem(noise_head_kernel, kTileSmemBytes);
    if (rc) return rc;
    configured.done[dev_] = true;
  }
  NoiseHeadParams p{h, v, (long long)n_nodes, w, n_gvps, n_out, eps_h, eps_x};
  const long long tiles = (n_nodes + kTileRows - 1) / kTileRows;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  noise_head_kernel<<<grid, kThreads, kTileSmemBytes, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_noise_head");
  return PF_OK;
}
