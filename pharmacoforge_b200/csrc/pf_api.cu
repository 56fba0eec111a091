// Error state, weight layout query, K5b posterior/COM kernels and the host-side step / loop drivers.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <string.h>

#include "pf_common.cuh"

namespace pf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Process-wide state of the library is limited to DIAGNOSTICS (nothing on the compute path reads it): the launch counter
// (atomic), the per-site event recorder below (armed only by bench.py's kernel-timing pass, guarded by a mutex) and the
// device-timeline pointer of pf_tc_trace.  Every compute entry point is a pure function of its arguments.
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::mutex g_prof_mu;

// ---- per-site event timing: pairs are recorded on the launching stream and resolved in pf_profile_collect
struct ProfPair {
  cudaEvent_t a, b;
  int site;
};
static ProfPair* g_pairs = nullptr;
static int g_pair_cap = 0, g_pair_n = 0, g_pair_open[kNumSites];
static volatile bool g_prof_on = false;

void prof_begin(int site, cudaStream_t st) {
  if (!g_prof_on) return;   // the common case: one relaxed read, no lock
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_on || g_pair_n >= g_pair_cap) {
    if (g_prof_on) g_pair_open[site] = -1;
    return;
  }
  g_pairs[g_pair_n].site = site;
  cudaEventRecord(g_pairs[g_pair_n].a, st);
  g_pair_open[site] = g_pair_n++;
}
void prof_end(int site, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_on || g_pair_open[site] < 0) return;
  cudaEventRecord(g_pairs[g_pair_open[site]].b, st);
  g_pair_open[site] = -1;
}

// ------------------------------------------------------------------------------------------------
// K5b.  One CTA per graph.  Every operation is rounded separately in the reference's order
// (pharmacodiff.py:416-426: mu = z/alpha - var*eps; z_s = mu + sigma*noise; com = sum/count; z -= com), so
// that with identical eps the step is bit-identical to the CPU oracle.
// ------------------------------------------------------------------------------------------------
// ---- counter-based Gaussian noise for throughput runs ("in-kernel Philox allowed for throughput runs").
// Philox4x32-10 keyed by the 64-bit seed (read from DEVICE memory, so a captured CUDA graph replays with a new seed),
// counter = (element quad, step, stream, 0); Box-Muller on the four words.  Element e of stream s at step k is a pure
// function of (seed, e, s, k): the draw does not depend on the launch geometry.  Parity runs inject explicit noise.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned stream, unsigned step, unsigned long long e) {
  const uint4 w = philox4x32_10(make_uint4((unsigned)(e >> 2), (unsigned)(e >> 34), step, stream),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const unsigned a = (e & 2) ? w.z : w.x, b = (e & 2) ? w.w : w.y;
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return (e & 1) ? rad * sn : rad * cs;
}
constexpr unsigned kNoiseStreamX = 0, kNoiseStreamH = 1;

__global__ void philox_normal_kernel(float* __restrict__ out, long long n, const unsigned long long* __restrict__ seed,
                                     unsigned stream_id, unsigned step) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_normal(*seed, stream_id, step, (unsigned long long)i);
}

__global__ void __launch_bounds__(128) posterior_kernel(float* __restrict__ pharm_x, float* __restrict__ pharm_h,
                                                        int nh, const float* __restrict__ eps_x,
                                                        const float* __restrict__ eps_h,
                                                        const float* __restrict__ noise_x,
                                                        const float* __restrict__ noise_h,
                                                        const int* __restrict__ pharm_ptr,
                                                        float* __restrict__ prot_x, const int* __restrict__ prot_ptr,
                                                        int n_graphs, float alpha_ts, float var_terms,
                                                        float sigma_q, const unsigned long long* __restrict__ seed_dev,
                                                        unsigned noise_step, float ep_c1, float ep_c2, int ep_mode,
                                                        float* __restrict__ com_out) {
  // com_out != nullptr (the sampling loop): the graph's centre of mass goes to com_out[3 g ..] and the protein is shifted by
  // prot_shift_kernel afterwards -- a streaming pass with one warp per graph -- instead of by this CTA after its two
  // barriers; same arithmetic per element either way.
  const unsigned long long seed = noise_x == nullptr ? *seed_dev : 0ull;
  // The new coordinates of a graph are kept in shared memory between the update and the centre-of-mass pass (up to
  // kPostCache / 3 pharmacophore centres; larger graphs go through global memory as before): the three sequential sums
  // -- node order, as index_add_ does on the CPU -- were a chain of dependent global loads of just-written values.
  // The kernel stays latency-bound (one CTA per graph, ~25-28 % of the copy bandwidth, 0.1 % of a step).
  constexpr int kPostCache = 384;
  __shared__ float s_z[kPostCache];
  __shared__ float s_com[3];
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int fa = pharm_ptr[g], fb = pharm_ptr[g + 1];
    const int nf = fb - fa;
    const bool cached = nf * 3 <= kPostCache;
    for (int i = threadIdx.x; i < nf * 3; i += blockDim.x) {
      const size_t o = (size_t)fa * 3 + i;
      // eps parameterisation: z / alpha_ts - var_terms * eps; endpoint parameterisation (pharmacodiff.py:413-414, the
      // network output is the predicted x_0): c1 * z + c2 * pred -- every product and sum rounded, as the tensor ops do
      const float mu = (ep_mode & PF_EP_COORD) ? __fadd_rn(__fmul_rn(ep_c1, pharm_x[o]), __fmul_rn(ep_c2, eps_x[o]))
                                               : __fsub_rn(__fdiv_rn(pharm_x[o], alpha_ts), __fmul_rn(var_terms, eps_x[o]));
      const float nz = noise_x != nullptr ? noise_x[o] : philox_normal(seed, kNoiseStreamX, noise_step, o);
      const float z = __fadd_rn(mu, __fmul_rn(sigma_q, nz));
      if (cached)
        s_z[i] = z;
      else
        pharm_x[o] = z;
    }
    for (int i = threadIdx.x; i < nf * nh; i += blockDim.x) {
      const size_t o = (size_t)fa * nh + i;
      const float mu = (ep_mode & PF_EP_FEAT) ? __fadd_rn(__fmul_rn(ep_c1, pharm_h[o]), __fmul_rn(ep_c2, eps_h[o]))
                                              : __fsub_rn(__fdiv_rn(pharm_h[o], alpha_ts), __fmul_rn(var_terms, eps_h[o]));
      const float nz = noise_x != nullptr ? noise_h[o] : philox_normal(seed, kNoiseStreamH, noise_step, o);
      pharm_h[o] = __fadd_rn(mu, __fmul_rn(sigma_q, nz));
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      float acc = 0.f;
      if (cached) {
        for (int i = 0; i < nf; ++i) acc = __fadd_rn(acc, s_z[i * 3 + threadIdx.x]);
      } else {
        for (int i = 0; i < nf; ++i) acc = __fadd_rn(acc, pharm_x[(size_t)(fa + i) * 3 + threadIdx.x]);
      }
      s_com[threadIdx.x] = __fdiv_rn(acc, (float)nf);
    }
    __syncthreads();
    if (nf > 0) {
      for (int i = threadIdx.x; i < nf * 3; i += blockDim.x) {
        const size_t o = (size_t)fa * 3 + i;
        pharm_x[o] = __fsub_rn(cached ? s_z[i] : pharm_x[o], s_com[i % 3]);
      }
      if (com_out != nullptr) {
        if (threadIdx.x < 3) com_out[(size_t)g * 3 + threadIdx.x] = s_com[threadIdx.x];
      } else {
        const int pa = prot_ptr[g], pb = prot_ptr[g + 1];
        for (int i = threadIdx.x; i < (pb - pa) * 3; i += blockDim.x) {
          const size_t o = (size_t)pa * 3 + i;
          prot_x[o] = __fsub_rn(prot_x[o], s_com[i % 3]);
        }
      }
    }
    __syncthreads();
  }
}

// x[n] -= com[graph(n)] for the protein atoms of every graph that has pharmacophore nodes (com_removal, pharmacodiff.py:106-107):
// one WARP per graph, four independent 128-byte rows of loads in flight per lane, grid-stride over the graphs.
__global__ void __launch_bounds__(256) prot_shift_kernel(float* __restrict__ prot_x, const int* __restrict__ prot_ptr,
                                                         const int* __restrict__ pharm_ptr, int n_graphs,
                                                         const float* __restrict__ com) {
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_graphs; g += n_warps) {
    if (pharm_ptr[g + 1] == pharm_ptr[g]) continue;   // no pharmacophore nodes: nothing was removed
    const int pa = prot_ptr[g], n3 = (prot_ptr[g + 1] - pa) * 3;
    float* x = prot_x + (size_t)pa * 3;
    const float c0 = com[(size_t)g * 3], c1 = com[(size_t)g * 3 + 1], c2 = com[(size_t)g * 3 + 2];
    for (int i0 = lane; i0 < n3; i0 += 128) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = i0 + 32 * k < n3 ? x[i0 + 32 * k] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + 32 * k, m = i % 3;
        if (i < n3) x[i] = __fsub_rn(v[k], m == 0 ? c0 : (m == 1 ? c1 : c2));
      }
    }
  }
}

__global__ void __launch_bounds__(128) segment_mean3_kernel(const float* __restrict__ x, const int* __restrict__ ptr,
                                                            int n_graphs, float* __restrict__ com) {
  // sequential sum in node order, as index_add_ on CPU does (dgl.readout_nodes 'mean')
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_graphs * 3) return;
  const int g = t / 3, c = t - 3 * g;
  const int a = ptr[g], b = ptr[g + 1];
  float acc = 0.f;
  for (int i = a; i < b; ++i) acc = __fadd_rn(acc, x[(size_t)i * 3 + c]);
  com[t] = __fdiv_rn(acc, (float)(b - a));
}

__global__ void __launch_bounds__(128) segment_shift3_kernel(float* __restrict__ x, const int* __restrict__ ptr,
                                                             int n_graphs, const float* __restrict__ com,
                                                             float sign) {
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int a = ptr[g], b = ptr[g + 1];
    for (int i = threadIdx.x; i < (b - a) * 3; i += blockDim.x) {
      const size_t o = (size_t)a * 3 + i;
      x[o] = __fadd_rn(x[o], __fmul_rn(sign, com[g * 3 + i % 3]));
    }
  }
}

// rows seg_dst[s] of (h, v) <- 0 for every non-empty segment s of a segment list
__global__ void __launch_bounds__(256) zero_listed_rows_kernel(float* __restrict__ h, float* __restrict__ v,
                                                               const int* __restrict__ seg_cnt, const int* __restrict__ seg_dst,
                                                               long long n_seg) {
  const int lane = threadIdx.x & 31;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_seg; s += n_warps) {
    if (seg_cnt[s] == 0) continue;
    const long long d = seg_dst[s];
    *reinterpret_cast<float4*>(h + d * kHidden + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < kVRow / 4) *reinterpret_cast<float4*>(v + d * kVRow + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Numeric message_norm (gvp.py:386-389, 512-517: SUM over the in-edges divided by a constant instead of the mean): the edge
// kernels leave the per-destination MEAN of one edge type in tmp; agg[d] (+)= tmp[d] * count * inv_norm turns it into the scaled
// sum.  One warp per segment s (destination d = seg_dst ? seg_dst[s] : s); empty segments contribute zero.
__global__ void __launch_bounds__(256) scaled_accumulate_kernel(const float* __restrict__ tmp_h, const float* __restrict__ tmp_v,
                                                                const int* __restrict__ seg_cnt, const int* __restrict__ seg_dst,
                                                                long long n_seg, float inv_norm,
                                                                const float* __restrict__ inv_norm_node, float* __restrict__ agg_h,
                                                                float* __restrict__ agg_v, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_seg; s += n_warps) {
    const int cnt = seg_cnt[s];
    if (seg_dst != nullptr && cnt == 0) continue;   // an unused slot of a segment list: no destination
    const long long d = seg_dst != nullptr ? seg_dst[s] : s;
    const float sc = (float)cnt * (inv_norm_node != nullptr ? inv_norm_node[d] : inv_norm);
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), v = h;
    if (cnt > 0) {
      h = *reinterpret_cast<const float4*>(tmp_h + d * kHidden + 4 * lane);
      if (lane < kVRow / 4) v = *reinterpret_cast<const float4*>(tmp_v + d * kVRow + 4 * lane);
    }
    float4* oh = reinterpret_cast<float4*>(agg_h + d * kHidden + 4 * lane);
    float4 a = accumulate ? *oh : make_float4(0.f, 0.f, 0.f, 0.f);
    *oh = make_float4(fmaf(h.x, sc, a.x), fmaf(h.y, sc, a.y), fmaf(h.z, sc, a.z), fmaf(h.w, sc, a.w));
    if (lane < kVRow / 4) {
      float4* ov = reinterpret_cast<float4*>(agg_v + d * kVRow + 4 * lane);
      a = accumulate ? *ov : make_float4(0.f, 0.f, 0.f, 0.f);
      *ov = make_float4(fmaf(v.x, sc, a.x), fmaf(v.y, sc, a.y), fmaf(v.z, sc, a.z), fmaf(v.w, sc, a.w));
    }
  }
}

// message_norm = 0 (gvp.py:504-507): per graph, (edges of every type into the node type) / (nodes of the type) + 1; the
// reciprocal is written per node.  Edge counts per graph as add_pharm_edges records them (dynamics_gvp.py:219-221): ff and pp
// are the true counts; pf (and fp, which copies it) is counted through prot_batch_idx[pf_idxs[0]] -- protein atoms with radius
// edges (true counts), but PHARMACOPHORE node indices with kNN edges, so there the edges of pharmacophore node i go to the graph
// that owns protein atom i.  Reproduced as is (the oracle is pinned to the reference's own output).  One CTA per graph.
__global__ void __launch_bounds__(128) degree_norms_kernel(const int* __restrict__ prot_ptr, const int* __restrict__ pharm_ptr,
                                                           int n_graphs, int n_pharm, const int* __restrict__ ff_cnt,
                                                           const int* __restrict__ pf_cnt, const int* __restrict__ pp_cnt,
                                                           int radius_mode, float* __restrict__ inv_pharm,
                                                           float* __restrict__ inv_prot) {
  __shared__ int s_sum[3];
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int pa = prot_ptr[g], pb = prot_ptr[g + 1], fa = pharm_ptr[g], fb = pharm_ptr[g + 1];
    __syncthreads();
    if (threadIdx.x < 3) s_sum[threadIdx.x] = 0;
    __syncthreads();
    int e_ff = 0, e_pf = 0, e_pp = 0;
    for (int i = fa + threadIdx.x; i < fb; i += blockDim.x) {
      e_ff += ff_cnt[i];
      if (radius_mode) e_pf += pf_cnt[i];
    }
    if (!radius_mode)
      for (int i = pa + threadIdx.x; i < pb && i < n_pharm; i += blockDim.x) e_pf += pf_cnt[i];
    for (int c = pa + threadIdx.x; c < pb; c += blockDim.x) e_pp += pp_cnt[c];
    atomicAdd(&s_sum[0], e_ff);   // integer sums: order-independent
    atomicAdd(&s_sum[1], e_pf);
    atomicAdd(&s_sum[2], e_pp);
    __syncthreads();
    const float nf = __fadd_rn(__fdiv_rn((float)(s_sum[0] + s_sum[1]), (float)(fb - fa)), 1.0f);
    const float np_ = __fadd_rn(__fdiv_rn((float)(s_sum[1] + s_sum[2]), (float)(pb - pa)), 1.0f);
    for (int i = fa + threadIdx.x; i < fb; i += blockDim.x) inv_pharm[i] = __fdiv_rn(1.0f, nf);
    for (int c = pa + threadIdx.x; c < pb; c += blockDim.x) inv_prot[c] = __fdiv_rn(1.0f, np_);
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_abi_version(void) { return PF_ABI_VERSION; }
extern "C" const char* pf_last_error(void) { return g_err; }
extern "C" size_t pf_sample_args_size(void) { return sizeof(PfSampleArgs); }
extern "C" int64_t pf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int pf_profile_enable(int32_t max_pairs) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (int i = 0; i < g_pair_cap; ++i) {
    cudaEventDestroy(g_pairs[i].a);
    cudaEventDestroy(g_pairs[i].b);
  }
  delete[] g_pairs;
  g_pairs = nullptr;
  g_pair_cap = g_pair_n = 0;
  g_prof_on = max_pairs > 0;
  for (int i = 0; i < kNumSites; ++i) g_pair_open[i] = -1;
  if (!g_prof_on) return PF_OK;
  g_pairs = new ProfPair[max_pairs];
  for (int i = 0; i < max_pairs; ++i) {
    if (cudaEventCreate(&g_pairs[i].a) != cudaSuccess || cudaEventCreate(&g_pairs[i].b) != cudaSuccess) {
      set_error("pf_profile_enable: cudaEventCreate failed");
      return PF_ERR_LAUNCH;
    }
  }
  g_pair_cap = max_pairs;
  return PF_OK;
}

extern "C" int pf_profile_collect(double* total_ms_host, int32_t* count_host, int32_t n_sites) {
  PF_CHECK_ARG(total_ms_host && count_host && n_sites >= kNumSites, "pf_profile_collect: need >= 11 sites");
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (int i = 0; i < n_sites; ++i) {
    total_ms_host[i] = 0.0;
    count_host[i] = 0;
  }
  for (int i = 0; i < g_pair_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_pairs[i].a, g_pairs[i].b) == cudaSuccess) {
      total_ms_host[g_pairs[i].site] += ms;
      count_host[g_pairs[i].site] += 1;
    }
  }
  g_pair_n = 0;
  return PF_OK;
}

extern "C" int64_t pf_gvp_layout(int vi, int vo, int si, int so, int64_t offsets_out_host[6]) {
  const GvpLayout L = gvp_layout(vi, vo, si, so);
  if (offsets_out_host) {
    offsets_out_host[0] = L.wh;
    offsets_out_host[1] = L.wu;
    offsets_out_host[2] = L.wf;
    offsets_out_host[3] = L.bf;
    offsets_out_host[4] = L.wg;
    offsets_out_host[5] = L.bg;
  }
  return L.total;
}

static int posterior_launch(float* pharm_x, float* pharm_h, int32_t nh, const float* eps_x, const float* eps_h,
                            const float* noise_x, const float* noise_h, const int32_t* pharm_ptr, float* prot_x,
                            const int32_t* prot_ptr, int32_t n_graphs, float alpha_ts, float var_terms, float sigma_q,
                            const uint64_t* seed_dev, uint32_t noise_step, void* stream, float ep_c1 = 0.f,
                            float ep_c2 = 0.f, int ep_mode = 0, float* com_scratch = nullptr) {
  PF_CHECK_ARG(pharm_x && pharm_h && eps_x && eps_h && pharm_ptr && prot_x && prot_ptr, "pf_posterior_step: null pointer");
  PF_CHECK_ARG((noise_x && noise_h) || (!noise_x && !noise_h && seed_dev),
               "pf_posterior_step: pass both noise arrays, or neither and a device seed");
  if (n_graphs <= 0) return PF_OK;
  if (com_scratch != nullptr) {
    // two launches: a 64-thread CTA per graph for the pharmacophore update + centre of mass (at most 16 x 9 values per graph),
    // then the protein shift as a streaming pass
    const int grid = n_graphs < 64 * num_sms() ? n_graphs : 64 * num_sms();
    posterior_kernel<<<grid, 64, 0, as_stream(stream)>>>(pharm_x, pharm_h, nh, eps_x, eps_h, noise_x, noise_h, pharm_ptr,
                                                         prot_x, prot_ptr, n_graphs, alpha_ts, var_terms, sigma_q,
                                                         reinterpret_cast<const unsigned long long*>(seed_dev), noise_step,
                                                         ep_c1, ep_c2, ep_mode, com_scratch);
    PF_CHECK_LAUNCH("pf_posterior_step");
    const int warps = n_graphs, blocks = (warps + 7) / 8;
    prot_shift_kernel<<<blocks < 16 * num_sms() ? blocks : 16 * num_sms(), 256, 0, as_stream(stream)>>>(
        prot_x, prot_ptr, pharm_ptr, n_graphs, com_scratch);
    PF_CHECK_LAUNCH("pf_posterior_step(prot shift)");
    return PF_OK;
  }
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  posterior_kernel<<<grid, 128, 0, as_stream(stream)>>>(pharm_x, pharm_h, nh, eps_x, eps_h, noise_x, noise_h, pharm_ptr,
                                                        prot_x, prot_ptr, n_graphs, alpha_ts, var_terms, sigma_q,
                                                        reinterpret_cast<const unsigned long long*>(seed_dev), noise_step,
                                                        ep_c1, ep_c2, ep_mode, nullptr);
  PF_CHECK_LAUNCH("pf_posterior_step");
  return PF_OK;
}

extern "C" int pf_posterior_step(float* pharm_x, float* pharm_h, int32_t nh, const float* eps_x, const float* eps_h,
                                 const float* noise_x, const float* noise_h, const int32_t* pharm_ptr, float* prot_x,
                                 const int32_t* prot_ptr, int32_t n_graphs, float alpha_ts, float var_terms,
                                 float sigma_q, void* stream) {
  PF_CHECK_ARG(noise_x && noise_h, "pf_posterior_step: null noise (use pf_posterior_step_philox for in-kernel noise)");
  return posterior_launch(pharm_x, pharm_h, nh, eps_x, eps_h, noise_x, noise_h, pharm_ptr, prot_x, prot_ptr, n_graphs,
                          alpha_ts, var_terms, sigma_q, nullptr, 0, stream);
}

extern "C" int pf_posterior_step_philox(float* pharm_x, float* pharm_h, int32_t nh, const float* eps_x, const float* eps_h,
                                        const uint64_t* seed_dev, uint32_t noise_step, const int32_t* pharm_ptr,
                                        float* prot_x, const int32_t* prot_ptr, int32_t n_graphs, float alpha_ts,
                                        float var_terms, float sigma_q, void* stream) {
  return posterior_launch(pharm_x, pharm_h, nh, eps_x, eps_h, nullptr, nullptr, pharm_ptr, prot_x, prot_ptr, n_graphs,
                          alpha_ts, var_terms, sigma_q, seed_dev, noise_step, stream);
}

extern "C" int pf_posterior_step_ep(float* pharm_x, float* pharm_h, int32_t nh, const float* pred_x, const float* pred_h,
                                    const float* noise_x, const float* noise_h, const uint64_t* seed_dev,
                                    uint32_t noise_step, const int32_t* pharm_ptr, float* prot_x, const int32_t* prot_ptr,
                                    int32_t n_graphs, float alpha_ts, float var_terms, float sigma_q, float ep_c1,
                                    float ep_c2, int32_t ep_mode, void* stream) {
  PF_CHECK_ARG((ep_mode & ~(PF_EP_COORD | PF_EP_FEAT)) == 0, "pf_posterior_step_ep: unknown mode bits");
  return posterior_launch(pharm_x, pharm_h, nh, pred_x, pred_h, noise_x, noise_h, pharm_ptr, prot_x, prot_ptr, n_graphs,
                          alpha_ts, var_terms, sigma_q, seed_dev, noise_step, stream, ep_c1, ep_c2, ep_mode);
}

extern "C" int pf_philox_normal(float* out, int64_t n, const uint64_t* seed_dev, uint32_t stream_id, uint32_t step,
                                void* stream) {
  PF_CHECK_ARG(out && seed_dev && n >= 0, "pf_philox_normal: null pointer");
  if (n == 0) return PF_OK;
  philox_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      out, n, reinterpret_cast<const unsigned long long*>(seed_dev), stream_id, step);
  PF_CHECK_LAUNCH("pf_philox_normal");
  return PF_OK;
}

extern "C" int pf_segment_mean3(const float* x, const int32_t* ptr, int32_t n_graphs, float* com, void* stream) {
  PF_CHECK_ARG(x && ptr && com, "pf_segment_mean3: null pointer");
  if (n_graphs <= 0) return PF_OK;
  segment_mean3_kernel<<<(n_graphs * 3 + 127) / 128, 128, 0, as_stream(stream)>>>(x, ptr, n_graphs, com);
  PF_CHECK_LAUNCH("pf_segment_mean3");
  return PF_OK;
}

extern "C" int pf_segment_shift3(float* x, const int32_t* ptr, int32_t n_graphs, const float* com, float sign,
                                 void* stream) {
  PF_CHECK_ARG(x && ptr && com, "pf_segment_shift3: null pointer");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  segment_shift3_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, ptr, n_graphs, com, sign);
  PF_CHECK_LAUNCH("pf_segment_shift3");
  return PF_OK;
}

extern "C" int pf_fill_f32(float* p, int64_t n, float v, void* stream) {
  PF_CHECK_ARG(p && n >= 0, "pf_fill_f32: null pointer");
  if (n == 0) return PF_OK;
  fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p, n, v);
  PF_CHECK_LAUNCH("pf_fill_f32");
  return PF_OK;
}


extern "C" int pf_degree_norms(const int32_t* prot_ptr, const int32_t* pharm_ptr, int32_t n_graphs, int32_t n_pharm,
                               const int32_t* ff_cnt, const int32_t* pf_cnt, const int32_t* pp_cnt, int32_t radius_mode,
                               float* inv_norm_pharm, float* inv_norm_prot, void* stream) {
  PF_CHECK_ARG(prot_ptr && pharm_ptr && ff_cnt && pf_cnt && pp_cnt && inv_norm_pharm && inv_norm_prot,
               "pf_degree_norms: null pointer");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  degree_norms_kernel<<<grid, 128, 0, as_stream(stream)>>>(prot_ptr, pharm_ptr, n_graphs, n_pharm, ff_cnt, pf_cnt, pp_cnt,
                                                           radius_mode, inv_norm_pharm, inv_norm_prot);
  PF_CHECK_LAUNCH("pf_degree_norms");
  return PF_OK;
}

extern "C" int pf_scaled_accumulate(const float* tmp_h, const float* tmp_v, const int32_t* seg_cnt, const int32_t* seg_dst,
                                    int64_t n_seg, float inv_norm, const float* inv_norm_node, float* agg_h, float* agg_v,
                                    int32_t accumulate, void* stream) {
  PF_CHECK_ARG(tmp_h && tmp_v && seg_cnt && agg_h && agg_v, "pf_scaled_accumulate: null pointer");
  if (n_seg <= 0) return PF_OK;
  const long long blocks = (n_seg + 7) / 8;
  const int grid = (int)(blocks < 32LL * num_sms() ? blocks : 32LL * num_sms());
  scaled_accumulate_kernel<<<grid, 256, 0, as_stream(stream)>>>(tmp_h, tmp_v, seg_cnt, seg_dst, n_seg, inv_norm, inv_norm_node,
                                                                agg_h, agg_v, accumulate);
  PF_CHECK_LAUNCH("pf_scaled_accumulate");
  return PF_OK;
}

#define PF_TRY(call)       \
  do {                     \
    int rc_ = (call);      \
    if (rc_ != PF_OK) return rc_; \
  } while (0)

static int edge_conv_any(bool tc, bool f16, const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                         const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst, const int32_t* col,
                         const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles, const float* w,
                         const void* w_tc, int32_t n_gvps, float* agg_h, float* agg_v, int32_t accumulate,
                         void* stream) {
  if (tc)
    return (f16 ? pf_edge_conv_tc_f16 : pf_edge_conv_tc)(src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col,
                                                         tiles, n_tiles, max_tiles, w_tc, agg_h, agg_v, accumulate,
                                                         stream);
  return pf_edge_conv(src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles, max_tiles, w,
                      n_gvps, agg_h, agg_v, accumulate, stream);
}

static int node_update_any(bool f16, const void* w_tc, const float* h_in, const float* v_in, const float* agg_h,
                           const float* agg_v, int64_t n_nodes, const float* w, int32_t n_gvps, float* h_out,
                           float* v_out, void* stream) {
  if (w_tc != nullptr)
    return (f16 ? pf_node_update_tc_f16 : pf_node_update_tc)(h_in, v_in, agg_h, agg_v, n_nodes, w_tc, h_out, v_out,
                                                             stream);
  return pf_node_update(h_in, v_in, agg_h, agg_v, n_nodes, w, n_gvps, h_out, v_out, stream);
}

// The opt-in shared-pocket denoiser (PF_FLAG_SHARE_POCKET_MESSAGES, see pf_share.cu): same kernels as the nominal path, run
// on the distinct pockets (first-layer pp messages) and on the compact protein rows (everything else on the protein side).
static int denoiser_shared(const PfSampleArgs* a, void* stream) {
  const bool f16 = (a->flags & PF_FLAG_FP16_SINGLE_PASS) != 0;
  PF_CHECK_ARG(a->tile_rows == PF_TC_TILE_ROWS && a->n_convs == 2 && a->n_msg_gvps == 3 && a->n_upd_gvps == 2 &&
                   (a->flags & PF_FLAG_SKIP_DEAD_WORK),
               "pf_denoiser: PF_FLAG_SHARE_POCKET_MESSAGES needs the tcgen05 path, n_convs == 2 and PF_FLAG_SKIP_DEAD_WORK");
  PF_CHECK_ARG(a->seed_row && a->seed_table && a->pk_x && a->pk_start && a->pk_cnt && a->pk_col && a->pk_tiles && a->pk_n_tiles &&
                   a->pk_seed_row && a->pk_node0 && a->enc_feats && a->enc_ptr && a->enc_rep && a->enc_table && a->aggd_h &&
                   a->aggd_v && a->c_x && a->c_h && a->c_v && a->c_agg_h && a->c_agg_v && a->c_seg_id && a->pf_col_c,
               "pf_denoiser: incomplete shared-pocket buffers");
  const int64_t n_c = (int64_t)a->pf_k * a->n_pharm;
  const int n_rows = a->n_graphs * a->n_prot_feats;
  // encoder output per (graph, atom type) instead of per protein node; pharmacophore encoder as usual
  PF_TRY(pf_encode(a->pharm_h, a->n_pharm_feats, a->pharm_ptr, a->n_graphs, a->t_graph, a->w_pharm_enc, a->pharm_hh, stream));
  PF_TRY(pf_encode(a->enc_feats, a->n_prot_feats, a->enc_ptr, a->n_graphs, a->t_graph, a->w_prot_enc, a->enc_table, stream));
  // compact protein rows of this step: coordinates + encoder rows of the fp destinations, compact source ids of the pf edges
  PF_TRY(pf_share_index(a->pharm_ptr, a->n_graphs, a->pf_k, a->pf_cnt, a->pf_col, a->fp_seg_dst, a->fp_seg_cnt, a->pf_col_c, stream));
  PF_TRY(pf_share_gather(a->pharm_ptr, a->prot_ptr, a->pk_node0, a->n_graphs, a->pf_k, a->fp_seg_dst, a->fp_seg_cnt, a->prot_x, a->seed_row,
                         a->enc_table, nullptr, nullptr, a->c_x, a->c_h, nullptr, nullptr, 0, stream));
  auto conv = f16 ? pf_edge_conv_tc_f16 : pf_edge_conv_tc;
  auto upd = f16 ? pf_node_update_tc_f16 : pf_node_update_tc;
  // ---- layer 0.  pharm <- ff (store) + pf (accumulate, sources = compact rows)
  prof_begin(kSiteFF, as_stream(stream));
  PF_TRY(conv(a->pharm_hh, nullptr, a->pharm_x, a->pharm_x, a->ff_start, a->ff_cnt, nullptr, a->ff_col, a->ff_tiles,
              a->dyn_n_tiles + 0, a->dyn_max_tiles, a->w_msg_tc[0][0], a->pharm_agg_h, a->pharm_agg_v, 0, stream));
  prof_end(kSiteFF, as_stream(stream));
  prof_begin(kSitePF, as_stream(stream));
  PF_TRY(conv(a->c_h, nullptr, a->c_x, a->pharm_x, a->pf_start, a->pf_cnt, nullptr, a->pf_col_c, a->pf_tiles,
              a->dyn_n_tiles + 1, a->dyn_max_tiles, a->w_msg_tc[0][1], a->pharm_agg_h, a->pharm_agg_v, 1, stream));
  prof_end(kSitePF, as_stream(stream));
  // pp once per DISTINCT pocket (input coordinates: x_src - x_dst is translation invariant)
  prof_begin(kSitePP, as_stream(stream));
  PF_TRY(pf_seed_table(a->enc_table, a->enc_rep, n_rows, a->w_msg[0][3], a->seed_table, stream));
  PF_TRY(pf_edge_conv_tc_seeded(a->pk_seed_row, a->seed_table, a->pk_x, a->pk_x, a->pk_start, a->pk_cnt, nullptr, a->pk_col,
                                a->pk_tiles, a->pk_n_tiles, a->pk_max_tiles, a->w_msg_tc[0][3], a->aggd_h, a->aggd_v, 0,
                                f16 ? 1 : 0, stream));
  prof_end(kSitePP, as_stream(stream));
  // fp means land in the compact rows (segment slot s -> row s), then the shared pp means join them
  prof_begin(kSiteFP, as_stream(stream));
  PF_TRY(conv(a->pharm_hh, nullptr, a->pharm_x, a->c_x, a->fp_seg_start, a->fp_seg_cnt, a->c_seg_id, a->fp_col, a->fp_tiles,
              a->dyn_n_tiles + 2, a->dyn_max_tiles, a->w_msg_tc[0][2], a->c_agg_h, a->c_agg_v, 0, stream));
  PF_TRY(pf_share_gather(a->pharm_ptr, a->prot_ptr, a->pk_node0, a->n_graphs, a->pf_k, a->fp_seg_dst, a->fp_seg_cnt, nullptr, nullptr, nullptr,
                         a->aggd_h, a->aggd_v, nullptr, nullptr, a->c_agg_h, a->c_agg_v, 1, stream));
  prof_end(kSiteFP, as_stream(stream));
  prof_begin(kSiteUpdPharm, as_stream(stream));
  PF_TRY(upd(a->pharm_hh, nullptr, a->pharm_agg_h, a->pharm_agg_v, a->n_pharm, a->w_upd_tc[0][0], a->pharm_hh, a->pharm_v, stream));
  prof_end(kSiteUpdPharm, as_stream(stream));
  prof_begin(kSiteUpdProt, as_stream(stream));
  PF_TRY(upd(a->c_h, nullptr, a->c_agg_h, a->c_agg_v, n_c, a->w_upd_tc[0][1], a->c_h, a->c_v, stream));
  prof_end(kSiteUpdProt, as_stream(stream));
  // ---- layer 1 (last): pharmacophore side only
  prof_begin(kSiteFF, as_stream(stream));
  PF_TRY(conv(a->pharm_hh, a->pharm_v, a->pharm_x, a->pharm_x, a->ff_start, a->ff_cnt, nullptr, a->ff_col, a->ff_tiles,
              a->dyn_n_tiles + 0, a->dyn_max_tiles, a->w_msg_tc[1][0], a->pharm_agg_h, a->pharm_agg_v, 0, stream));
  prof_end(kSiteFF, as_stream(stream));
  prof_begin(kSitePF, as_stream(stream));
  PF_TRY(conv(a->c_h, a->c_v, a->c_x, a->pharm_x, a->pf_start, a->pf_cnt, nullptr, a->pf_col_c, a->pf_tiles, a->dyn_n_tiles + 1,
              a->dyn_max_tiles, a->w_msg_tc[1][1], a->pharm_agg_h, a->pharm_agg_v, 1, stream));
  prof_end(kSitePF, as_stream(stream));
  prof_begin(kSiteUpdPharm, as_stream(stream));
  PF_TRY(upd(a->pharm_hh, a->pharm_v, a->pharm_agg_h, a->pharm_agg_v, a->n_pharm, a->w_upd_tc[1][0], a->pharm_hh, a->pharm_v, stream));
  prof_end(kSiteUpdPharm, as_stream(stream));
  prof_begin(kSiteNoise, as_stream(stream));
  PF_TRY(pf_noise_head(a->pharm_hh, a->pharm_v, a->n_pharm, a->w_noise, a->n_noise_gvps, a->n_pharm_feats, a->eps_h, a->eps_x,
                       stream));
  prof_end(kSiteNoise, as_stream(stream));
  return PF_OK;
}

// One eps prediction: PharmRecDynamicsGVP.forward (dynamics_gvp.py:131-185) with a->t_graph already set.
extern "C" int pf_denoiser(const PfSampleArgs* a, void* stream) {
  PF_CHECK_ARG(a != nullptr, "pf_denoiser: null args");
  PF_CHECK_ARG(a->n_convs >= 1 && a->n_convs <= 8, "pf_denoiser: n_convs out of range (1..8)");
  // graph of this step: ff radius + pf kNN + fp reverse (dynamics_gvp.py:176-177)
  // pf_k == 0: pf / fp edges from radius(pharm, prot, r_pf) (dynamics_gvp.py:210-216).  A pharmacophore node then has more
  // in-edges than a tile holds: the pf edge kernels run on sub-segments of at most tile_rows edges and write one mean per
  // sub-segment to sub_agg_*, pf_combine_subsegments folds them into the node's aggregate; fp has one segment per protein atom.
  const bool radius = a->pf_k == 0;
  PF_CHECK_ARG(!radius || (a->pf_sub_ptr && a->fp_base && a->pf_sub_start && a->pf_sub_cnt && a->pf_sub_chunk_ptr &&
                           a->sub_agg_h && a->sub_agg_v && a->pf_sub_x && a->pf_r > 0.f && a->pf_max_nbrs >= 1),
               "pf_denoiser: pf_k == 0 needs the radius-graph buffers (pf_sub_*, fp_base, sub_agg_*, pf_r, pf_max_nbrs)");
  PF_CHECK_ARG(!radius || !(a->flags & PF_FLAG_SHARE_POCKET_MESSAGES), "pf_denoiser: the shared-pocket mode is built for pf_k >= 1");
  prof_begin(kSiteGraph, as_stream(stream));
  if (radius)
    PF_TRY(pf_dyn_graph_radius(a->prot_x, a->prot_ptr, a->pharm_x, a->pharm_ptr, a->n_graphs, a->ff_r, a->ff_max_nbrs, a->ff_k,
                               a->pf_r, a->pf_max_nbrs, a->tile_rows, a->ff_start, a->ff_cnt, a->ff_col, a->pf_start,
                               a->pf_sub_ptr, a->fp_base, a->pf_cnt, a->pf_col, a->pf_sub_start, a->pf_sub_cnt, a->pf_sub_x,
                               a->fp_seg_start,
                               a->fp_seg_cnt, a->fp_col, a->dev_status, stream));
  else
  PF_TRY(pf_dyn_graph_ffk(a->prot_x, a->prot_ptr, a->pharm_x, a->pharm_ptr, a->n_graphs, a->ff_r, a->ff_max_nbrs, a->ff_k, a->pf_k,
                      a->ff_start, a->ff_cnt, a->ff_col, a->pf_cnt, a->pf_col, a->fp_seg_dst, a->fp_seg_start,
                      a->fp_seg_cnt, a->fp_col, a->dev_status, stream));
  prof_end(kSiteGraph, as_stream(stream));
  const int32_t* const pf_seg_start = radius ? a->pf_sub_start : a->pf_start;   // what the pf edge kernels run on
  const int32_t* const pf_seg_cnt = radius ? a->pf_sub_cnt : a->pf_cnt;
  const float* const pf_dst_x = radius ? a->pf_sub_x : a->pharm_x;              // destination coordinates, indexed like the output rows
  const int32_t* const fp_dst = radius ? nullptr : a->fp_seg_dst;               // radius: one fp segment per protein atom
  const int64_t n_fp_seg = radius ? (int64_t)a->n_prot : (int64_t)a->pf_k * a->n_pharm;
  PF_TRY(pf_zero_i32(a->dyn_n_tiles, 3, stream));
  PF_CHECK_ARG(a->tile_rows == PF_TILE_ROWS || a->tile_rows == PF_TC_TILE_ROWS, "pf_denoiser: tile_rows must be 64 or 128");
  const bool tc = a->tile_rows == PF_TC_TILE_ROWS;
  const bool f16 = (a->flags & PF_FLAG_FP16_SINGLE_PASS) != 0;
  PF_CHECK_ARG(!f16 || tc, "pf_denoiser: PF_FLAG_FP16_SINGLE_PASS needs the tcgen05 path (tile_rows = 128)");
  if (tc) {
    PF_CHECK_ARG(a->n_msg_gvps == 3, "pf_denoiser: the tcgen05 message kernel is built for n_message_gvps == 3");
    for (int l = 0; l < a->n_convs; ++l)
      for (int e = 0; e < 4; ++e) PF_CHECK_ARG(a->w_msg_tc[l][e] != nullptr, "pf_denoiser: missing tcgen05 weight blob");
  }
  {   // the per-step tile plans of ff, pf (or its sub-segments) and fp in one launch
    const int32_t* const cnts[3] = {a->ff_cnt, pf_seg_cnt, a->fp_seg_cnt};
    const int32_t* const chunks[3] = {a->pharm_chunk_ptr, radius ? a->pf_sub_chunk_ptr : a->pharm_chunk_ptr, a->fp_chunk_ptr};
    const int32_t n_chunks[3] = {a->n_pharm_chunks, radius ? a->n_pf_sub_chunks : a->n_pharm_chunks, a->n_fp_chunks};
    const int32_t skip[3] = {0, 1, 1};
    int32_t* const tiles[3] = {a->ff_tiles, a->pf_tiles, a->fp_tiles};
    PF_TRY(pf_plan_tiles3(cnts, chunks, n_chunks, skip, a->tile_rows, tiles, a->dyn_max_tiles, a->dyn_n_tiles, a->dev_status,
                          stream));
  }
  // numeric message_norm: every edge type's means go to tmp_agg_* and are folded into the aggregate as count / norm * mean
  // message_norm = 0 (msg_norm_degree): the divisor is per graph, edges per node + 1 (pf_degree_norms), read per destination
  const bool norm0 = a->msg_norm_degree != 0;
  const bool summode = norm0 || a->msg_norm_pharm > 0.f || a->msg_norm_prot > 0.f;
  PF_CHECK_ARG(!summode || ((norm0 || (a->msg_norm_pharm > 0.f && a->msg_norm_prot > 0.f)) && a->tmp_agg_h && a->tmp_agg_v),
               "pf_denoiser: numeric message_norm needs both norms and the tmp_agg buffers");
  PF_CHECK_ARG(!norm0 || (a->inv_norm_pharm && a->inv_norm_prot), "pf_denoiser: message_norm = 0 needs the inv_norm_* arrays");
  if (norm0)
    PF_TRY(pf_degree_norms(a->prot_ptr, a->pharm_ptr, a->n_graphs, a->n_pharm, a->ff_cnt, a->pf_cnt, a->pp_cnt, radius ? 1 : 0,
                           a->inv_norm_pharm, a->inv_norm_prot, stream));
  const float inv_f = summode && !norm0 ? 1.0f / a->msg_norm_pharm : 0.f, inv_p = summode && !norm0 ? 1.0f / a->msg_norm_prot : 0.f;
  const float* const node_f = norm0 ? a->inv_norm_pharm : nullptr;
  const float* const node_p = norm0 ? a->inv_norm_prot : nullptr;
  PF_CHECK_ARG(!summode || !(a->flags & PF_FLAG_SHARE_POCKET_MESSAGES), "pf_denoiser: the shared-pocket mode is built for message_norm = 'mean'");
  float* const fagg_h = summode ? a->tmp_agg_h : a->pharm_agg_h;   // where the pharm-side / prot-side edge kernels write
  float* const fagg_v = summode ? a->tmp_agg_v : a->pharm_agg_v;
  float* const pagg_h = summode ? a->tmp_agg_h : a->prot_agg_h;
  float* const pagg_v = summode ? a->tmp_agg_v : a->prot_agg_v;
  const int acc1 = summode ? 0 : 1;                                // second edge type of a node type: accumulate (mean mode)
  float* const pf_out_h = radius ? a->sub_agg_h : fagg_h;          // where the pf edge kernels write, and how
  float* const pf_out_v = radius ? a->sub_agg_v : fagg_v;
  const int pf_acc = radius ? 0 : acc1;
  if (a->flags & PF_FLAG_SHARE_POCKET_MESSAGES) return denoiser_shared(a, stream);
  // encoders (dynamics_gvp.py:143-151); node vectors start at zero (:162-173) and are never materialised
  PF_TRY(pf_encode(a->pharm_h, a->n_pharm_feats, a->pharm_ptr, a->n_graphs, a->t_graph, a->w_pharm_enc, a->pharm_hh,
                   stream));
  // First-layer encoder table (see PF_FLAG_NO_LAYER0_TABLE): one encoder row per (graph, atom type) instead of one per node
  const bool table0 = tc && a->seed_row != nullptr && a->enc_feats && a->enc_ptr && a->enc_rep && a->enc_table &&
                      a->n_upd_gvps == 2 && a->w_upd_tc[0][1] != nullptr &&
                      !(a->flags & (PF_FLAG_NO_LAYER0_SEED | PF_FLAG_NO_LAYER0_TABLE));
  if (table0)
    PF_TRY(pf_encode(a->enc_feats, a->n_prot_feats, a->enc_ptr, a->n_graphs, a->t_graph, a->w_prot_enc, a->enc_table, stream));
  else
    PF_TRY(pf_encode(a->prot_feats, a->n_prot_feats, a->prot_ptr, a->n_graphs, a->t_graph, a->w_prot_enc, a->prot_h,
                     stream));
  for (int l = 0; l < a->n_convs; ++l) {
    const float* fv = l == 0 ? nullptr : a->pharm_v;
    const float* pv = l == 0 ? nullptr : a->prot_v;
    // pharm <- ff (store) + pf (accumulate); prot <- pp (store) + fp (accumulate)  (gvp.py:484-497)
    prof_begin(kSiteFF, as_stream(stream));
    PF_TRY(edge_conv_any(tc, f16, a->pharm_hh, fv, a->pharm_x, a->pharm_x, a->ff_start, a->ff_cnt, nullptr, a->ff_col,
                         a->ff_tiles, a->dyn_n_tiles + 0, a->dyn_max_tiles, a->w_msg[l][0], a->w_msg_tc[l][0],
                         a->n_msg_gvps, fagg_h, fagg_v, 0, stream));
    if (summode)
      PF_TRY(pf_scaled_accumulate(fagg_h, fagg_v, a->ff_cnt, nullptr, a->n_pharm, inv_f, node_f, a->pharm_agg_h,
                                  a->pharm_agg_v, 0, stream));
  prof_end(kSiteFF, as_stream(stream));
    prof_begin(kSitePF, as_stream(stream));
    if (l == 0 && table0)
      PF_TRY(pf_edge_conv_tc_mapped(a->enc_table, a->seed_row, nullptr, a->prot_x, pf_dst_x, pf_seg_start, pf_seg_cnt, nullptr,
                                    a->pf_col, a->pf_tiles, a->dyn_n_tiles + 1, a->dyn_max_tiles, a->w_msg_tc[l][1],
                                    pf_out_h, pf_out_v, pf_acc, f16 ? 1 : 0, stream));
    else
    PF_TRY(edge_conv_any(tc, f16, a->prot_h, pv, a->prot_x, pf_dst_x, pf_seg_start, pf_seg_cnt, nullptr, a->pf_col,
                         a->pf_tiles, a->dyn_n_tiles + 1, a->dyn_max_tiles, a->w_msg[l][1], a->w_msg_tc[l][1],
                         a->n_msg_gvps, pf_out_h, pf_out_v, pf_acc, stream));
    if (radius)   // mean over ALL in-edges of the node (or SUM / norm) from the sub-segment means, added to the ff aggregate
      PF_TRY(pf_combine_subsegments(a->sub_agg_h, a->sub_agg_v, a->pf_sub_cnt, a->pf_sub_ptr, a->pf_cnt, a->n_pharm, inv_f,
                                    node_f, a->pharm_agg_h, a->pharm_agg_v, 1, stream));
    else if (summode)
      PF_TRY(pf_scaled_accumulate(fagg_h, fagg_v, a->pf_cnt, nullptr, a->n_pharm, inv_f, node_f, a->pharm_agg_h,
                                  a->pharm_agg_v, 1, stream));
  prof_end(kSitePF, as_stream(stream));
    // exact dead-work elimination (opt-in): nothing reads the protein side of the last layer
    const bool prot_side = !((a->flags & PF_FLAG_SKIP_DEAD_WORK) && l == a->n_convs - 1);
    if (prot_side) {
    prof_begin(kSitePP, as_stream(stream));
    if (l == 0 && tc && a->seed_row != nullptr && !(a->flags & PF_FLAG_NO_LAYER0_SEED)) {
      // first layer: the per-node part of GVP 0 from the (graph, atom type) table, the rest per edge (timed with the site)
      PF_CHECK_ARG(a->seed_rep && a->seed_table && a->n_seed_rows > 0, "pf_denoiser: incomplete seed arrays");
      if (table0)   // table row r is its own representative
        PF_TRY(pf_seed_table(a->enc_table, a->enc_rep, a->n_seed_rows, a->w_msg[l][3], a->seed_table, stream));
      else
      PF_TRY(pf_seed_table(a->prot_h, a->seed_rep, a->n_seed_rows, a->w_msg[l][3], a->seed_table, stream));
      PF_TRY(pf_edge_conv_tc_seeded(a->seed_row, a->seed_table, a->prot_x, a->prot_x, a->pp_start, a->pp_cnt, nullptr,
                                    a->pp_col, a->pp_tiles, a->pp_n_tiles, a->pp_max_tiles, a->w_msg_tc[l][3],
                                    pagg_h, pagg_v, 0, f16 ? 1 : 0, stream));
    } else {
    PF_TRY(edge_conv_any(tc, f16, a->prot_h, pv, a->prot_x, a->prot_x, a->pp_start, a->pp_cnt, nullptr, a->pp_col,
                         a->pp_tiles, a->pp_n_tiles, a->pp_max_tiles, a->w_msg[l][3], a->w_msg_tc[l][3],
                         a->n_msg_gvps, pagg_h, pagg_v, 0, stream));
    }
    if (summode)
      PF_TRY(pf_scaled_accumulate(pagg_h, pagg_v, a->pp_cnt, nullptr, a->n_prot, inv_p, node_p, a->prot_agg_h,
                                  a->prot_agg_v, 0, stream));
  prof_end(kSitePP, as_stream(stream));
    prof_begin(kSiteFP, as_stream(stream));
    if (summode) {
      // the fp segment list has unused slots (count 0, destination = the graph's first atom): in store mode the edge kernel
      // would write zeros through them, so the listed rows of tmp are cleared and the kernel accumulates as always
      const long long n_slots = (long long)a->pf_k * a->n_pharm;
      if (radius) {   // identity destinations: every row of tmp is a destination
        if (cudaMemsetAsync(pagg_h, 0, (size_t)a->n_prot * kHidden * sizeof(float), as_stream(stream)) != cudaSuccess ||
            cudaMemsetAsync(pagg_v, 0, (size_t)a->n_prot * kVRow * sizeof(float), as_stream(stream)) != cudaSuccess) {
          set_error("pf_denoiser: cudaMemsetAsync failed");
          return PF_ERR_LAUNCH;
        }
      } else if (n_slots > 0) {
        const long long blocks = (n_slots + 7) / 8;
        zero_listed_rows_kernel<<<(int)(blocks < 32LL * num_sms() ? blocks : 32LL * num_sms()), 256, 0, as_stream(stream)>>>(
            pagg_h, pagg_v, a->fp_seg_cnt, a->fp_seg_dst, n_slots);
        PF_CHECK_LAUNCH("pf_denoiser(zero fp rows)");
      }
    }
    PF_TRY(edge_conv_any(tc, f16, a->pharm_hh, fv, a->pharm_x, a->prot_x, a->fp_seg_start, a->fp_seg_cnt, fp_dst,
                         a->fp_col, a->fp_tiles, a->dyn_n_tiles + 2, a->dyn_max_tiles, a->w_msg[l][2],
                         a->w_msg_tc[l][2], a->n_msg_gvps, pagg_h, pagg_v, 1, stream));
    if (summode)
      PF_TRY(pf_scaled_accumulate(pagg_h, pagg_v, a->fp_seg_cnt, fp_dst, n_fp_seg, inv_p, node_p, a->prot_agg_h,
                                  a->prot_agg_v, 1, stream));
  prof_end(kSiteFP, as_stream(stream));
    }
    // node updates, in place (gvp.py:501-536)
    prof_begin(kSiteUpdPharm, as_stream(stream));
    PF_TRY(node_update_any(f16, tc && a->n_upd_gvps == 2 ? a->w_upd_tc[l][0] : nullptr, a->pharm_hh, fv, a->pharm_agg_h,
                           a->pharm_agg_v, a->n_pharm, a->w_upd[l][0], a->n_upd_gvps, a->pharm_hh, a->pharm_v, stream));
  prof_end(kSiteUpdPharm, as_stream(stream));
    if (prot_side) {
    prof_begin(kSiteUpdProt, as_stream(stream));
    if (l == 0 && table0)
      PF_TRY(pf_node_update_tc_mapped(a->enc_table, a->seed_row, nullptr, a->prot_agg_h, a->prot_agg_v, a->n_prot,
                                      a->w_upd_tc[l][1], a->prot_h, a->prot_v, f16 ? 1 : 0, stream));
    else
    PF_TRY(node_update_any(f16, tc && a->n_upd_gvps == 2 ? a->w_upd_tc[l][1] : nullptr, a->prot_h, pv, a->prot_agg_h,
                           a->prot_agg_v, a->n_prot, a->w_upd[l][1], a->n_upd_gvps, a->prot_h, a->prot_v, stream));
  prof_end(kSiteUpdProt, as_stream(stream));
    }
  }
  prof_begin(kSiteNoise, as_stream(stream));
  PF_TRY(pf_noise_head(a->pharm_hh, a->pharm_v, a->n_pharm, a->w_noise, a->n_noise_gvps, a->n_pharm_feats, a->eps_h,
                       a->eps_x, stream));
  prof_end(kSiteNoise, as_stream(stream));
  return PF_OK;
}

// sample_given_receptor's loop (pharmacodiff.py:466-472).
extern "C" int pf_sample_loop(const PfSampleArgs* a, void* stream) {
  PF_CHECK_ARG(a != nullptr, "pf_sample_loop: null args");
  PF_CHECK_ARG(a->t_host && a->alpha_ts_host && a->var_terms_host && a->sigma_q_host, "pf_sample_loop: missing schedule");
  const bool philox = a->noise_x == nullptr;
  PF_CHECK_ARG(philox ? (a->noise_h == nullptr && a->noise_seed != nullptr) : a->noise_h != nullptr,
               "pf_sample_loop: pass noise_x and noise_h, or neither and noise_seed (in-kernel Philox)");
  PF_CHECK_ARG(a->ep_mode == 0 || (a->ep_c1_host && a->ep_c2_host), "pf_sample_loop: endpoint mode without its coefficient tables");
  const size_t fx = (size_t)a->n_pharm * 3, fh = (size_t)a->n_pharm * a->n_pharm_feats;
  for (int i = 0; i < a->n_steps; ++i) {
    PF_TRY(pf_fill_f32(a->t_graph, a->n_graphs, a->t_host[i], stream));
    PF_TRY(pf_denoiser(a, stream));
    prof_begin(kSitePosterior, as_stream(stream));
  PF_TRY(posterior_launch(a->pharm_x, a->pharm_h, a->n_pharm_feats, a->eps_x, a->eps_h,
                            philox ? nullptr : a->noise_x + i * fx, philox ? nullptr : a->noise_h + i * fh, a->pharm_ptr,
                            a->prot_x, a->prot_ptr, a->n_graphs, a->alpha_ts_host[i], a->var_terms_host[i],
                            a->sigma_q_host[i], a->noise_seed, (uint32_t)(a->noise_step0 + i), stream,
                            a->ep_mode ? a->ep_c1_host[i] : 0.f, a->ep_mode ? a->ep_c2_host[i] : 0.f, a->ep_mode,
                            // per-graph centres of mass in a buffer that is dead between two denoiser calls
                            (a->prot_agg_v != nullptr && (int64_t)a->n_prot * 16 >= a->n_graphs) ? a->prot_agg_v : nullptr));
  prof_end(kSitePosterior, as_stream(stream));
  }
  return PF_OK;
}
