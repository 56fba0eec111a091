// Opt-in exact work elimination for SAMPLING (SURVEY.md hard part 5b + 5c; never the nominal / headline path).
//
// With n_convs = 2 and the same timestep for every graph of a batch (the reverse-diffusion loop), three facts hold:
//   (5a) nothing reads the protein side of the last conv layer (dynamics_gvp.py:84-92);
//   (5b) the first layer's pp messages depend only on (pocket, t): protein scalars are enc(one-hot, t), protein vectors are
//        zero and x_src - x_dst is translation invariant -- they are identical for all samples of a pocket;
//   (5c) the only protein rows anything reads after the first layer are the <= k per pharmacophore centre that the pf edges
//        of the last layer gather, i.e. exactly the destinations of the fp segments.
// The denoiser then (PF_FLAG_SHARE_POCKET_MESSAGES, pf_api.cu) runs the seeded pp kernel once per DISTINCT pocket, and
// updates only a COMPACT set of protein rows: one row per fp segment slot s (slot s of graph g = protein node
// fp_seg_dst[s]).  The two kernels here build that compact view; the message / update kernels are the nominal ones, called
// on the compact arrays.  Results equal the nominal path up to the fp32 rounding of x_src - x_dst (the pocket's input
// coordinates instead of each copy's shifted frame): ~1e-7, tested at 2e-5 against the nominal kernels.
#include "pf_common.cuh"

namespace pf {

// pf edge (pharm node i, slot j) has source protein node p = pf_col[k i + j]; its row in the compact arrays is the fp
// segment slot of graph g whose destination is p.  One CTA per graph (<= k nf <= 2048 slots).
__global__ void __launch_bounds__(128) share_index_kernel(const int* __restrict__ pharm_ptr, int n_graphs, int k,
                                                          const int* __restrict__ pf_cnt, const int* __restrict__ pf_col,
                                                          const int* __restrict__ fp_seg_dst, const int* __restrict__ fp_seg_cnt,
                                                          int* __restrict__ pf_col_c) {
  extern __shared__ int s_dst[];   // destinations of the graph's live segment slots, -1 for empty slots
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int fa = pharm_ptr[g], nf = pharm_ptr[g + 1] - fa;
    const int base = k * fa, nslot = k * nf;
    __syncthreads();
    for (int s = threadIdx.x; s < nslot; s += blockDim.x) s_dst[s] = fp_seg_cnt[base + s] > 0 ? fp_seg_dst[base + s] : -1;
    __syncthreads();
    for (int e = threadIdx.x; e < nslot; e += blockDim.x) {
      const int i = e / k, j = e - i * k;
      int out = base;   // unused pf slots point at a valid row
      if (j < pf_cnt[fa + i]) {
        const int p = pf_col[base + e];
        for (int s = 0; s < nslot; ++s)
          if (s_dst[s] == p) {
            out = base + s;
            break;
          }
      }
      pf_col_c[base + e] = out;
    }
  }
}

// Compact rows.  stage 0: c_x[s] = prot_x[dst], c_h[s] = enc_table[seed_row[dst]] (the encoder output of that atom).
// stage 1: c_agg[s] += aggd[distinct node of dst] (the shared pp means join the fp means already stored in c_agg).
// One warp per slot; empty slots carry dst = first node of the graph (K2's convention) and are processed like any other.
__global__ void __launch_bounds__(256) share_gather_kernel(const int* __restrict__ pharm_ptr, const int* __restrict__ prot_ptr,
                                                           const int* __restrict__ pk_node0, int n_graphs, int k,
                                                           const int* __restrict__ fp_seg_dst, const int* __restrict__ fp_seg_cnt,
                                                           const float* __restrict__ prot_x,
                                                           const int* __restrict__ seed_row, const float* __restrict__ enc_table,
                                                           const float* __restrict__ aggd_h, const float* __restrict__ aggd_v,
                                                           float* __restrict__ c_x, float* __restrict__ c_h,
                                                           float* __restrict__ c_agg_h, float* __restrict__ c_agg_v, int stage) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int fa = pharm_ptr[g], nf = pharm_ptr[g + 1] - fa, base = k * fa;
    const int pa = prot_ptr[g], shift = pk_node0[g] - pa;
    for (int s = warp; s < k * nf; s += nwarp) {
      const int dst = fp_seg_dst[base + s];
      const size_t row = (size_t)(base + s);
      if (stage == 0) {
        if (lane < 3) c_x[row * 3 + lane] = prot_x[(size_t)dst * 3 + lane];
        const float4 v = __ldg(reinterpret_cast<const float4*>(enc_table + (size_t)seed_row[dst] * kHidden) + lane);
        reinterpret_cast<float4*>(c_h + row * kHidden)[lane] = v;
      } else if (fp_seg_cnt[base + s] > 0) {   // empty slots are never read downstream: leave their rows alone
        const size_t d = (size_t)(dst + shift);
        float4* oh = reinterpret_cast<float4*>(c_agg_h + row * kHidden) + lane;
        const float4 a = __ldg(reinterpret_cast<const float4*>(aggd_h + d * kHidden) + lane);
        float4 o = *oh;
        o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
        *oh = o;
        if (lane < kVRow / 4) {
          float4* ov = reinterpret_cast<float4*>(c_agg_v + row * kVRow) + lane;
          const float4 b = __ldg(reinterpret_cast<const float4*>(aggd_v + d * kVRow) + lane);
          float4 w = *ov;
          w.x += b.x, w.y += b.y, w.z += b.z, w.w += b.w;
          *ov = w;
        }
      }
    }
  }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_share_index(const int32_t* pharm_ptr, int32_t n_graphs, int32_t pf_k, const int32_t* pf_cnt,
                              const int32_t* pf_col, const int32_t* fp_seg_dst, const int32_t* fp_seg_cnt, int32_t* pf_col_c,
                              void* stream) {
  PF_CHECK_ARG(pharm_ptr && pf_cnt && pf_col && fp_seg_dst && fp_seg_cnt && pf_col_c, "pf_share_index: null pointer");
  PF_CHECK_ARG(pf_k >= 1 && pf_k <= PF_MAX_KNN, "pf_share_index: pf_k out of range");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 16 * num_sms() ? n_graphs : 16 * num_sms();
  const size_t smem = (size_t)pf_k * PF_MAX_PHARM_PER_GRAPH * sizeof(int);
  share_index_kernel<<<grid, 128, smem, as_stream(stream)>>>(pharm_ptr, n_graphs, pf_k, pf_cnt, pf_col, fp_seg_dst, fp_seg_cnt,
                                                             pf_col_c);
  PF_CHECK_LAUNCH("pf_share_index");
  return PF_OK;
}

extern "C" int pf_share_gather(const int32_t* pharm_ptr, const int32_t* prot_ptr, const int32_t* pk_node0, int32_t n_graphs,
                               int32_t pf_k, const int32_t* fp_seg_dst, const int32_t* fp_seg_cnt, const float* prot_x,
                               const int32_t* seed_row,
                               const float* enc_table, const float* aggd_h, const float* aggd_v, float* c_x, float* c_h,
                               float* c_agg_h, float* c_agg_v, int32_t stage, void* stream) {
  PF_CHECK_ARG(pharm_ptr && prot_ptr && pk_node0 && fp_seg_dst && fp_seg_cnt, "pf_share_gather: null pointer");
  PF_CHECK_ARG(stage == 0 ? (prot_x && seed_row && enc_table && c_x && c_h) : (aggd_h && aggd_v && c_agg_h && c_agg_v),
               "pf_share_gather: null pointer for this stage");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 16 * num_sms() ? n_graphs : 16 * num_sms();
  share_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(pharm_ptr, prot_ptr, pk_node0, n_graphs, pf_k, fp_seg_dst, fp_seg_cnt, prot_x,
                                                           seed_row, enc_table, aggd_h, aggd_v, c_x, c_h, c_agg_h, c_agg_v, stage);
  PF_CHECK_LAUNCH("pf_share_gather");
  return PF_OK;
}
