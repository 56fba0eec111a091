/*
 * pharmacoforge_b200 -- C ABI of the B200-native PharmacoForge denoising hot path.
 *
 * The reference (eflynn8/pharmacophore-diffusion) is pure Python; its native compute is reached through
 * torch_cluster, DGL and ATen.  There is no FFI layer upstream, so every entry point below cites the
 * reference call site (file:line under /root/reference) whose work it replaces.  The Python host in
 * pharmacoforge_b200/ binds these with ctypes and exposes them as torch.library custom ops.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never allocates or frees
 *     device memory and keeps no pointer after a call returns;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation;
 *   - return value: 0 on success, a negative PfStatus on bad arguments or a launch failure; the text is
 *     available from pf_last_error() (thread-local);
 *   - conditions only the device can detect (in-degree above the tile capacity, more pharmacophore nodes
 *     in one graph than the graph builder stages) set bits in the caller-provided `dev_status` word;
 *   - node feature rows are fp32: scalars h[N][128], vectors v[N][3][16] (component-major; the reference
 *     uses [N][16][3]), coordinates x[N][3]; indices are int32.
 *   - an edge type is described by a SEGMENT TABLE sorted by destination: segment s owns the edge slots
 *     [seg_start[s], seg_start[s]+seg_cnt[s]) of `col` (source node ids) and aggregates onto node
 *     seg_dst[s] (seg_dst == NULL means segment s is node s).  A TILE is a run of consecutive segments
 *     whose edges fit one CTA tile; tiles own whole segments, so aggregation needs no atomics.
 */
#ifndef PHARMACOFORGE_B200_H
#define PHARMACOFORGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 5

enum PfStatus {
  PF_OK = 0,
  PF_ERR_BAD_ARG = -1,
  PF_ERR_LAUNCH = -2,
  PF_ERR_UNSUPPORTED = -3,
  PF_ERR_WORKSPACE = -4,
};

/* bits of the device status word */
#define PF_DEV_DEGREE_OVERFLOW 1u /* a destination has more in-edges than PF_TILE_ROWS */
#define PF_DEV_GRAPH_TOO_LARGE 2u /* a graph has more pharmacophore nodes than PF_MAX_PHARM_PER_GRAPH */
#define PF_DEV_TILE_OVERFLOW 4u   /* the tile list capacity was exceeded */
#define PF_DEV_EDGE_OVERFLOW 8u   /* an edge buffer capacity was exceeded */

#define PF_HIDDEN 128            /* dynamics.n_hidden_scalars (configs/dev.yml:82) */
#define PF_VEC 16                /* dynamics.vector_size (configs/dev.yml:80) */
#define PF_RBF 16                /* GVPMultiEdgeConv rbf_dim default (gvp.py:350) */
#define PF_TILE_ROWS 64          /* edge / node rows per CTA tile, fp32 FFMA kernels */
#define PF_TC_TILE_ROWS 128      /* edge rows per tile of the tcgen05 kernels (= TMEM lanes) */
#define PF_MAX_PHARM_PER_GRAPH 128
#define PF_MAX_KNN 16

int pf_abi_version(void);
const char* pf_last_error(void);

/* ---- packed GVP weights -------------------------------------------------------------------------
 * One GVP (gvp.py:43-116) with vi/vo vector channels in/out (hidden vh = max(vi,vo)), si/so scalar
 * features in/out is packed as six fp32 sections, each padded to a multiple of 4 floats:
 *   Wh[vi][vh] | Wu[vh][vo] | WfT[K4][so] (to_feats_out.0.weight transposed, K4 = roundup(si+vh,4),
 *   padding rows zero) | bf[so] | WgT[so][vo] (scalar_to_vector_gates.weight transposed) | bg[vo].
 * pf_gvp_layout writes the six section offsets (in floats) and returns the packed size in floats. */
int64_t pf_gvp_layout(int vi, int vo, int si, int so, int64_t offsets_out_host[6]);

/* ---- exclusive scan (plumbing for CSR construction) ---------------------------------------------*/
size_t pf_scan_workspace_bytes(int64_t n);
/* out[i] = sum(in[0..i)), i in [0, n]; out has n+1 entries. */
int pf_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- K1: static radius graph -> destination-sorted CSR --------------------------------------------
 * Replaces torch_cluster.radius_graph(prot_x, r=3.5, max_num_neighbors=100) at
 * protein_pharm_dataset.py:235 and its per-copy replication (unorganized_utils.py:28-81, dgl.batch).
 * Nodes of segment g are [seg_ptr[g], seg_ptr[g+1]).  Edge (src j -> dst i) exists iff same segment,
 * i != j, ((dx*dx+dy*dy)+dz*dz) < r*r in fp32 without FMA contraction, keeping the first max_nbrs
 * neighbours in ascending j.  Pass 1 writes deg[i]; the caller scans deg into rowptr; pass 2 writes
 * col[rowptr[i] .. rowptr[i]+deg[i]) in ascending j. */
int pf_radius_count(const float* x, const int32_t* seg_ptr, int32_t n_seg, float r, int32_t max_nbrs,
                    int32_t* deg, void* stream);
int pf_radius_fill(const float* x, const int32_t* seg_ptr, int32_t n_seg, float r, int32_t max_nbrs,
                   const int32_t* rowptr, int32_t* col, void* stream);

/* K1 as a cell list over DISTINCT pockets + replication (the graph is built once per pocket and copied per sample, as
 * protein_pharm_dataset.py:234-236 + unorganized_utils.py:28-50 + dgl.batch do).  Same edge rule and the same output as
 * pf_radius_count / pf_radius_fill (bit-identical CSR), O(N) instead of O(N^2 / pocket).  Both calls take the same
 * workspace of pf_cell_radius_workspace_bytes(n_nodes, n_seg) bytes: count builds the cell lists in it (cell edge
 * >= r, at most 2 n + 8 cells per pocket), fill reads them.  Rows with more than 160 hits fall back to the ordered scan. */
size_t pf_cell_radius_workspace_bytes(int64_t n_nodes, int32_t n_seg);
int pf_cell_radius_count(const float* x, const int32_t* seg_ptr, int32_t n_seg, int64_t n_nodes, float r, int32_t max_nbrs,
                         void* workspace, size_t workspace_bytes, int32_t* deg, void* stream);
int pf_cell_radius_fill(const float* x, const int32_t* seg_ptr, int32_t n_seg, int64_t n_nodes, float r, int32_t max_nbrs,
                        void* workspace, size_t workspace_bytes, const int32_t* rowptr, int32_t* col, void* stream);
/* Graph g of the batch is a copy of the pocket whose first node is pk_node0[g] in the pocket arrays: its nodes are
 * [prot_ptr[g], prot_ptr[g+1]), its edges start at edge0[g] (exclusive scan of the copies' edge counts).  Writes
 * rowptr [N+1], cnt [N], col [E] of the batched graph: 4 B per edge + 8 B per node, coalesced. */
int pf_replicate_csr(const int32_t* pk_rowptr, const int32_t* pk_col, const int32_t* pk_node0, const int32_t* prot_ptr,
                     const int32_t* edge0, int32_t n_graphs, int32_t* rowptr, int32_t* cnt, int32_t* col, void* stream);

/* ---- K2: per-step dynamic graph ------------------------------------------------------------------
 * Replaces, per reverse-diffusion step, dynamics_gvp.py:187-246: radius_graph(pharm x_t, r=ff_r,
 * max 200) (:196), knn(prot x_0, pharm x_t, k) (:202) and the six DGL add/remove_edges mutations.
 * One CTA per graph.  Outputs (all int32, caller-allocated):
 *   ff: row i (a pharm node of graph g with nf nodes) owns slots [ff_start[i], ff_start[i]+nf-1);
 *       ff_start is an INPUT (static, computed once per batch); ff_cnt[i] and ff_col are written.
 *   pf: pharm node i owns slots [k*i, k*i+k) of pf_col (prot ids, ascending distance, ties to the
 *       lower id); pf_cnt[i] = min(k, prot atoms in the graph).
 *   fp: the reverse edges sorted by (prot id, pharm id).  Graph g owns segment slots
 *       [k*pharm_ptr[g], k*pharm_ptr[g+1]): fp_seg_dst / fp_seg_start / fp_seg_cnt (unused slots have
 *       cnt 0), and the same range of fp_col (pharm ids). */
int pf_dyn_graph(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x, const int32_t* pharm_ptr,
                 int32_t n_graphs, float ff_r, int32_t ff_max_nbrs, int32_t pf_k, const int32_t* ff_start,
                 int32_t* ff_cnt, int32_t* ff_col, int32_t* pf_cnt, int32_t* pf_col, int32_t* fp_seg_dst,
                 int32_t* fp_seg_start, int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status, void* stream);

/* The same with the ff edges taken from knn_graph(pharm x_t, k = ff_k) when ff_k > 0 (dynamics_gvp.py:193-194): per centre
 * the ff_k + 1 nearest nodes of its graph INCLUDING itself, ordered by (distance, index), minus the self pair; stored in
 * ascending source index in the same ff slots.  ff_k == 0 is pf_dyn_graph. */
int pf_dyn_graph_ffk(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x, const int32_t* pharm_ptr,
                     int32_t n_graphs, float ff_r, int32_t ff_max_nbrs, int32_t ff_k, int32_t pf_k, const int32_t* ff_start,
                     int32_t* ff_cnt, int32_t* ff_col, int32_t* pf_cnt, int32_t* pf_col, int32_t* fp_seg_dst,
                     int32_t* fp_seg_start, int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status, void* stream);

/* The pf_k == 0 branch of the same call site (dynamics_gvp.py:210-216, the reference constructor's default; configs/dev.yml
 * uses pf_k = 5): pf / fp edges from radius(x = pharm x_t, y = prot x_0, r = pf_r, max_num_neighbors = pf_max_nbrs) -- every
 * protein atom keeps the pharmacophore nodes of its graph with squared distance < r * r in ascending index, at most
 * pf_max_nbrs; pf = (prot -> pharm), fp = the reverse.  ff as in pf_dyn_graph_ffk.  Static inputs (per batch): pf_start[i] =
 * first slot of pharmacophore node i in pf_col (capacity = the atoms of its graph), sub_ptr[i] = first of its
 * ceil(atoms / sub_rows) sub-segment slots, fp_base[g] = first fp_col slot of graph g (min(nf, pf_max_nbrs) per atom).
 * Written: pf_cnt[i] (in-degree), pf_col (atom ids, ascending), sub_start / sub_cnt (the node's pf segment cut into pieces of
 * at most sub_rows edges: what the edge kernels run on, one mean per piece -> pf_combine_subsegments), sub_x[s][3] (the
 * coordinates of the piece's destination node: the edge kernels' dst_x, indexed by piece like their output), and one fp segment per
 * protein atom: fp_seg_start / fp_seg_cnt [n_prot], fp_col (pharm ids, ascending). */
int pf_dyn_graph_radius(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x, const int32_t* pharm_ptr,
                        int32_t n_graphs, float ff_r, int32_t ff_max_nbrs, int32_t ff_k, float pf_r, int32_t pf_max_nbrs,
                        int32_t sub_rows, const int32_t* ff_start, int32_t* ff_cnt, int32_t* ff_col, const int32_t* pf_start,
                        const int32_t* sub_ptr, const int32_t* fp_base, int32_t* pf_cnt, int32_t* pf_col, int32_t* sub_start,
                        int32_t* sub_cnt, float* sub_x, int32_t* fp_seg_start, int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status,
                        void* stream);

/* agg[d] (+)= sum over s in [sub_ptr[d], sub_ptr[d+1]) of sub[s] * sub_cnt[s] * w with w = 1 / tot_cnt[d] (inv_norm == 0: the
 * mean over all in-edges of d, fn.mean of gvp.py:488-497), w = inv_norm (numeric message_norm) or, when inv_norm_node is
 * given, w = inv_norm_node[d] (message_norm = 0).  Rows as pf_scaled_accumulate. */
int pf_combine_subsegments(const float* sub_h, const float* sub_v, const int32_t* sub_cnt, const int32_t* sub_ptr,
                           const int32_t* tot_cnt, int64_t n_dst, float inv_norm, const float* inv_norm_node, float* agg_h,
                           float* agg_v, int32_t accumulate, void* stream);

/* ---- tile planner --------------------------------------------------------------------------------
 * Greedily packs consecutive segments of each chunk [chunk_ptr[c], chunk_ptr[c+1]) into tiles of at
 * most tile_rows edges and tile_rows segments (PF_TILE_ROWS for the FFMA kernels, PF_TC_TILE_ROWS for tcgen05).  tiles[2*t], tiles[2*t+1] = first / one-past-last
 * segment.  *n_tiles must be zeroed by the caller (pf_zero_i32).  With skip_empty != 0 tiles without
 * edges are not emitted (accumulate-mode edge types). */
int pf_plan_tiles(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                  int32_t tile_rows, int32_t* tiles, int32_t max_tiles, int32_t* n_tiles, uint32_t* dev_status,
                  void* stream);
/* Three plans (the per-step ff / pf / fp plans of pf_denoiser) in one launch: arrays of three HOST-side entries each,
 * n_tiles3[i] is plan i's counter (zeroed by the caller). */
int pf_plan_tiles3(const int32_t* const seg_cnt[3], const int32_t* const chunk_ptr[3], const int32_t n_chunks[3],
                   const int32_t skip_empty[3], int32_t tile_rows, int32_t* const tiles[3], int32_t max_tiles,
                   int32_t* n_tiles3, uint32_t* dev_status, void* stream);
int pf_zero_i32(int32_t* p, int64_t n, void* stream);
/* The same plan with the tiles in CHUNK ORDER (a chunk's tiles contiguous, chunks ascending), for the static pp plan: the
 * persistent kernels process consecutive tiles concurrently, so the source rows of a graph are fetched from DRAM once and
 * reused from L2.  Two passes around a caller-side exclusive scan: pf_plan_tiles_count writes the number of tiles of every
 * chunk; pf_plan_tiles_fill writes tiles[2*(chunk_tile_off[c] + i)] and *n_tiles (chunk_tile_off = exclusive scan). */
int pf_plan_tiles_count(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                        int32_t tile_rows, int32_t* chunk_tiles, uint32_t* dev_status, void* stream);
int pf_plan_tiles_fill(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                       int32_t tile_rows, const int32_t* chunk_tile_off, int32_t* tiles, int32_t max_tiles,
                       int32_t* n_tiles, uint32_t* dev_status, void* stream);

/* ---- K0: time-conditioned scalar encoders --------------------------------------------------------
 * h[n] = LayerNorm(SiLU(W [feats[n], t[graph(n)]] + b)) (dynamics_gvp.py:107-117,143-151).
 * w = Wt[(nf+1)][128] (Linear weight transposed) | b[128] | ln_w[128] | ln_b[128]. */
int pf_encode(const float* feats, int32_t n_feats, const int32_t* node_ptr, int32_t n_graphs, const float* t,
              const float* w, float* h_out, void* stream);

/* ---- K3: fused edge message + mean aggregation for one edge type ----------------------------------
 * Replaces, for one edge type of one GVPMultiEdgeConv layer: u_sub_v + normalise + RBF (gvp.py:472-480),
 * the gather of edges.src['h'|'v'] and the n_gvps-GVP message chain (gvp.py:540-551 -> 89-116), and
 * multi_update_all(copy_e, mean) (gvp.py:488-497).  src_v == NULL means all-zero source vectors (first
 * layer, dynamics_gvp.py:162-173).  accumulate == 0: every destination covered by a tile is written
 * (zero rows for segments without edges); accumulate != 0: out += mean for segments with edges only
 * (the cross-etype 'sum' reducer).  w = n_gvps packed GVPs back to back: (17,16,144,128) then
 * (16,16,128,128) each. */
int pf_edge_conv(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                 const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst, const int32_t* col,
                 const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles, const float* w, int32_t n_gvps,
                 float* agg_h, float* agg_v, int32_t accumulate, void* stream);

/* K3 on the tensor cores (tcgen05.mma, bf16 hi/lo split operands, fp32 accumulation in TMEM; fp32-accurate).
 * Same contract as pf_edge_conv for n_gvps == 3 with tiles planned at PF_TC_TILE_ROWS; `wblob` is the
 * pf_tc_msg_blob_bytes()-byte image built by pharmacoforge_b200/weights.py:pack_message_tc (weight slabs in the
 * UMMA SWIZZLE_NONE K-major layout, split into bf16 hi and lo parts), 16-byte aligned. */
size_t pf_tc_msg_blob_bytes(void);
int pf_edge_conv_tc(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                    const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst, const int32_t* col,
                    const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles, const void* wblob, float* agg_h,
                    float* agg_v, int32_t accumulate, void* stream);
/* Single-pass fp16 variant -- the reduced-precision edge-MLP path of BASELINE.json configs[3]: same arguments and weight
 * image; every contraction is ONE tcgen05.mma pass over the fp16 hi parts (11-bit operands, fp32 accumulation) and SiLU
 * runs on packed fp16 pairs.  Tolerance: eps within 2e-2 of max|eps| per call (tests/test_gpu_tcgen05.py), not the
 * 1e-4 fp32 bar -- never used unless asked for (PF_FLAG_FP16_SINGLE_PASS / dynamics.edge_mlp_precision = "fp16"). */
int pf_edge_conv_tc_f16(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                        const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst, const int32_t* col,
                        const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles, const void* wblob,
                        float* agg_h, float* agg_v, int32_t accumulate, void* stream);

/* First conv layer with one-hot source features: the exact split of SURVEY.md hard part 2,
 *   Wf0 [h_src; rbf; sh] = Wf0[:, 0:128] h_src  (per source NODE)  +  Wf0[:, 128:161] [rbf; sh]  (per edge).
 * pf_seed_table: table[r][:] = k * Wf0[:, 0:128] h[rep_node[r]][:] for r < n_rows (rows with rep_node[r] < 0 are left
 * untouched), fp32 FFMA; w_msg is the fp32 packed message chain of the edge type (pf_edge_conv's `w`), k = -log2(e) the
 * factor the tcgen05 weight images carry.  pf_edge_conv_tc_seeded: pf_edge_conv_tc / _f16 without source scalars and
 * vectors: edge e starts GVP 0's accumulator from table row seed_row[col[e]] and contracts only [rbf; sh] per edge
 * (3 of the 11 K-steps of GVP 0); no 512-byte row gather, no fp16 split of gathered features.  Replaces the same
 * reference lines as pf_edge_conv (gvp.py:472-497, 540-551) for conv layer 0, where edges.src['v'] is zero and
 * edges.src['h'] is the encoder output (dynamics_gvp.py:143-173). */
int pf_seed_table(const float* h, const int32_t* rep_node, int32_t n_rows, const float* w_msg, float* table, void* stream);
int pf_edge_conv_tc_seeded(const int32_t* seed_row, const float* seed_table, const float* src_x, const float* dst_x,
                           const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst, const int32_t* col,
                           const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles, const void* wblob,
                           float* agg_h, float* agg_v, int32_t accumulate, int32_t fp16_single_pass, void* stream);

/* Debug timeline of CTA 0 of the next pf_edge_conv_tc launches: device_buf = int64[4][4096][2] of (tag, clock64)
 * for the epilogue of tile slot 0 / 1 and the two MMA issuers, followed by int64[148][2] = (begin, end) globaltimer ns of
 * every CTA; NULL disarms.  Not used on the product path. */
int pf_tc_trace(long long* device_buf);

/* ---- K4: node update ------------------------------------------------------------------------------
 * gvp.py:511-532 in eval mode: (h,v) <- GVPLayerNorm_msg(h + agg_h, v + agg_v); (rh,rv) = GVP x n_gvps;
 * (h,v) <- GVPLayerNorm_upd(h + rh, v + rv).  v_in == NULL means zero.  In place is allowed
 * (h_out == h_in).  w = ln_msg_w[128] | ln_msg_b[128] | ln_upd_w[128] | ln_upd_b[128] | n_gvps packed GVPs
 * (16,16,128,128). */
int pf_node_update(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v, int64_t n_nodes,
                   const float* w, int32_t n_gvps, float* h_out, float* v_out, void* stream);

/* K4 on the tensor cores (same engine as pf_edge_conv_tc, two plain GVPs): n_gvps == 2 only; `wblob` is the
 * pf_tc_upd_blob_bytes()-byte image built by weights.py:pack_update_tc.  In place is allowed. */
size_t pf_tc_upd_blob_bytes(void);
int pf_node_update_tc(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v, int64_t n_nodes,
                      const void* wblob, float* h_out, float* v_out, void* stream);
int pf_node_update_tc_f16(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v, int64_t n_nodes,
                          const void* wblob, float* h_out, float* v_out, void* stream); /* single fp16 pass, see above */

/* ---- K5a: noise head -------------------------------------------------------------------------------
 * NoisePredictionBlock (dynamics_gvp.py:10-42): (n_gvps-1) x GVP(16,16,128,128) + GVP(16,1,128,64) with
 * identity gate activation + Linear(64 -> n_out).  w = packed GVPs | Wt[64][n_out4] | b[n_out4]
 * (n_out4 = roundup(n_out,4)).  eps_h [n][n_out], eps_x [n][3]. */
int pf_noise_head(const float* h, const float* v, int64_t n_nodes, const float* w, int32_t n_gvps, int32_t n_out,
                  float* eps_h, float* eps_x, void* stream);

/* ---- K5b: DDPM posterior step + centre-of-mass removal, in place ------------------------------------
 * sample_p_zs_given_zt (pharmacodiff.py:413-429) for the eps parameterisation:
 *   z_s = z_t/alpha_ts - var_terms*eps + sigma_q*noise   for x (3) and h (nh),
 * then the per-graph pharmacophore COM is subtracted from pharm x AND prot x (com_removal, :88-108).
 * The three coefficients are the same for every graph of a sampling batch. */
int pf_posterior_step(float* pharm_x, float* pharm_h, int32_t nh, const float* eps_x, const float* eps_h,
                      const float* noise_x, const float* noise_h, const int32_t* pharm_ptr, float* prot_x,
                      const int32_t* prot_ptr, int32_t n_graphs, float alpha_ts, float var_terms, float sigma_q,
                      void* stream);

/* The same step with in-kernel Gaussian noise (SURVEY.md 8d allows Philox for throughput runs; parity runs inject noise):
 * Philox4x32-10 keyed by the 64-bit *seed_dev (device memory), counter = (element index, noise_step, stream x / h),
 * Box-Muller.  pf_philox_normal fills out[0..n) with the draws of (stream_id, step) -- the initial z_T uses step 0 with
 * stream 0 for x and 1 for h (pharmacodiff.py:455-456), loop iteration i uses step i + 1. */
int pf_posterior_step_philox(float* pharm_x, float* pharm_h, int32_t nh, const float* eps_x, const float* eps_h,
                             const uint64_t* seed_dev, uint32_t noise_step, const int32_t* pharm_ptr, float* prot_x,
                             const int32_t* prot_ptr, int32_t n_graphs, float alpha_ts, float var_terms, float sigma_q,
                             void* stream);
int pf_philox_normal(float* out, int64_t n, const uint64_t* seed_dev, uint32_t stream_id, uint32_t step, void* stream);

/* The same step for models trained with the ENDPOINT parameterisation (pharmacodiff.py:413-418; `endpoint_param_coord` /
 * `endpoint_param_feat`: the network output is the predicted x_0 / h_0 instead of eps):
 *   mu = ep_c1 * z_t + ep_c2 * pred,  ep_c1 = alpha_{t|s} sigma_s^2 / sigma_t^2,  ep_c2 = alpha_s sigma^2_{t|s} / sigma_t^2
 * for the parts selected by ep_mode (PF_EP_COORD: x, PF_EP_FEAT: h); the other part keeps the eps form.  Noise: pass
 * noise_x and noise_h (injected), or both NULL and seed_dev / noise_step (in-kernel Philox). */
#define PF_EP_COORD 1
#define PF_EP_FEAT 2
int pf_posterior_step_ep(float* pharm_x, float* pharm_h, int32_t nh, const float* pred_x, const float* pred_h,
                         const float* noise_x, const float* noise_h, const uint64_t* seed_dev, uint32_t noise_step,
                         const int32_t* pharm_ptr, float* prot_x, const int32_t* prot_ptr, int32_t n_graphs,
                         float alpha_ts, float var_terms, float sigma_q, float ep_c1, float ep_c2, int32_t ep_mode,
                         void* stream);

/* per-graph mean of x over [ptr[g], ptr[g+1]) -> com[g][3]; and x[n] += sign * com[graph(n)]
 * (dgl.readout_nodes + broadcast subtract/add, pharmacodiff.py:442-452,483-486) */
int pf_segment_mean3(const float* x, const int32_t* ptr, int32_t n_graphs, float* com, void* stream);
int pf_segment_shift3(float* x, const int32_t* ptr, int32_t n_graphs, const float* com, float sign, void* stream);

/* ---- whole reverse-diffusion loop ---------------------------------------------------------------------
 * sample_given_receptor's loop (pharmacodiff.py:466-472): enqueues all `n_steps` steps (graph build,
 * tile plans, encoders, n_convs x (4 edge types + 2 node updates), noise head, posterior + COM) on
 * `stream` without returning to Python.  All buffers are described by PfSampleArgs. */
typedef struct PfSampleArgs {
  int32_t n_graphs, n_prot, n_pharm, n_prot_feats, n_pharm_feats;
  int32_t n_convs, n_msg_gvps, n_upd_gvps, n_noise_gvps, pf_k, ff_max_nbrs;
  float ff_r;
  /* batch (device) */
  float* prot_x;              /* [n_prot][3], shifted in place every step */
  const float* prot_feats;    /* [n_prot][n_prot_feats] */
  const int32_t* prot_ptr;    /* [n_graphs+1] */
  float* pharm_x;             /* [n_pharm][3] */
  float* pharm_h;             /* [n_pharm][n_pharm_feats] */
  const int32_t* pharm_ptr;   /* [n_graphs+1] */
  /* static pp graph + its tile plan */
  const int32_t *pp_start, *pp_cnt, *pp_col, *pp_tiles, *pp_n_tiles;
  int32_t pp_max_tiles;
  /* dynamic graph buffers */
  const int32_t* ff_start;    /* [n_pharm] static */
  int32_t *ff_cnt, *ff_col, *pf_start, *pf_cnt, *pf_col, *fp_seg_dst, *fp_seg_start, *fp_seg_cnt, *fp_col;
  const int32_t *pharm_chunk_ptr, *fp_chunk_ptr; /* planner chunks over pharm nodes / fp segment slots */
  int32_t n_pharm_chunks, n_fp_chunks;
  int32_t *ff_tiles, *pf_tiles, *fp_tiles, *dyn_n_tiles; /* dyn_n_tiles[3] = ff, pf, fp */
  int32_t dyn_max_tiles;
  /* features */
  float *prot_h, *prot_v, *prot_agg_h, *prot_agg_v;     /* [n_prot][128], [n_prot][48] */
  float *pharm_hh, *pharm_v, *pharm_agg_h, *pharm_agg_v; /* [n_pharm][128], [n_pharm][48] */
  float *eps_h, *eps_x;                                  /* [n_pharm][n_pharm_feats], [n_pharm][3] */
  float* t_graph;                                        /* [n_graphs] scratch: timestep value per graph */
  /* packed weights (device) */
  const float *w_pharm_enc, *w_prot_enc;
  const float* w_msg[8][4];   /* [conv][etype: ff, pf, fp, pp] */
  const float* w_upd[8][2];   /* [conv][ntype: pharm, prot] */
  const float* w_noise;
  /* tcgen05 path: tile_rows = PF_TC_TILE_ROWS and every w_msg_tc[conv][etype] set -> K3 runs on the tensor cores;
   * tile_rows = PF_TILE_ROWS -> fp32 FFMA kernels */
  const void* w_msg_tc[8][4];
  const void* w_upd_tc[8][2]; /* optional (n_upd_gvps == 2): K4 on the tensor cores when tile_rows = PF_TC_TILE_ROWS */
  int32_t tile_rows;
  /* schedule (host): step i uses t = t_host[i], coefficients alpha_ts_host[i], ... */
  const float *t_host, *alpha_ts_host, *var_terms_host, *sigma_q_host;
  const float* noise_x;       /* device [n_steps][n_pharm][3] */
  const float* noise_h;       /* device [n_steps][n_pharm][n_pharm_feats] */
  int32_t n_steps;
  uint32_t* dev_status;
  /* PF_FLAG_SKIP_DEAD_WORK: do not compute what PharmRecGVP.forward (dynamics_gvp.py:84-92) never reads -- the
   * protein-side outputs of the LAST conv layer (its pp and fp messages and its protein node update): only
   * node_data['pharm'] of the last layer feeds the noise head.  eps_h / eps_x are bit-identical either way; the
   * nominal (reference-equivalent) work is done when the flag is clear, which is the default. */
  uint32_t flags;
  /* First-layer seeding (tcgen05 path; seed_row == NULL disables it): when the protein features are one-hot, the encoder
   * output of a protein node depends only on (graph, atom type), so the per-node part of the first message GVP's scalar
   * contraction is one table row per (graph, type) -- see pf_seed_table / pf_edge_conv_tc_seeded.  seed_row and seed_rep are
   * static per batch; seed_table is scratch rewritten by every pf_denoiser call. */
  const int32_t* seed_row;  /* [n_prot] table row of every protein node: graph * n_prot_feats + type */
  const int32_t* seed_rep;  /* [n_seed_rows] one protein node with that (graph, type), or -1 if the graph has none */
  float* seed_table;        /* [n_seed_rows][128] */
  int32_t n_seed_rows;
  /* In-kernel noise for throughput runs (noise_x == noise_h == NULL): Philox4x32-10 keyed by *noise_seed (DEVICE memory:
   * a captured CUDA graph replays with a new seed), loop iteration i draws with step number noise_step0 + i. */
  const uint64_t* noise_seed;
  int32_t noise_step0;
  int32_t ff_k; /* > 0: ff edges are the kNN graph of the pharmacophore nodes (pf_dyn_graph_ffk) instead of the radius graph */
  /* PF_FLAG_SHARE_POCKET_MESSAGES: buffers of the opt-in shared-pocket mode (csrc/pf_share.cu); NULL when unused */
  const float* pk_x;           /* [n_distinct][3] input coordinates of the DISTINCT pockets of the batch */
  const int32_t *pk_start, *pk_cnt, *pk_col, *pk_tiles, *pk_n_tiles; /* their pp CSR + tile plan */
  int32_t pk_max_tiles, n_distinct;
  const int32_t* pk_seed_row;  /* [n_distinct] seed-table row of every distinct protein node */
  const int32_t* pk_node0;     /* [n_graphs] first distinct node of the graph's pocket */
  const float* enc_feats;      /* [n_graphs * n_prot_feats][n_prot_feats] identity blocks: input of the encoder table */
  const int32_t* enc_ptr;      /* [n_graphs + 1] = n_prot_feats * g */
  const int32_t* enc_rep;      /* [n_graphs * n_prot_feats] identity (every table row is its own representative) */
  float* enc_table;            /* [n_graphs * n_prot_feats][128] encoder output per (graph, atom type) */
  float *aggd_h, *aggd_v;      /* [n_distinct][128], [n_distinct][48] pp means per distinct node */
  float *c_x, *c_h, *c_v, *c_agg_h, *c_agg_v; /* compact protein rows, one per fp segment slot: [pf_k * n_pharm][...] */
  const int32_t* c_seg_id;     /* [pf_k * n_pharm] identity */
  int32_t* pf_col_c;           /* [pf_k * n_pharm] compact source row of every pf edge */
  /* endpoint parameterisation (pf_posterior_step_ep): per-step coefficient tables (host) and the PF_EP_* mode bits;
   * ep_mode == 0 (configs/dev.yml) leaves the eps form of every step */
  const float *ep_c1_host, *ep_c2_host;
  int32_t ep_mode;
  /* numeric message_norm (gvp.py:375-389, 512-517; configs/dev.yml uses 'mean' = both zero): aggregation is the SUM over the
   * in-edges divided by the norm of the destination node type.  The edge kernels then write each edge type's means to
   * tmp_agg_* ([max(n_prot, n_pharm)][128] / [..][48], scratch) and pf_scaled_accumulate folds them in as count / norm * mean. */
  float msg_norm_pharm, msg_norm_prot;
  float *tmp_agg_h, *tmp_agg_v;
  /* pf_k == 0: pf / fp edges from the radius graph (pf_dyn_graph_radius).  pf_start / pf_cnt / pf_col then describe whole pf
   * segments (capacity = atoms of the graph per pharmacophore node), fp_seg_start / fp_seg_cnt are [n_prot] with identity
   * destinations (fp_seg_dst unused), fp_chunk_ptr / n_fp_chunks chunk the protein atoms, and the pf edge kernels run on the
   * sub-segment list below, one mean per sub-segment into sub_agg_*, folded in by pf_combine_subsegments. */
  float pf_r;
  int32_t pf_max_nbrs;
  const int32_t *pf_sub_ptr, *fp_base;       /* [n_pharm + 1], [n_graphs] static */
  int32_t *pf_sub_start, *pf_sub_cnt;        /* [n_pf_sub] */
  const int32_t* pf_sub_chunk_ptr;           /* planner chunks over the sub-segment slots */
  int32_t n_pf_sub_chunks, n_pf_sub;
  float *sub_agg_h, *sub_agg_v;              /* [n_pf_sub][128], [n_pf_sub][48] scratch */
  float* pf_sub_x;                           /* [n_pf_sub][3] destination coordinates per sub-segment */
  /* message_norm = 0: SUM / (edges per node of the graph + 1), pf_degree_norms; tmp_agg_* as for a numeric norm */
  int32_t msg_norm_degree;
  float *inv_norm_pharm, *inv_norm_prot;     /* [n_pharm], [n_prot] */
} PfSampleArgs;
#define PF_FLAG_SKIP_DEAD_WORK 1u
/* PF_FLAG_FP16_SINGLE_PASS: K3 / K4 run pf_edge_conv_tc_f16 / pf_node_update_tc_f16 (tcgen05 path only); the graph
 * kernels, noise head and posterior step are unchanged (fp32).  Off by default: the default is the fp32-parity mode. */
#define PF_FLAG_FP16_SINGLE_PASS 2u

/* PF_FLAG_NO_LAYER0_SEED: run the first conv layer's pp messages through the general kernel (row gather + all 11 K-steps of
 * GVP 0) even when the seed arrays are bound -- the A/B switch of the seeded path (results agree to fp32 rounding). */
#define PF_FLAG_NO_LAYER0_SEED 4u

/* PF_FLAG_SHARE_POCKET_MESSAGES (with PF_FLAG_SKIP_DEAD_WORK, n_convs == 2, tcgen05 path, seed arrays bound, and THE SAME
 * TIMESTEP FOR EVERY GRAPH -- the reverse-diffusion loop): exact work elimination of SURVEY.md hard part 5b + 5c.  The
 * first layer's pp messages are computed once per distinct pocket (they depend only on (pocket, t)), and only the protein
 * rows the last layer's pf edges read (the fp segment destinations) are encoded and updated, in a compact buffer.  Equal to
 * the nominal path up to fp32 rounding of x_src - x_dst; never the default, reported separately by bench.py. */
#define PF_FLAG_SHARE_POCKET_MESSAGES 8u
/* First-layer encoder TABLE (tcgen05 path, one-hot protein features, enc_* buffers bound): the encoder output of a
 * protein node depends only on (graph, atom type), so the first conv layer reads the protein scalars from the
 * [n_graphs * n_prot_feats][128] table enc_table through the row map seed_row instead of a materialised [n_prot][128]
 * array -- the pf message gather (pf_edge_conv_tc_mapped) and the protein node update (pf_node_update_tc_mapped, which
 * writes the full prot_h) -- and the per-node encoder pass disappears.  Same arithmetic on the same values: results are
 * bit-identical.  PF_FLAG_NO_LAYER0_TABLE is the A/B switch. */
#define PF_FLAG_NO_LAYER0_TABLE 16u
int pf_edge_conv_tc_mapped(const float* src_h, const int32_t* src_map, const float* src_v, const float* src_x,
                           const float* dst_x, const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                           const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                           const void* wblob, float* agg_h, float* agg_v, int32_t accumulate, int32_t f16, void* stream);
int pf_node_update_tc_mapped(const float* h_in, const int32_t* h_map, const float* v_in, const float* agg_h,
                             const float* agg_v, int64_t n_nodes, const void* wblob, float* h_out, float* v_out,
                             int32_t f16, void* stream);
int pf_share_index(const int32_t* pharm_ptr, int32_t n_graphs, int32_t pf_k, const int32_t* pf_cnt, const int32_t* pf_col,
                   const int32_t* fp_seg_dst, const int32_t* fp_seg_cnt, int32_t* pf_col_c, void* stream);
/* stage 0: c_x / c_h from prot_x and the encoder table; stage 1: c_agg += pp means of the row's distinct node */
int pf_share_gather(const int32_t* pharm_ptr, const int32_t* prot_ptr, const int32_t* pk_node0, int32_t n_graphs,
                    int32_t pf_k, const int32_t* fp_seg_dst, const int32_t* fp_seg_cnt, const float* prot_x,
                    const int32_t* seed_row, const float* enc_table, const float* aggd_h, const float* aggd_v, float* c_x, float* c_h,
                    float* c_agg_h, float* c_agg_v, int32_t stage, void* stream);

/* agg[d] (+)= tmp[d] * seg_cnt[s] * inv_norm for every segment s (d = seg_dst ? seg_dst[s] : s): mean -> scaled sum (numeric
 * message_norm, see PfSampleArgs.msg_norm_*); with inv_norm_node != NULL the factor is inv_norm_node[d] (message_norm = 0).
 * Rows: agg_h / tmp_h [n][128], agg_v / tmp_v [n][48]. */
int pf_scaled_accumulate(const float* tmp_h, const float* tmp_v, const int32_t* seg_cnt, const int32_t* seg_dst, int64_t n_seg,
                         float inv_norm, const float* inv_norm_node, float* agg_h, float* agg_v, int32_t accumulate,
                         void* stream);

/* message_norm = 0 (gvp.py:504-507): 1 / ((edges of every type into the node type) / (nodes of the type) + 1) per graph, written
 * per node.  Per-graph edge counts as add_pharm_edges records them (dynamics_gvp.py:219-221): ff / pp true counts; pf (= fp)
 * through prot_batch_idx[pf_idxs[0]] -- true counts with radius edges (radius_mode = 1), and with kNN edges the edges of
 * pharmacophore node i counted for the graph that owns protein atom i (the reference's behaviour, kept). */
int pf_degree_norms(const int32_t* prot_ptr, const int32_t* pharm_ptr, int32_t n_graphs, int32_t n_pharm, const int32_t* ff_cnt,
                    const int32_t* pf_cnt, const int32_t* pp_cnt, int32_t radius_mode, float* inv_norm_pharm,
                    float* inv_norm_prot, void* stream);

/* One eps prediction, PharmRecDynamicsGVP.forward (dynamics_gvp.py:131-185); a->t_graph[g] must hold the
 * timestep value of graph g.  Results in a->eps_h / a->eps_x. */
int pf_denoiser(const PfSampleArgs* a, void* stream);
int pf_sample_loop(const PfSampleArgs* a, void* stream);
int pf_fill_f32(float* p, int64_t n, float v, void* stream);
size_t pf_sample_args_size(void); /* sizeof(PfSampleArgs), for binding self-checks */

/* ---- training path (PharmacophoreDiff.forward + backward, pharmacodiff.py:162-243) --------------------------------
 * Differentiable primitives the reference's GVP (gvp.py:89-116), GVPLayerNorm (gvp.py:159-166) and GVPMultiEdgeConv
 * (gvp.py:459-551) are composed of, forward and backward, fp32.  The host registers them as torch.library custom ops
 * with register_autograd (pharmacoforge_b200/train_ops.py).  This path is unfused: per-edge tensors are materialised.
 * Vectors are component-major [rows][3][channels].  Kernels whose name says `accumulate` / `+=` add into the output. */
/* C[M][N] (ldc) (+)= A(m,k) B(k,n) (+ bias[n]); A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs]: nn.Linear forward
 * (gvp.py:76-79, 84), its dgrad and wgrad, and the Wh / Wu contractions (gvp.py:99-100).  split_k > 1 splits K over
 * CTAs (the wgrad reductions over all edges); this entry point has no workspace and accumulates the splits with atomicAdd
 * (order-dependent rounding) -- pf_train_gemm with a training workspace reduces them deterministically. */
int pf_train_sgemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                   int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate, int32_t split_k,
                   void* stream);
/* The same contraction on the tensor cores (csrc/pf_tc_gemm.cu): operands split on the fly into fp16 (hi, lo), three
 * tcgen05.mma passes, fp32 accumulation in TMEM -- the numerics of the fused sampling kernels (22 significand bits).
 * N <= 176 and either K <= 176 (forward / input-gradient shapes: B stays resident in shared memory, persistent CTAs over
 * the rows) or M <= 128 with a long contraction (weight-gradient shapes: the contraction is split over CTAs, partials go
 * to `workspace` (pf_tc_gemm_workspace_bytes) and are summed in CTA order: deterministic, no atomics).
 * pf_train_gemm picks this kernel when the shape fits and a workspace is given, pf_train_sgemm (fp32 FFMA) otherwise. */
size_t pf_tc_gemm_workspace_bytes(int32_t M, int32_t N, int64_t K);
int pf_tc_gemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int64_t K, int64_t a_rs,
               int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate, void* workspace,
               size_t workspace_bytes, void* stream);
int pf_train_gemm(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                  int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int32_t ldc, int32_t accumulate, int32_t split_k,
                  void* workspace, size_t workspace_bytes, void* stream);
/* The TRAINING WORKSPACE (one per device and stream, caller-owned): pf_train_workspace_bytes() bytes, ZERO-FILLED once
 * before its first use.  Its front holds the partials of every reduction that is split over CTAs -- split-K weight
 * gradients (tensor-core and FFMA kernels: a second launch adds the splits up in split order), bias column sums and
 * LayerNorm affine gradients (the last CTA to arrive adds the CTAs' partials in CTA order; ticket counters in the last
 * PF_TRAIN_WS_TAIL bytes, which every call leaves at zero) -- so the gradients are bit-identical from run to run: no
 * atomics on floating-point data.  workspace == NULL selects the atomicAdd fallbacks. */
#define PF_TRAIN_WS_TAIL 16384
size_t pf_train_workspace_bytes(void);
/* out[n] += sum_m x[m][n] (bias gradient) */
int pf_train_colsum(const float* x, float* out, int64_t M, int32_t N, void* workspace, size_t workspace_bytes, void* stream);
/* SiLU (gvp.py:78): dy == NULL -> out = silu(x); else out = dy * silu'(x) */
int pf_train_silu(const float* x, const float* dy, float* out, int64_t n, void* stream);
/* vector gating (gvp.py:108-114): dout == NULL -> out = act(gate) * Vu; else dgate (in out_or_dgate) and dVu */
int pf_train_gate(const float* gate, const float* vu, const float* dout, float* out_or_dgate, float* dvu, int64_t rows,
                  int32_t U, int32_t act_sigmoid, void* stream);
/* _norm_no_nan over the 3 components (gvp.py:12-19, 102): dsh == NULL -> out = sh [rows][H]; else out = dVh */
int pf_train_vecnorm(const float* vh, const float* dsh, float* out, int64_t rows, int32_t H, void* stream);
/* nn.LayerNorm(D), eps 1e-5 (gvp.py:157, 161; encoders dynamics_gvp.py:107-117); stats = (mean, rstd) per row */
int pf_train_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* stats, int64_t rows,
                           int32_t D, void* stream);
int pf_train_layernorm_bwd(const float* x, const float* w, const float* stats, const float* dy, float* dx, float* dw,
                           float* db, int64_t rows, int32_t D, void* workspace, size_t workspace_bytes,
                           void* stream); /* dw, db += */
/* vector half of GVPLayerNorm (gvp.py:163-165): dout == NULL -> out = v / vn; else out = dv */
int pf_train_vecln(const float* v, const float* dout, float* out, int64_t rows, int32_t U, void* stream);
/* edges.src[...] (gvp.py:543-545): backward == 0 -> out[e] = x[idx[e]]; else dx[idx[e]] += dout[e] (atomic) */
int pf_train_gather(const float* x_or_dout, const int32_t* idx, float* out_or_dx, int64_t E, int32_t D, int32_t backward,
                    void* stream);
/* the backward without atomics: perm = the edges grouped by source row (a stable sort of idx), row n owns
 * perm[ptr[n] .. ptr[n+1]); dx[n] = sum of dout[perm[e]] in that order (dx is written, not accumulated) */
int pf_train_gather_bwd_sorted(const float* dout, const int32_t* perm, const int32_t* ptr, float* dx, int64_t n_rows,
                               int32_t D, void* stream);
/* fn.mean per destination + cross-etype sum (gvp.py:488-497) over destination-sorted message rows: segment s = rows
 * [ptr[s], ptr[s+1]) -> node seg_dst[s] (NULL: s).  backward == 0 -> out[node] += mean; else dmsg[row] = dout[node]/cnt */
int pf_train_segmean(const float* msg_or_dout, const int32_t* ptr, const int32_t* seg_dst, float* out_or_dmsg,
                     int32_t n_seg, int32_t D, int32_t backward, void* stream);
/* x_diff [E][3] and rbf [E][16] of every edge (gvp.py:472-480); inputs are data: no backward */
int pf_train_edge_geom(const float* src_x, const float* dst_x, const int32_t* src, const int32_t* dst, float* xdiff,
                       float* rbf, int64_t E, void* stream);

/* One whole GVP (gvp.py:89-116) per host call: the primitives above enqueued back to back, forward and backward.
 * feats [M][n], vec [M][3][vi], Wh [vi][h], Wu [h][vo], Wf [no][n+h], Wg [vo][no]; the forward also returns what the
 * backward needs (Vh, Vu, s = [feats | sh], z, f, gates).  dbf / dbg must come zeroed; dgates, dVu, dfz, ds, dVh are scratch. */
int pf_train_gvp_fwd(const float* feats, const float* vec, const float* Wh, const float* Wu, const float* Wf, const float* bf,
                     const float* Wg, const float* bg, int64_t M, int32_t n, int32_t vi, int32_t h, int32_t vo, int32_t no,
                     int32_t act_sigmoid, float* Vh, float* Vu, float* s, float* z, float* f, float* gates, float* vout,
                     void* workspace, size_t workspace_bytes, void* stream);
int pf_train_gvp_bwd(const float* vec, const float* Wh, const float* Wu, const float* Wf, const float* Wg, const float* Vh,
                     const float* Vu, const float* s, const float* z, const float* f, const float* gates, const float* df_out,
                     const float* dvout, int64_t M, int32_t n, int32_t vi, int32_t h, int32_t vo, int32_t no,
                     int32_t act_sigmoid, float* dgates, float* dVu, float* dfz, float* ds, float* dVh, float* dfeats,
                     float* dvec, float* dWh, float* dWu, float* dWf, float* dbf, float* dWg, float* dbg, void* workspace,
                     size_t workspace_bytes, void* stream);
/* workspace / workspace_bytes of the two calls above: scratch of pf_tc_gemm_workspace_bytes(128, 176, 0) bytes for the tensor-
 * core GEMMs (NULL: every contraction runs on the fp32 FFMA kernels). */

/* ---- measurement hooks (bench.py) ----------------------------------------------------------------------
 * pf_launch_count: kernels this library has launched in this process so far.
 * pf_profile_enable(n): n > 0 arms CUDA-event pairs around every kernel site of pf_denoiser /
 * pf_sample_loop (recorded on the launching stream); 0 disarms.  After synchronising the stream,
 * pf_profile_collect sums the elapsed ms and launch count per site and resets the recorder.  Sites:
 * 0 dyn_graph, 1 (unused), 2 (unused), 3 edge ff, 4 edge pf, 5 edge pp, 6 edge fp, 7 update pharm,
 * 8 update prot, 9 noise head, 10 posterior+COM. */
/* tcgen05 building-block self-test: D[128][N] = A[128][K] * B[N][K]^T on the tensor cores, A split into bf16
 * (hi, lo) in TMEM, B given as packed bf16 shared-memory images (layout: pf_tc.cuh); npass 1 (hi*hi) or 3
 * (hi*hi + hi*lo + lo*hi, fp32-accurate). */
int pf_tc_selftest(const float* A, const void* b_hi, const void* b_lo, float* D, int32_t K, int32_t N, int32_t npass,
                   void* stream);
int64_t pf_launch_count(void);
int pf_profile_enable(int32_t max_pairs);
int pf_profile_collect(double* total_ms_host, int32_t* count_host, int32_t n_sites);

#ifdef __cplusplus
}
#endif
#endif
