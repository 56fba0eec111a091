// Graph construction: exclusive scan, K1 static radius graph, K2 per-step dynamic graph, tile planner.
// Integer / byte work bound by memory latency; distances use the canonical fp32 form
// ((dx*dx + dy*dy) + dz*dz) with separately rounded operations so that edge membership and kNN order
// are bit-identical to the CPU oracle.
#include <limits.h>

#include "pf_common.cuh"

#define PF_TRY_RC(call)           \
  do {                            \
    int rc_ = (call);             \
    if (rc_ != PF_OK) return rc_; \
  } while (0)

namespace pf {

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ------------------------------------------------------------------------------------------------ scan
constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanBlock = kScanThreads * kScanItems;  // 2048 elements per block

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    int w = lane < nw ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  total = s_warp[32];
  const int res = s_warp[warp] + inc - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_local_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                  long long n, int* __restrict__ block_sums) {
  __shared__ int s_warp[33];
  const long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    sum += v[i];
  }
  int total;
  int off = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int* __restrict__ block_sums, int n_blocks) {
  __shared__ int s_warp[33];
  int carry = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
    const int i = b0 + threadIdx.x;
    const int v = i < n_blocks ? block_sums[i] : 0;
    int total;
    const int off = block_exclusive_scan(v, s_warp, total);
    if (i < n_blocks) block_sums[i] = carry + off;
    carry += total;
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, long long n,
                                                                const int* __restrict__ block_sums, int n_blocks) {
  const long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanItems;
  const int add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[n_blocks];
}

// ------------------------------------------------------------------------------------------------ K1
// One CTA per segment (graph); thread i owns centre i and walks every candidate j of the segment in
// ascending order through a shared-memory staging tile (all lanes read the same j: broadcast).
constexpr int kRadThreads = 256;
constexpr int kRadStage = 2048;

template <bool FILL>
__global__ void __launch_bounds__(kRadThreads) radius_kernel(const float* __restrict__ x,
                                                             const int* __restrict__ seg_ptr, int n_seg, float r2,
                                                             int max_nbrs, int* __restrict__ deg,
                                                             const int* __restrict__ rowptr, int* __restrict__ col) {
  __shared__ float sx[kRadStage], sy[kRadStage], sz[kRadStage];
  for (int g = blockIdx.x; g < n_seg; g += gridDim.x) {
    const int a = seg_ptr[g], b = seg_ptr[g + 1];
    for (int i0 = a; i0 < b; i0 += kRadThreads) {
      const int i = i0 + threadIdx.x;
      const bool live = i < b;
      float xi = 0.f, yi = 0.f, zi = 0.f;
      if (live) {
        xi = x[3 * (size_t)i];
        yi = x[3 * (size_t)i + 1];
        zi = x[3 * (size_t)i + 2];
      }
      int rank = 0, kept = 0;
      const int out0 = (FILL && live) ? rowptr[i] : 0;
      for (int j0 = a; j0 < b; j0 += kRadStage) {
        const int nj = min(kRadStage, b - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nj; t += kRadThreads) {
          sx[t] = x[3 * (size_t)(j0 + t)];
          sy[t] = x[3 * (size_t)(j0 + t) + 1];
          sz[t] = x[3 * (size_t)(j0 + t) + 2];
        }
        __syncthreads();
        if (live) {
          for (int t = 0; t < nj; ++t) {
            // torch_cluster: first max_nbrs+1 hits in ascending index INCLUDING the centre, then drop the centre
            if (sqdist3(xi, yi, zi, sx[t], sy[t], sz[t]) < r2) {
              ++rank;
              if (rank <= max_nbrs + 1 && j0 + t != i) {
                if (FILL) col[out0 + kept] = j0 + t;
                ++kept;
              }
            }
          }
        }
      }
      if (!FILL && live) deg[i] = kept;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K1, cell list
// The static pp graph is a property of the POCKET, not of its copies: it is built once per distinct pocket with a cell
// list (cell edge c >= r, so the neighbours of a centre live in its 27 surrounding cells) and then replicated per graph
// with node offsets (protein_pharm_dataset.py:234-236 builds it once per pocket; copy_graph + dgl.batch replicate it,
// unorganized_utils.py:28-50).  One CTA per pocket.  Workspace per pocket (ints), carved from one caller buffer:
//   grid[8]           : min x, y, z (float bits), 1/c (float bits), nx, ny, nz, n_cells
//   cell_end[2n + 8]  : after the build, atoms of cell k are sorted[(k ? cell_end[k-1] : 0) .. cell_end[k])
//   atom_cell[n], sorted[n]
// The cell edge starts at r (1 + 1e-4) -- the margin keeps a neighbour at |dx| -> r from landing two cells away under
// fp32 rounding of (x - min) / c -- and grows by 2^(1/3) until the grid has at most 2n + 8 cells (sparse point clouds).
// Edge membership uses the same canonical squared distance as everywhere else; a row's hits are sorted ascending and cut
// as torch_cluster does (first max_nbrs + 1 hits INCLUDING the centre, then the centre is dropped), so the CSR is
// bit-identical to the brute-force kernel's and to the reference's.
constexpr int kCellThreads = 256;
constexpr int kCellRowCap = 160;   // hits of one centre sorted in local memory; denser rows fall back to the ordered scan

__host__ __device__ inline size_t cell_ws_ints(long long n_nodes, int n_seg) { return (size_t)(4 * n_nodes) + (size_t)n_seg * 16; }
struct CellWs {
  int *grid, *cell_end, *atom_cell, *sorted;
};
__device__ __forceinline__ CellWs cell_ws(int* ws, const int* seg_ptr, int n_seg, long long n_nodes, int g) {
  CellWs w;
  const int a = seg_ptr[g];
  w.grid = ws + (size_t)g * 8;
  w.cell_end = ws + (size_t)n_seg * 8 + (size_t)2 * a + (size_t)8 * g;
  w.atom_cell = ws + (size_t)n_seg * 16 + (size_t)2 * n_nodes + a;
  w.sorted = w.atom_cell + n_nodes;
  return w;
}

__global__ void __launch_bounds__(kCellThreads) cell_build_kernel(const float* __restrict__ x, const int* __restrict__ seg_ptr,
                                                                  int n_seg, long long n_nodes, float r, int* __restrict__ ws) {
  __shared__ float s_red[6][kCellThreads / 32];
  __shared__ float s_min[3], s_inv;
  __shared__ int s_dim[4];
  __shared__ int s_warp[33];
  for (int g = blockIdx.x; g < n_seg; g += gridDim.x) {
    const int a = seg_ptr[g], n = seg_ptr[g + 1] - a;
    const CellWs w = cell_ws(ws, seg_ptr, n_seg, n_nodes, g);
    // ---- bounding box
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = threadIdx.x; i < n; i += kCellThreads)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = x[3 * (size_t)(a + i) + c];
        lo[c] = fminf(lo[c], v);
        hi[c] = fmaxf(hi[c], v);
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
        hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
      }
      if ((threadIdx.x & 31) == 0) {
        s_red[c][threadIdx.x >> 5] = lo[c];
        s_red[3 + c][threadIdx.x >> 5] = hi[c];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float mn[3], mx[3];
      for (int c = 0; c < 3; ++c) {
        mn[c] = s_red[c][0];
        mx[c] = s_red[3 + c][0];
        for (int k = 1; k < kCellThreads / 32; ++k) {
          mn[c] = fminf(mn[c], s_red[c][k]);
          mx[c] = fmaxf(mx[c], s_red[3 + c][k]);
        }
        if (n == 0) mn[c] = mx[c] = 0.f;
      }
      float cedge = r * 1.0001f;
      int d[3];
      long long cells;
      for (;;) {
        cells = 1;
        for (int c = 0; c < 3; ++c) {
          const float q = floorf((mx[c] - mn[c]) / cedge);
          d[c] = q < 1048575.f ? (int)q + 1 : 1048576;   // clamp: the loop below grows the cell until the grid is small
          cells *= d[c];
        }
        if (cells <= 2LL * n + 8) break;
        cedge *= 1.2599211f;
      }
      for (int c = 0; c < 3; ++c) {
        s_min[c] = mn[c];
        s_dim[c] = d[c];
        w.grid[c] = __float_as_int(mn[c]);
        w.grid[4 + c] = d[c];
      }
      s_inv = 1.0f / cedge;
      s_dim[3] = (int)cells;
      w.grid[3] = __float_as_int(s_inv);
      w.grid[7] = (int)cells;
    }
    __syncthreads();
    const int n_cells = s_dim[3];
    for (int k = threadIdx.x; k < n_cells; k += kCellThreads) w.cell_end[k] = 0;
    __syncthreads();
    // ---- histogram
    for (int i = threadIdx.x; i < n; i += kCellThreads) {
      int cc[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int v = (int)floorf((x[3 * (size_t)(a + i) + c] - s_min[c]) * s_inv);
        cc[c] = v < 0 ? 0 : (v >= s_dim[c] ? s_dim[c] - 1 : v);
      }
      const int cell = (cc[2] * s_dim[1] + cc[1]) * s_dim[0] + cc[0];
      w.atom_cell[i] = cell;
      atomicAdd(&w.cell_end[cell], 1);
    }
    __syncthreads();
    // ---- exclusive scan of the counts, in place (cell_end[k] = first slot of cell k), block by block
    int carry = 0;
    for (int k0 = 0; k0 < n_cells; k0 += kCellThreads) {
      const int k = k0 + threadIdx.x;
      const int v = k < n_cells ? w.cell_end[k] : 0;
      int total;
      const int off = block_exclusive_scan(v, s_warp, total);
      if (k < n_cells) w.cell_end[k] = carry + off;
      carry += total;
    }
    __syncthreads();
    // ---- fill: the cursor of a cell advances to its end (order inside a cell is arbitrary: rows are sorted later)
    for (int i = threadIdx.x; i < n; i += kCellThreads) w.sorted[atomicAdd(&w.cell_end[w.atom_cell[i]], 1)] = i;
    __syncthreads();
  }
}

template <bool FILL>
__global__ void __launch_bounds__(kCellThreads) cell_radius_kernel(const float* __restrict__ x, const int* __restrict__ seg_ptr,
                                                                   int n_seg, long long n_nodes, float r2, int max_nbrs,
                                                                   int* __restrict__ ws, int* __restrict__ deg,
                                                                   const int* __restrict__ rowptr, int* __restrict__ col) {
  for (int g = blockIdx.x; g < n_seg; g += gridDim.x) {
    const int a = seg_ptr[g], n = seg_ptr[g + 1] - a;
    const CellWs w = cell_ws(ws, seg_ptr, n_seg, n_nodes, g);
    const int nx = w.grid[4], ny = w.grid[5], nz = w.grid[6];
    for (int i = threadIdx.x; i < n; i += kCellThreads) {
      const float xi = x[3 * (size_t)(a + i)], yi = x[3 * (size_t)(a + i) + 1], zi = x[3 * (size_t)(a + i) + 2];
      const int cell = w.atom_cell[i];
      const int cx = cell % nx, cy = (cell / nx) % ny, cz = cell / (nx * ny);
      int hits = 0, lower = 0;          // hits including the centre itself; hits with a smaller index
      int buf[FILL ? kCellRowCap : 1];
      for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = cy + dy;
          if (y < 0 || y >= ny) continue;
          // the three cells of one x-run are contiguous in memory: one range per (dy, dz)
          const int k0 = (z * ny + y) * nx + (cx > 0 ? cx - 1 : 0), k1 = (z * ny + y) * nx + (cx + 1 < nx ? cx + 1 : nx - 1);
          const int beg = k0 > 0 ? w.cell_end[k0 - 1] : 0, end = w.cell_end[k1];
          for (int t = beg; t < end; ++t) {
            const int j = w.sorted[t];
            if (sqdist3(xi, yi, zi, x[3 * (size_t)(a + j)], x[3 * (size_t)(a + j) + 1], x[3 * (size_t)(a + j) + 2]) < r2) {
              if (FILL && hits < kCellRowCap) buf[hits] = j;
              ++hits;
              lower += j < i;
            }
          }
        }
      }
      // torch_cluster: the first max_nbrs + 1 hits in ascending index INCLUDING the centre, then the centre is dropped
      const int kept = (hits < max_nbrs + 1 ? hits : max_nbrs + 1) - (lower < max_nbrs + 1 ? 1 : 0);
      if (!FILL) {
        deg[a + i] = kept;
      } else {
        int* out = col + rowptr[a + i];
        if (hits <= kCellRowCap) {
          for (int u = 1; u < hits; ++u) {   // insertion sort, ascending
            const int v = buf[u];
            int q = u - 1;
            while (q >= 0 && buf[q] > v) {
              buf[q + 1] = buf[q];
              --q;
            }
            buf[q + 1] = v;
          }
          int o = 0;
          for (int u = 0; u < hits && u < max_nbrs + 1; ++u)
            if (buf[u] != i) out[o++] = a + buf[u];
        } else {                             // a very dense row: ordered scan of the whole pocket (the brute-force rule)
          int rank = 0, o = 0;
          for (int j = 0; j < n && rank < max_nbrs + 1; ++j)
            if (sqdist3(xi, yi, zi, x[3 * (size_t)(a + j)], x[3 * (size_t)(a + j) + 1], x[3 * (size_t)(a + j) + 2]) < r2) {
              ++rank;
              if (j != i) out[o++] = a + j;
            }
        }
      }
    }
  }
}

// Replication of the per-pocket CSR over the graphs of a batch (copy_graph + dgl.batch): graph g is a copy of the pocket
// whose nodes start at pk_node0[g] in the pocket arrays; its nodes are [prot_ptr[g], prot_ptr[g+1]) and its edges start at
// edge0[g] (exclusive scan of the copies' edge counts).  One CTA per graph, coalesced stores: 4 B per edge + 8 B per node.
__global__ void __launch_bounds__(256) replicate_csr_kernel(const int* __restrict__ pk_rowptr, const int* __restrict__ pk_col,
                                                            const int* __restrict__ pk_node0, const int* __restrict__ prot_ptr,
                                                            const int* __restrict__ edge0, int n_graphs,
                                                            int* __restrict__ rowptr, int* __restrict__ cnt,
                                                            int* __restrict__ col) {
  for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
    const int p0 = pk_node0[g], n0 = prot_ptr[g], n = prot_ptr[g + 1] - n0, e0 = edge0[g];
    const int pe0 = pk_rowptr[p0], ne = pk_rowptr[p0 + n] - pe0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int rs = pk_rowptr[p0 + i];
      rowptr[n0 + i] = e0 + rs - pe0;
      cnt[n0 + i] = pk_rowptr[p0 + i + 1] - rs;
    }
    const int shift = n0 - p0;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) col[e0 + e] = pk_col[pe0 + e] + shift;
    if (g == n_graphs - 1 && threadIdx.x == 0) rowptr[n0 + n] = e0 + ne;
  }
}

// ------------------------------------------------------------------------------------------------ K2
constexpr int kDynThreads = 128;
constexpr int kDynWarps = kDynThreads / 32;
constexpr int kMaxF = PF_MAX_PHARM_PER_GRAPH;
constexpr int kMaxK = PF_MAX_KNN;

struct DynGraphParams {
  const float *prot_x, *pharm_x;
  const int *prot_ptr, *pharm_ptr;
  int n_graphs;
  float ff_r2;
  int ff_max, k;
  int ff_k;  // > 0: ff edges from knn_graph(pharm x_t, k = ff_k) instead of the radius graph (dynamics_gvp.py:193-194)
  const int* ff_start;
  int *ff_cnt, *ff_col, *pf_cnt, *pf_col, *fp_seg_dst, *fp_seg_start, *fp_seg_cnt, *fp_col;
  unsigned* status;
};

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

// ff edges of pharmacophore node i of a graph (one warp): radius_graph(pharm x_t, r, max) or, with ff_k > 0, the kNN graph
// (dynamics_gvp.py:193-196).  fx / fy / fz hold the graph's nf pharmacophore coordinates; writes ff_col / ff_cnt.
__device__ __forceinline__ void ff_edges_of_node(const DynGraphParams& p, const float* fx, const float* fy, const float* fz,
                                                 const int nf, const int fa, const int i, const int lane) {
  const float qx = fx[i], qy = fy[i], qz = fz[i];
  int rank_run = 0, kept_run = 0;
  const int out0 = p.ff_start[fa + i];
  if (p.ff_k > 0) {
    // ---- ff: knn_graph(pharm x_t, k = ff_k) (dynamics_gvp.py:194) = the ff_k + 1 nearest nodes INCLUDING the centre,
    // ordered by (distance, index), then the self pair is dropped.  nf <= 128: four candidates per lane; the key
    // (distance bits, index) is unique, so each round of the warp-wide minimum selects exactly one candidate.
    unsigned long long key[kMaxF / 32];
    unsigned sel = 0;
#pragma unroll
    for (int t = 0; t < kMaxF / 32; ++t) {
      const int j = lane + 32 * t;
      key[t] = j < nf ? (((unsigned long long)__float_as_uint(sqdist3(qx, qy, qz, fx[j], fy[j], fz[j])) << 32) | (unsigned)j)
                      : ~0ull;
    }
    const int take = p.ff_k + 1 < nf ? p.ff_k + 1 : nf;
    for (int r = 0; r < take; ++r) {
      unsigned long long mine = ~0ull;
#pragma unroll
      for (int t = 0; t < kMaxF / 32; ++t)
        if (!(sel & (1u << t)) && key[t] < mine) mine = key[t];
      unsigned long long best = mine;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = shfl_xor_u64(best, o);
        best = other < best ? other : best;
      }
      if (mine == best && best != ~0ull) {
#pragma unroll
        for (int t = 0; t < kMaxF / 32; ++t)
          if (key[t] == best) sel |= 1u << t;
      }
    }
#pragma unroll
    for (int t = 0; t < kMaxF / 32; ++t) {   // neighbours in ascending index
      const int j = lane + 32 * t;
      const bool keep = (sel & (1u << t)) && j != i;
      const unsigned kb = __ballot_sync(0xffffffffu, keep);
      if (keep) p.ff_col[out0 + kept_run + __popc(kb & ((1u << lane) - 1u))] = fa + j;
      kept_run += __popc(kb);
    }
  } else
  // ---- ff: radius_graph(pharm x_t, r, max) (dynamics_gvp.py:196); centre i, neighbours ascending
  for (int j0 = 0; j0 < nf; j0 += 32) {
    const int j = j0 + lane;
    const bool hit = j < nf && sqdist3(qx, qy, qz, fx[j < nf ? j : 0], fy[j < nf ? j : 0], fz[j < nf ? j : 0]) < p.ff_r2;
    const unsigned hb = __ballot_sync(0xffffffffu, hit);
    const int rank = rank_run + __popc(hb & ((1u << lane) - 1u)) + 1;
    const bool keep = hit && rank <= p.ff_max + 1 && j != i;
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    if (keep) p.ff_col[out0 + kept_run + __popc(kb & ((1u << lane) - 1u))] = fa + j;
    rank_run += __popc(hb);
    kept_run += __popc(kb);
  }
  if (lane == 0) p.ff_cnt[fa + i] = kept_run;
}

template <int K>
__global__ void __launch_bounds__(kDynThreads) dyn_graph_kernel(const DynGraphParams p) {
  __shared__ float fx[kMaxF], fy[kMaxF], fz[kMaxF];
  __shared__ int e_dst[kMaxF * K];   // prot id of pf edge (i, j), -1 if absent
  __shared__ int s_dst[kMaxF * K];   // fp edges sorted by (dst, src)
  __shared__ int s_src[kMaxF * K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = p.k;
  for (int g = blockIdx.x; g < p.n_graphs; g += gridDim.x) {
    const int pa = p.prot_ptr[g], pb = p.prot_ptr[g + 1];
    const int fa = p.pharm_ptr[g], fb = p.pharm_ptr[g + 1];
    const int nf = fb - fa, np_ = pb - pa;
    if (nf > kMaxF) {
      if (threadIdx.x == 0) atomicOr(p.status, PF_DEV_GRAPH_TOO_LARGE);
      continue;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nf; i += kDynThreads) {
      fx[i] = p.pharm_x[3 * (size_t)(fa + i)];
      fy[i] = p.pharm_x[3 * (size_t)(fa + i) + 1];
      fz[i] = p.pharm_x[3 * (size_t)(fa + i) + 2];
    }
    __syncthreads();
    const int kk = min(k, np_);

    for (int i = warp; i < nf; i += kDynWarps) {
      // ---- pf: k nearest prot atoms of pharm node i (knn(prot, pharm, k), dynamics_gvp.py:202)
      float bd[K];
      int bi[K];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        bd[j] = __int_as_float(0x7f800000);
        bi[j] = INT_MAX;
      }
      const float qx = fx[i], qy = fy[i], qz = fz[i];
      for (int c = pa + lane; c < pb; c += 32) {
        const float d = sqdist3(p.prot_x[3 * (size_t)c], p.prot_x[3 * (size_t)c + 1], p.prot_x[3 * (size_t)c + 2],
                                qx, qy, qz);
        if (d < bd[K - 1]) {  // strict: an equal distance never displaces an earlier (lower) index
          bd[K - 1] = d;
          bi[K - 1] = c;
#pragma unroll
          for (int j = K - 1; j > 0; --j) {
            if (bd[j] < bd[j - 1]) {
              const float td = bd[j];
              bd[j] = bd[j - 1];
              bd[j - 1] = td;
              const int ti = bi[j];
              bi[j] = bi[j - 1];
              bi[j - 1] = ti;
            }
          }
        }
      }
      // k-way merge across lanes on the key (distance bits, index); distances are >= 0 so the bit
      // pattern orders like the value
      for (int j = 0; j < k; ++j) {
        int sel = -1;
        if (j < kk) {
          const unsigned long long mine =
              ((unsigned long long)__float_as_uint(bd[0]) << 32) | (unsigned)bi[0];
          unsigned long long best = mine;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = shfl_xor_u64(best, o);
            best = other < best ? other : best;
          }
          sel = (int)(unsigned)(best & 0xffffffffull);
          if (mine == best) {  // indices are unique, so exactly one lane pops its head
#pragma unroll
            for (int q = 0; q + 1 < K; ++q) {
              bd[q] = bd[q + 1];
              bi[q] = bi[q + 1];
            }
            bd[K - 1] = __int_as_float(0x7f800000);
            bi[K - 1] = INT_MAX;
          }
        }
        if (lane == 0) {
          p.pf_col[(size_t)k * (fa + i) + j] = sel;
          e_dst[i * k + j] = sel;
        }
      }
      if (lane == 0) p.pf_cnt[fa + i] = kk;

      ff_edges_of_node(p, fx, fy, fz, nf, fa, i, lane);
    }
    __syncthreads();

    // ---- fp: the pf edges reversed, sorted by (prot id, pharm id) by rank counting (E <= k*nf is small)
    const int E = nf * kk;
    for (int e = threadIdx.x; e < E; e += kDynThreads) {
      const int i = e / kk, j = e - i * kk;
      const int d = e_dst[i * k + j];
      int rank = 0;
      for (int i2 = 0; i2 < nf; ++i2)
        for (int j2 = 0; j2 < kk; ++j2) {
          const int d2 = e_dst[i2 * k + j2];
          rank += (d2 < d) || (d2 == d && i2 < i);
        }
      s_dst[rank] = d;
      s_src[rank] = fa + i;
    }
    __syncthreads();
    const size_t base = (size_t)k * fa;
    for (int e = threadIdx.x; e < nf * k; e += kDynThreads) {
      if (e < E) p.fp_col[base + e] = s_src[e];
      // segment heads: position e starts a segment iff its dst differs from its predecessor's
      int seg = -1, cnt = 0;
      if (e < E && (e == 0 || s_dst[e] != s_dst[e - 1])) {
        seg = 0;
        for (int q = 1; q <= e; ++q) seg += s_dst[q] != s_dst[q - 1];
        cnt = 1;
        while (e + cnt < E && s_dst[e + cnt] == s_dst[e]) ++cnt;
        p.fp_seg_dst[base + seg] = s_dst[e];
        p.fp_seg_start[base + seg] = (int)base + e;
        p.fp_seg_cnt[base + seg] = cnt;
      }
    }
    __syncthreads();
    // unused segment slots: count the segments, clear the rest
    if (threadIdx.x == 0) {
      int nseg = E > 0 ? 1 : 0;
      for (int q = 1; q < E; ++q) nseg += s_dst[q] != s_dst[q - 1];
      s_src[0] = nseg;  // reuse as broadcast slot (fp_col already written)
    }
    __syncthreads();
    const int nseg = s_src[0];
    for (int s = nseg + threadIdx.x; s < nf * k; s += kDynThreads) {
      p.fp_seg_dst[base + s] = pa;
      p.fp_seg_start[base + s] = (int)base;
      p.fp_seg_cnt[base + s] = 0;
    }
  }
}

// K2 variant for pf_k == 0 (dynamics_gvp.py:210-216): pf / fp edges from radius(x = pharm, y = prot, r, max_num_neighbors) --
// every protein atom (the query) keeps the pharmacophore nodes of its graph with squared distance < r * r, ascending index, at
// most max_nbrs of them; pf = (prot -> pharm), fp = the reverse.  A pharmacophore node then has up to n_prot(graph) in-edges,
// more than one edge tile holds, so its pf segment is cut into sub-segments of PF_TC_TILE_ROWS rows: the edge kernels write
// one mean per sub-segment and pf_combine_subsegments folds them into the node's aggregate.  fp segments are one per protein
// atom (identity destination).  One CTA per graph.  Capacities are static per batch: pharmacophore node i of a graph with np
// atoms owns pf_col [pf_start[i], + np) and the sub-segment slots [sub_ptr[i], sub_ptr[i + 1]) = ceil(np / rows) of them;
// protein atom c of graph g owns fp_col [fp_base[g] + c * min(nf, max_nbrs), ...).
struct DynRadiusParams {
  DynGraphParams d;       // ff part and the coordinate / ptr arrays; the kNN outputs of d are unused
  float pf_r2;
  int pf_max, sub_rows;
  const int *pf_start, *sub_ptr;  // [n_pharm], [n_pharm + 1] static
  const int* fp_base;             // [n_graphs] static
  int *sub_start, *sub_cnt;
  float* sub_x;  // [n_sub][3] coordinates of the sub-segment's destination node: the dst_x array of the pf edge kernels
};

__global__ void __launch_bounds__(kDynThreads) dyn_graph_radius_kernel(const DynRadiusParams q) {
  __shared__ float fx[kMaxF], fy[kMaxF], fz[kMaxF];
  const DynGraphParams& p = q.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = blockIdx.x; g < p.n_graphs; g += gridDim.x) {
    const int pa = p.prot_ptr[g], pb = p.prot_ptr[g + 1];
    const int fa = p.pharm_ptr[g], fb = p.pharm_ptr[g + 1];
    const int nf = fb - fa, np_ = pb - pa;
    if (nf > kMaxF) {
      if (threadIdx.x == 0) atomicOr(p.status, PF_DEV_GRAPH_TOO_LARGE);
      continue;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nf; i += kDynThreads) {
      fx[i] = p.pharm_x[3 * (size_t)(fa + i)];
      fy[i] = p.pharm_x[3 * (size_t)(fa + i) + 1];
      fz[i] = p.pharm_x[3 * (size_t)(fa + i) + 2];
    }
    __syncthreads();
    const int n_sub = (np_ + q.sub_rows - 1) / q.sub_rows;
    for (int i = warp; i < nf; i += kDynWarps) {
      ff_edges_of_node(p, fx, fy, fz, nf, fa, i, lane);
      // ---- pf: in-edges of pharmacophore node i = the protein atoms that keep it, ascending atom index
      const float qx = fx[i], qy = fy[i], qz = fz[i];
      const int out0 = q.pf_start[fa + i];
      int run = 0;
      for (int c0 = 0; c0 < np_; c0 += 32) {
        const int c = c0 + lane;
        bool hit = false;
        if (c < np_) {
          const float px = p.prot_x[3 * (size_t)(pa + c)], py = p.prot_x[3 * (size_t)(pa + c) + 1],
                      pz = p.prot_x[3 * (size_t)(pa + c) + 2];
          hit = sqdist3(px, py, pz, qx, qy, qz) < q.pf_r2;
          if (hit && i >= q.pf_max) {   // the atom's cap: node i is kept iff fewer than max_nbrs lower-index nodes hit
            int rank = 0;
            for (int j = 0; j < i; ++j) rank += sqdist3(px, py, pz, fx[j], fy[j], fz[j]) < q.pf_r2;
            hit = rank < q.pf_max;
          }
        }
        const unsigned hb = __ballot_sync(0xffffffffu, hit);
        if (hit) p.pf_col[out0 + run + __popc(hb & ((1u << lane) - 1u))] = pa + c;
        run += __popc(hb);
      }
      const int sub0 = q.sub_ptr[fa + i];
      if (lane == 0) p.pf_cnt[fa + i] = run;
      for (int s = lane; s < n_sub; s += 32) {
        const int left = run - s * q.sub_rows;
        q.sub_start[sub0 + s] = out0 + s * q.sub_rows;
        q.sub_cnt[sub0 + s] = left < 0 ? 0 : (left > q.sub_rows ? q.sub_rows : left);
        q.sub_x[3 * (size_t)(sub0 + s)] = qx;
        q.sub_x[3 * (size_t)(sub0 + s) + 1] = qy;
        q.sub_x[3 * (size_t)(sub0 + s) + 2] = qz;
      }
    }
    // ---- fp: in-edges of protein atom c = its kept pharmacophore nodes, ascending node index
    const int cap = nf < q.pf_max ? nf : q.pf_max;
    for (int c = threadIdx.x; c < np_; c += kDynThreads) {
      const float px = p.prot_x[3 * (size_t)(pa + c)], py = p.prot_x[3 * (size_t)(pa + c) + 1],
                  pz = p.prot_x[3 * (size_t)(pa + c) + 2];
      const int out0 = q.fp_base[g] + c * cap;
      int o = 0;
      for (int j = 0; j < nf && o < q.pf_max; ++j)
        if (sqdist3(px, py, pz, fx[j], fy[j], fz[j]) < q.pf_r2) p.fp_col[out0 + o++] = fa + j;
      p.fp_seg_start[pa + c] = out0;
      p.fp_seg_cnt[pa + c] = o;
    }
  }
}

// agg[d] (+)= sum over the sub-segments s in [sub_ptr[d], sub_ptr[d+1]) of sub[s] * sub_cnt[s] * w, w = 1 / tot_cnt[d] (the
// mean over all in-edges, inv_norm == 0), inv_norm (numeric message_norm: SUM / norm) or inv_norm_node[d] (message_norm = 0).  One warp per destination; empty
// sub-segments are skipped (their rows of `sub` may never have been written), a destination without in-edges gets zero.
__global__ void __launch_bounds__(256) combine_subsegments_kernel(const float* __restrict__ sub_h, const float* __restrict__ sub_v,
                                                                  const int* __restrict__ sub_cnt, const int* __restrict__ sub_ptr,
                                                                  const int* __restrict__ tot_cnt, long long n_dst,
                                                                  float inv_norm, const float* __restrict__ inv_norm_node,
                                                                  float* __restrict__ agg_h,
                                                                  float* __restrict__ agg_v, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; d < n_dst; d += n_warps) {
    const int tot = tot_cnt[d];
    const float w = inv_norm_node != nullptr ? inv_norm_node[d]
                                             : (inv_norm != 0.f ? inv_norm : (tot > 0 ? __fdividef(1.0f, (float)tot) : 0.f));
    float4* oh = reinterpret_cast<float4*>(agg_h + d * kHidden + 4 * lane);
    float4* ov = reinterpret_cast<float4*>(agg_v + d * kVRow + 4 * lane);
    float4 h = accumulate ? *oh : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = (accumulate && lane < kVRow / 4) ? *ov : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = sub_ptr[d]; s < sub_ptr[d + 1]; ++s) {
      const int cnt = sub_cnt[s];
      if (cnt == 0) continue;
      const float sc = (float)cnt * w;
      const float4 a = *reinterpret_cast<const float4*>(sub_h + (long long)s * kHidden + 4 * lane);
      h = make_float4(fmaf(a.x, sc, h.x), fmaf(a.y, sc, h.y), fmaf(a.z, sc, h.z), fmaf(a.w, sc, h.w));
      if (lane < kVRow / 4) {
        const float4 b = *reinterpret_cast<const float4*>(sub_v + (long long)s * kVRow + 4 * lane);
        v = make_float4(fmaf(b.x, sc, v.x), fmaf(b.y, sc, v.y), fmaf(b.z, sc, v.z), fmaf(b.w, sc, v.w));
      }
    }
    *oh = h;
    if (lane < kVRow / 4) *ov = v;
  }
}


// ------------------------------------------------------------------------------------------------ planner
__global__ void __launch_bounds__(128) plan_tiles_kernel(const int* __restrict__ seg_cnt,
                                                         const int* __restrict__ chunk_ptr, int n_chunks,
                                                         int skip_empty, int tile_rows,
                                                         int* __restrict__ tiles, int max_tiles,
                                                         int* __restrict__ n_tiles, unsigned* __restrict__ status) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  int s = chunk_ptr[c];
  const int s_end = chunk_ptr[c + 1];
  while (s < s_end) {
    int rows = 0, e = s;
    while (e < s_end && e - s < tile_rows) {
      const int cnt = seg_cnt[e];
      if (rows + cnt > tile_rows) break;
      rows += cnt;
      ++e;
    }
    if (e == s) {  // a single destination with more in-edges than a tile holds
      atomicOr(status, PF_DEV_DEGREE_OVERFLOW);
      s = s + 1;
      continue;
    }
    if (!(skip_empty && rows == 0)) {
      const int slot = atomicAdd(n_tiles, 1);
      if (slot < max_tiles) {
        tiles[2 * slot] = s;
        tiles[2 * slot + 1] = e;
      } else {
        atomicOr(status, PF_DEV_TILE_OVERFLOW);
      }
    }
    s = e;
  }
}

// The three per-step plans (ff, pf, fp) in ONE launch: blockIdx.y selects the plan, ONE WARP per chunk.  The greedy packing of
// plan_tiles_kernel (consecutive segments while they fit tile_rows edges and tile_rows segments) is a prefix-sum question: the
// warp loads a window of tile_rows segment sizes (tile_rows / 32 consecutive ones per lane), scans it, and the segments whose
// inclusive prefix stays within tile_rows form the tile -- a handful of instructions per tile instead of one dependent global
// load per segment (the one-thread-per-chunk walk took ~47 us per plan whatever the batch size and forced small chunks, i.e.
// a poorly filled last tile per chunk).  Same tiles as the sequential walk; their order in the list is arbitrary, as before.
struct PlanDesc {
  const int *seg_cnt, *chunk_ptr;
  int n_chunks, skip_empty;
  int *tiles, *n_tiles;
};
struct Plan3 {
  PlanDesc d[3];
};
__global__ void __launch_bounds__(128) plan_tiles3_kernel(const Plan3 p, int tile_rows, int max_tiles, unsigned* __restrict__ status) {
  const PlanDesc& d = p.d[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= d.n_chunks) return;
  int s = d.chunk_ptr[c];
  const int s_end = d.chunk_ptr[c + 1];
  const int per = tile_rows >> 5;   // 2 or 4 segments per lane
  while (s < s_end) {
    int incl[4], tot = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = s + lane * per + j;
      // past the chunk (or the window): a size that can never fit, so nothing behind it fits either
      const int cnt = (j < per && idx < s_end) ? d.seg_cnt[idx] : tile_rows + 1;
      tot += j < per ? cnt : 0;
      incl[j] = tot;
    }
    int base = tot;   // exclusive scan of the lanes' totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, base, o);
      if (lane >= o) base += t;
    }
    base -= tot;
    int n_fit = 0, rows = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < per && base + incl[j] <= tile_rows) {
        ++n_fit;
        rows = base + incl[j];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_fit += __shfl_xor_sync(0xffffffffu, n_fit, o);
      rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, o));
    }
    if (n_fit == 0) {  // a single destination with more in-edges than a tile holds
      if (lane == 0) atomicOr(status, PF_DEV_DEGREE_OVERFLOW);
      s = s + 1;
      continue;
    }
    if (lane == 0 && !(d.skip_empty && rows == 0)) {
      const int slot = atomicAdd(d.n_tiles, 1);
      if (slot < max_tiles) {
        d.tiles[2 * slot] = s;
        d.tiles[2 * slot + 1] = s + n_fit;
      } else {
        atomicOr(status, PF_DEV_TILE_OVERFLOW);
      }
    }
    s += n_fit;
  }
}

// Ordered variant for the static pp plan: tiles come out in chunk (graph) order, a graph's tiles contiguous.  The persistent
// edge kernels hand tile k to CTA k mod grid, so with this order the ~25 tiles of a 400-atom graph run on 25 SMs AT THE SAME
// TIME and the graph's source rows (282 KB) are fetched from DRAM once and then hit L2 for their other ~6.5 uses.  With the
// atomic planner above the list interleaves all graphs of the batch (threads claim slots in lockstep): ncu showed 697 B of
// DRAM reads per edge at 23 M edges -- every gather a DRAM access, 4x the compulsory traffic (profiles/r02_*).
// Pass 1 (count != nullptr): tiles per chunk.  Pass 2: fill at tile_off[c] (exclusive scan of the counts).
__global__ void __launch_bounds__(128) plan_tiles_ordered_kernel(const int* __restrict__ seg_cnt, const int* __restrict__ chunk_ptr,
                                                                 int n_chunks, int skip_empty, int tile_rows,
                                                                 int* __restrict__ count, const int* __restrict__ tile_off,
                                                                 int* __restrict__ tiles, int max_tiles, int* __restrict__ n_tiles,
                                                                 unsigned* __restrict__ status) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  int s = chunk_ptr[c];
  const int s_end = chunk_ptr[c + 1];
  int n = 0;
  const int base = count ? 0 : tile_off[c];
  while (s < s_end) {
    int rows = 0, e = s;
    while (e < s_end && e - s < tile_rows) {
      const int cnt = seg_cnt[e];
      if (rows + cnt > tile_rows) break;
      rows += cnt;
      ++e;
    }
    if (e == s) {  // a single destination with more in-edges than a tile holds
      if (count) atomicOr(status, PF_DEV_DEGREE_OVERFLOW);
      s = s + 1;
      continue;
    }
    if (!(skip_empty && rows == 0)) {
      if (!count) {
        if (base + n < max_tiles) {
          tiles[2 * (base + n)] = s;
          tiles[2 * (base + n) + 1] = e;
        } else {
          atomicOr(status, PF_DEV_TILE_OVERFLOW);
        }
      }
      ++n;
    }
    s = e;
  }
  if (count) count[c] = n;
  else if (c == n_chunks - 1) *n_tiles = base + n < max_tiles ? base + n : max_tiles;
}

__global__ void zero_i32_kernel(int* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

}  // namespace pf

using namespace pf;

extern "C" size_t pf_scan_workspace_bytes(int64_t n) {
  const int64_t blocks = (n + kScanBlock - 1) / kScanBlock;
  return (size_t)(blocks + 1) * sizeof(int32_t);
}

extern "C" int pf_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  PF_CHECK_ARG(in && out && n >= 0, "pf_exclusive_scan_i32: null pointer or negative n");
  if (workspace_bytes < pf_scan_workspace_bytes(n) || !workspace) {
    set_error("pf_exclusive_scan_i32: workspace too small (%zu < %zu)", workspace_bytes, pf_scan_workspace_bytes(n));
    return PF_ERR_WORKSPACE;
  }
  const int blocks = (int)((n + kScanBlock - 1) / kScanBlock);
  int* sums = static_cast<int*>(workspace);
  cudaStream_t st = as_stream(stream);
  if (blocks == 0) {
    zero_i32_kernel<<<1, 32, 0, st>>>(out, 1);
    PF_CHECK_LAUNCH("pf_exclusive_scan_i32");
    return PF_OK;
  }
  scan_local_kernel<<<blocks, kScanThreads, 0, st>>>(in, out, n, sums);
  scan_sums_kernel<<<1, kScanThreads, 0, st>>>(sums, blocks);
  scan_add_kernel<<<blocks, kScanThreads, 0, st>>>(out, n, sums, blocks);
  PF_CHECK_LAUNCH("pf_exclusive_scan_i32");
  return PF_OK;
}

extern "C" int pf_radius_count(const float* x, const int32_t* seg_ptr, int32_t n_seg, float r, int32_t max_nbrs,
                               int32_t* deg, void* stream) {
  PF_CHECK_ARG(x && seg_ptr && deg && max_nbrs >= 0, "pf_radius_count: null pointer");
  if (n_seg <= 0) return PF_OK;
  const int grid = n_seg < 16 * num_sms() ? n_seg : 16 * num_sms();
  radius_kernel<false><<<grid, kRadThreads, 0, as_stream(stream)>>>(x, seg_ptr, n_seg, r * r, max_nbrs, deg, nullptr,
                                                                    nullptr);
  PF_CHECK_LAUNCH("pf_radius_count");
  return PF_OK;
}

extern "C" int pf_radius_fill(const float* x, const int32_t* seg_ptr, int32_t n_seg, float r, int32_t max_nbrs,
                              const int32_t* rowptr, int32_t* col, void* stream) {
  PF_CHECK_ARG(x && seg_ptr && rowptr && col && max_nbrs >= 0, "pf_radius_fill: null pointer");
  if (n_seg <= 0) return PF_OK;
  const int grid = n_seg < 16 * num_sms() ? n_seg : 16 * num_sms();
  radius_kernel<true><<<grid, kRadThreads, 0, as_stream(stream)>>>(x, seg_ptr, n_seg, r * r, max_nbrs, nullptr,
                                                                   rowptr, col);
  PF_CHECK_LAUNCH("pf_radius_fill");
  return PF_OK;
}

extern "C" size_t pf_cell_radius_workspace_bytes(int64_t n_nodes, int32_t n_seg) {
  return cell_ws_ints(n_nodes, n_seg) * sizeof(int);
}

static int cell_radius_common(const float* x, const int32_t* seg_ptr, int32_t n_seg, int64_t n_nodes, float r, int32_t max_nbrs,
                              void* workspace, size_t workspace_bytes, const char* who) {
  if (!(x && seg_ptr && workspace && max_nbrs >= 0 && r > 0.f)) {
    set_error("bad argument: %s: null pointer / bad radius", who);
    return PF_ERR_BAD_ARG;
  }
  if (workspace_bytes < pf_cell_radius_workspace_bytes(n_nodes, n_seg)) {
    set_error("%s: workspace too small (%zu < %zu bytes)", who, workspace_bytes, pf_cell_radius_workspace_bytes(n_nodes, n_seg));
    return PF_ERR_WORKSPACE;
  }
  return PF_OK;
}

extern "C" int pf_cell_radius_count(const float* x, const int32_t* seg_ptr, int32_t n_seg, int64_t n_nodes, float r,
                                    int32_t max_nbrs, void* workspace, size_t workspace_bytes, int32_t* deg, void* stream) {
  PF_TRY_RC(cell_radius_common(x, seg_ptr, n_seg, n_nodes, r, max_nbrs, workspace, workspace_bytes, "pf_cell_radius_count"));
  PF_CHECK_ARG(deg != nullptr, "pf_cell_radius_count: null deg");
  if (n_seg <= 0) return PF_OK;
  const int grid = n_seg < 16 * num_sms() ? n_seg : 16 * num_sms();
  cell_build_kernel<<<grid, kCellThreads, 0, as_stream(stream)>>>(x, seg_ptr, n_seg, n_nodes, r, static_cast<int*>(workspace));
  PF_CHECK_LAUNCH("pf_cell_radius_count (build)");
  cell_radius_kernel<false><<<grid, kCellThreads, 0, as_stream(stream)>>>(x, seg_ptr, n_seg, n_nodes, r * r, max_nbrs,
                                                                          static_cast<int*>(workspace), deg, nullptr, nullptr);
  PF_CHECK_LAUNCH("pf_cell_radius_count");
  return PF_OK;
}

extern "C" int pf_cell_radius_fill(const float* x, const int32_t* seg_ptr, int32_t n_seg, int64_t n_nodes, float r,
                                   int32_t max_nbrs, void* workspace, size_t workspace_bytes, const int32_t* rowptr,
                                   int32_t* col, void* stream) {
  PF_TRY_RC(cell_radius_common(x, seg_ptr, n_seg, n_nodes, r, max_nbrs, workspace, workspace_bytes, "pf_cell_radius_fill"));
  PF_CHECK_ARG(rowptr && col, "pf_cell_radius_fill: null pointer");
  if (n_seg <= 0) return PF_OK;
  const int grid = n_seg < 16 * num_sms() ? n_seg : 16 * num_sms();
  cell_radius_kernel<true><<<grid, kCellThreads, 0, as_stream(stream)>>>(x, seg_ptr, n_seg, n_nodes, r * r, max_nbrs,
                                                                         static_cast<int*>(workspace), nullptr, rowptr, col);
  PF_CHECK_LAUNCH("pf_cell_radius_fill");
  return PF_OK;
}

extern "C" int pf_replicate_csr(const int32_t* pk_rowptr, const int32_t* pk_col, const int32_t* pk_node0,
                                const int32_t* prot_ptr, const int32_t* edge0, int32_t n_graphs, int32_t* rowptr,
                                int32_t* cnt, int32_t* col, void* stream) {
  PF_CHECK_ARG(pk_rowptr && pk_col && pk_node0 && prot_ptr && edge0 && rowptr && cnt && col, "pf_replicate_csr: null pointer");
  if (n_graphs <= 0) return PF_OK;
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  replicate_csr_kernel<<<grid, 256, 0, as_stream(stream)>>>(pk_rowptr, pk_col, pk_node0, prot_ptr, edge0, n_graphs, rowptr,
                                                            cnt, col);
  PF_CHECK_LAUNCH("pf_replicate_csr");
  return PF_OK;
}

extern "C" int pf_dyn_graph(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x,
                            const int32_t* pharm_ptr, int32_t n_graphs, float ff_r, int32_t ff_max_nbrs,
                            int32_t pf_k, const int32_t* ff_start, int32_t* ff_cnt, int32_t* ff_col,
                            int32_t* pf_cnt, int32_t* pf_col, int32_t* fp_seg_dst, int32_t* fp_seg_start,
                            int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status, void* stream) {
  return pf_dyn_graph_ffk(prot_x, prot_ptr, pharm_x, pharm_ptr, n_graphs, ff_r, ff_max_nbrs, 0, pf_k, ff_start, ff_cnt,
                          ff_col, pf_cnt, pf_col, fp_seg_dst, fp_seg_start, fp_seg_cnt, fp_col, dev_status, stream);
}

extern "C" int pf_dyn_graph_ffk(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x,
                                const int32_t* pharm_ptr, int32_t n_graphs, float ff_r, int32_t ff_max_nbrs, int32_t ff_k,
                                int32_t pf_k, const int32_t* ff_start, int32_t* ff_cnt, int32_t* ff_col,
                                int32_t* pf_cnt, int32_t* pf_col, int32_t* fp_seg_dst, int32_t* fp_seg_start,
                                int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status, void* stream) {
  PF_CHECK_ARG(ff_k >= 0, "pf_dyn_graph: ff_k < 0");
  PF_CHECK_ARG(prot_x && prot_ptr && pharm_x && pharm_ptr && ff_start && ff_cnt && ff_col && pf_cnt && pf_col &&
                   fp_seg_dst && fp_seg_start && fp_seg_cnt && fp_col && dev_status,
               "pf_dyn_graph: null pointer");
  if (pf_k < 1 || pf_k > kMaxK) {
    set_error("pf_dyn_graph: pf_k=%d unsupported (1..%d; pf_k = 0, the radius variant of the pf edges, is pf_dyn_graph_radius)",
              pf_k, kMaxK);
    return PF_ERR_UNSUPPORTED;
  }
  if (n_graphs <= 0) return PF_OK;
  DynGraphParams p{prot_x, pharm_x, prot_ptr, pharm_ptr,  n_graphs,     ff_r * ff_r, ff_max_nbrs, pf_k,      ff_k, ff_start,
                   ff_cnt, ff_col,  pf_cnt,   pf_col,     fp_seg_dst,   fp_seg_start, fp_seg_cnt, fp_col, dev_status};
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  if (pf_k <= 8)
    dyn_graph_kernel<8><<<grid, kDynThreads, 0, as_stream(stream)>>>(p);
  else
    dyn_graph_kernel<16><<<grid, kDynThreads, 0, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_dyn_graph");
  return PF_OK;
}

extern "C" int pf_dyn_graph_radius(const float* prot_x, const int32_t* prot_ptr, const float* pharm_x, const int32_t* pharm_ptr,
                                   int32_t n_graphs, float ff_r, int32_t ff_max_nbrs, int32_t ff_k, float pf_r,
                                   int32_t pf_max_nbrs, int32_t sub_rows, const int32_t* ff_start, int32_t* ff_cnt,
                                   int32_t* ff_col, const int32_t* pf_start, const int32_t* sub_ptr, const int32_t* fp_base,
                                   int32_t* pf_cnt, int32_t* pf_col, int32_t* sub_start, int32_t* sub_cnt, float* sub_x,
                                   int32_t* fp_seg_start, int32_t* fp_seg_cnt, int32_t* fp_col, uint32_t* dev_status,
                                   void* stream) {
  PF_CHECK_ARG(ff_k >= 0 && pf_max_nbrs >= 1 && pf_r > 0.f, "pf_dyn_graph_radius: ff_k < 0, pf_max_nbrs < 1 or pf_r <= 0");
  PF_CHECK_ARG(sub_rows == PF_TILE_ROWS || sub_rows == PF_TC_TILE_ROWS, "pf_dyn_graph_radius: sub_rows must be 64 or 128");
  PF_CHECK_ARG(prot_x && prot_ptr && pharm_x && pharm_ptr && ff_start && ff_cnt && ff_col && pf_start && sub_ptr && fp_base &&
                   pf_cnt && pf_col && sub_start && sub_cnt && sub_x && fp_seg_start && fp_seg_cnt && fp_col && dev_status,
               "pf_dyn_graph_radius: null pointer");
  if (n_graphs <= 0) return PF_OK;
  DynRadiusParams q{{prot_x, pharm_x, prot_ptr, pharm_ptr, n_graphs, ff_r * ff_r, ff_max_nbrs, 0, ff_k, ff_start, ff_cnt, ff_col,
                     pf_cnt, pf_col, nullptr, fp_seg_start, fp_seg_cnt, fp_col, dev_status},
                    pf_r * pf_r, pf_max_nbrs, sub_rows, pf_start, sub_ptr, fp_base, sub_start, sub_cnt, sub_x};
  const int grid = n_graphs < 32 * num_sms() ? n_graphs : 32 * num_sms();
  dyn_graph_radius_kernel<<<grid, kDynThreads, 0, as_stream(stream)>>>(q);
  PF_CHECK_LAUNCH("pf_dyn_graph_radius");
  return PF_OK;
}

extern "C" int pf_combine_subsegments(const float* sub_h, const float* sub_v, const int32_t* sub_cnt, const int32_t* sub_ptr,
                                      const int32_t* tot_cnt, int64_t n_dst, float inv_norm, const float* inv_norm_node,
                                      float* agg_h, float* agg_v, int32_t accumulate, void* stream) {
  PF_CHECK_ARG(sub_h && sub_v && sub_cnt && sub_ptr && tot_cnt && agg_h && agg_v, "pf_combine_subsegments: null pointer");
  PF_CHECK_ARG(inv_norm >= 0.f, "pf_combine_subsegments: negative inv_norm");
  if (n_dst <= 0) return PF_OK;
  const long long blocks = (n_dst + 7) / 8;
  const int grid = (int)(blocks < 32LL * num_sms() ? blocks : 32LL * num_sms());
  combine_subsegments_kernel<<<grid, 256, 0, as_stream(stream)>>>(sub_h, sub_v, sub_cnt, sub_ptr, tot_cnt, n_dst, inv_norm,
                                                                  inv_norm_node, agg_h, agg_v, accumulate);
  PF_CHECK_LAUNCH("pf_combine_subsegments");
  return PF_OK;
}

extern "C" int pf_plan_tiles(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                             int32_t tile_rows, int32_t* tiles, int32_t max_tiles, int32_t* n_tiles,
                             uint32_t* dev_status, void* stream) {
  PF_CHECK_ARG(seg_cnt && chunk_ptr && tiles && n_tiles && dev_status, "pf_plan_tiles: null pointer");
  PF_CHECK_ARG(tile_rows == PF_TILE_ROWS || tile_rows == PF_TC_TILE_ROWS, "pf_plan_tiles: tile_rows must be 64 or 128");
  if (n_chunks <= 0) return PF_OK;
  plan_tiles_kernel<<<(n_chunks + 127) / 128, 128, 0, as_stream(stream)>>>(seg_cnt, chunk_ptr, n_chunks, skip_empty,
                                                                          tile_rows, tiles, max_tiles, n_tiles,
                                                                          dev_status);
  PF_CHECK_LAUNCH("pf_plan_tiles");
  return PF_OK;
}

extern "C" int pf_plan_tiles3(const int32_t* const seg_cnt[3], const int32_t* const chunk_ptr[3], const int32_t n_chunks[3],
                              const int32_t skip_empty[3], int32_t tile_rows, int32_t* const tiles[3], int32_t max_tiles,
                              int32_t* n_tiles3, uint32_t* dev_status, void* stream) {
  PF_CHECK_ARG(seg_cnt && chunk_ptr && n_chunks && skip_empty && tiles && n_tiles3 && dev_status, "pf_plan_tiles3: null pointer");
  PF_CHECK_ARG(tile_rows == PF_TILE_ROWS || tile_rows == PF_TC_TILE_ROWS, "pf_plan_tiles3: tile_rows must be 64 or 128");
  Plan3 p;
  int most = 0;
  for (int i = 0; i < 3; ++i) {
    PF_CHECK_ARG(seg_cnt[i] && chunk_ptr[i] && tiles[i] && n_chunks[i] >= 0, "pf_plan_tiles3: null plan array");
    p.d[i] = PlanDesc{seg_cnt[i], chunk_ptr[i], n_chunks[i], skip_empty[i], tiles[i], n_tiles3 + i};
    most = n_chunks[i] > most ? n_chunks[i] : most;
  }
  if (most == 0) return PF_OK;
  plan_tiles3_kernel<<<dim3((most + 3) / 4, 3), 128, 0, as_stream(stream)>>>(p, tile_rows, max_tiles, dev_status);   // 4 warps = 4 chunks per CTA
  PF_CHECK_LAUNCH("pf_plan_tiles3");
  return PF_OK;
}

extern "C" int pf_plan_tiles_count(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                                   int32_t tile_rows, int32_t* chunk_tiles, uint32_t* dev_status, void* stream) {
  PF_CHECK_ARG(seg_cnt && chunk_ptr && chunk_tiles && dev_status, "pf_plan_tiles_count: null pointer");
  PF_CHECK_ARG(tile_rows == PF_TILE_ROWS || tile_rows == PF_TC_TILE_ROWS, "pf_plan_tiles_count: tile_rows must be 64 or 128");
  if (n_chunks <= 0) return PF_OK;
  plan_tiles_ordered_kernel<<<(n_chunks + 127) / 128, 128, 0, as_stream(stream)>>>(
      seg_cnt, chunk_ptr, n_chunks, skip_empty, tile_rows, chunk_tiles, nullptr, nullptr, 0, nullptr, dev_status);
  PF_CHECK_LAUNCH("pf_plan_tiles_count");
  return PF_OK;
}

extern "C" int pf_plan_tiles_fill(const int32_t* seg_cnt, const int32_t* chunk_ptr, int32_t n_chunks, int32_t skip_empty,
                                  int32_t tile_rows, const int32_t* chunk_tile_off, int32_t* tiles, int32_t max_tiles,
                                  int32_t* n_tiles, uint32_t* dev_status, void* stream) {
  PF_CHECK_ARG(seg_cnt && chunk_ptr && chunk_tile_off && tiles && n_tiles && dev_status, "pf_plan_tiles_fill: null pointer");
  PF_CHECK_ARG(tile_rows == PF_TILE_ROWS || tile_rows == PF_TC_TILE_ROWS, "pf_plan_tiles_fill: tile_rows must be 64 or 128");
  if (n_chunks <= 0) return PF_OK;
  plan_tiles_ordered_kernel<<<(n_chunks + 127) / 128, 128, 0, as_stream(stream)>>>(
      seg_cnt, chunk_ptr, n_chunks, skip_empty, tile_rows, nullptr, chunk_tile_off, tiles, max_tiles, n_tiles, dev_status);
  PF_CHECK_LAUNCH("pf_plan_tiles_fill");
  return PF_OK;
}

extern "C" int pf_zero_i32(int32_t* p, int64_t n, void* stream) {
  PF_CHECK_ARG(p && n >= 0, "pf_zero_i32: null pointer");
  if (n == 0) return PF_OK;
  zero_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p, n);
  PF_CHECK_LAUNCH("pf_zero_i32");
  return PF_OK;
}
