// K3 on the 5th-generation tensor cores: gather -> edge features -> 3-GVP message chain -> segmented mean, one
// persistent CTA per SM, two 128-edge tiles in flight (ping-pong between the tensor pipe and the CUDA cores).
//
// Replaces, per edge type of one GVPMultiEdgeConv layer, gvp.py:472-497 + 540-551 -> 89-116 of the reference.
//
// Numerics: every dense contraction runs as tcgen05.mma kind::f16 with both operands split into fp16 (hi, lo)
// pairs and three passes hi*hi + hi*lo + lo*hi accumulated in fp32 TMEM (|x - hi - lo| <= 2^-22 |x|: 22 of the
// 24 significand bits), which keeps the fp32 1e-4 parity bar of BASELINE.json with margin (a bf16 split, 2^-16,
// does not); SiLU / sigmoid / norms / means run in fp32 on the CUDA cores.
//
// Per tile and GVP g (A = TMEM region holding the scalar operand, D = the other region; they swap every GVP):
//   V_g : Vh|Vu[128 x 32] (x3 components) = V[128 x 16] . [Wh | Wh.Wu]     A from smem (staging), B resident
//   S_g : D[128 x 128] = [f | rbf | sh][128 x K] . Wf^T                    A from TMEM (f) + smem (rbf, sh),
//                                                                           B streamed through a 12-slab ring
//   G_g : gate[128 x 16] = f'[128 x 128] . Wg^T                            A from TMEM, B resident
// with CUDA-core stages between them: EPI-A (vector norms -> sh), EPI-B (bias + SiLU, re-split in place into the
// next A operand), EPI-C (sigmoid gate * Vu -> next vector operand).  A TMEM lane is one edge row.
//
// Warp roles (608 threads): warps 0-7 epilogue of slot 0, warps 8-15 epilogue of slot 1 (two threads per edge row:
// column halves), warps 16 / 17 issue the MMAs of slot 0 / 1 (converged, under elect.sync), warp 18 lane 0 streams
// weight slabs with cp.async.bulk.
//
// Template switches of every kernel in this file: HAS_V (source vectors present: layers > 0), FAST (single fp16 pass,
// pf_*_tc_f16: hi images only, SiLU on packed fp16 pairs; tolerance 2e-2 instead of 1e-4), TRACE (device timeline,
// launched only while pf_tc_trace is armed).  What the epilogue warps are short of is registers (96 per thread, one
// CTA of 19 warps per SM, L1 reduced to ~30 KB by the 217 KB of shared memory): values that must survive a long stage
// are parked in dead TMEM columns, the row gather runs in two rounds of eight loads, and the next tile's source rows
// are pulled into L2 one tile ahead ("r01c").
#include "pf_common.cuh"
#include "pf_tc.cuh"
#include <type_traits>

namespace pf {
namespace tcc {

constexpr int kRows = PF_TC_TILE_ROWS;  // 128
constexpr int kRing = 12;               // weight slabs resident in shared memory
constexpr int kSlab = 8192;             // one K=16 slab of Wf^T [128 x 16]: hi image 4 KB | lo image 4 KB
// Packed weight blobs (pharmacoforge_b200/weights.py: pack_message_tc / pack_update_tc): Wf^T K-slabs of every GVP,
// then the "small" block kept resident in shared memory: gate images (8 KB per GVP) | vector images (2 KB per GVP)
// | fp32 constants.
template <int MODE>
struct Cfg;
template <>
struct Cfg<0> {  // 3-GVP edge message chain; GVP 0: K = 128 h + 16 rbf + 17 sh (+15 zero) = 176, GVP 1, 2: K = 144
  static constexpr int kGvps = 3, kSlabsPerTile = 29, kVecOff = 24576, kConstOff = 30720, kSmallBytes = 32768;
  static constexpr int kBlobSlab0 = 0, kBlobSlabs = 29;  // first slab of the tile sequence / slabs in the blob
  __host__ __device__ static constexpr int nslab(int g) { return g == 0 ? 11 : 9; }
};
template <>
struct Cfg<1> {  // 2-GVP node update chain; K = 128 f + 16 sh = 144
  static constexpr int kGvps = 2, kSlabsPerTile = 18, kVecOff = 16384, kConstOff = 20480, kSmallBytes = 24576;
  static constexpr int kBlobSlab0 = 0, kBlobSlabs = 18;
  __host__ __device__ static constexpr int nslab(int) { return 9; }
};
template <>
struct Cfg<2> {  // seeded first-layer message chain (same blob as Cfg<0>): the h part of GVP 0 (its first 8 K-steps) is
                 // precomputed per source node and arrives as the accumulator seed, so a tile streams slabs 8 .. 28 only
  static constexpr int kGvps = 3, kSlabsPerTile = 21, kVecOff = 24576, kConstOff = 30720, kSmallBytes = 32768;
  static constexpr int kBlobSlab0 = 8, kBlobSlabs = 29;
  __host__ __device__ static constexpr int nslab(int g) { return g == 0 ? 3 : 9; }
};
template <int MODE>
constexpr int blob_small_off() { return Cfg<MODE>::kBlobSlabs * kSlab; }
template <int MODE>
constexpr int blob_bytes() { return blob_small_off<MODE>() + Cfg<MODE>::kSmallBytes; }
constexpr int kGateOff = 0;
constexpr int kSmallMax = 32768;
constexpr int kConstOff = Cfg<0>::kConstOff;  // edge kernel
// node-update constants (floats from Cfg<1>::kConstOff): per GVP g at 144 g: bf[128] | bg[16]; LayerNorm rows at 512
constexpr int kCLnMsgW = 512, kCLnMsgB = 640, kCLnUpdW = 768, kCLnUpdB = 896;
// fp32 constants (floats): per GVP g at 144 g: bf[128] | bg[16]; then GVP 0 extras
constexpr int kCWh0 = 432;    // Wh0[0][h], h = 0..16  (x_diff row)
constexpr int kCWhu0 = 452;   // (Wh0.Wu0)[0][u], u = 0..15
constexpr int kCWhc16 = 468;  // Wh0[1+u][16], u = 0..15 (17th hidden channel)

constexpr int kStage = 36864;  // per-slot staging: smem A operands (3 x 8 KB) / 8 transpose buffers / mean buffers
constexpr int kMetaInts = 2064; // tile metadata (520 ints) + per-row exchange between the column halves (1024) + segment records (512)
constexpr int kOffRing = 0;
constexpr int kOffSmall = kRing * kSlab;              // 98,304
constexpr int kOffStage = kOffSmall + kSmallMax;      // 131,072
constexpr int kOffMeta = kOffStage + 2 * kStage;      // 186,368
constexpr int kOffBars = kOffMeta + 2 * kMetaInts * 4;  // 190,720
constexpr int kNumBars = 3 * kRing + 12 + 2;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
constexpr int kThreadsTc = 608;  // 16 epilogue warps (2 slots x 2 column halves x 4 lane quarters) + 2 MMA + producer

struct SlotBars {
  uint64_t vecA, vecD, A, D, F, gate;
};

struct Params {
  const float *src_h, *src_v, *src_x, *dst_x;
  const int *seg_start, *seg_cnt, *seg_dst, *col, *tiles, *n_tiles;
  const uint8_t* wblob;
  float *agg_h, *agg_v;
  int accumulate;
  long long* trace;  // optional timeline of CTA 0 (pf_tc_trace): [4 roles][kTraceCap][2] = (tag, clock64)
  // SEED kernels only (first conv layer, one-hot source features): row seed_row[src] of `seed` ([rows][128] fp32) is
  // k Wf0[:, 0:128] h_src, the per-node part of GVP 0's scalar contraction (pf_seed_table); src_h is not read
  const int* seed_row;
  const float* seed;
  // optional row map of src_h (general kernels): source node n reads scalar row src_map[n] instead of row n -- the first
  // conv layer's protein scalars are one encoder row per (graph, atom type), read from that table instead of a
  // materialised [n_prot][128] array (src_v / src_x are indexed by the node as always)
  const int* src_map;
};

// sigma(y) = 1 / (1 + 2^(-y log2 e)) on the two MUFU ops with no range fix-up code: ex2.approx overflows to +inf for
// y << 0 and rcp.approx(+inf) = +0, which is the correct limit (relative error ~2^-22, inside the fp32 parity bar)
__device__ __forceinline__ float sigmoid_fast(float y) { return tc::rcp_approx(1.0f + tc::ex2_approx(y * -1.4426950408889634f)); }
// sqrt(max(x, 1e-8)) as x' * rsqrt.approx(x'): 3 instructions instead of the ~8 of sqrtf's IEEE path (2 ulp)
__device__ __forceinline__ float sqrt_clamped(float x) {
  const float c = fmaxf(x, 1e-8f);
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(c));
  return c * r;
}
__device__ __forceinline__ float silu_fast(float y) { return y * sigmoid_fast(y); }

// 16 fp32 values of row m -> fp16 (hi, lo) in the K-major SWIZZLE_NONE image of a [128 x 16] A operand:
// byte (k/8)*2048 + (m/8)*128 + (m%8)*16 + (k%8)*2, hi image at +0, lo image at +4096.
__device__ __forceinline__ void stage_store16(uint8_t* slab, int m, const float* x) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tc::split_pack_h(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
  uint8_t* a = slab + (m >> 3) * 128 + (m & 7) * 16;
  *reinterpret_cast<uint4*>(a) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(a + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  *reinterpret_cast<uint4*>(a + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(a + 6144) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

// Experiment switches of the node-update kernel: tile-wise alternation of the two slots' S jobs,
// and a start-up delay of the odd CTAs (ns) so that half of the grid runs half a tile behind the other half.
#ifndef PF_SILU_SHARED_RCP
#define PF_SILU_SHARED_RCP 1   // 1 (default, measured -2.5 .. -3 % per pp launch): one reciprocal per four SiLU elements in
                               // EPI-B (1.25 MUFU + 8 instructions per element); 0: one per element (2 MUFU + 6): the XU pipe,
                               // not the issue slots, is the scarcer resource of that stage
#endif
#ifndef PF_MEAN_SPLIT
#define PF_MEAN_SPLIT 1   // 1 (default, measured -1.1 .. -1.5 % per pp launch, bit-identical): one column quad per thread in
                          // the segmented mean (16 threads per segment), see segment_means1; 0: two quads, 8 threads
#endif
#ifndef PF_K3_TILE_ALT
#define PF_K3_TILE_ALT 0   // the same tile-wise order for the edge kernels (every weight slab then comes twice per tile pair from
                           // L2: 232 KB per 128-edge tile instead of per pair).  Measured at 23.2 M edges: general kernel
                           // without / with source vectors 14.22 -> 16.56 / 16.65 -> 17.11 ms: worse, stays off
#endif
#ifndef PF_K4_TILE_ALT
#define PF_K4_TILE_ALT 1   // node update: S jobs of the two slots alternate TILE-wise and the weight ring is loaded in that order
                           // (one consumer per slab): the slots sit ~a quarter period apart instead of marching in lockstep
                           // through load phase, GVP chain and store phase.  Measured 3.01 / 3.02 -> 2.87 / 2.85 ms at 3.07 M
                           // nodes, identical outputs.  0: job-wise alternation on a ring shared by both slots (K3's scheme)
#endif
#ifndef PF_K4_PREFETCH_AT
#define PF_K4_PREFETCH_AT 6   // where / how a slot pulls its NEXT tile's rows into L2: 0 never, 1 per-thread prefetches at the
                              // start of the GVP chain, 2 at tile start, 3 at the start of the back end, 4 at 1 and 3, 5 / 6 =
                              // 1 / 3 as four bulk-copy prefetches (cp.async.bulk.prefetch.L2, one per array).  Measured at
                              // 3.07 M nodes with the tile-wise ring order (layer-1 launch): 0: 3.05 ms, 1: 2.84, 2: 3.29,
                              // 3: 2.81, 4: 2.99, 5: 2.77, 6: 2.72 (default)
#endif
#ifndef PF_K4_L2HINT
#define PF_K4_L2HINT 15  // (default: all four, measured 3.09 / 3.12 -> 2.99 / 3.00 ms at 3.07 M nodes, identical outputs)
                         // node update, L2 eviction hints (bits): 1 = the normalised rows written by the front end and re-read
                         // by the back end are stored evict_last; 2 = the input rows (h_in, agg_h, v_in, agg_v) and the
                         // residual re-reads are loaded evict_first; 4 = the final rows are stored evict_first; 8 = the
                         // next tile's L2 prefetch asks for evict_last
#endif
#ifndef PF_K4_CTA_STAGGER_NS
#define PF_K4_CTA_STAGGER_NS 0
#endif
constexpr int kTraceCap = 4096;
#ifndef PF_DSLEEP
#define PF_DSLEEP 100
#endif
constexpr unsigned kDSleepNs = PF_DSLEEP;  // sleep between polls of the S-job barrier
// TRACE is a compile-time switch (the traced kernels are separate instantiations, launched only while pf_tc_trace is
// armed): the product kernels carry neither the event counter nor the pointer -- at 96 registers per thread both spilled.
template <bool TRACE>
__device__ __forceinline__ void trace_ev(long long* trace, int role, int& n, int tag) {
  if constexpr (!TRACE) return;
  if (trace != nullptr && blockIdx.x == 0 && n < kTraceCap) {
    trace[((size_t)role * kTraceCap + n) * 2] = tag;
    trace[((size_t)role * kTraceCap + n) * 2 + 1] = clock64();
    ++n;
  }
}

__device__ __forceinline__ void slot_barrier(int T) { tc::named_bar_sync(1 + T, 256); }
// The two warps that share a TMEM lane quarter (column halves hh = 0 / 1 of the same 32 rows) exchange per-row values
// through s_xch: a 64-thread barrier (ids 3..10) instead of the slot's 256-thread one wherever nothing else is shared.
__device__ __forceinline__ void pair_barrier(int T, int q) { tc::named_bar_sync(3 + 4 * T + q, 64); }

// ------------------------------------------------------------------------------------------------ producer
template <int MODE>
__host__ __device__ constexpr bool tile_alt() { return MODE == 1 ? (PF_K4_TILE_ALT != 0) : (PF_K3_TILE_ALT != 0); }
template <int MODE>
__host__ __device__ constexpr int ring_uses_p(int pos) { return (Cfg<MODE>::kSlabsPerTile - pos + kRing - 1) / kRing; }

// Slab i of the tile sequence goes to ring position i % kRing (see ring_uses / full_parity below).  Before the copy,
// the previous occupant of the position must have been consumed by both tile slots: the `empty` barrier of a slot
// flips once per use, so the n-th use overall waits for parity (n - 1) & 1.
template <int MODE>
__device__ void producer_role(const uint8_t* wblob, uint8_t* smem, uint64_t* bar_full, uint64_t* bar_empty,
                              uint64_t* bar_small, int my_tiles) {
  constexpr int kSlabsPerTile = Cfg<MODE>::kSlabsPerTile;
  if (my_tiles == 0) return;
  tc::mbar_expect_tx(bar_small, Cfg<MODE>::kSmallBytes);
#pragma unroll
  for (int i = 0; i < Cfg<MODE>::kSmallBytes / 8192; ++i)
    tc::bulk_g2s(smem + kOffSmall + i * 8192, wblob + blob_small_off<MODE>() + i * 8192, 8192, bar_small);
  if constexpr (tile_alt<MODE>()) {
    // Tile-wise order (node update, PF_K4_TILE_ALT): the slab sequence of tile q (slot q & 1) follows that of tile q - 1,
    // every slab has ONE consumer (the slot of its tile) and one `empty` barrier per ring position; the S jobs are issued
    // in the same order (mma_role), so loads and consumption walk the ring in lockstep and the two slots can sit half a
    // tile apart (one in its memory phases while the other runs its GVP chain).  Costs each slab twice per tile pair.
#pragma unroll 1
    for (int q = 0; q < my_tiles; ++q) {
#pragma unroll 1
      for (int i = 0; i < kSlabsPerTile; ++i) {
        const int n = kSlabsPerTile * q + i, pos = n % kRing, u = n / kRing;
        if (u >= 1) tc::mbar_wait(&bar_empty[pos], (uint32_t)(u - 1) & 1u);
        tc::mbar_expect_tx(&bar_full[pos], kSlab);
        tc::bulk_g2s(smem + kOffRing + pos * kSlab, wblob + (size_t)(i + Cfg<MODE>::kBlobSlab0) * kSlab, kSlab, &bar_full[pos]);
      }
    }
    return;
  }
  const int pairs = (my_tiles + 1) >> 1;
#pragma unroll 1
  for (int t = 0; t < pairs; ++t) {
#pragma unroll
    for (int i = 0; i < kSlabsPerTile; ++i) {
      const int pos = i % kRing, u = i / kRing;
      const int uses = ring_uses_p<MODE>(pos);
      const int n = uses * t + u;  // uses of this position before this one
      if (n >= 1) {
        const uint32_t par = (uint32_t)(n - 1) & 1u;
        const int tprev = u > 0 ? t : t - 1;  // tile pair of the previous use
        tc::mbar_wait(&bar_empty[pos], par);
        if (2 * tprev + 1 < my_tiles) tc::mbar_wait(&bar_empty[kRing + pos], par);
      }
      tc::mbar_expect_tx(&bar_full[pos], kSlab);
      tc::bulk_g2s(smem + kOffRing + pos * kSlab, wblob + (size_t)(i + Cfg<MODE>::kBlobSlab0) * kSlab, kSlab, &bar_full[pos]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ MMA issuer
// One issuing WARP per tile slot (warps 16, 17).  The whole warp runs the control flow converged and the tcgen05
// instructions are issued under elect.sync with descriptors advanced by adds: a single diverged lane that rebuilds its
// descriptors per MMA is issue-bound at ~160 cycles per MMA whatever N is (scratch/mma_bench.cu), the converged form
// reaches 77 (N = 128) / 41 (N = 32) / 36 (N = 16) cycles.
// Ring bookkeeping shared by the producer and the issuers: slab i of a tile's sequence (i < kSlabsPerTile) always
// lives in ring position i % kRing, so every position, TMEM column and barrier address in the unrolled issue code
// is a compile-time constant; only the barrier parity depends on the tile-pair counter t, and only for positions
// that are used an odd number of times per tile.
template <int MODE>
__host__ __device__ constexpr int ring_uses(int pos) {  // uses of ring position `pos` per tile
  return (Cfg<MODE>::kSlabsPerTile - pos + kRing - 1) / kRing;
}
template <int MODE>
__host__ __device__ constexpr int slab_base(int g) {  // first slab of GVP g in the tile's sequence
  return g == 0 ? 0 : slab_base<MODE>(g - 1) + Cfg<MODE>::nslab(g - 1);
}
// parity of the `full` barrier for the use of slab i in tile pair t
template <int MODE>
__device__ __forceinline__ uint32_t full_parity(int i, uint32_t t) {
  return ((ring_uses<MODE>(i % kRing) & 1 ? t : 0u) + (uint32_t)(i / kRing)) & 1u;
}

template <int MODE, bool HAS_V, bool FAST, bool TRACE>
__device__ void mma_role(const int T, uint8_t* smem, uint32_t tmem, uint64_t* bar_full, uint64_t* bar_empty,
                         SlotBars* sb, uint64_t* bar_small, uint64_t* bar_stagger, volatile int* turn, int my_tiles,
                         long long* trace_) {
  const int n_mine = (my_tiles + 1 - T) >> 1;
  if (n_mine == 0) return;
  const bool lane0 = (threadIdx.x & 31) == 0;
  long long* trace = TRACE && lane0 ? trace_ : nullptr;
  // S jobs (the long N = 128 MMA runs) of the two tile slots are issued strictly alternately -- slot 0 job j, slot 1
  // job j, slot 0 job j + 1, ... -- so that they execute back to back on the tensor pipe instead of interleaved:
  // one slot's accumulator completes a full job ahead of the other's, and the slots settle half a GVP apart (one on
  // the CUDA cores while the other owns the tensor pipe) instead of marching in lockstep.  The order also keeps the
  // two consumers of the shared weight ring within one job (<= 11 of 12 slabs) of each other.
  const int n_other_jobs = ((my_tiles + T) >> 1) * Cfg<MODE>::kGvps;  // S jobs of the other slot
  const int n_my_jobs = n_mine * Cfg<MODE>::kGvps;
  // Global issue order of the S jobs: *turn holds the sequence number that may issue next.  Edge kernels: job-wise
  // alternation (slot 0 job j, slot 1 job j, ...).  Node update (kTileAlt): TILE-wise alternation (both jobs of slot 0's
  // tile, then both of slot 1's): a slot's GVP chain then runs under the other slot's memory phases (back end + front end)
  // instead of both slots loading, computing and storing in lockstep.
  constexpr bool kTileAlt = tile_alt<MODE>();
  constexpr int kG = Cfg<MODE>::kGvps;
  auto seq_of = [&](int slot, int j) { return kTileAlt ? (2 * (j / kG) + slot) * kG + (j % kG) : 2 * j + slot; };
  auto seq_valid = [&](int sq) {   // does the job with this sequence number exist?
    int slot, j;
    if (kTileAlt) {
      const int blk = sq / kG;
      slot = blk & 1;
      j = (blk >> 1) * kG + sq % kG;
    } else {
      slot = sq & 1;
      j = sq >> 1;
    }
    return j < (slot == T ? n_my_jobs : n_other_jobs);
  };
  const int seq_end = kTileAlt ? 2 * kG * ((my_tiles + 1) >> 1) + 2 * kG : 2 * (n_my_jobs > n_other_jobs ? n_my_jobs : n_other_jobs) + 2;
  int job = 0;
  int tn = 0;
  constexpr uint32_t kI128 = tc::make_idesc_f16(128, 128);
  constexpr uint32_t kI32 = tc::make_idesc_f16(128, 32);
  constexpr uint32_t kI16 = tc::make_idesc_f16(128, 16);
  const uint32_t ring_a = tc::smem_u32(smem + kOffRing);
  const uint32_t small_a = tc::smem_u32(smem + kOffSmall);
  const uint32_t stage = tc::smem_u32(smem + kOffStage) + T * kStage;
  const uint32_t regP = tmem + 256 * T, regQ = regP + 128;
  // descriptors of the staged A operands and of the ring (16-byte units in the low word: advancing by `bytes >> 4`)
  const uint64_t stage_hi = tc::make_smem_desc(stage, 2048, 128), stage_lo = tc::make_smem_desc(stage + 4096, 2048, 128);
  const uint64_t ring_hi = tc::make_smem_desc(ring_a, 2048, 128), ring_lo = tc::make_smem_desc(ring_a + 4096, 2048, 128);
  const uint64_t vec_hi = tc::make_smem_desc(small_a + Cfg<MODE>::kVecOff, 512, 128);
  const uint64_t vec_lo = tc::make_smem_desc(small_a + Cfg<MODE>::kVecOff + 1024, 512, 128);
  const uint64_t gate_hi = tc::make_smem_desc(small_a + kGateOff, 256, 128);
  const uint64_t gate_lo = tc::make_smem_desc(small_a + kGateOff + 512, 256, 128);
  SlotBars& B = sb[T];
  uint64_t* my_empty = bar_empty + T * kRing;
  uint32_t p_vecA = 0, p_A = 0, p_F = 0;
  tc::mbar_wait(bar_small, 0);

  auto gvp = [&](auto gc, const uint32_t t) {
    constexpr int g = decltype(gc)::value;
    const uint32_t Areg = g == 1 ? regQ : regP;
    const uint32_t Dreg = g == 1 ? regP : regQ;
    if (MODE == 1 || HAS_V || g > 0) {  // ---- V_g: vector channels, A = staged V (hi, lo), B = [Wh | Wh.Wu] image
      tc::mbar_wait(&B.vecA, p_vecA);
      p_vecA ^= 1;
      tc::fence_after_sync();
      trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x10);
      if (tc::elect_one()) {
        const uint64_t b_hi = vec_hi + (uint64_t)(g * (2048 >> 4)), b_lo = vec_lo + (uint64_t)(g * (2048 >> 4));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint64_t a_hi = stage_hi + (uint64_t)(c * (8192 >> 4)), a_lo = stage_lo + (uint64_t)(c * (8192 >> 4));
          tc::mma_ss(Dreg + 32 * c, a_hi, b_hi, kI32, 0);
          if constexpr (!FAST) {
            tc::mma_ss(Dreg + 32 * c, a_hi, b_lo, kI32, 1);
            tc::mma_ss(Dreg + 32 * c, a_lo, b_hi, kI32, 1);
          }
        }
        tc::mma_commit(&B.vecD);
      }
      __syncwarp();
      trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x11);
    }
    // ---- S_g: scalar features, one weight slab (K = 16) at a time
    tc::mbar_wait(&B.A, p_A);
    p_A ^= 1;
    {
      const int my_seq = seq_of(T, job);
      while (*turn != my_seq) __nanosleep(20);
    }
    tc::fence_after_sync();
    trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x20);
    constexpr int nslab = Cfg<MODE>::nslab(g);
    // Two weight slabs per trip through the issue code (wait for both, one elected issue block, two commits): the
    // per-slab overhead -- barrier poll, elect, warp re-convergence -- is paid once per pair.
    // MODE 2, GVP 0: the accumulator was seeded by the epilogue warps (K-steps 0 .. 7 precomputed per node); the job is
    // the three staged K-steps (rbf, sh, 17th hidden channel) and accumulates from its first MMA
    constexpr int kstep0 = (MODE == 2 && g == 0) ? 8 : 0;
    auto issue_slab = [&](const int kslab, const int pos) {
      const int k = kslab + kstep0;
      const uint64_t b_hi = ring_hi + (uint64_t)(pos * (kSlab >> 4)), b_lo = ring_lo + (uint64_t)(pos * (kSlab >> 4));
      if (k < 8) {
        const uint32_t a_hi = Areg + 16 * k, a_lo = a_hi + 8;
        tc::mma_ts(Dreg, a_hi, b_hi, kI128, k > 0);
        if constexpr (!FAST) {
          tc::mma_ts(Dreg, a_hi, b_lo, kI128, 1);
          tc::mma_ts(Dreg, a_lo, b_hi, kI128, 1);
        }
      } else {
        const uint64_t a_hi = stage_hi + (uint64_t)((k - 8) * (8192 >> 4)), a_lo = stage_lo + (uint64_t)((k - 8) * (8192 >> 4));
        tc::mma_ss(Dreg, a_hi, b_hi, kI128, 1);
        if constexpr (!FAST) {
          tc::mma_ss(Dreg, a_hi, b_lo, kI128, 1);
          tc::mma_ss(Dreg, a_lo, b_hi, kI128, 1);
        }
      }
      tc::mma_commit(kTileAlt ? &bar_empty[pos] : &my_empty[pos]);
    };
#pragma unroll
    for (int k = 0; k < nslab; k += 2) {
      constexpr int base = slab_base<MODE>(g);
      const int i0 = base + k;
      const bool two = k + 1 < nslab;
      const int i1 = i0 + 1;
      // tile-wise order: slab i of this slot's t-th tile is slab (2 t + T) kSlabsPerTile + i of the CTA's sequence
      // (tile_n0 = slabs of the CTA's sequence before this tile, tile 2 t + T: position and lap of slab i follow from it)
      const uint32_t tile_n0 = (uint32_t)Cfg<MODE>::kSlabsPerTile * (2u * t + (uint32_t)T);
      const uint32_t n0 = kTileAlt ? tile_n0 % kRing + (uint32_t)i0 : (uint32_t)i0, n1 = n0 + 1;
      const int pos0 = (int)(n0 % kRing), pos1 = (int)(n1 % kRing);
      const uint32_t tile_uses = tile_n0 / kRing;   // ring laps before this tile
      const uint32_t par0 = kTileAlt ? (tile_uses + n0 / kRing) & 1u : full_parity<MODE>(i0, t);
      const uint32_t par1 = kTileAlt ? (tile_uses + n1 / kRing) & 1u : full_parity<MODE>(i1, t);
      tc::mbar_wait(&bar_full[pos0], par0);
      if (two) tc::mbar_wait(&bar_full[pos1], par1);
      if (tc::elect_one()) {
        issue_slab(k, pos0);
        if (two) issue_slab(k + 1, pos1);
        if (k + 2 >= nslab) tc::mma_commit(&B.D);
      }
      __syncwarp();
    }
    {
      // hand the turn to the next job that exists (an unpartnered last tile leaves gaps in the other slot's numbering)
      int nxt = seq_of(T, job) + 1;
      while (nxt < seq_end && !seq_valid(nxt)) ++nxt;
      if (lane0) {
        __threadfence_block();
        *turn = nxt;
      }
      ++job;
      __syncwarp();
    }
    if (T == 0 && t == 0 && g == 0 && lane0) tc::mbar_arrive(bar_stagger);  // slot 1 starts half a phase behind slot 0
    trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x21);
    // ---- G_g: vector gates from the new scalars (split in place in Dreg), output -> Areg[0:16)
    tc::mbar_wait(&B.F, p_F);
    p_F ^= 1;
    tc::fence_after_sync();
    trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x30);
    if (tc::elect_one()) {
      const uint64_t g_hi = gate_hi + (uint64_t)(g * (8192 >> 4)), g_lo = gate_lo + (uint64_t)(g * (8192 >> 4));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t b_hi = g_hi + (uint64_t)(k * (1024 >> 4)), b_lo = g_lo + (uint64_t)(k * (1024 >> 4));
        const uint32_t a_hi = Dreg + 16 * k, a_lo = a_hi + 8;
        tc::mma_ts(Areg, a_hi, b_hi, kI16, k > 0);
        if constexpr (!FAST) {
          tc::mma_ts(Areg, a_hi, b_lo, kI16, 1);
          tc::mma_ts(Areg, a_lo, b_hi, kI16, 1);
        }
      }
      tc::mma_commit(&B.gate);
    }
    __syncwarp();
    trace_ev<TRACE>(trace, 2 + T, tn, (g << 8) | 0x31);
  };

#pragma unroll 1
  for (uint32_t t = 0; t < (uint32_t)n_mine; ++t) {
    gvp(std::integral_constant<int, 0>{}, t);
    gvp(std::integral_constant<int, 1>{}, t);
    if constexpr (Cfg<MODE>::kGvps == 3) gvp(std::integral_constant<int, 2>{}, t);
  }
}

// ------------------------------------------------------------------------------------------------ epilogue warps
// 8 bits of headroom below the fp16 maximum: the row's largest |v| lands in [2^13, 2^14).  Message vectors shrink
// by ~10x per GVP (1e-4 after three); without the per-row power-of-two scale their fp16 (hi, lo) parts fall into
// the subnormal range and lose the 22-bit accuracy the scalars get.
__device__ __forceinline__ void row_scale(float m, float& sc, float& inv) {
  int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
  e = e < 24 ? 24 : (e > 240 ? 240 : e);
  sc = __uint_as_float((uint32_t)(267 - e) << 23);
  inv = __uint_as_float((uint32_t)(e - 13) << 23);
}

// 8 fp32 values of row m, K positions [8 kc, 8 kc + 8) of a [128 x 16] A operand slab (see stage layout above)
template <bool FAST = false>
__device__ __forceinline__ void stage_store8(uint8_t* slab, int m, int kc, const float* x) {
  uint8_t* a = slab + kc * 2048 + (m >> 3) * 128 + (m & 7) * 16;
  if constexpr (FAST) {  // single fp16 pass: the lo image is never read
    *reinterpret_cast<uint4*>(a) = make_uint4(tc::pack_f16x2_sat(x[0], x[1]), tc::pack_f16x2_sat(x[2], x[3]),
                                              tc::pack_f16x2_sat(x[4], x[5]), tc::pack_f16x2_sat(x[6], x[7]));
  } else {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::split_pack_h(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    *reinterpret_cast<uint4*>(a) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Segmented mean over the rows of a tile staged in shared memory as buf[row][pitch] (fp32).  Work item = (segment,
// 8 columns): the 256 threads of a tile slot are 8 column groups x 32 segments in flight; a segment's rows are
// loaded eight at a time (all loads issued before the first add: shared-memory latency is paid once, not per row)
// and summed in row order, so the result does not depend on where the segment sits in the tile -- the means are
// bit-identical under any batch composition.  Segments are whole destinations: exactly one thread writes a
// destination's columns, no atomics race.  With `accumulate` the mean is added to what earlier edge types left in the
// row (one red.add per address and launch, hence still deterministic).  s_rec[j] = (first row, end row, dst, 1/count).
constexpr int kMeanPitch = 68;   // 64 columns + 4: 16-byte row stores and quad loads are bank-conflict free
constexpr int kMeanPitchV = 52;  // 48 vector entries + 4
// predicated 16-byte shared-memory load of row I of a batch into pre-zeroed registers: one compare against an immediate
// and one load at an immediate offset from the batch's base address (the C++ form costs an address computation and a
// branch per row; this phase runs at ~12 cycles per instruction and warp, so the integer work was half of its time)
template <int I, int PITCH>
__device__ __forceinline__ void lds128_row(float4& v, uint32_t base, int rows_left) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.gt.s32 p, %5, %6;\n\t"
      "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%7];\n\t"
      "}\n"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "r"(base), "r"(rows_left), "n"(I), "n"(I * PITCH * 4));
}
template <int PITCH>
__device__ __forceinline__ void lds128_rows8(float4 (&v)[8], uint32_t base, int rows_left) {
  lds128_row<0, PITCH>(v[0], base, rows_left);
  lds128_row<1, PITCH>(v[1], base, rows_left);
  lds128_row<2, PITCH>(v[2], base, rows_left);
  lds128_row<3, PITCH>(v[3], base, rows_left);
  lds128_row<4, PITCH>(v[4], base, rows_left);
  lds128_row<5, PITCH>(v[5], base, rows_left);
  lds128_row<6, PITCH>(v[6], base, rows_left);
  lds128_row<7, PITCH>(v[7], base, rows_left);
}
// A thread owns two column quads, `bx` and `by` (buffer addresses of row 0) -> `ox` and `oy` (output addresses of
// destination 0): the eight threads of a segment read 128 contiguous bytes per quad and row, i.e. one shared-memory
// wavefront per quarter warp (8 consecutive columns per thread cost two).
template <int PITCH>
__device__ __forceinline__ void segment_means(const float* bx, const float* by, const int jfirst, const int nseg,
                                              const int4* s_rec, float* ox, float* oy, const int out_pitch,
                                              const int accumulate) {
  const uint32_t ax = tc::smem_u32(bx), ay = tc::smem_u32(by);
  for (int j = jfirst; j < nseg; j += 32) {
    const int4 rec = s_rec[j];
    const int r0 = rec.x, r1 = rec.y;
    if (accumulate && r1 == r0) continue;
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;  // four packed fp32 pairs: columns (0,1) (2,3) (4,5) (6,7) of the group
    for (int r = r0; r < r1; r += 8) {
      float4 x[8], y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        y[i] = x[i];
      }
      lds128_rows8<PITCH>(x, ax + r * (PITCH * 4), r1 - r);
      lds128_rows8<PITCH>(y, ay + r * (PITCH * 4), r1 - r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // row order, two columns per FADD2
        s0 = tc::add2(s0, tc::pack2(x[i].x, x[i].y));
        s1 = tc::add2(s1, tc::pack2(x[i].z, x[i].w));
        s2 = tc::add2(s2, tc::pack2(y[i].x, y[i].y));
        s3 = tc::add2(s3, tc::pack2(y[i].z, y[i].w));
      }
    }
    const float rc = __int_as_float(rec.w);
    const uint64_t rc2 = tc::pack2(rc, rc);
    float4 a0, a1;
    tc::unpack2(tc::mul2(s0, rc2), a0.x, a0.y);
    tc::unpack2(tc::mul2(s1, rc2), a0.z, a0.w);
    tc::unpack2(tc::mul2(s2, rc2), a1.x, a1.y);
    tc::unpack2(tc::mul2(s3, rc2), a1.z, a1.w);
    const size_t o = (size_t)rec.z * out_pitch;
    if (accumulate) {
      atomicAdd(reinterpret_cast<float4*>(ox + o), a0);
      atomicAdd(reinterpret_cast<float4*>(oy + o), a1);
    } else {
      *reinterpret_cast<float4*>(ox + o) = a0;
      *reinterpret_cast<float4*>(oy + o) = a1;
    }
  }
}

// The same reduction with ONE column quad per thread (PF_MEAN_SPLIT): 16 threads per segment, 16 segments in flight per
// slot.  A tile of the pp graph holds ~17 segments, so with 32 segment slots x 2 quads per thread half of the slot's threads
// idle while the others do twice the work; per (segment, column) the rows are summed in the same order: identical results.
template <int PITCH>
__device__ __forceinline__ void segment_means1(const float* bx, const int jfirst, const int jstep, const int nseg,
                                               const int4* s_rec, float* ox, const int out_pitch, const int accumulate) {
  const uint32_t ax = tc::smem_u32(bx);
  for (int j = jfirst; j < nseg; j += jstep) {
    const int4 rec = s_rec[j];
    const int r0 = rec.x, r1 = rec.y;
    if (accumulate && r1 == r0) continue;
    uint64_t s0 = 0, s1 = 0;
    for (int r = r0; r < r1; r += 8) {
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      lds128_rows8<PITCH>(x, ax + r * (PITCH * 4), r1 - r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s0 = tc::add2(s0, tc::pack2(x[i].x, x[i].y));
        s1 = tc::add2(s1, tc::pack2(x[i].z, x[i].w));
      }
    }
    const float rc = __int_as_float(rec.w);
    const uint64_t rc2 = tc::pack2(rc, rc);
    float4 a0;
    tc::unpack2(tc::mul2(s0, rc2), a0.x, a0.y);
    tc::unpack2(tc::mul2(s1, rc2), a0.z, a0.w);
    const size_t o = (size_t)rec.z * out_pitch;
    if (accumulate) {
      atomicAdd(reinterpret_cast<float4*>(ox + o), a0);
    } else {
      *reinterpret_cast<float4*>(ox + o) = a0;
    }
  }
}

// Two threads per edge row: half hh owns scalar columns [64 hh, 64 hh + 64) and vector channels [8 hh, 8 hh + 8).
template <bool HAS_V, bool FAST, bool TRACE, bool SEED>
__device__ void epilogue_role(const Params& p, const int T, uint8_t* smem, uint32_t tmem, SlotBars* sb,
                              uint64_t* bar_small, uint64_t* bar_stagger, int my_tiles) {
  static_assert(!(SEED && HAS_V), "the seeded chain is the first conv layer: no source vectors");
  const int stid = threadIdx.x & 255;
  const int hh = stid >> 7;          // column half
  const int et = stid & 127;         // edge row of the tile == TMEM lane
  const int q = et >> 5, lane = et & 31;
  const int wslot = stid >> 5;       // warp within the slot, 0..7
  const uint32_t P = tmem + 256 * T + ((uint32_t)(q * 32) << 16), Q = P + 128;
  uint8_t* stage = smem + kOffStage + T * kStage;
  int* s_off = reinterpret_cast<int*>(smem + kOffMeta) + T * kMetaInts;  // [129]
  int* s_start = s_off + 132;                                            // [128]
  int* s_dst = s_start + 128;                                            // [128]
  int* s_rowseg = s_dst + 128;                                           // [128]
  int* s_wsum = s_rowseg + 128;                                          // [4]
  float4* s_xch = reinterpret_cast<float4*>(s_wsum + 4);                 // [128][2]
  int4* s_rec = reinterpret_cast<int4*>(s_xch + 256);                    // [128]
  const float* cst = reinterpret_cast<const float*>(smem + kOffSmall + kConstOff);
  SlotBars& B = sb[T];
  uint32_t par_vecD = 0, par_D = 0, par_gate = 0;
  if (T < my_tiles) tc::mbar_wait(bar_small, 0);
  // ping-pong: slot 1 starts once slot 0's first scalar job is issued, so that one slot's CUDA-core stages run
  // under the other slot's MMAs instead of both slots marching in lockstep
  if (T == 1 && T < my_tiles) tc::mbar_wait(bar_stagger, 0);
  long long* trace = TRACE && stid == 0 ? p.trace : nullptr;
  int tn = 0;

  // Software prefetch of the tile descriptors: (first, end) segment of the slot's next-but-one tile and the segment
  // records of its next tile are loaded one tile ahead, so the per-tile start-up does not wait on two dependent
  // global round trips.
  int pf_s0 = 0, pf_nseg = 0, pf_c = 0, pf_start = 0, pf_dst = 0, pf2_s0 = 0, pf2_s1 = 0;
  int pf_off = 0, pf_next = 0;  // start - start of the tile's first segment; start of the next segment
  // Gather pipeline, one tile ahead (contiguous tiles only): pf_e0 / pf_rows = first edge / row count of the slot's
  // next tile (uniform), nx_src = the source row of edge row `et` in it -- loaded in the middle of the current tile, its
  // feature rows are then pulled into L2 before the next tile's gather, which would otherwise wait on DRAM twice
  // (col -> rows) with nothing to overlap.
  int pf_e0 = 0, pf_rows = 0, nx_src = -1, cur_pre_src = -1;
  int nx_srow = -1, cur_pre_srow = -1;  // SEED: seed-table row of nx_src, loaded one tile ahead as well
  auto load_segs = [&](int s0_, int nseg_) {
    pf_c = 0;
    pf_e0 = __ldg(p.seg_start + s0_);
    pf_rows = __ldg(p.seg_start + s0_ + nseg_ - 1) + __ldg(p.seg_cnt + s0_ + nseg_ - 1) - pf_e0;
    if (hh == 0 && et < nseg_) {
      pf_c = __ldg(p.seg_cnt + s0_ + et);
      pf_start = __ldg(p.seg_start + s0_ + et);
      pf_off = pf_start - __ldg(p.seg_start + s0_);
      pf_next = et + 1 < nseg_ ? __ldg(p.seg_start + s0_ + et + 1) : pf_start + pf_c;
      pf_dst = p.seg_dst ? __ldg(p.seg_dst + s0_ + et) : s0_ + et;
    }
  };
  if (T < my_tiles) {
    const int tile = blockIdx.x + T * gridDim.x;
    pf_s0 = __ldg(p.tiles + 2 * tile);
    pf_nseg = __ldg(p.tiles + 2 * tile + 1) - pf_s0;
    load_segs(pf_s0, pf_nseg);
    if (T + 2 < my_tiles) {
      const int tile2 = blockIdx.x + (T + 2) * gridDim.x;
      pf2_s0 = __ldg(p.tiles + 2 * tile2);
      pf2_s1 = __ldg(p.tiles + 2 * tile2 + 1);
    }
  }

  for (int it = T; it < my_tiles; it += 2) {
    const int s0 = pf_s0, nseg = pf_nseg;
    const int cur_c = pf_c, cur_start = pf_start, cur_dst = pf_dst, cur_off = pf_off, cur_next = pf_next;
    cur_pre_src = nx_src;  // meaningful only if this tile turns out contiguous (it was loaded as col[first edge + et])
    nx_src = -1;
    cur_pre_srow = nx_srow;
    nx_srow = -1;
    const bool have_next = it + 2 < my_tiles;
    if (it + 2 < my_tiles) {  // prefetch for the next tile of this slot
      pf_s0 = pf2_s0;
      pf_nseg = pf2_s1 - pf2_s0;
      load_segs(pf_s0, pf_nseg);
      if (it + 4 < my_tiles) {
        const int tile4 = blockIdx.x + (it + 4) * gridDim.x;
        pf2_s0 = __ldg(p.tiles + 2 * tile4);
        pf2_s1 = __ldg(p.tiles + 2 * tile4 + 1);
      }
    }
    slot_barrier(T);  // everyone is done with the previous tile's metadata and staging
    trace_ev<TRACE>(trace, T, tn, 0x01);
    // ---- tile metadata.  Short path: when the tile's segments are stored back to back (start[j] + cnt[j] ==
    // start[j + 1], e.g. the pp CSR, 78 % of a step), a segment's first row is start[j] - start[0] -- no scan, no
    // row -> segment table (rows find their segment by binary search) and ONE barrier, which also carries the vote.
    bool contig_ok = true;
    if (hh == 0 && et < nseg) {
      s_start[et] = cur_start;
      s_dst[et] = cur_dst;
      s_off[et] = cur_off;
      s_rec[et] = make_int4(cur_off, cur_off + cur_c, cur_dst, __float_as_int(1.0f / (float)(cur_c > 0 ? cur_c : 1)));
      if (et == nseg - 1) s_off[nseg] = cur_off + cur_c;
      contig_ok = cur_start + cur_c == cur_next;
    }
    const bool contig = tc::named_bar_and(1 + T, 256, contig_ok);
    // ---- general path (half 0): exclusive scan of the segment sizes (<= 128 segments)
    if (!contig) {
      const int c = cur_c;
      int inc = c;
      if (hh == 0) {
        if (et < nseg) {
          s_start[et] = cur_start;
          s_dst[et] = cur_dst;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[q] = inc;
      }
      slot_barrier(T);
      if (hh == 0) {
        int base = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) base += w < q ? s_wsum[w] : 0;
        s_off[et] = base + inc - c;
        if (et == 127) s_off[128] = base + inc;
        if (et < nseg)
          s_rec[et] = make_int4(base + inc - c, base + inc, cur_dst, __float_as_int(1.0f / (float)(c > 0 ? c : 1)));
      }
      slot_barrier(T);
      if (hh == 0 && et < nseg)
        for (int r = s_off[et]; r < s_off[et + 1]; ++r) s_rowseg[r] = et;
      slot_barrier(T);
    }
    const int nrows = s_off[nseg];
    int src = -1;
    float xd[3] = {0.f, 0.f, 0.f}, dist = 0.f;
    float gx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // coordinates of the edge's source and destination
    if (et < nrows) {
      int j = 0;
      if (contig) {  // largest j with s_off[j] <= et (empty segments share their successor's offset and lose)
#pragma unroll
        for (int step = 64; step >= 1; step >>= 1) {
          const int c = j + step;
          if (c < nseg && s_off[c] <= et) j = c;
        }
      } else {
        j = s_rowseg[et];
      }
      src = contig && cur_pre_src >= 0 ? cur_pre_src : __ldg(p.col + s_start[j] + (et - s_off[j]));
      const int dst = s_dst[j];
#pragma unroll
      for (int c = 0; c < 3; ++c) {  // loads only: the geometry is computed under the first round of row loads below
        gx[c] = __ldg(p.src_x + (size_t)src * 3 + c);
        gx[3 + c] = __ldg(p.dst_x + (size_t)dst * 3 + c);
      }
    }

    trace_ev<TRACE>(trace, T, tn, 0x02);
    // ---- gather h[src] (coalesced: 8 lanes x 16 B per row chunk), transpose through smem to one thread per row,
    //      split into fp16 (hi, lo) and store as the TMEM A operand of S_0 in region P
    float* tb = reinterpret_cast<float*>(stage + wslot * 4608);  // private to the warp: [32][36] / [32][28]
    float4 vq[HAS_V ? 6 : 1];  // this half's 24 entries of v[src] ([3][16] component-major rows), cooperative layout
    if constexpr (SEED) {
      // ---- seeded layer: no row gather, no fp16 split.  The accumulator of S_0 (region Q) is initialised with this edge's
      // row of the per-node table (k Wf0[:, 0:128] h_src); the table has one row per (graph, atom type), so the rows of
      // a tile are a handful of 512-byte lines that stay in L1 / L2.  Two rounds of 32 columns (8 loads in flight).
      const bool pre = contig && cur_pre_src >= 0 && cur_pre_srow >= 0;
      const int srow = src >= 0 ? (pre ? cur_pre_srow : __ldg(p.seed_row + src)) : -1;
      const float4* prow = reinterpret_cast<const float4*>(p.seed + (size_t)(srow < 0 ? 0 : srow) * kHidden + 64 * hh);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float4 r8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r8[i] = srow >= 0 ? __ldg(prow + 8 * c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c == 0 && src >= 0) {  // edge geometry (gvp.py:474-479) while the first eight loads are in flight
          const float dx = gx[0] - gx[3], dy = gx[1] - gx[4], dz = gx[2] - gx[5];
          const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          dist = sqrtf(fmaxf(d2, 1e-8f)) + 1e-8f;
          xd[0] = dx / dist;
          xd[1] = dy / dist;
          xd[2] = dz / dist;
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            w[4 * i] = __float_as_uint(r8[4 * h2 + i].x);
            w[4 * i + 1] = __float_as_uint(r8[4 * h2 + i].y);
            w[4 * i + 2] = __float_as_uint(r8[4 * h2 + i].z);
            w[4 * i + 3] = __float_as_uint(r8[4 * h2 + i].w);
          }
          tc::tmem_st16(Q + 64 * hh + 32 * c + 16 * h2, w);
        }
      }
    } else
    {
      // Two rounds of 8 row loads (32 registers each) instead of 16 at once: the second round is issued right after the
      // first one's staging stores and flies under its fp16 split.  With all 16 loads in flight the 64 data registers did
      // not fit next to the tile state (96 registers per thread): 23 of the loaded values went through local memory.
      float4 ga[8], gb[8];
      const int srch = (p.src_map != nullptr && src >= 0) ? __ldg(p.src_map + src) : src;   // scalar row of the source
      auto load8 = [&](float4 (&g)[8], const int c2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int sr = __shfl_sync(0xffffffffu, srch, 4 * i + (lane >> 3));
          g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sr >= 0)
            g[i] = __ldg(reinterpret_cast<const float4*>(p.src_h + (size_t)sr * kHidden + 32 * (2 * hh + c2)) + (lane & 7));
        }
      };
      auto stage8 = [&](const float4 (&g)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(tb + (4 * i + (lane >> 3)) * 36 + 4 * (lane & 7)) = g[i];
        __syncwarp();
      };
      auto split32 = [&](const int c) {  // 32 staged columns of the own row -> two K-steps in TMEM: hi (8 cols) | lo (8 cols)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(tb + lane * 36 + 16 * ks + 4 * j);
            if constexpr (FAST) {
              hi[2 * j] = tc::pack_f16x2_sat(v.x, v.y);
              hi[2 * j + 1] = tc::pack_f16x2_sat(v.z, v.w);
            } else {
              tc::split_pack_h(v.x, v.y, hi[2 * j], lo[2 * j]);
              tc::split_pack_h(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
            }
          }
          tc::tmem_st8(P + 32 * c + 16 * ks, hi);
          if constexpr (!FAST) tc::tmem_st8(P + 32 * c + 16 * ks + 8, lo);
        }
        __syncwarp();
      };
      load8(ga, 0);
      if (src >= 0) {  // edge geometry (gvp.py:474-479) while the first eight row loads are in flight
        const float dx = gx[0] - gx[3], dy = gx[1] - gx[4], dz = gx[2] - gx[5];
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        dist = sqrtf(fmaxf(d2, 1e-8f)) + 1e-8f;
        xd[0] = dx / dist;
        xd[1] = dy / dist;
        xd[2] = dz / dist;
      }
      stage8(ga);
      load8(gb, 1);
      split32(2 * hh);
      stage8(gb);
      if constexpr (HAS_V) {  // the vector rows fly under the second split (they used to cost a round trip of their own)
#pragma unroll
        for (int ps = 0; ps < 6; ++ps) {
          const int qq = 32 * ps + lane;
          const int rr = qq / 6, pc = qq - 6 * rr;
          const int sr = __shfl_sync(0xffffffffu, src, rr);
          vq[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sr >= 0)
            vq[ps] = __ldg(reinterpret_cast<const float4*>(p.src_v + (size_t)sr * kVRow) + 4 * (pc >> 1) + 2 * hh + (pc & 1));
        }
      }
      split32(2 * hh + 1);
    }

    float Vu[24];                       // vector channels [8 hh, 8 hh + 8) of the 3 components, index 8 c + u'
    float vsc = 1.f, vinv = 1.f;        // power-of-two scale of the staged vector operand and its inverse
    float vh16[3] = {0.f, 0.f, 0.f};    // 17th hidden channel of GVP 0 (used by half 0)
    if constexpr (HAS_V) {
      // ---- transpose the vector rows (loaded above) the same way
#pragma unroll
      for (int ps = 0; ps < 6; ++ps) {
        const int qq = 32 * ps + lane;
        const int rr = qq / 6, pc = qq - 6 * rr;
        *reinterpret_cast<float4*>(tb + rr * 28 + 4 * pc) = vq[ps];
      }
      __syncwarp();
      float pm = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(tb + lane * 28 + 4 * j);
        Vu[4 * j] = v.x;
        Vu[4 * j + 1] = v.y;
        Vu[4 * j + 2] = v.z;
        Vu[4 * j + 3] = v.w;
        pm = fmaxf(fmaxf(pm, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
      float pv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) a = fmaf(Vu[8 * c + u], cst[kCWhc16 + 8 * hh + u], a);
        pv[c] = a;
      }
      s_xch[et * 2 + hh] = make_float4(pm, pv[0], pv[1], pv[2]);
    }
    trace_ev<TRACE>(trace, T, tn, 0x03);
    // transposes done (the staging writes below overlap other warps' buffers); exchange visible.  The seeded kernel has
    // neither: its first staging write comes after the tile-start barrier, which orders it behind the previous tile's mean.
    if constexpr (!SEED) slot_barrier(T);
    if constexpr (HAS_V) {
      const float4 o = s_xch[et * 2 + (1 - hh)];
      const float4 m = s_xch[et * 2 + hh];
      row_scale(fmaxf(m.x, o.x), vsc, vinv);
      const float w16 = cst[kCWh0 + 16];
      vh16[0] = fmaf(xd[0], w16, m.y + o.y);
      vh16[1] = fmaf(xd[1], w16, m.z + o.z);
      vh16[2] = fmaf(xd[2], w16, m.w + o.w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t8[u] = Vu[8 * c + u] * vsc;
        stage_store8<FAST>(stage + c * 8192, et, hh, t8);
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&B.vecA);
    }

#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      const uint32_t Areg = g == 1 ? Q : P, Dreg = g == 1 ? P : Q;
      if (have_next) {  // gather pipeline for the slot's next tile (see above); pf_* arrived long ago
        if (g == 1) {
          if (et < pf_rows && pf_rows <= kRows) nx_src = __ldg(p.col + pf_e0 + et);
        } else if (g == 2 && nx_src >= 0) {
          if constexpr (SEED) {
            nx_srow = __ldg(p.seed_row + nx_src);
          } else if (p.src_map == nullptr) {   // (a mapped source is a small table that lives in L2 anyway)
            const char* hrow = reinterpret_cast<const char*>(p.src_h + (size_t)nx_src * kHidden) + 256 * hh;
            tc::prefetch_l2(hrow);
            tc::prefetch_l2(hrow + 128);
          }
          if (hh == 0) {
            tc::prefetch_l2(p.src_x + (size_t)nx_src * 3);
          } else if constexpr (HAS_V) {
            const char* vrow = reinterpret_cast<const char*>(p.src_v + (size_t)nx_src * kVRow);
            tc::prefetch_l2(vrow);
            tc::prefetch_l2(vrow + 128);
          }
        }
      }
      // ================= EPI-A: hidden vector channels -> norms sh (scalar operand tail), Vu kept in registers
      trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x10);
      {
        float sh[8];
        float sh16 = 0.f;
        if (g == 0 && !HAS_V) {
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const float w = cst[kCWh0 + 8 * hh + h];
            const float a0 = xd[0] * w, a1 = xd[1] * w, a2 = xd[2] * w;
            sh[h] = sqrt_clamped(a0 * a0 + a1 * a1 + a2 * a2);
          }
          {
            const float w = cst[kCWh0 + 16];
            const float a0 = xd[0] * w, a1 = xd[1] * w, a2 = xd[2] * w;
            sh16 = sqrt_clamped(a0 * a0 + a1 * a1 + a2 * a2);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u) Vu[8 * c + u] = xd[c] * cst[kCWhu0 + 8 * hh + u];
          vsc = vinv = 1.f;
        } else {
          tc::mbar_wait(&B.vecD, par_vecD);
          par_vecD ^= 1;
          tc::fence_after_sync();
          trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x11);
#pragma unroll
          for (int h = 0; h < 8; ++h) sh[h] = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint32_t ra[8], rb[8];
            tc::tmem_ld8(Dreg + 32 * c + 8 * hh, ra);
            tc::tmem_ld8(Dreg + 32 * c + 16 + 8 * hh, rb);
            tc::wait_ld();
            const float xs = xd[c] * vsc;  // GVP 0: the x_diff channel joins in the scaled domain
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              float vh = __uint_as_float(ra[h]);
              if (g == 0) vh = fmaf(xs, cst[kCWh0 + 8 * hh + h], vh);
              sh[h] = fmaf(vh, vh, sh[h]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float vu = __uint_as_float(rb[u]);
              if (g == 0) vu = fmaf(xs, cst[kCWhu0 + 8 * hh + u], vu);
              Vu[8 * c + u] = vu;
            }
          }
          const float inv2 = vinv * vinv;
#pragma unroll
          for (int h = 0; h < 8; ++h) sh[h] = sqrt_clamped(sh[h] * inv2);
          if (g == 0) sh16 = sqrt_clamped(vh16[0] * vh16[0] + vh16[1] * vh16[1] + vh16[2] * vh16[2]);
        }
        if (g == 0) {
          float t8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float z = (dist - (float)(8 * hh + k)) * (1.0f / 0.9375f);  // 1 ulp from the division, no IEEE div sequence
            t8[k] = __expf(-(z * z));
          }
          stage_store8<FAST>(stage, et, hh, t8);
          stage_store8<FAST>(stage + 8192, et, hh, sh);
#pragma unroll
          for (int k = 0; k < 8; ++k) t8[k] = 0.f;
          if (hh == 0) t8[0] = sh16;
          stage_store8<FAST>(stage + 16384, et, hh, t8);
        } else {
          stage_store8<FAST>(stage, et, hh, sh);
        }
        tc::fence_proxy_async();
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&B.A);
        trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x12);
      }

      // ================= EPI-B: f = SiLU(D + b), split in place into the next A operand; last GVP: mean of f
      {
        tc::mbar_wait_sleep(&B.D, par_D, kDSleepNs);
        par_D ^= 1;
        tc::fence_after_sync();
        trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x21);
        // Park the 24 vector channels in TMEM while this stage runs: the S job is done, so its operand region (Areg) is
        // dead except for columns [0, 16), where the gate will land; half hh uses [32 + 32 hh, 56 + 32 hh).  At 96
        // registers per thread the compiler otherwise spills them to local memory around the SiLU loop, and with
        // 217 KB of the SM's 256 KB configured as shared memory those reloads come from L2.
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t pk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) pk[u] = __float_as_uint(Vu[8 * c + u]);
          tc::tmem_st8(Areg + 32 + 32 * hh + 8 * c, pk);
        }
        const float* bf = cst + 144 * g;
        // half hh owns the 16-column chunks j = 4 hh .. 4 hh + 3 (one K-step of the next scalar operand each);
        // the TMEM load of chunk j + 1 is in flight while chunk j is processed
        uint32_t r[2][16];
        float keep[FAST ? 1 : 32];      // last GVP: fp32 copy of chunks 2, 3 for the second mean pass
        uint32_t keeph[FAST ? 16 : 1];  // single-pass mode: the same as packed fp16 pairs
        tc::tmem_ld16(Dreg + 16 * (4 * hh), r[0]);
        tc::wait_ld();
        float* ab = reinterpret_cast<float*>(stage);  // mean staging [128][kMeanPitch]: 32 columns of each half per pass
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int j = 4 * hh + j4;
          if (j4 < 3) tc::tmem_ld16(Dreg + 16 * (j + 1), r[(j4 + 1) & 1]);
          uint32_t hi[8], lo[8];
          float fv[16];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            if constexpr (FAST) {
              hi[i] = tc::silu_h2(r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i));
              hi[i + 1] = tc::silu_h2(r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3],
                                      *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2));
            } else {
#if PF_SILU_SHARED_RCP
              const uint32_t a4[4] = {r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3]};
              float f4[4];
              tc::silu_pre_split4(a4, *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i),
                                  *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2), f4, hi[i], hi[i + 1], lo[i], lo[i + 1]);
              fv[2 * i] = f4[0], fv[2 * i + 1] = f4[1], fv[2 * i + 2] = f4[2], fv[2 * i + 3] = f4[3];
#else
              tc::silu_pre_split2(r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i),
                                  fv[2 * i], fv[2 * i + 1], hi[i], lo[i]);
              tc::silu_pre_split2(r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3],
                                  *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2), fv[2 * i + 2], fv[2 * i + 3],
                                  hi[i + 1], lo[i + 1]);
#endif
            }
          }
          if (g == 2) {
            if (j4 < 2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if constexpr (FAST) {
                  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi[2 * i]));
                  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi[2 * i + 1]));
                  *reinterpret_cast<float4*>(ab + et * kMeanPitch + 32 * hh + 16 * j4 + 4 * i) = make_float4(a.x, a.y, b.x, b.y);
                } else {
                  *reinterpret_cast<float4*>(ab + et * kMeanPitch + 32 * hh + 16 * j4 + 4 * i) =
                      make_float4(fv[4 * i], fv[4 * i + 1], fv[4 * i + 2], fv[4 * i + 3]);
                }
              }
            } else {
              if constexpr (FAST) {
#pragma unroll
                for (int i = 0; i < 8; ++i) keeph[8 * (j4 - 2) + i] = hi[i];
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) keep[16 * (j4 - 2) + i] = fv[i];
              }
            }
          }
          tc::tmem_st8(Dreg + 16 * j, hi);
          if constexpr (!FAST) tc::tmem_st8(Dreg + 16 * j + 8, lo);
          if (j4 < 3) tc::wait_ld();
        }
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&B.F);
        trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x22);
        if (g == 2) {  // segmented mean of the scalar messages, two passes of 64 columns through shared memory
#pragma unroll 1
          for (int ps = 0; ps < 2; ++ps) {
            if (ps == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if constexpr (FAST) {
                  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&keeph[2 * i]));
                  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&keeph[2 * i + 1]));
                  *reinterpret_cast<float4*>(ab + et * kMeanPitch + 32 * hh + 4 * i) = make_float4(a.x, a.y, b.x, b.y);
                } else {
                  *reinterpret_cast<float4*>(ab + et * kMeanPitch + 32 * hh + 4 * i) =
                      make_float4(keep[4 * i], keep[4 * i + 1], keep[4 * i + 2], keep[4 * i + 3]);
                }
              }
            }
            slot_barrier(T);
            trace_ev<TRACE>(trace, T, tn, (g << 8) | (0x23 + 3 * ps));
            {
              // staging columns [0, 32) = columns 32 ps .. of half 0, [32, 64) = the same of half 1
#if PF_MEAN_SPLIT
              const int c8 = stid & 7, hf = (stid >> 3) & 1;
              segment_means1<kMeanPitch>(ab + 32 * hf + 4 * c8, stid >> 4, 16, nseg, s_rec,
                                         p.agg_h + 32 * ps + 4 * c8 + 64 * hf, kHidden, p.accumulate);
#else
              const int c8 = stid & 7;
              float* out = p.agg_h + 32 * ps + 4 * c8;
              segment_means<kMeanPitch>(ab + 4 * c8, ab + 32 + 4 * c8, stid >> 3, nseg, s_rec, out, out + 64, kHidden,
                                        p.accumulate);
#endif
            }
            trace_ev<TRACE>(trace, T, tn, (g << 8) | (0x24 + 3 * ps));
            slot_barrier(T);
            trace_ev<TRACE>(trace, T, tn, (g << 8) | (0x25 + 3 * ps));
          }
        }
      }

      // ================= EPI-C: V_out = sigmoid(gate) * Vu -> next vector operand, or the vector message mean
      {
        {  // vector channels back from their TMEM parking columns (the loads fly while the gate MMAs finish)
          uint32_t pk[3][8];
#pragma unroll
          for (int c = 0; c < 3; ++c) tc::tmem_ld8(Areg + 32 + 32 * hh + 8 * c, pk[c]);
          tc::wait_ld();
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u) Vu[8 * c + u] = __uint_as_float(pk[c][u]);
        }
        tc::mbar_wait(&B.gate, par_gate);
        par_gate ^= 1;
        tc::fence_after_sync();
        trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x31);
        uint32_t r[8];
        tc::tmem_ld8(Areg + 8 * hh, r);
        tc::wait_ld();
        tc::fence_before_sync();
        const float* bg = cst + 144 * g + 128 + 8 * hh;
        float pm = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float gt = sigmoid_fast(__uint_as_float(r[u]) + bg[u]) * vinv;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            Vu[8 * c + u] *= gt;
            pm = fmaxf(pm, fabsf(Vu[8 * c + u]));
          }
        }
        if (g < 2) {
          s_xch[et * 2 + hh].x = pm;
          pair_barrier(T, q);
          row_scale(fmaxf(pm, s_xch[et * 2 + (1 - hh)].x), vsc, vinv);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float t8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t8[u] = Vu[8 * c + u] * vsc;
            stage_store8<FAST>(stage + c * 8192, et, hh, t8);
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(&B.vecA);
          trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x32);
        } else {
          float* ab = reinterpret_cast<float*>(stage);  // [128][kMeanPitchV]
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int u4 = 0; u4 < 2; ++u4)
              *reinterpret_cast<float4*>(ab + et * kMeanPitchV + 16 * c + 8 * hh + 4 * u4) =
                  make_float4(Vu[8 * c + 4 * u4], Vu[8 * c + 4 * u4 + 1], Vu[8 * c + 4 * u4 + 2], Vu[8 * c + 4 * u4 + 3]);
          slot_barrier(T);
          {
#if PF_MEAN_SPLIT
            const int c16 = stid & 15;
            if (c16 < kVRow / 4)
              segment_means1<kMeanPitchV>(ab + 4 * c16, stid >> 4, 16, nseg, s_rec, p.agg_v + 4 * c16, kVRow, p.accumulate);
#else
            const int c8 = stid & 7;
            if (c8 < kVRow / 8)
              segment_means<kMeanPitchV>(ab + 4 * c8, ab + 24 + 4 * c8, stid >> 3, nseg, s_rec, p.agg_v + 4 * c8,
                                         p.agg_v + 24 + 4 * c8, kVRow, p.accumulate);
#endif
          }
        }
      }
    }
  }
}

template <bool HAS_V, bool FAST, bool TRACE, bool SEED = false>
__global__ void __launch_bounds__(kThreadsTc, 1) edge_conv_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_full = bars;             // [kRing]
  uint64_t* bar_empty = bars + kRing;    // [2][kRing]
  SlotBars* sb = reinterpret_cast<SlotBars*>(bars + 3 * kRing);  // [2]
  uint64_t* bar_small = bars + 3 * kRing + 12;
  uint64_t* bar_stagger = bars + 3 * kRing + 13;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);
  volatile int* s_turn = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = *p.n_tiles;
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  long long t_begin = 0;
  if (TRACE && p.trace != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));

  if (warp == 16) {
    tc::tmem_alloc(s_tmem, 512);
    if (lane == 0) {
      for (int i = 0; i < kRing; ++i) {
        tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_empty[i], 1);
        tc::mbar_init(&bar_empty[kRing + i], 1);
      }
      for (int T = 0; T < 2; ++T) {
        tc::mbar_init(&sb[T].vecA, 256);
        tc::mbar_init(&sb[T].vecD, 1);
        tc::mbar_init(&sb[T].A, 256);
        tc::mbar_init(&sb[T].D, 1);
        tc::mbar_init(&sb[T].F, 256);
        tc::mbar_init(&sb[T].gate, 1);
      }
      tc::mbar_init(bar_small, 1);
      tc::mbar_init(bar_stagger, 1);
      *s_turn = 0;
      tc::fence_mbar_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp < 16) {
    epilogue_role<HAS_V, FAST, TRACE, SEED>(p, warp >> 3, smem, tmem, sb, bar_small, bar_stagger, my_tiles);
  } else if (warp < 18) {
    mma_role<SEED ? 2 : 0, HAS_V, FAST, TRACE>(warp - 16, smem, tmem, bar_full, bar_empty, sb, bar_small, bar_stagger, s_turn, my_tiles, p.trace);
  } else {
    if (lane == 0) producer_role<SEED ? 2 : 0>(p.wblob, smem, bar_full, bar_empty, bar_small, my_tiles);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 16) tc::tmem_dealloc(tmem, 512);
  if (TRACE && p.trace != nullptr && threadIdx.x == 0) {  // debug: (begin, end) of every CTA in ns after the timeline block
    long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    p.trace[(size_t)4 * kTraceCap * 2 + 2 * blockIdx.x] = t_begin;
    p.trace[(size_t)4 * kTraceCap * 2 + 2 * blockIdx.x + 1] = t_end;
  }
}


// =====================================================================================================================
// K4 on the tensor cores: (h, v) <- GVPLayerNorm_msg(h + agg_h, v + agg_v); (rh, rv) = GVP x 2; (h, v) <-
// GVPLayerNorm_upd(h + rh, v + rv)   (gvp.py:511-532, eval mode).  Same tile engine as the message kernel: a tile is
// 128 consecutive nodes, the chain is two plain GVPs (K = 144).  The normalised input rows are written to h_out /
// v_out by the front end and read back as the residual by the back end (same rows, same CTA), so in-place updates
// (h_out == h_in) are allowed.
// =====================================================================================================================
struct NodeParams {
  const float *h_in, *v_in, *agg_h, *agg_v;
  long long n_nodes;
  const uint8_t* wblob;
  float *h_out, *v_out;
  long long* trace;
  const int* h_map;   // optional: node n reads its input scalars from row h_map[n] of h_in (see Params::src_map)
};

__device__ __forceinline__ float xor8_sum(float v) {  // sum over the 8 lanes that share one row of a 4-row pass
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <bool HAS_V, bool FAST, bool TRACE, bool MAPPED>
__device__ void node_epilogue_role(const NodeParams& p, const int T, uint8_t* smem, uint32_t tmem, SlotBars* sb,
                                   uint64_t* bar_small, uint64_t* bar_stagger, int my_tiles) {
  const int stid = threadIdx.x & 255;
  const int hh = stid >> 7;          // column half
  const int et = stid & 127;         // node row of the tile == TMEM lane
  const int q = et >> 5, lane = et & 31;
  const int wslot = stid >> 5;
  const uint32_t P = tmem + 256 * T + ((uint32_t)(q * 32) << 16), Q = P + 128;
  uint8_t* stage = smem + kOffStage + T * kStage;
  float4* s_xch = reinterpret_cast<float4*>(reinterpret_cast<int*>(smem + kOffMeta) + T * kMetaInts + 520);  // [128][2]
  const float* cst = reinterpret_cast<const float*>(smem + kOffSmall + Cfg<1>::kConstOff);
  SlotBars& B = sb[T];
  uint32_t par_vecD = 0, par_D = 0, par_gate = 0;
  if (T < my_tiles) tc::mbar_wait(bar_small, 0);
  if (T == 1 && T < my_tiles) tc::mbar_wait(bar_stagger, 0);
  if (PF_K4_CTA_STAGGER_NS > 0 && (blockIdx.x & 1) && T < my_tiles) {
    // de-phase the grid: all CTAs have identical work, so without this every SM is in its load phase, its GVP chain and
    // its store phase at the same time and DRAM alternates between a burst and idling
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(1000);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < PF_K4_CTA_STAGGER_NS);
  }
  float* tb = reinterpret_cast<float*>(stage + wslot * 4608);  // private to the warp: [32][36]
  const int prow = lane >> 3, piece = lane & 7;                 // cooperative layout: 8 lanes x 16 B per row chunk
  long long* trace = TRACE && stid == 0 ? p.trace : nullptr;
  int tn = 0;

  for (int it = T; it < my_tiles; it += 2) {
    const long long n0 = (long long)(blockIdx.x + (long long)it * gridDim.x) * kRows;
    const long long rem = p.n_nodes - n0;
    const int nrows = rem < kRows ? (int)rem : kRows;
    slot_barrier(T);  // everyone is done with the previous tile's staging / exchange buffers
    trace_ev<TRACE>(trace, T, tn, 0x01);
    auto prefetch_next = [&]() {
      if (it + 2 < my_tiles) {
        const long long m0 = n0 + 2LL * gridDim.x * kRows;
        const long long mrem = p.n_nodes - m0;
        const int mrows = mrem < kRows ? (int)mrem : kRows;
        const int lines_h = mrows * 4, lines_v = (mrows * kVRow * 4 + 127) / 128;  // 128-byte lines
#if PF_K4_L2HINT & 8
#define PF_K4_PREFETCH tc::prefetch_l2_evict_last
#else
#define PF_K4_PREFETCH tc::prefetch_l2
#endif
        for (int l = stid; l < lines_h; l += 256) {
          if constexpr (!MAPPED) PF_K4_PREFETCH(reinterpret_cast<const char*>(p.h_in + m0 * kHidden) + (size_t)l * 128);
          PF_K4_PREFETCH(reinterpret_cast<const char*>(p.agg_h + m0 * kHidden) + (size_t)l * 128);
        }
        for (int l = stid; l < lines_v; l += 256) {
          PF_K4_PREFETCH(reinterpret_cast<const char*>(p.agg_v + m0 * kVRow) + (size_t)l * 128);
          if constexpr (HAS_V) PF_K4_PREFETCH(reinterpret_cast<const char*>(p.v_in + m0 * kVRow) + (size_t)l * 128);
        }
#undef PF_K4_PREFETCH
      }
    };
    if (PF_K4_PREFETCH_AT == 2) prefetch_next();
    auto prefetch_next_bulk = [&]() {   // the same ranges by the bulk-copy engine: one instruction per array and tile
      if (it + 2 < my_tiles && stid < 4) {
        const long long m0 = n0 + 2LL * gridDim.x * kRows;
        const long long mrem = p.n_nodes - m0;
        const int mrows = mrem < kRows ? (int)mrem : kRows;
        if (stid == 0 && !MAPPED) tc::prefetch_l2_bulk(p.h_in + m0 * kHidden, (uint32_t)mrows * kHidden * 4);
        if (stid == 1) tc::prefetch_l2_bulk(p.agg_h + m0 * kHidden, (uint32_t)mrows * kHidden * 4);
        if (stid == 2) tc::prefetch_l2_bulk(p.agg_v + m0 * kVRow, (uint32_t)mrows * kVRow * 4);
        if (stid == 3 && HAS_V) tc::prefetch_l2_bulk(p.v_in + m0 * kVRow, (uint32_t)mrows * kVRow * 4);
      }
    };

    // ---- scalars: x = h_in + agg_h (own 64 columns, cooperative layout), LayerNorm_msg over the full row
    {
      // Row statistics in ONE exchange between the column halves: sum and sum of squares (var = E[x^2] - mean^2; the
      // rows are O(1) with |mean| << std, so the cancellation costs ~1e-7 relative, far inside the parity bar) -- the
      // two-pass form needed a second slot barrier.
      float4 g4[2][8];
      float rs[8], rq[8];
      int hrow[8];   // table rows of this thread's eight node rows (h_map mode only)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rs[i] = rq[i] = 0.f;
        const int row = 32 * q + 4 * i + prow;
        hrow[i] = (MAPPED && row < nrows) ? __ldg(p.h_map + n0 + row) : 0;
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 32 * q + 4 * i + prow;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < nrows) {
            const size_t o = (size_t)(n0 + row) * kHidden + 32 * (2 * hh + c2) + 4 * piece;
            float4 a;
            if constexpr (MAPPED) {   // input scalars from a (small, L2-resident) row table
              x = __ldg(reinterpret_cast<const float4*>(p.h_in + (size_t)hrow[i] * kHidden + 32 * (2 * hh + c2) + 4 * piece));
#if PF_K4_L2HINT & 2
              a = tc::ldg_hint(reinterpret_cast<const float4*>(p.agg_h + o), tc::l2_policy_evict_first());
#else
              a = __ldg(reinterpret_cast<const float4*>(p.agg_h + o));
#endif
            } else {
#if PF_K4_L2HINT & 2
            x = tc::ld_global_hint(reinterpret_cast<const float4*>(p.h_in + o), tc::l2_policy_evict_first());
            a = tc::ldg_hint(reinterpret_cast<const float4*>(p.agg_h + o), tc::l2_policy_evict_first());
#else
            x = *reinterpret_cast<const float4*>(p.h_in + o);  // plain load: h_out may alias h_in
            a = __ldg(reinterpret_cast<const float4*>(p.agg_h + o));
#endif
            }
            x.x += a.x;
            x.y += a.y;
            x.z += a.z;
            x.w += a.w;
          }
          g4[c2][i] = x;
          rs[i] += (x.x + x.y) + (x.z + x.w);
          rq[i] += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        }
      trace_ev<TRACE>(trace, T, tn, 0x4a);   // all 32 row loads of this thread have arrived
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rs[i] = xor8_sum(rs[i]);
        rq[i] = xor8_sum(rq[i]);
        if (piece == 0) {
          float4& e = s_xch[(32 * q + 4 * i + prow) * 2 + hh];
          e.x = rs[i];
          e.y = rq[i];
        }
      }
      pair_barrier(T, q);
      trace_ev<TRACE>(trace, T, tn, 0x4b);   // row statistics exchanged
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = 2 * hh + c2;
        const float4 lw = *reinterpret_cast<const float4*>(cst + kCLnMsgW + 32 * c + 4 * piece);
        const float4 lb = *reinterpret_cast<const float4*>(cst + kCLnMsgB + 32 * c + 4 * piece);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 32 * q + 4 * i + prow;
          const float mean = (s_xch[row * 2].x + s_xch[row * 2 + 1].x) * (1.0f / kHidden);
          const float var = fmaxf((s_xch[row * 2].y + s_xch[row * 2 + 1].y) * (1.0f / kHidden) - mean * mean, 0.f);
          const float rstd = rsqrtf(var + 1e-5f);
          float4 y = g4[c2][i];
          y.x = (y.x - mean) * rstd * lw.x + lb.x;
          y.y = (y.y - mean) * rstd * lw.y + lb.y;
          y.z = (y.z - mean) * rstd * lw.z + lb.z;
          y.w = (y.w - mean) * rstd * lw.w + lb.w;
#if PF_K4_L2HINT & 1
          if (row < nrows)
            tc::st_global_hint(reinterpret_cast<float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece), y,
                               tc::l2_policy_evict_last());
#else
          if (row < nrows) *reinterpret_cast<float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece) = y;
#endif
          *reinterpret_cast<float4*>(tb + (4 * i + prow) * 36 + 4 * piece) = y;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(tb + lane * 36 + 16 * ks + 4 * j);
            if constexpr (FAST) {
              hi[2 * j] = tc::pack_f16x2_sat(v.x, v.y);
              hi[2 * j + 1] = tc::pack_f16x2_sat(v.z, v.w);
            } else {
              tc::split_pack_h(v.x, v.y, hi[2 * j], lo[2 * j]);
              tc::split_pack_h(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
            }
          }
          tc::tmem_st8(P + 32 * c + 16 * ks, hi);
          if constexpr (!FAST) tc::tmem_st8(P + 32 * c + 16 * ks + 8, lo);
        }
        __syncwarp();
      }
    }
    trace_ev<TRACE>(trace, T, tn, 0x41);
    slot_barrier(T);  // scalar transposes done: the vector buffers below overlap them
    trace_ev<TRACE>(trace, T, tn, 0x42);

    // ---- vectors: v = v_in + agg_v, vector LayerNorm (all 16 channels), own 8 channels kept
    float Vu[24];
    float vsc = 1.f, vinv = 1.f;
    {
      float* tv = reinterpret_cast<float*>(stage + q * 6656);  // [32][52], filled by the two warps of the quarter
      {
        // Loads first, stores after, in small batches of passes: a plain (possibly aliased: v_out may be v_in) global
        // load cannot be hoisted above a store through the generic staging pointer, and one pass at a time this loop was
        // 12 dependent L2 round trips -- 30 k of the tile's 95 k cycles on the device timeline.
        // (half hh takes passes 6 hh .. 6 hh + 5, three at a time)
#pragma unroll
        for (int pb = 0; pb < 6; pb += 3) {
          float4 xs[3], as[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int qq = 32 * (6 * hh + pb + k) + lane;
            const int rr = qq / 12, pc = qq - 12 * rr;
            const int row = 32 * q + rr;
            xs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            as[k] = xs[k];
            if (row < nrows) {
              const size_t o = (size_t)(n0 + row) * kVRow + 4 * pc;
#if PF_K4_L2HINT & 2
              xs[k] = tc::ldg_hint(reinterpret_cast<const float4*>(p.agg_v + o), tc::l2_policy_evict_first());
              if constexpr (HAS_V) as[k] = tc::ld_global_hint(reinterpret_cast<const float4*>(p.v_in + o), tc::l2_policy_evict_first());
#else
              xs[k] = __ldg(reinterpret_cast<const float4*>(p.agg_v + o));
              if constexpr (HAS_V) as[k] = *reinterpret_cast<const float4*>(p.v_in + o);
#endif
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int qq = 32 * (6 * hh + pb + k) + lane;
            const int rr = qq / 12, pc = qq - 12 * rr;
            *reinterpret_cast<float4*>(tv + rr * 52 + 4 * pc) =
                make_float4(xs[k].x + as[k].x, xs[k].y + as[k].y, xs[k].z + as[k].z, xs[k].w + as[k].w);
          }
        }
      }
      trace_ev<TRACE>(trace, T, tn, 0x46);
      pair_barrier(T, q);
      trace_ev<TRACE>(trace, T, tn, 0x47);
      float nrm = 0.f;
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const float4 a = *reinterpret_cast<const float4*>(tv + lane * 52 + 4 * u4);
        const float4 b = *reinterpret_cast<const float4*>(tv + lane * 52 + 16 + 4 * u4);
        const float4 c = *reinterpret_cast<const float4*>(tv + lane * 52 + 32 + 4 * u4);
        nrm += fmaxf(a.x * a.x + b.x * b.x + c.x * c.x, 1e-8f) + fmaxf(a.y * a.y + b.y * b.y + c.y * c.y, 1e-8f) +
               fmaxf(a.z * a.z + b.z * b.z + c.z * c.z, 1e-8f) + fmaxf(a.w * a.w + b.w * b.w + c.w * c.w, 1e-8f);
      }
      const float ivn = 1.0f / (sqrtf(nrm * (1.0f / kVec) + 1e-5f) + 1e-5f);  // one division, then multiplies (1 ulp)
      float pm = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int u4 = 0; u4 < 2; ++u4) {
          float4 x = *reinterpret_cast<const float4*>(tv + lane * 52 + 16 * c + 8 * hh + 4 * u4);
          x.x *= ivn;
          x.y *= ivn;
          x.z *= ivn;
          x.w *= ivn;
          Vu[8 * c + 4 * u4] = x.x;
          Vu[8 * c + 4 * u4 + 1] = x.y;
          Vu[8 * c + 4 * u4 + 2] = x.z;
          Vu[8 * c + 4 * u4 + 3] = x.w;
#if PF_K4_L2HINT & 1
          if (et < nrows)
            tc::st_global_hint(reinterpret_cast<float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4), x,
                               tc::l2_policy_evict_last());
#else
          if (et < nrows) *reinterpret_cast<float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4) = x;
#endif
          pm = fmaxf(fmaxf(pm, fabsf(x.x)), fmaxf(fabsf(x.y), fmaxf(fabsf(x.z), fabsf(x.w))));
        }
      s_xch[et * 2 + hh].z = pm;
      trace_ev<TRACE>(trace, T, tn, 0x48);
      slot_barrier(T);  // all rows read (the staging writes below overlap tv); exchange visible
      trace_ev<TRACE>(trace, T, tn, 0x49);
      row_scale(fmaxf(pm, s_xch[et * 2 + (1 - hh)].z), vsc, vinv);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t8[u] = Vu[8 * c + u] * vsc;
        stage_store8<FAST>(stage + c * 8192, et, hh, t8);
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&B.vecA);
    }

    // The device timeline shows every CTA of the grid in the same phase: a front end that saturates DRAM (all 296 tile
    // slots load 180 KB and store 88 KB at once: ~60 k cycles) followed by ~35 k cycles of GVP chain and back end
    // during which DRAM idles.  Pull the rows of the slot's NEXT tile into L2 now, under the GVP chain (per-thread
    // prefetches: the bulk-copy engine would queue them in front of the weight slabs).  Issued at tile start instead,
    // the same prefetch made the kernel 5 % slower (it joins the demand burst and doubles the L2 footprint).
    if (PF_K4_PREFETCH_AT == 1 || PF_K4_PREFETCH_AT == 4) prefetch_next();
    if (PF_K4_PREFETCH_AT == 5) prefetch_next_bulk();
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      const uint32_t Areg = g == 1 ? Q : P, Dreg = g == 1 ? P : Q;
      trace_ev<TRACE>(trace, T, tn, (g << 8) | 0x10);
      // ================= EPI-A: hidden vector channels -> norms sh, Vu kept in registers (scaled by vsc)
      {
        float sh[8];
        tc::mbar_wait(&B.vecD, par_vecD);
        par_vecD ^= 1;
        tc::fence_after_sync();
#pragma unroll
        for (int h = 0; h < 8; ++h) sh[h] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t ra[8], rb[8];
          tc::tmem_ld8(Dreg + 32 * c + 8 * hh, ra);
          tc::tmem_ld8(Dreg + 32 * c + 16 + 8 * hh, rb);
          tc::wait_ld();
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const float vh = __uint_as_float(ra[h]);
            sh[h] = fmaf(vh, vh, sh[h]);
            Vu[8 * c + h] = __uint_as_float(rb[h]);
          }
        }
        const float inv2 = vinv * vinv;
#pragma unroll
        for (int h = 0; h < 8; ++h) sh[h] = sqrt_clamped(sh[h] * inv2);
        stage_store8<FAST>(stage, et, hh, sh);
        tc::fence_proxy_async();
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&B.A);
      }
      // ================= EPI-B: f = SiLU(D + b), split in place into the next A operand
      {
        tc::mbar_wait_sleep(&B.D, par_D, kDSleepNs);
        par_D ^= 1;
        tc::fence_after_sync();
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // park the vector channels in the dead operand region (see the edge kernel)
          uint32_t pk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) pk[u] = __float_as_uint(Vu[8 * c + u]);
          tc::tmem_st8(Areg + 32 + 32 * hh + 8 * c, pk);
        }
        const float* bf = cst + 144 * g;
        uint32_t r[2][16];
        tc::tmem_ld16(Dreg + 16 * (4 * hh), r[0]);
        tc::wait_ld();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int j = 4 * hh + j4;
          if (j4 < 3) tc::tmem_ld16(Dreg + 16 * (j + 1), r[(j4 + 1) & 1]);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            if constexpr (FAST) {
              hi[i] = tc::silu_h2(r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i));
              hi[i + 1] = tc::silu_h2(r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3],
                                      *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2));
            } else {
#if PF_SILU_SHARED_RCP
              const uint32_t a4[4] = {r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3]};
              float f4[4];
              tc::silu_pre_split4(a4, *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i),
                                  *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2), f4, hi[i], hi[i + 1], lo[i], lo[i + 1]);
#else
              float f0, f1;
              tc::silu_pre_split2(r[j4 & 1][2 * i], r[j4 & 1][2 * i + 1], *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i),
                                  f0, f1, hi[i], lo[i]);
              tc::silu_pre_split2(r[j4 & 1][2 * i + 2], r[j4 & 1][2 * i + 3],
                                  *reinterpret_cast<const uint64_t*>(bf + 16 * j + 2 * i + 2), f0, f1, hi[i + 1], lo[i + 1]);
#endif
            }
          }
          tc::tmem_st8(Dreg + 16 * j, hi);
          if constexpr (!FAST) tc::tmem_st8(Dreg + 16 * j + 8, lo);
          if (j4 < 3) tc::wait_ld();
        }
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&B.F);
      }
      // ================= EPI-C: V_out = sigmoid(gate) * Vu
      {
        {
          uint32_t pk[3][8];
#pragma unroll
          for (int c = 0; c < 3; ++c) tc::tmem_ld8(Areg + 32 + 32 * hh + 8 * c, pk[c]);
          tc::wait_ld();
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u) Vu[8 * c + u] = __uint_as_float(pk[c][u]);
        }
        tc::mbar_wait(&B.gate, par_gate);
        par_gate ^= 1;
        tc::fence_after_sync();
        uint32_t r[8];
        tc::tmem_ld8(Areg + 8 * hh, r);
        tc::wait_ld();
        tc::fence_before_sync();
        const float* bg = cst + 144 * g + 128 + 8 * hh;
        float pm = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float gt = sigmoid_fast(__uint_as_float(r[u]) + bg[u]) * vinv;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            Vu[8 * c + u] *= gt;
            pm = fmaxf(pm, fabsf(Vu[8 * c + u]));
          }
        }
        if (g == 0) {
          s_xch[et * 2 + hh].z = pm;
          pair_barrier(T, q);
          row_scale(fmaxf(pm, s_xch[et * 2 + (1 - hh)].z), vsc, vinv);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float t8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t8[u] = Vu[8 * c + u] * vsc;
            stage_store8<FAST>(stage + c * 8192, et, hh, t8);
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(&B.vecA);
        }
      }
    }

    if (PF_K4_PREFETCH_AT == 3 || PF_K4_PREFETCH_AT == 4) prefetch_next();
    if (PF_K4_PREFETCH_AT == 6) prefetch_next_bulk();
    // ================= back end: residual + GVPLayerNorm_upd.  Last GVP was g = 1: f (hi, lo) sits in region P,
    // region Q is free (its gate columns were read above).
    trace_ev<TRACE>(trace, T, tn, 0x43);
    {
      // ---- vectors: t = v_res + V_out, one norm over all 16 channels (partial over the own 8, exchanged)
      float nrm = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int u4 = 0; u4 < 2; ++u4) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
#if PF_K4_L2HINT & 2
          if (et < nrows)
            x = tc::ld_global_hint(reinterpret_cast<const float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4),
                                   tc::l2_policy_evict_first());
#else
          if (et < nrows) x = *reinterpret_cast<const float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4);
#endif
          Vu[8 * c + 4 * u4] += x.x;
          Vu[8 * c + 4 * u4 + 1] += x.y;
          Vu[8 * c + 4 * u4 + 2] += x.z;
          Vu[8 * c + 4 * u4 + 3] += x.w;
        }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        nrm += fmaxf(Vu[u] * Vu[u] + Vu[8 + u] * Vu[8 + u] + Vu[16 + u] * Vu[16 + u], 1e-8f);
      s_xch[et * 2 + hh].z = nrm;
      pair_barrier(T, q);  // also orders the gate reads of both halves before region Q is overwritten below
      {
        const float tot = nrm + s_xch[et * 2 + (1 - hh)].z;
        const float ivn = 1.0f / (sqrtf(tot * (1.0f / kVec) + 1e-5f) + 1e-5f);
        if (et < nrows) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int u4 = 0; u4 < 2; ++u4) {
              float4 x;
              x.x = Vu[8 * c + 4 * u4] * ivn;
              x.y = Vu[8 * c + 4 * u4 + 1] * ivn;
              x.z = Vu[8 * c + 4 * u4 + 2] * ivn;
              x.w = Vu[8 * c + 4 * u4 + 3] * ivn;
#if PF_K4_L2HINT & 4
              tc::st_global_hint(reinterpret_cast<float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4), x,
                                 tc::l2_policy_evict_first());
#else
              *reinterpret_cast<float4*>(p.v_out + (size_t)(n0 + et) * kVRow + 16 * c + 8 * hh + 4 * u4) = x;
#endif
            }
        }
      }
      trace_ev<TRACE>(trace, T, tn, 0x44);
      // ---- scalars, pass 1: y = f + h_res (f rebuilt from its fp16 hi + lo parts), kept in region Q as fp32
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = 2 * hh + c2;
        float4 xr[8];  // all eight loads before the first staging store (see the front end: aliasing serialises them)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 32 * q + 4 * i + prow;
          xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#if PF_K4_L2HINT & 2
          if (row < nrows)
            xr[i] = tc::ld_global_hint(reinterpret_cast<const float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece),
                                       tc::l2_policy_evict_first());
#else
          if (row < nrows) xr[i] = *reinterpret_cast<const float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece);
#endif
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(tb + (4 * i + prow) * 36 + 4 * piece) = xr[i];
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int j = 2 * c + ks;
          uint32_t hi[8], lo[8], y[16];
          tc::tmem_ld8(P + 16 * j, hi);
          if constexpr (!FAST) tc::tmem_ld8(P + 16 * j + 8, lo);
          tc::wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
            const float2 b = FAST ? make_float2(0.f, 0.f) : __half22float2(*reinterpret_cast<const __half2*>(&lo[i]));
            const float y0 = (a.x + b.x) + tb[lane * 36 + 16 * ks + 2 * i];
            const float y1 = (a.y + b.y) + tb[lane * 36 + 16 * ks + 2 * i + 1];
            sum += y0 + y1;
            sq = fmaf(y0, y0, fmaf(y1, y1, sq));
            y[2 * i] = __float_as_uint(y0);
            y[2 * i + 1] = __float_as_uint(y1);
          }
          tc::tmem_st16(Q + 16 * j, y);
        }
        __syncwarp();
      }
      tc::wait_st();
      s_xch[et * 2 + hh].x = sum;
      s_xch[et * 2 + hh].y = sq;
      pair_barrier(T, q);  // one exchange of (sum, sum of squares), as in the front end
      const float mean = (sum + s_xch[et * 2 + (1 - hh)].x) * (1.0f / kHidden);
      const float var = fmaxf((sq + s_xch[et * 2 + (1 - hh)].y) * (1.0f / kHidden) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      trace_ev<TRACE>(trace, T, tn, 0x45);
      // ---- pass 3: normalise, transpose back to the cooperative layout, coalesced store
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = 2 * hh + c2;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t y[16];
          tc::tmem_ld16(Q + 16 * (2 * c + ks), y);
          tc::wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int col = 32 * c + 16 * ks + i;
            const float4 w = *reinterpret_cast<const float4*>(cst + kCLnUpdW + col);
            const float4 b = *reinterpret_cast<const float4*>(cst + kCLnUpdB + col);
            *reinterpret_cast<float4*>(tb + lane * 36 + 16 * ks + i) =
                make_float4((__uint_as_float(y[i]) - mean) * rstd * w.x + b.x, (__uint_as_float(y[i + 1]) - mean) * rstd * w.y + b.y,
                            (__uint_as_float(y[i + 2]) - mean) * rstd * w.z + b.z, (__uint_as_float(y[i + 3]) - mean) * rstd * w.w + b.w);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 32 * q + 4 * i + prow;
          const float4 x = *reinterpret_cast<const float4*>(tb + (4 * i + prow) * 36 + 4 * piece);
#if PF_K4_L2HINT & 4
          if (row < nrows)
            tc::st_global_hint(reinterpret_cast<float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece), x,
                               tc::l2_policy_evict_first());
#else
          if (row < nrows) *reinterpret_cast<float4*>(p.h_out + (size_t)(n0 + row) * kHidden + 32 * c + 4 * piece) = x;
#endif
        }
        __syncwarp();
      }
      tc::fence_before_sync();  // TMEM reads of this tile are complete before the next tile's stores
    }
  }
}

template <bool HAS_V, bool FAST, bool TRACE, bool MAPPED = false>
__global__ void __launch_bounds__(kThreadsTc, 1) node_update_tc_kernel(const NodeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kRing;
  SlotBars* sb = reinterpret_cast<SlotBars*>(bars + 3 * kRing);
  uint64_t* bar_small = bars + 3 * kRing + 12;
  uint64_t* bar_stagger = bars + 3 * kRing + 13;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);
  volatile int* s_turn = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (p.n_nodes + kRows - 1) / kRows;
  const int my_tiles = n_tiles > (long long)blockIdx.x ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp == 16) {
    tc::tmem_alloc(s_tmem, 512);
    if (lane == 0) {
      for (int i = 0; i < kRing; ++i) {
        tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_empty[i], 1);
        tc::mbar_init(&bar_empty[kRing + i], 1);
      }
      for (int T = 0; T < 2; ++T) {
        tc::mbar_init(&sb[T].vecA, 256);
        tc::mbar_init(&sb[T].vecD, 1);
        tc::mbar_init(&sb[T].A, 256);
        tc::mbar_init(&sb[T].D, 1);
        tc::mbar_init(&sb[T].F, 256);
        tc::mbar_init(&sb[T].gate, 1);
      }
      tc::mbar_init(bar_small, 1);
      tc::mbar_init(bar_stagger, 1);
      *s_turn = 0;
      tc::fence_mbar_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp < 16) {
    node_epilogue_role<HAS_V, FAST, TRACE, MAPPED>(p, warp >> 3, smem, tmem, sb, bar_small, bar_stagger, my_tiles);
  } else if (warp < 18) {
    mma_role<1, true, FAST, TRACE>(warp - 16, smem, tmem, bar_full, bar_empty, sb, bar_small, bar_stagger, s_turn, my_tiles, p.trace);
  } else {
    if (lane == 0) producer_role<1>(p.wblob, smem, bar_full, bar_empty, bar_small, my_tiles);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 16) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tcc
}  // namespace pf

using namespace pf;

extern "C" size_t pf_tc_msg_blob_bytes(void) { return (size_t)tcc::blob_bytes<0>(); }
extern "C" size_t pf_tc_upd_blob_bytes(void) { return (size_t)tcc::blob_bytes<1>(); }

static long long* g_tc_trace = nullptr;
extern "C" int pf_tc_trace(long long* device_buf) {  // (4 * 4096 + 148) * 2 int64; nullptr disarms
  g_tc_trace = device_buf;
  return PF_OK;
}

static int launch_edge_conv_tc(bool fast, const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                               const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                               const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                               const void* wblob, float* agg_h, float* agg_v, int32_t accumulate, void* stream,
                               const int32_t* seed_row = nullptr, const float* seed = nullptr,
                               const int32_t* src_map = nullptr) {
  const bool seeded = seed_row != nullptr;
  PF_CHECK_ARG((src_h || seeded) && src_x && dst_x && seg_start && seg_cnt && col && tiles && n_tiles && wblob && agg_h && agg_v,
               "pf_edge_conv_tc: null pointer");
  PF_CHECK_ARG(!seeded || (seed != nullptr && src_v == nullptr), "pf_edge_conv_tc_seeded: needs the seed table and no source vectors");
  PF_CHECK_ARG((reinterpret_cast<uintptr_t>(wblob) & 15) == 0, "pf_edge_conv_tc: weight blob must be 16-byte aligned");
  if (max_tiles <= 0) return PF_OK;
  using KernelFn = void (*)(tcc::Params);
  static const KernelFn fns[8] = {  // index = HAS_V + 2 FAST + 4 TRACE
      tcc::edge_conv_tc_kernel<false, false, false>, tcc::edge_conv_tc_kernel<true, false, false>,
      tcc::edge_conv_tc_kernel<false, true, false>,  tcc::edge_conv_tc_kernel<true, true, false>,
      tcc::edge_conv_tc_kernel<false, false, true>,  tcc::edge_conv_tc_kernel<true, false, true>,
      tcc::edge_conv_tc_kernel<false, true, true>,   tcc::edge_conv_tc_kernel<true, true, true>};
  static const KernelFn seeded_fns[4] = {  // index = FAST + 2 TRACE (no source vectors)
      tcc::edge_conv_tc_kernel<false, false, false, true>, tcc::edge_conv_tc_kernel<false, true, false, true>,
      tcc::edge_conv_tc_kernel<false, false, true, true>,  tcc::edge_conv_tc_kernel<false, true, true, true>};
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    for (int i = 0; i < 12; ++i) {
      const KernelFn f = i < 8 ? fns[i] : seeded_fns[i - 8];
      const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, tcc::kSmemBytes);
      if (e != cudaSuccess) {
        set_error("pf_edge_conv_tc: cudaFuncSetAttribute(smem=%d): %s", tcc::kSmemBytes, cudaGetErrorString(e));
        return PF_ERR_LAUNCH;
      }
    }
    configured.done[dev_] = true;
  }
  tcc::Params p{src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles,
                static_cast<const uint8_t*>(wblob), agg_h, agg_v, accumulate, g_tc_trace, seed_row, seed, src_map};
  const int grid = max_tiles < num_sms() ? max_tiles : num_sms();
  const int which = (src_v != nullptr ? 1 : 0) + (fast ? 2 : 0) + (g_tc_trace != nullptr ? 4 : 0);
  const KernelFn fn = seeded ? seeded_fns[(fast ? 1 : 0) + (g_tc_trace != nullptr ? 2 : 0)] : fns[which];
  fn<<<grid, tcc::kThreadsTc, tcc::kSmemBytes, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_edge_conv_tc");
  return PF_OK;
}

extern "C" int pf_edge_conv_tc(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                               const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                               const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                               const void* wblob, float* agg_h, float* agg_v, int32_t accumulate, void* stream) {
  return launch_edge_conv_tc(false, src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles,
                             max_tiles, wblob, agg_h, agg_v, accumulate, stream);
}

// General kernel with a row map on the source scalars: node n reads row src_map[n] of src_h (a table with one row per
// distinct encoder output) instead of row n; src_v / src_x are indexed by the node.  f16 != 0: single-pass mode.
extern "C" int pf_edge_conv_tc_mapped(const float* src_h, const int32_t* src_map, const float* src_v, const float* src_x,
                                      const float* dst_x, const int32_t* seg_start, const int32_t* seg_cnt,
                                      const int32_t* seg_dst, const int32_t* col, const int32_t* tiles, const int32_t* n_tiles,
                                      int32_t max_tiles, const void* wblob, float* agg_h, float* agg_v, int32_t accumulate,
                                      int32_t f16, void* stream) {
  PF_CHECK_ARG(src_map != nullptr, "pf_edge_conv_tc_mapped: null row map");
  return launch_edge_conv_tc(f16 != 0, src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles,
                             max_tiles, wblob, agg_h, agg_v, accumulate, stream, nullptr, nullptr, src_map);
}

// Single-pass fp16 variant (the "bf16 edge-MLP path" of BASELINE.json configs[3]): same arguments and weight image,
// products hi x hi only (11-bit operands, fp32 accumulation), SiLU on packed fp16 pairs.
extern "C" int pf_edge_conv_tc_f16(const float* src_h, const float* src_v, const float* src_x, const float* dst_x,
                                   const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                                   const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                                   const void* wblob, float* agg_h, float* agg_v, int32_t accumulate, void* stream) {
  return launch_edge_conv_tc(true, src_h, src_v, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles, n_tiles,
                             max_tiles, wblob, agg_h, agg_v, accumulate, stream);
}

// ---- first conv layer, one-hot source features (SURVEY.md hard part 2: W [h_src; rbf; sh] = W_h h_src per NODE + the per-edge
// rest).  pf_seed_table computes table[r] = k Wf0[:, 0:128] h[rep[r]] for every distinct source row r (one per (graph,
// atom type): the encoder output of a one-hot row depends on nothing else), fp32 FFMA; pf_edge_conv_tc_seeded initialises
// GVP 0's accumulator with row seed_row[src] and runs only the per-edge K-steps (rbf, sh) of that contraction.
namespace pf {
namespace tcc {
constexpr int kSeedRowsPerBlock = 8;
__global__ void __launch_bounds__(128) seed_table_kernel(const float* __restrict__ h, const int* __restrict__ rep, int n_rows,
                                                         const float* __restrict__ wfT, float scale, float* __restrict__ table) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                      // [128 (j)][128 (n)]: Wf0^T rows 0 .. 127 (the h part), as packed for the FFMA kernels
  float* hs = sm + kHidden * kHidden;  // [kSeedRowsPerBlock][128]
  const int n = threadIdx.x;
  for (int i = n; i < kHidden * kHidden / 4; i += 128)
    reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wfT) + i);
  for (int r0 = blockIdx.x * kSeedRowsPerBlock; r0 < n_rows; r0 += gridDim.x * kSeedRowsPerBlock) {
    __syncthreads();
    int node[kSeedRowsPerBlock];
#pragma unroll
    for (int r = 0; r < kSeedRowsPerBlock; ++r) {
      node[r] = r0 + r < n_rows ? __ldg(rep + r0 + r) : -1;
      hs[r * kHidden + n] = node[r] >= 0 ? __ldg(h + (size_t)node[r] * kHidden + n) : 0.f;
    }
    __syncthreads();
    float acc[kSeedRowsPerBlock];
#pragma unroll
    for (int r = 0; r < kSeedRowsPerBlock; ++r) acc[r] = 0.f;
    for (int j = 0; j < kHidden; j += 4) {
      const float w0 = ws[(j + 0) * kHidden + n], w1 = ws[(j + 1) * kHidden + n];
      const float w2 = ws[(j + 2) * kHidden + n], w3 = ws[(j + 3) * kHidden + n];
#pragma unroll
      for (int r = 0; r < kSeedRowsPerBlock; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(hs + r * kHidden + j);
        acc[r] = fmaf(x.w, w3, fmaf(x.z, w2, fmaf(x.y, w1, fmaf(x.x, w0, acc[r]))));
      }
    }
#pragma unroll
    for (int r = 0; r < kSeedRowsPerBlock; ++r)
      if (node[r] >= 0) table[(size_t)(r0 + r) * kHidden + n] = acc[r] * scale;
  }
}
}  // namespace tcc
}  // namespace pf

extern "C" int pf_seed_table(const float* h, const int32_t* rep_node, int32_t n_rows, const float* w_msg, float* table,
                             void* stream) {
  PF_CHECK_ARG(h && rep_node && w_msg && table, "pf_seed_table: null pointer");
  if (n_rows <= 0) return PF_OK;
  const size_t smem = (size_t)(kHidden * kHidden + tcc::kSeedRowsPerBlock * kHidden) * sizeof(float);
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    const cudaError_t e = cudaFuncSetAttribute(tcc::seed_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pf_seed_table: cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
      return PF_ERR_LAUNCH;
    }
    configured.done[dev_] = true;
  }
  const GvpLayout L = gvp_layout(17, 16, 144, 128);  // GVP 0 of a message chain (pharmacoforge_b200.h: packed GVP sections)
  const int blocks = (n_rows + tcc::kSeedRowsPerBlock - 1) / tcc::kSeedRowsPerBlock;
  const int grid = blocks < 2 * num_sms() ? blocks : 2 * num_sms();
  tcc::seed_table_kernel<<<grid, 128, smem, as_stream(stream)>>>(h, rep_node, n_rows, w_msg + L.wf, -1.4426950408889634f, table);
  PF_CHECK_LAUNCH("pf_seed_table");
  return PF_OK;
}

extern "C" int pf_edge_conv_tc_seeded(const int32_t* seed_row, const float* seed_table, const float* src_x, const float* dst_x,
                                      const int32_t* seg_start, const int32_t* seg_cnt, const int32_t* seg_dst,
                                      const int32_t* col, const int32_t* tiles, const int32_t* n_tiles, int32_t max_tiles,
                                      const void* wblob, float* agg_h, float* agg_v, int32_t accumulate,
                                      int32_t fp16_single_pass, void* stream) {
  PF_CHECK_ARG(seed_row && seed_table, "pf_edge_conv_tc_seeded: null seed");
  return launch_edge_conv_tc(fp16_single_pass != 0, nullptr, nullptr, src_x, dst_x, seg_start, seg_cnt, seg_dst, col, tiles,
                             n_tiles, max_tiles, wblob, agg_h, agg_v, accumulate, stream, seed_row, seed_table);
}

static int launch_node_update_tc(bool fast, const float* h_in, const float* v_in, const float* agg_h, const float* agg_v,
                                 int64_t n_nodes, const void* wblob, float* h_out, float* v_out, void* stream,
                                 const int32_t* h_map = nullptr) {
  PF_CHECK_ARG(h_in && agg_h && agg_v && wblob && h_out && v_out, "pf_node_update_tc: null pointer");
  PF_CHECK_ARG(h_map == nullptr || h_in != h_out, "pf_node_update_tc_mapped: the row table cannot be the output");
  PF_CHECK_ARG((reinterpret_cast<uintptr_t>(wblob) & 15) == 0, "pf_node_update_tc: weight blob must be 16-byte aligned");
  if (n_nodes <= 0) return PF_OK;
  using KernelFn = void (*)(tcc::NodeParams);
  static const KernelFn fns[8] = {  // index = HAS_V + 2 FAST + 4 TRACE
      tcc::node_update_tc_kernel<false, false, false>, tcc::node_update_tc_kernel<true, false, false>,
      tcc::node_update_tc_kernel<false, true, false>,  tcc::node_update_tc_kernel<true, true, false>,
      tcc::node_update_tc_kernel<false, false, true>,  tcc::node_update_tc_kernel<true, false, true>,
      tcc::node_update_tc_kernel<false, true, true>,   tcc::node_update_tc_kernel<true, true, true>};
  static const KernelFn mapped_fns[4] = {  // index = FAST + 2 TRACE (first layer: no input vectors)
      tcc::node_update_tc_kernel<false, false, false, true>, tcc::node_update_tc_kernel<false, true, false, true>,
      tcc::node_update_tc_kernel<false, false, true, true>,  tcc::node_update_tc_kernel<false, true, true, true>};
  PF_CHECK_ARG(h_map == nullptr || v_in == nullptr, "pf_node_update_tc_mapped: built for the first layer (no input vectors)");
  static PerDeviceFlag configured = {};
  const int dev_ = current_device();
  if (!configured.done[dev_]) {
    for (int i = 0; i < 12; ++i) {
      const KernelFn f = i < 8 ? fns[i] : mapped_fns[i - 8];
      const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, tcc::kSmemBytes);
      if (e != cudaSuccess) {
        set_error("pf_node_update_tc: cudaFuncSetAttribute(smem=%d): %s", tcc::kSmemBytes, cudaGetErrorString(e));
        return PF_ERR_LAUNCH;
      }
    }
    configured.done[dev_] = true;
  }
  tcc::NodeParams p{h_in, v_in, agg_h, agg_v, (long long)n_nodes, static_cast<const uint8_t*>(wblob), h_out, v_out,
                    g_tc_trace, h_map};
  const long long tiles = (n_nodes + tcc::kRows - 1) / tcc::kRows;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  const int which = (v_in != nullptr ? 1 : 0) + (fast ? 2 : 0) + (g_tc_trace != nullptr ? 4 : 0);
  const KernelFn fn = h_map != nullptr ? mapped_fns[(fast ? 1 : 0) + (g_tc_trace != nullptr ? 2 : 0)] : fns[which];
  fn<<<grid, tcc::kThreadsTc, tcc::kSmemBytes, as_stream(stream)>>>(p);
  PF_CHECK_LAUNCH("pf_node_update_tc");
  return PF_OK;
}

extern "C" int pf_node_update_tc(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v,
                                 int64_t n_nodes, const void* wblob, float* h_out, float* v_out, void* stream) {
  return launch_node_update_tc(false, h_in, v_in, agg_h, agg_v, n_nodes, wblob, h_out, v_out, stream);
}

// Node update whose input scalars come from a row table: node n reads row h_map[n] of h_in (see pf_edge_conv_tc_mapped).
extern "C" int pf_node_update_tc_mapped(const float* h_in, const int32_t* h_map, const float* v_in, const float* agg_h,
                                        const float* agg_v, int64_t n_nodes, const void* wblob, float* h_out, float* v_out,
                                        int32_t f16, void* stream) {
  PF_CHECK_ARG(h_map != nullptr, "pf_node_update_tc_mapped: null row map");
  return launch_node_update_tc(f16 != 0, h_in, v_in, agg_h, agg_v, n_nodes, wblob, h_out, v_out, stream, h_map);
}

extern "C" int pf_node_update_tc_f16(const float* h_in, const float* v_in, const float* agg_h, const float* agg_v,
                                     int64_t n_nodes, const void* wblob, float* h_out, float* v_out, void* stream) {
  return launch_node_update_tc(true, h_in, v_in, agg_h, agg_v, n_nodes, wblob, h_out, v_out, stream);
}
