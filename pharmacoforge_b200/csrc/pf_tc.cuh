// Blackwell (sm_100a) tensor-core primitives used by the tcgen05 kernels: mbarrier, 1-D bulk async copy,
// TMEM allocation / load / store, UMMA shared-memory + instruction descriptors, tcgen05.mma issue / commit.
// Everything is inline PTX; no CUTLASS types.
//
// Operand conventions of this code base
//   D  : fp32 accumulator in TMEM, lane = tile row (0..127), one 32-bit column per output column.
//   A  : bf16 in TMEM (".kind::f16" TS form): lane = tile row, column j holds K elements (2j, 2j+1) in its
//        (low, high) halves; one MMA consumes K=16 = 8 columns.
//   B  : bf16 in shared memory, K-major, SWIZZLE_NONE "interleaved" canonical layout: 8x8 core matrices
//        (8 rows of N x 16 bytes of K) stored as 128 contiguous bytes; byte offset of element (n, k) is
//        (k/8)*LBO + (n/8)*128 + (n%8)*16 + (k%8)*2 with LBO = (N/8)*128, SBO = 128.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase flips (or the hint expires)
  // instead of burning issue slots in a poll loop next to the epilogue warps that share its scheduler
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

// ---------------------------------------------------------------- 1-D bulk copy global -> shared (TMA engine)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* gmem) { asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem)); }
// L2 prefetch of a contiguous global range by the bulk-copy engine (one instruction, no destination): 16-byte aligned
// address, size a multiple of 16
__device__ __forceinline__ void prefetch_l2_bulk(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// L2 eviction-priority hints (createpolicy + .L2::cache_hint).  evict_last: lines that are re-read soon (a tile's normalised
// rows between the front and the back end of the node update); evict_first: streaming data that is dead after this access.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_global_hint(float4* ptr, const float4 v, const uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 ld_global_hint(const float4* ptr, const uint64_t pol) {  // plain (coherent) load
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg_hint(const float4* ptr, const uint64_t pol) {  // read-only path
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void prefetch_l2_evict_last(const void* gmem) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(gmem));
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns per thread
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// K-major SWIZZLE_NONE shared-memory operand: start address, leading (K) and stride (M/N) byte offsets in 16-byte
// units, descriptor version 1 (Blackwell), layout type 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16, A = B = bf16 (K-major), D = fp32, M x N
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f16, A = B = fp16 (K-major), D = fp32, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on `bar` once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- fp32 -> (hi, lo) fp16 split, packed pairs
// x ~= hi + lo with |x - hi - lo| <= max(2^-22 |x|, 2^-25): 11 + 11 significand bits; lo is subnormal for
// |x| < 2^-3 (absolute spacing 2^-24).  Conversions saturate: |x| > 65504 (never reached by LayerNorm / SiLU /
// sigmoid outputs) degrades to a finite value instead of inf - inf = NaN.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float x0, float x1) {  // {low half: x0, high half: x1}, saturating
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
__device__ __forceinline__ void split_pack_h(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2_sat(x0, x1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16x2_sat(x0 - hf.x, x1 - hf.y);
}

// ---------------------------------------------------------------- packed fp32 pairs (Blackwell FADD2 / FMUL2)
// Two fp32 lanes per instruction: the epilogues are issue-bound, not FP32-pipe-bound, so halving the instruction count
// of the elementwise chains is a direct win.  A pair lives in one 64-bit register (even-aligned register pair).
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// (f0, f1) = SiLU((a0, a1) + (b0, b1)) and its fp16 (hi, lo) split: 6.5 instructions per element
// (3 FADD2 + 2 FMUL2 + 4 MUFU + 2 F2FP + 2 HADD2.F32 per pair).  sigma(y) = 1 / (1 + 2^(-y log2 e)): ex2.approx
// overflows to +inf for y << 0 and rcp.approx(+inf) = +0, the correct limit, so no range fix-up code is needed.
__device__ __forceinline__ void silu_split2(uint32_t a0, uint32_t a1, uint64_t bias, float& f0, float& f1, uint32_t& hi,
                                            uint32_t& lo) {
  const uint64_t v = add2(pack2(__uint_as_float(a0), __uint_as_float(a1)), bias);
  const uint64_t t = mul2(v, pack2(-1.4426950408889634f, -1.4426950408889634f));
  float t0, t1;
  unpack2(t, t0, t1);
  const uint64_t d = add2(pack2(ex2_approx(t0), ex2_approx(t1)), pack2(1.0f, 1.0f));
  float d0, d1;
  unpack2(d, d0, d1);
  const uint64_t f = mul2(v, pack2(rcp_approx(d0), rcp_approx(d1)));
  unpack2(f, f0, f1);
  hi = pack_f16x2_sat(f0, f1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float l0, l1;
  unpack2(sub2(f, pack2(hf.x, hf.y)), l0, l1);
  lo = pack_f16x2_sat(l0, l1);
}

// Pre-scaled form used by the tensor-core kernels since r02: the host folds k = -log2(e) into Wf and bf
// (weights.py: TC_PRESCALE), so the accumulator already holds t = k (D + b) and
//   SiLU(y) = y / (1 + 2^t) = t * rcp(k (2^t + 1)),  y = t / k
// costs 3.5 instructions per element (FADD2 bias, 2 MUFU.EX2, FFMA2, 2 MUFU.RCP, FMUL2 per pair) + 2.5 for the fp16
// (hi, lo) split, against 8.5 for the shared-reciprocal form below.  The XU pipe was 20 % busy there: the kernels are
// bound by instruction issue and latency, not by MUFU throughput, so two MUFU per element is the cheaper trade.
// Limits: t -> +inf (y << 0): 2^t = +inf, k * inf = -inf, rcp = -0, f = -0;  t -> -inf (y >> 0): 2^t = 0, f = t / k = y.
constexpr float kSiluK = -1.4426950408889634f;  // -log2(e)
__device__ __forceinline__ uint64_t silu_pre2(uint32_t a0, uint32_t a1, uint64_t bias) {
  const uint64_t t = add2(pack2(__uint_as_float(a0), __uint_as_float(a1)), bias);
  float t0, t1;
  unpack2(t, t0, t1);
  const uint64_t kk = pack2(kSiluK, kSiluK);
  float d0, d1;
  unpack2(fma2(pack2(ex2_approx(t0), ex2_approx(t1)), kk, kk), d0, d1);
  return mul2(t, pack2(rcp_approx(d0), rcp_approx(d1)));
}
__device__ __forceinline__ void silu_pre_split2(uint32_t a0, uint32_t a1, uint64_t bias, float& f0, float& f1,
                                                uint32_t& hi, uint32_t& lo) {
  const uint64_t f = silu_pre2(a0, a1, bias);
  unpack2(f, f0, f1);
  hi = pack_f16x2_sat(f0, f1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float l0, l1;
  unpack2(sub2(f, pack2(hf.x, hf.y)), l0, l1);
  lo = pack_f16x2_sat(l0, l1);
}

// Pre-scaled form with ONE reciprocal per four elements (PF_SILU_SHARED_RCP=1, A/B variant): 1.25 MUFU and 8 instructions
// per element instead of 2 and 6 -- trades XU-pipe time (16 lanes / clk / SM) for issue slots.  2^t is clamped at 2^30 so the
// product of four denominators stays finite (absolute error below 3e-8 for y < -20.8).
__device__ __forceinline__ void silu_pre_split4(const uint32_t (&a)[4], uint64_t bias01, uint64_t bias23, float (&f)[4],
                                                uint32_t& hi01, uint32_t& hi23, uint32_t& lo01, uint32_t& lo23) {
  const uint64_t kk = pack2(kSiluK, kSiluK);
  const uint64_t t01 = add2(pack2(__uint_as_float(a[0]), __uint_as_float(a[1])), bias01);
  const uint64_t t23 = add2(pack2(__uint_as_float(a[2]), __uint_as_float(a[3])), bias23);
  float t0, t1, t2, t3;
  unpack2(t01, t0, t1);
  unpack2(t23, t2, t3);
  const float e0 = fminf(ex2_approx(t0), 1073741824.0f), e1 = fminf(ex2_approx(t1), 1073741824.0f);
  const float e2 = fminf(ex2_approx(t2), 1073741824.0f), e3 = fminf(ex2_approx(t3), 1073741824.0f);
  float d0, d1, d2, d3;                       // d = k (2^t + 1): f = t / d
  unpack2(fma2(pack2(e0, e1), kk, kk), d0, d1);
  unpack2(fma2(pack2(e2, e3), kk, kk), d2, d3);
  const float p01 = d0 * d1, p23 = d2 * d3;
  const float r = rcp_approx(p01 * p23);
  const float r01 = r * p23, r23 = r * p01;   // 1 / (d0 d1), 1 / (d2 d3)
  const uint64_t f01 = mul2(t01, mul2(pack2(r01, r01), pack2(d1, d0)));
  const uint64_t f23 = mul2(t23, mul2(pack2(r23, r23), pack2(d3, d2)));
  unpack2(f01, f[0], f[1]);
  unpack2(f23, f[2], f[3]);
  hi01 = pack_f16x2_sat(f[0], f[1]);
  hi23 = pack_f16x2_sat(f[2], f[3]);
  const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi01));
  const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi23));
  float l0, l1, l2, l3;
  unpack2(sub2(f01, pack2(h01.x, h01.y)), l0, l1);
  unpack2(sub2(f23, pack2(h23.x, h23.y)), l2, l3);
  lo01 = pack_f16x2_sat(l0, l1);
  lo23 = pack_f16x2_sat(l2, l3);
}

// Four elements per call with ONE reciprocal: 1 / d_k = (1 / (d0 d1 d2 d3)) * (product of the other three).  EPI-B is
// bound by the MUFU pipe (16 lanes / clk / SM: 8 cycles per warp instruction and scheduler), so this trades 3 of every
// 8 MUFU operations for 5 multiplies: 5 MUFU and 34 instructions per 4 elements instead of 8 and 26.  2^t is clamped at
// 2^30 so that the product of four denominators stays finite: for y < -20.8 the result is y * 2^-30 instead of
// y * e^y, an absolute error below 3e-8.
__device__ __forceinline__ void silu_split4(const uint32_t (&a)[4], uint64_t bias01, uint64_t bias23, float (&f)[4],
                                            uint32_t& hi01, uint32_t& hi23, uint32_t& lo01, uint32_t& lo23) {
  const uint64_t kNegLog2e = pack2(-1.4426950408889634f, -1.4426950408889634f), kOne = pack2(1.0f, 1.0f);
  const uint64_t v01 = add2(pack2(__uint_as_float(a[0]), __uint_as_float(a[1])), bias01);
  const uint64_t v23 = add2(pack2(__uint_as_float(a[2]), __uint_as_float(a[3])), bias23);
  float t0, t1, t2, t3;
  unpack2(mul2(v01, kNegLog2e), t0, t1);
  unpack2(mul2(v23, kNegLog2e), t2, t3);
  const float e0 = fminf(ex2_approx(t0), 1073741824.0f), e1 = fminf(ex2_approx(t1), 1073741824.0f);
  const float e2 = fminf(ex2_approx(t2), 1073741824.0f), e3 = fminf(ex2_approx(t3), 1073741824.0f);
  float d0, d1, d2, d3;
  unpack2(add2(pack2(e0, e1), kOne), d0, d1);
  unpack2(add2(pack2(e2, e3), kOne), d2, d3);
  const float p01 = d0 * d1, p23 = d2 * d3;
  const float r = rcp_approx(p01 * p23);
  const float r01 = r * p23, r23 = r * p01;  // 1 / (d0 d1), 1 / (d2 d3)
  const uint64_t f01 = mul2(v01, mul2(pack2(r01, r01), pack2(d1, d0)));
  const uint64_t f23 = mul2(v23, mul2(pack2(r23, r23), pack2(d3, d2)));
  unpack2(f01, f[0], f[1]);
  unpack2(f23, f[2], f[3]);
  hi01 = pack_f16x2_sat(f[0], f[1]);
  hi23 = pack_f16x2_sat(f[2], f[3]);
  const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi01));
  const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi23));
  float l0, l1, l2, l3;
  unpack2(sub2(f01, pack2(h01.x, h01.y)), l0, l1);
  unpack2(sub2(f23, pack2(h23.x, h23.y)), l2, l3);
  lo01 = pack_f16x2_sat(l0, l1);
  lo23 = pack_f16x2_sat(l2, l3);
}

// ---------------------------------------------------------------- single-pass fp16 mode (PF_FLAG_FP16_SINGLE_PASS)
// SiLU on packed fp16 pairs: silu(y) = h + h tanh(h), h = y / 2, with ONE MUFU op per PAIR of elements
// (tanh.approx.f16x2, relative error ~2^-11 -- the precision of the fp16 operand it feeds) instead of the four of the
// fp32-parity form: 5 instructions per pair (FADD2, F2FP, HMUL2, MUFU.TANH, HFMA2) against 13.
__device__ __forceinline__ uint32_t tanh_h2(uint32_t x) {
  uint32_t r;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t silu_h2(uint32_t a0, uint32_t a1, uint64_t bias) {
  // the accumulator and the bias carry the host-side factor k = -log2(e) (see silu_pre2): h = y / 2 = t * (0.5 / k)
  float y0, y1;
  const uint64_t hk = pack2(0.5f / kSiluK, 0.5f / kSiluK);
  unpack2(mul2(add2(pack2(__uint_as_float(a0), __uint_as_float(a1)), bias), hk), y0, y1);
  const uint32_t h = pack_f16x2_sat(y0, y1);
  return hfma2_u(h, tanh_h2(h), h);
}

// ---------------------------------------------------------------- fp32 -> (hi, lo) bf16 split, packed pairs
// x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|; word = {low half: element 2j, high half: element 2j+1}
__device__ __forceinline__ void split_pack(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}


// ---------------------------------------------------------------- additions used by the fused GVP-chain kernels
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread; the MMA scheduler must never sleep on one tile slot)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait with a sleep between polls, for waits that are known to be long (an epilogue warp waiting for its slot's S job:
// >= 1.7 k cycles).  mbar_wait's try_wait returns every ~50 cycles on this part: its polls were 14 % of all issued
// instructions of the edge kernel, taken from the other slot's working warps on the same scheduler.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_try(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma / bulk copies reading smem)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// one lane of a converged warp (elect.sync): the compiler keeps the guarded tcgen05 issue on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// named barrier + AND-reduction of a predicate over its participants (bar.red): one barrier that is also a vote
__device__ __forceinline__ bool named_bar_and(int id, int threads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "bar.red.and.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(r)
      : "r"(id), "r"(threads), "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace tc
}  // namespace pf
