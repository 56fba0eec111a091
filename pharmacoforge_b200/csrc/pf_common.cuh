// Shared helpers for the pharmacoforge_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pharmacoforge_b200.h"

namespace pf {

void set_error(const char* fmt, ...);
void count_launch();

// Optional per-site CUDA-event timing (bench.py's roofline leg).  Sites index the kernels of one denoiser call.
enum ProfSite { kSiteGraph = 0, kSitePlan, kSiteEncode, kSiteFF, kSitePF, kSitePP, kSiteFP, kSiteUpdPharm,
                kSiteUpdProt, kSiteNoise, kSitePosterior, kNumSites };
void prof_begin(int site, cudaStream_t st);
void prof_end(int site, cudaStream_t st);

#define PF_CHECK_ARG(cond, msg)                  \
  do {                                           \
    if (!(cond)) {                               \
      pf::set_error("bad argument: %s", msg);    \
      return PF_ERR_BAD_ARG;                     \
    }                                            \
  } while (0)

#define PF_CHECK_LAUNCH(name)                                              \
  do {                                                                     \
    cudaError_t e_ = cudaPeekAtLastError();                                \
    if (e_ != cudaSuccess) {                                               \
      pf::set_error("%s: %s", name, cudaGetErrorString(e_));               \
      (void)cudaGetLastError();                                            \
      return PF_ERR_LAUNCH;                                                \
    }                                                                      \
    pf::count_launch();                                                    \
  } while (0)

constexpr int kHidden = PF_HIDDEN;
constexpr int kVec = PF_VEC;
constexpr int kRbf = PF_RBF;
constexpr int kVRow = 3 * kVec;  // 48 floats per node vector row, layout [c][u]
constexpr int kNumSms = 148;  // B200; the persistent kernels size their grids with num_sms() (queried per device)

// Per-device launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a property of the (function, device) pair,
// so a process-wide "configured" flag breaks the second GPU of a process; the flag is kept per device ordinal.  Races are
// benign (the attribute call is idempotent), the flags are only ever set.
constexpr int kMaxDevices = 64;
struct PerDeviceFlag {
  volatile bool done[kMaxDevices];
};
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < kMaxDevices ? d : 0;
}
inline int num_sms() {  // SM count of the current device (148 on B200), cached per ordinal
  static volatile int cache[kMaxDevices];
  const int d = current_device();
  if (cache[d] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = kNumSms;
    cache[d] = n;
  }
  return cache[d];
}

__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }

// Section offsets of one packed GVP (see pharmacoforge_b200.h).
struct GvpLayout {
  int wh, wu, wf, bf, wg, bg, total;
};
__host__ __device__ inline GvpLayout gvp_layout(int vi, int vo, int si, int so) {
  const int vh = vi > vo ? vi : vo;
  GvpLayout L;
  L.wh = 0;
  L.wu = L.wh + round_up4(vi * vh);
  L.wf = L.wu + round_up4(vh * vo);
  L.bf = L.wf + round_up4(si + vh) * so;
  L.wg = L.bf + round_up4(so);
  L.bg = L.wg + round_up4(so * vo);
  L.total = L.bg + round_up4(vo);
  return L;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace pf
