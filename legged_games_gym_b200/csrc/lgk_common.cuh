// lgk_common.cuh -- shared helpers for the liblgk kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/lgk.h"

#if defined(__CUDACC__)
#define LGK_HD __host__ __device__ __forceinline__
#define LGK_D __device__ __forceinline__
#define LGK_COLD static __device__ __noinline__
#else
#define LGK_HD inline
#define LGK_D inline
#define LGK_COLD static inline
#endif

// ---- individually-rounded fp32 ops (no FMA contraction).  On the device these are the _rn intrinsics;
// on the host (tests/hostcheck builds with -ffp-contract=off) plain operators are already exact.
#if defined(__CUDA_ARCH__)
LGK_D float f_mul(float a, float b) { return __fmul_rn(a, b); }
LGK_D float f_add(float a, float b) { return __fadd_rn(a, b); }
LGK_D float f_sub(float a, float b) { return __fsub_rn(a, b); }
LGK_D float f_div(float a, float b) { return __fdiv_rn(a, b); }
LGK_D float f_sqrt(float a) { return __fsqrt_rn(a); }
LGK_D float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
LGK_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
LGK_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
LGK_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
LGK_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
LGK_HD float f_sqrt(float a) { return sqrtf(a); }
LGK_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

#if defined(__CUDACC__)
// ---- Blackwell packed-fp32 (f32x2) helpers: one issue slot, two individually IEEE-rounded fp32 results
namespace lgk {
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
}  // namespace lgk
#endif

namespace lgk {

constexpr int kWarp = 32;
constexpr int kDof = LGK_NUM_DOF;

// host-side error plumbing (lgk_abi.cu)
int set_error(int code, const char* msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);
int ensure_func_attr(const void* func, int smem_bytes, const char* name, bool max_carveout = false);   // per device

// ---- programmatic dependent launch (PDL): the kernels of one env step form a chain on one stream; launched with
// programmatic stream serialization each may start (prologue, barrier init, constant staging) while its predecessor
// drains, and blocks in pdl_wait() until the predecessor's writes are visible.  Works under stream capture (the edges
// become programmatic graph edges).  lgk_set_pdl(0) switches back to plain launches.
extern int g_pdl;
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define LGK_REQUIRE(cond, msg) do { if (!(cond)) return lgk::set_error(LGK_ERR_ARG, msg); } while (0)
#define LGK_ALIGNED16(ptr, msg) do { if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return lgk::set_error(LGK_ERR_ALIGN, msg); } while (0)

}  // namespace lgk
