// lgk_common.cuh -- shared helpers for the liblgk kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/lgk.h"

#if defined(__CUDACC__)
#define LGK_HD __host__ __device__ __forceinline__
#define LGK_D __device__ __forceinline__
#else
#define LGK_HD inline
#define LGK_D inline
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

#define LGK_REQUIRE(cond, msg) do { if (!(cond)) return lgk::set_error(LGK_ERR_ARG, msg); } while (0)
#define LGK_ALIGNED16(ptr, msg) do { if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return lgk::set_error(LGK_ERR_ALIGN, msg); } while (0)

}  // namespace lgk
