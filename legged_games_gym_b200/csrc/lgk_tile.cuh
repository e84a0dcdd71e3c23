// lgk_tile.cuh -- pieces shared by the post-physics kernels: PTX helpers (TMA bulk copies, mbarriers, named barriers),
// warp reductions and the observation-noise formula.
#pragma once
#include "lgk_step_device.cuh"

namespace lgk {

constexpr int kTile = 32;          // envs per K1 tile (lane = env)
constexpr int kK1Threads = 4 * kTile;
// scan frame of one env: [zn, wn, root_x, root_y] (pre-reset yaw frame, LR:853-854) + [root_z_post_reset - 0.5] (LR:225)
constexpr int kFrameFloats = 8;
constexpr int kHeadCols = 48;      // proprioceptive observation columns (LR:212-222)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// L2 prefetch of a contiguous chunk (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// observation noise (LR:229-230): obs + (2u - 1) * noise_scale, as ONE fused multiply-add on the individually rounded
// observation -- written out so that every kernel rounds identically
__device__ __forceinline__ float noise_unit(uint32_t word) { return f_fma(2.0f, u32_to_uniform(word), -1.0f); }
__device__ __forceinline__ float noisy_obs(float v, uint32_t word, float scale) { return f_fma(noise_unit(word), scale, v); }

// host launchers of the two step kernels (lgk_post_k1.cu / lgk_post_physics.cu)
int launch_k1(const LgkStepParams* p, cudaStream_t st);

// internal phase_mask bit (never accepted from callers): the launch belongs to lgk_post_physics_finalize, whose finalize
// CTA rides in K2's grid.  K1 then publishes the step in step_counter_dev[1] for K2 and that CTA.
constexpr int kPhaseFusedFin = 0x100;

}  // namespace lgk
