// umma_rate_probe.cu -- how many cycles does one tcgen05.mma kind::tf32 (M = 128, K = 8) take on this part, as a function
// of N, of where the A operand lives (shared memory / tensor memory) and of how many B rows are re-read?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/umma_rate_probe profiles/umma_rate_probe.cu && profiles/umma_rate_probe
//
// One CTA per SM (148), one issuing thread, `reps` MMAs back to back into the same accumulator block, one commit, wait.
// Operands are zeros (only the rate matters).  Prints cycles per MMA and the implied dense TF32 rate of the whole chip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n, int kind_bits) {   // kind_bits: 2 = tf32, 1 = bf16 (A/B format field)
  return (1u << 4) | ((uint32_t)kind_bits << 7) | ((uint32_t)kind_bits << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int KIND>   // 0: tf32, 1: f16 (bf16 operands)
__global__ void __launch_bounds__(128, 1) probe(int n, int ts, int reps, int bstep, long long* out, int commit_every = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_addr(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t s_tmem;
  const uint32_t a_addr = smem_addr(smem), b_addr = a_addr + 16384;      // A: 128 rows x 128 B; B: up to 256 rows x 128 B x 4 chunks
  for (int i = threadIdx.x; i < (16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n, KIND == 0 ? 2 : 1);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint32_t ks = (uint32_t)(i & 3), chunk = (uint32_t)((i >> 2) % bstep);
      const uint64_t bd = make_smem_desc(b_addr + chunk * 32768 + ks * 32);
      const uint32_t acc = i ? 1u : 0u;
      if (ts) {
        const uint32_t a_t = tmem + 256 + (uint32_t)((i * 8) & 127);
        if (KIND == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                       ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                       ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      } else {
        const uint64_t ad = make_smem_desc(a_addr + ks * 32);
        if (KIND == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      if (commit_every && ((i + 1) & (commit_every - 1)) == 0)     // a ring-stage release: nobody waits on bar2
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar2)) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    uint32_t ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(smem_addr(&bar)) : "memory");
    } while (!ok);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}


// ---- issue-loop probe: groups of four N-wide TS MMAs (descriptors precomputed), optionally followed by a commit, a wait on an
// already completed barrier, and `spin` dependent integer instructions: does work done by the issuing thread between
// MMAs overlap with the tensor pipe (cycles per group stay at 4 x MMA time) or add to it?
__global__ void __launch_bounds__(128, 1) group_probe(int n, int groups, int do_commit, int do_wait, int spin, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_addr(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t s_tmem;
  const uint32_t b_addr = smem_addr(smem) + 16384;
  for (int i = threadIdx.x; i < (16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar3)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar3)) : "memory");     // phase 0 of bar3 is complete
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n, 2);
    uint64_t bd[4];
    for (int k = 0; k < 4; ++k) bd[k] = make_smem_desc(b_addr + k * 32);
    uint32_t junk = (uint32_t)groups;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (do_wait) {
        uint32_t ok;
        do {
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                       : "=r"(ok) : "r"(smem_addr(&bar3)) : "memory");
        } while (!ok);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                     ::"r"(tmem), "r"(tmem + 256 + 8 * k), "l"(bd[k]), "r"(idesc), "r"((g | k) ? 1u : 0u) : "memory");
      if (do_commit)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar2)) : "memory");
      for (int j = 0; j < spin; ++j) junk = junk * 1664525u + 1013904223u;       // dependent chain: ~4-5 cycles per step
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    uint32_t ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(smem_addr(&bar)) : "memory");
    } while (!ok);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0 + (junk == 77u); }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---- CTA-pair probe: cta_group::2, M = 256 (128 rows in each CTA's tensor memory), N wide, A from TMEM, each CTA holds half
// of B's rows in its own shared memory; the leader CTA's thread issues for both, one multicast commit ends the run.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_probe(int n, int groups, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_addr(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t b_addr = smem_addr(smem) + 16384;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const long long t0 = clock64();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    uint64_t bd[4];
    for (int k = 0; k < 4; ++k) bd[k] = make_smem_desc(b_addr + k * 32);
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                     ::"r"(tmem), "r"(tmem + 256 + 8 * k), "l"(bd[k]), "r"(idesc), "r"((g | k) ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_addr(&bar)), "h"((uint16_t)3) : "memory");
  }
  if (threadIdx.x == 0) {
    uint32_t ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(smem_addr(&bar)) : "memory");
    } while (!ok);
    if (blockIdx.x < 2) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---- weight-stream probe: one producer thread streams `total` bytes from an L2-resident buffer through a ring of `stages`
// stages of `stage_bytes` (cp.async.bulk + mbarrier tx); one consumer thread frees a stage as soon as it is full.
// Reports bytes per cycle per SM: what a weight ring of that geometry can deliver at best.
__global__ void __launch_bounds__(64, 1) stream_probe(const uint8_t* src, int total, int stage_bytes, int stages, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_addr(raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ntiles = total / stage_bytes;
  auto wait = [](uint32_t bar, uint32_t par) {
    uint32_t ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    } while (!ok);
  };
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % stages;
      wait(smem_addr(&empty[s]), ((i / stages) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&full[s])), "r"(stage_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(smem) + s * stage_bytes), "l"(src + (size_t)i * stage_bytes), "r"(stage_bytes), "r"(smem_addr(&full[s])) : "memory");
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % stages;
      wait(smem_addr(&full[s]), (i / stages) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&empty[s])) : "memory");
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 16384 + 4 * 32768 + 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int reps = 512;
  printf("kind  A-from  N    B-chunks  cycles/MMA(issue)  cycles/MMA(done)  chip TFLOP/s at %.3f GHz\n", clk_khz / 1e6);
  for (int kind = 0; kind < 2; ++kind)
    for (int ts = 0; ts < 2; ++ts)
      for (int n : {64, 128, 256})
        for (int bstep : {1, 4}) {
          for (int w = 0; w < 3; ++w) {
            if (kind == 0) probe<0><<<148, 128, smem>>>(n, ts, reps, bstep, d);
            else probe<1><<<148, 128, smem>>>(n, ts, reps, bstep, d);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          const double cyc = (double)h[1] / reps;
          const double k = kind == 0 ? 8.0 : 16.0;
          const double tflops = 2.0 * 128 * n * k / cyc * 148 * (clk_khz * 1e3) / 1e12;
          printf("%s  %s  %3d  %d  %8.1f  %8.1f  %8.1f\n", kind == 0 ? "tf32" : "bf16", ts ? "tmem" : "smem", n, bstep,
                 (double)h[0] / reps, cyc, tflops);
        }
  cudaFuncSetAttribute(group_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("\nissue loop (tf32, A from tmem, groups of 4 MMAs): N  commit  wait  spin  cycles/group(issue)  cycles/group(done)\n");
  for (int n : {128, 256})
    for (int cfg = 0; cfg < 7; ++cfg) {
      const int cm[7] = {0, 1, 1, 0, 0, 1, 1}, wt[7] = {0, 0, 1, 0, 0, 1, 1}, sp[7] = {0, 0, 0, 25, 50, 25, 50};
      for (int w = 0; w < 3; ++w) group_probe<<<148, 128, smem>>>(n, 128, cm[cfg], wt[cfg], sp[cfg], d);
      cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%3d  %d  %d  %2d  %8.1f  %8.1f\n", n, cm[cfg], wt[cfg], sp[cfg], (double)h[0] / 128, (double)h[1] / 128);
    }
  cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("\nCTA pair (cta_group::2, M=256, tf32, A from tmem, groups of 4 MMAs): N  cycles/MMA seen by CTA 0 / CTA 1\n");
  for (int n : {128, 256}) {
    for (int w = 0; w < 3; ++w) pair_probe<<<148, 128, 64 * 1024>>>(n, 128, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%3d  %8.1f  %8.1f   (%.1f TFLOP/s chip)\n", n, (double)h[0] / 512, (double)h[1] / 512,
           2.0 * 256 * n * 8 / ((double)h[0] / 512) * 74 * (clk_khz * 1e3) / 1e12);
  }
  // ---- weight stream
  uint8_t* w;
  const int total = 1152 * 1024;                // one net's packed weights
  cudaMalloc(&w, total);
  cudaMemset(w, 0, total);
  cudaFuncSetAttribute(stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\nweight stream (1.15 MB, L2-resident): CTAs  stage KB  stages  in flight KB  B/cycle/SM  us per pass\n");
  for (int ctas : {64, 148})
    for (int stage_kb : {8, 16, 32})
      for (int stages : {2, 3, 4, 5, 6, 8}) {
        if (stage_kb * stages > 192) continue;
        for (int wi = 0; wi < 3; ++wi) stream_probe<<<ctas, 64, 200 * 1024>>>(w, total, stage_kb * 1024, stages, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%4d  %3d  %d  %4d  %7.1f  %7.2f\n", ctas, stage_kb, stages, stage_kb * stages, (double)total / h[0], h[0] / (clk_khz * 1e-3));
      }
  return 0;
}
