// sysmem_read_probe.cu -- how fast do SM-originated reads of PINNED HOST memory go, by access shape?  The env step's kernels
// read the host-resident simulator state in place (unified addressing); this measures a 393 KB read (the dof state of 4096
// envs) and a 4 MB read as 8-byte / 16-byte ld.global.cv per thread and as 1 KB / 4 KB cp.async.bulk per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/sysmem_read_probe profiles/sysmem_read_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void rd8(const float2* src, float* sink, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float2 v = __ldcv(src + i); if (v.x == 123.456f) sink[0] = v.y; }
}
__global__ void rd16(const float4* src, float* sink, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float4 v = __ldcv(src + i); if (v.x == 123.456f) sink[0] = v.y; }
}
__global__ void rdbulk(const uint8_t* src, float* sink, int bytes_per_cta, long long total) {
  extern __shared__ __align__(128) uint8_t buf[];
  __shared__ uint64_t bar;
  const long long off = (long long)blockIdx.x * bytes_per_cta;
  if (off >= total) return;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes_per_cta) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src + off), "r"(bytes_per_cta), "r"(b) : "memory");
  }
  __syncthreads();
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
  } while (!ok);
  if (reinterpret_cast<float*>(buf)[threadIdx.x] == 123.456f) sink[0] = 1.f;
}

template <class F> float time_us(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / reps;
}

int main() {
  float* sink; cudaMalloc(&sink, 16);
  for (long long bytes : {393216LL, 4LL << 20}) {
    void* h; cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    for (long long i = 0; i < bytes / 4; ++i) reinterpret_cast<float*>(h)[i] = (float)i;
    const int n8 = (int)(bytes / 8), n16 = (int)(bytes / 16);
    const float t8 = time_us([&] { rd8<<<(n8 + 127) / 128, 128>>>((const float2*)h, sink, n8); }, 50);
    const float t16 = time_us([&] { rd16<<<(n16 + 127) / 128, 128>>>((const float4*)h, sink, n16); }, 50);
    const float tb1 = time_us([&] { rdbulk<<<(int)(bytes / 1024), 128, 1024>>>((const uint8_t*)h, sink, 1024, bytes); }, 50);
    const float tb4 = time_us([&] { rdbulk<<<(int)(bytes / 4096), 128, 4096>>>((const uint8_t*)h, sink, 4096, bytes); }, 50);
    const float tb16 = time_us([&] { rdbulk<<<(int)(bytes / 16384), 128, 16384>>>((const uint8_t*)h, sink, 16384, bytes); }, 50);
    void* d; cudaMalloc(&d, bytes);
    const float tc = time_us([&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); }, 50);
    printf("%7lld KB from pinned host: ld.cv 8 B %6.1f us (%5.1f GB/s) | ld.cv 16 B %6.1f us (%5.1f) | bulk 1 KB/CTA %6.1f us (%5.1f) | bulk 4 KB/CTA "
           "%6.1f us (%5.1f) | bulk 16 KB/CTA %6.1f us (%5.1f) | cudaMemcpyAsync %6.1f us (%5.1f)\n", bytes >> 10, t8, bytes / t8 / 1e3, t16,
           bytes / t16 / 1e3, tb1, bytes / tb1 / 1e3, tb4, bytes / tb4 / 1e3, tb16, bytes / tb16 / 1e3, tc, bytes / tc / 1e3);
    cudaFree(d); cudaFreeHost(h);
  }
  return 0;
}
