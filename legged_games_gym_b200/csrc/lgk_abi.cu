// lgk_abi.cu -- error plumbing, version, RNG tap dump, L2 flush (bench helper).
#include "lgk_math.cuh"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace lgk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
int g_pdl = 1;

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LGK_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return LGK_ERR_CUDA_BASE + (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// cudaFuncSetAttribute is per (device, function): remember the largest dynamic shared-memory size set for every pair so that
// a process driving several GPUs (or switching devices) configures each of them
int ensure_func_attr(const void* func, int smem_bytes, const char* name, bool max_carveout) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> done;
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  std::lock_guard<std::mutex> lk(mu);
  int& have = done[std::make_pair(dev, func)];
  if (smem_bytes <= have) return LGK_OK;
  if (int rc = check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), name)) return rc;
  if (max_carveout) cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  have = smem_bytes;
  return LGK_OK;
}

// kind 0: uniforms, 1: raw words, 2: Box-Muller normals on pairs (ACT stream convention)
__global__ void rng_dump_kernel(uint64_t seed, int step, long long env_off, int n, int stream_id, int count, int kind,
                                void* out) {
  const long long total = (long long)n * count;
  const RngKey key = make_key(seed, step);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / count), j = (int)(i % count);
    const uint32_t genv = (uint32_t)(env_off + e);
    if (kind == 2) {
      const int pr = j >> 1;      // pair index: uniforms 2pr, 2pr+1
      const U4 r = rng_block(key, genv, stream_id, (uint32_t)(pr >> 1));
      const uint32_t wa = (pr & 1) ? r.z : r.x, wb = (pr & 1) ? r.w : r.y;
      const float u1 = 1.0f - u32_to_uniform(wa), u2 = u32_to_uniform(wb);
      const float rad = sqrtf(-2.0f * logf(u1)), th = 6.283185307179586f * u2;
      reinterpret_cast<float*>(out)[i] = (j & 1) ? rad * sinf(th) : rad * cosf(th);
      continue;
    }
    uint32_t blk = (uint32_t)(j >> 2);
    int w = j & 3;
    if (stream_id == LGK_STREAM_OBS) { blk = (uint32_t)(32 * (j >> 7) + (j & 31)); w = (j >> 5) & 3; }
    const uint32_t word = pick(rng_block(key, genv, stream_id, blk), w);
    if (kind == 0) reinterpret_cast<float*>(out)[i] = u32_to_uniform(word);
    else reinterpret_cast<uint32_t*>(out)[i] = word;
  }
}

__global__ void l2_flush_kernel(uint4* buf, long long n16, uint32_t tag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    buf[i] = make_uint4(tag, tag, tag, tag);
}

// Host-sim pipeline (sim state in PINNED host memory, reference sim_device=cpu): pinned allocations are mapped into the
// device address space (unified addressing), so a kernel can pull them over PCIe with plain 16-byte loads -- `ld.cv`
// (never serve from a cache: the host rewrites the buffer every step) -- or push results with plain stores.  Inside a
// captured graph this is 2-3x faster than a memcpy node for the 0.2-1 MB tensors of an env step (scratch/pcie_probe.py).
__global__ void __launch_bounds__(256) pinned_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long n16,
                                                          int from_host) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {          // four loads in flight per thread: PCIe latency is ~1.5 us
    uint4 a, b, c, d;
    if (from_host) { a = __ldcv(src + i); b = __ldcv(src + i + stride); c = __ldcv(src + i + 2 * stride); d = __ldcv(src + i + 3 * stride); }
    else { a = src[i]; b = src[i + stride]; c = src[i + 2 * stride]; d = src[i + 3 * stride]; }
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = from_host ? __ldcv(src + i) : src[i];
}

// set_*_tensor_indexed of a host-resident sim (LR:409-412, 433-436): only the rows of the envs reset this step go back
// over PCIe.  ids / count are the device-side outputs of lgk_finalize_step; one warp per listed env.
__global__ void __launch_bounds__(128) pinned_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int row_floats,
                                                          const int32_t* __restrict__ ids, const int32_t* __restrict__ count,
                                                          int row_stride, int row_offset, int max_ids) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = min(*count, max_ids);
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 4 + (threadIdx.x >> 5); i < n; i += gridDim.x * 4) {
    const size_t row = ((size_t)ids[i] * row_stride + row_offset) * row_floats;
    for (int c = lane; c < row_floats; c += 32) dst[row + c] = src[row + c];
  }
}

}  // namespace lgk

using namespace lgk;

extern "C" int lgk_copy_rows_to_pinned(void* dst_pinned_host, const void* src_device, int32_t row_floats, const int32_t* ids,
                                       const int32_t* count, int32_t row_stride, int32_t row_offset, int32_t max_ids, void* stream) {
  LGK_REQUIRE(dst_pinned_host && src_device && ids && count && row_floats > 0 && row_stride > 0 && row_offset >= 0 && max_ids > 0,
              "copy_rows_to_pinned: bad arguments");
  int blocks = (max_ids + 3) / 4;
  blocks = blocks > 148 * 4 ? 148 * 4 : blocks;
  const cudaError_t e = launch_chained(pinned_rows_kernel, dim3(blocks), dim3(128), 0, (cudaStream_t)stream,
                                       reinterpret_cast<float*>(dst_pinned_host), reinterpret_cast<const float*>(src_device),
                                       (int)row_floats, ids, count, (int)row_stride, (int)row_offset, (int)max_ids);
  count_launch();
  return check_cuda(e, "pinned_rows_kernel launch");
}

static int pinned_copy(void* dst, const void* src, int64_t bytes, int from_host, void* stream) {
  LGK_REQUIRE(dst != nullptr && src != nullptr && bytes > 0 && bytes % 16 == 0, "pinned copy: bad arguments (bytes must be a multiple of 16)");
  LGK_ALIGNED16(dst, "pinned copy dst"); LGK_ALIGNED16(src, "pinned copy src");
  const long long n16 = bytes / 16;
  long long blocks = (n16 + 4 * 256 - 1) / (4 * 256);
  blocks = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  const cudaError_t e = launch_chained(pinned_copy_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                                       reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n16, from_host);
  count_launch();
  return check_cuda(e, "pinned_copy_kernel launch");
}
extern "C" int lgk_copy_from_pinned(void* dst_device, const void* src_pinned_host, int64_t bytes, void* stream) {
  return pinned_copy(dst_device, src_pinned_host, bytes, 1, stream);
}
extern "C" int lgk_copy_to_pinned(void* dst_pinned_host, const void* src_device, int64_t bytes, void* stream) {
  return pinned_copy(dst_pinned_host, src_device, bytes, 0, stream);
}

extern "C" const char* lgk_last_error_string(void) { return g_err; }
extern "C" int lgk_abi_version(void) { return LGK_ABI_VERSION; }
extern "C" int lgk_set_pdl(int enable) { const int prev = g_pdl; g_pdl = enable ? 1 : 0; return prev; }
extern "C" int64_t lgk_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int lgk_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(LgkTorqueParams);
    case 1: return (int)sizeof(LgkLstmWeights);
    case 2: return (int)sizeof(LgkStepParams);
    case 3: return (int)sizeof(LgkPolicyParams);
    case 4: return (int)sizeof(LgkGameParams);
    default: return -1;
  }
}

extern "C" int lgk_rng_dump(uint64_t seed, int32_t step, int64_t env_id_offset, int32_t num_envs, int32_t stream_id,
                            int32_t count, int32_t kind, void* out, void* stream) {
  LGK_REQUIRE(out != nullptr && num_envs > 0 && count > 0, "rng_dump: bad arguments");
  LGK_REQUIRE(kind >= 0 && kind <= 2, "rng_dump: kind must be 0, 1 or 2");
  const long long total = (long long)num_envs * count;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  rng_dump_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, step, env_id_offset, num_envs, stream_id, count, kind, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "rng_dump_kernel launch");
}

extern "C" int lgk_l2_flush(void* scratch, int64_t bytes, void* stream) {
  LGK_REQUIRE(scratch != nullptr && bytes >= 16, "l2_flush: bad arguments");
  LGK_ALIGNED16(scratch, "l2_flush scratch");
  static uint32_t tag = 0;
  l2_flush_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(scratch), bytes / 16, ++tag);
  return check_cuda(cudaGetLastError(), "l2_flush_kernel launch");
}
