// lgk_abi.cu -- error plumbing, version, RNG tap dump, L2 flush (bench helper).
#include "lgk_math.cuh"
#include <atomic>
#include <cstdio>
#include <cstring>

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

}  // namespace lgk

using namespace lgk;

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
