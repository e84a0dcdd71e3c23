// lgk_policy_tc.cu -- ActorCritic.act on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// rsl_rl ActorCritic (not vendored in the reference; call sites utils/task_registry.py:37-38,154):
//   actor / critic = Linear(O,h0) ELU Linear(h0,h1) ELU Linear(h1,h2) ELU Linear(h2,A|1)
//
// One CTA owns a tile of 128 environments (= the 128 TMEM lanes) of ONE network and carries it through all four layers
// without touching global memory in between.  Activations never leave TENSOR MEMORY: an accumulator block is drained
// tcgen05.ld -> bias -> ELU -> TF32 -> tcgen05.st back into the SAME columns, and the next layer's MMAs take their A
// operand from there (tcgen05.mma with A in TMEM: lane = row, one 32-bit column per k).  Only layer 1 reads its A
// operand (the observation tile) from shared memory.
//
// What the shape of the kernel follows from (profiles/umma_rate_probe.cu, measured on B200): one tcgen05.mma
// kind::tf32 M=128 K=8 takes 126 cycles for ANY N <= 128, 138 cycles for N = 256 with A in TMEM and 171 cycles for
// N = 256 with A in shared memory -- so every MMA is issued as wide as the layer allows (N = h0/2, h1, h2); and one
// cp.async.bulk costs its issuing thread ~270 cycles whatever its size, so weight tiles are as large as the ring allows
// (<= 256 rows x 32 k = 32 KB: 110 B/cycle/SM from L2, against 60 for 16 KB tiles).
//
//   warps 0-15 stage the observation tile into shared memory (TF32, 128-byte swizzled K-major chunks of 32 columns; the
//              first four chunks are released to the tensor pipe while the last four are still loading), later drain
//              accumulator blocks in place, group by group (warp w owns lane quadrant w%4, the four warps of a quadrant
//              split a group's columns); the last drain feeds the h2->A layer on FP32 FFMA2, then Normal sampling (Philox
//              ACT stream, drawn while layer 3 is still running) / log-prob / value, spread over all sixteen warps.
//   warp 16    streams pre-packed weight tiles (<=256 rows x 32 k, already swizzled + TF32-rounded by
//              policy_pack_kernel) through a 3-stage 32 KB ring with TMA bulk copies (cp.async.bulk + mbarrier tx).
//   warp 17    owns TMEM (512 columns) and walks the tile schedule the host built (TcPlan::sched, read from the kernel-
//              parameter constant bank) with warp-uniform control flow; one elected lane issues
//              tcgen05.mma.cta_group::1.kind::tf32 with uniform-register operands (one UTCHMMA per MMA) and the
//              tcgen05.commit that releases ring stages and signals finished accumulator blocks.
//
// TMEM columns (fp32 accumulators / TF32 activations), H = h0/2, drain groups of G columns (up to four per block);
// default 512-256-128 nets: H = 256, G = 64:
//   S = [0,H)  one half of layer 1's columns at a time      ACC2 = [H, H+h1)  layer-2 accumulators      layer 3 -> [0,h2)
//   tensor pipe: L1a>S | L2a(A=S g0) .. L2a(A=S g3) | L1b>S | L2b(g0) .. L2b(g3) | L3(A=ACC2 g0) .. L3(A=ACC2 g3)
//   epilogue   :        drain S g0 .. g3 (in place)           drain S g0 .. g3     drain ACC2 g0 .. g3           final layer
// Layer 2 accumulates one k-group at a time as its inputs appear, so a drain of G columns (not of the whole block) is what
// the tensor pipe waits for.  The MMAs of one thread execute in issue order, which is what makes the reuse of S safe
// (L1b overwrites S only after L2a has read it).
//
// Numerics: TF32 operands rounded to nearest (cvt.rna), FP32 accumulation in TMEM, biases / ELU / last layer / sampling
// in FP32: within the north_star's 1e-3 for policy outputs (tests/test_gpu_parity.py).
#include "lgk_policy_common.cuh"
#include "lgk_policy_tc_plan.h"
#include <mutex>
#include <unordered_map>

namespace lgk {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                     // fp32 per 128-byte swizzle row
constexpr int kChunkBytes = kTileM * 128;       // one A chunk: [128 rows][32 k] = 16 KB
constexpr int kMaxChunks = 8;                   // observation width <= 256
constexpr int kStageBytes = 32768;              // one weight tile: <= 256 rows x 128 B
constexpr int kStages = 3;
constexpr int kColSplit = 4;                     // warps per TMEM lane quadrant: they split the accumulator columns
constexpr int kEpiWarps = 4 * kColSplit, kEpiThreads = 32 * kEpiWarps;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kTcThreads = 32 * (kEpiWarps + 2);
constexpr int kMaxH2 = 128;
constexpr int kMaxOut = 16;
constexpr int kTmemCols = 512;

// shared-memory carve-up (offsets from a 1024-byte aligned base).  Biases are read from global memory (uniform 16-byte
// loads, L1-resident) and the transposed last layer is loaded into the observation buffer once layer 1 has consumed it, so
// that the tile + a three-stage ring of 32 KB weight tiles fit.
constexpr int kOffA = 0;
constexpr int kOffRing = kOffA + kMaxChunks * kChunkBytes;            // 131072
constexpr int kOffMisc = kOffRing + kStages * kStageBytes;            // 229376
constexpr int kOffBar = kOffMisc + 256;                              // b4[16] | std[16] 1/std[16] log std[16]
constexpr int kSmemBytes = kOffBar + 256 + 1024;                      // + barriers + alignment slack
static_assert(kSmemBytes <= 232448, "policy_tc_kernel: shared memory budget");
// inside the (dead) observation buffer, after layer 1: partial sums of the last layer [4][16][128] fp32 + partial log-probs
// [4][128], then W4^T [h2][16]
constexpr int kOffPartial = 0;
constexpr int kOffW4 = 36864;

// barrier ids of the schedule (TcTile::wait_bar / commit_bar are 1 + id)
enum { kWaitChunk0 = 0, kWaitAct = 8, kWaitAct2 = 12, kNumWaitBars = 16 };     // obs chunk c | S group g activated | ACC2 group g activated
enum { kCommitS = 0, kCommitAcc = 1 };

static int build_schedule(TcPlan* pl, int net) {
  TcTile* out = pl->sched[net];
  int n = 0;
  const int H = pl->q, h1 = pl->h1, h2 = pl->h2, O = pl->o[net], kc1 = pl->kc1[net];
  auto put = [&](int a, int d_col, int rows, int ksteps, bool ts, bool first, int wait_bar, int wait_par, int commit_bar,
                 int layer, int n0, int k0) {
    TcTile t;
    t.a = (uint16_t)a; t.d_col = (uint16_t)d_col; t.rows = (uint16_t)rows; t.ksteps = (uint8_t)ksteps;
    t.flags = (uint8_t)((ts ? 1 : 0) | (first ? 2 : 0));
    t.wait_bar = (uint8_t)wait_bar; t.wait_par = (uint8_t)wait_par; t.commit_bar = (uint8_t)commit_bar;
    t.layer = (uint8_t)layer; t.n0 = (uint16_t)n0; t.k0 = (uint16_t)k0;
    out[n++] = t;
  };
  auto l1 = [&](int half) {     // layer-1 columns [half*H, (half+1)*H) into S, A = observation chunks in shared memory
    for (int kc = 0; kc < kc1; ++kc) {
      const int left = O - kc * kChunkK;
      const int ks = left >= kChunkK ? 4 : (left + 7) / 8;
      put(kc, 0, H, ks, false, kc == 0, half == 0 ? 1 + kWaitChunk0 + kc : 0, 0, kc == kc1 - 1 ? 1 + kCommitS : 0, 0,
          half * H, kc * kChunkK);
    }
  };
  const int ng1 = pl->ng1, G1 = H / ng1, ng2 = pl->ng2, G2 = h1 / ng2;
  auto l2 = [&](int half) {     // layer-2 partial sums over k in [half*H, (half+1)*H), A = the activated groups of S in TMEM
    for (int g = 0; g < ng1; ++g)
      for (int kc = 0; kc < G1 / kChunkK; ++kc)
        put(g * G1 + kc * kChunkK, H, h1, 4, true, half == 0 && g == 0 && kc == 0, kc == 0 ? 1 + kWaitAct + g : 0, half,
            (half == 1 && g == ng1 - 1 && kc == G1 / kChunkK - 1) ? 1 + kCommitAcc : 0, 1, 0, half * H + g * G1 + kc * kChunkK);
  };
  l1(0); l2(0); l1(1); l2(1);
  for (int g = 0; g < ng2; ++g)   // layer 3: A = the activated groups of ACC2, accumulators at column 0 (S is free)
    for (int kc = 0; kc < G2 / kChunkK; ++kc)
      put(H + g * G2 + kc * kChunkK, 0, h2, 4, true, g == 0 && kc == 0, kc == 0 ? 1 + kWaitAct2 + g : 0, 0,
          (g == ng2 - 1 && kc == G2 / kChunkK - 1) ? 1 + kCommitAcc : 0, 2, 0, g * G2 + kc * kChunkK);
  return n;
}

// returns false when the shape does not fit this kernel (the FP32 path in lgk_policy.cu handles it)
bool policy_tc_plan(const LgkPolicyParams* p, TcPlan* pl) {
  const int h0 = p->hidden[0], h1 = p->hidden[1], h2 = p->hidden[2];
  if (h0 <= 0 || h1 <= 0 || h2 <= 0) return false;
  if (h0 != 128 && h0 != 256 && h0 != 512) return false;        // halves of 64 / 128 / 256 columns, drained in groups of 32 or 64
  const int H = h0 / 2;
  if ((h1 != 64 && h1 != 128 && h1 != 256) || H + h1 > kTmemCols) return false;
  if (h2 % 32 != 0 || h2 > kMaxH2 || h2 > H) return false;
  if (p->num_actions < 1 || p->num_actions > kMaxOut) return false;
  if (p->num_obs < 1 || p->num_obs > kMaxChunks * kChunkK || p->num_critic_obs < 1 || p->num_critic_obs > kMaxChunks * kChunkK) return false;
  for (int i = 0; i < 3; ++i)                                   // biases are read with 16-byte loads
    if ((reinterpret_cast<uintptr_t>(p->actor_b[i]) & 15u) || (reinterpret_cast<uintptr_t>(p->critic_b[i]) & 15u)) return false;
  pl->o[0] = p->num_obs; pl->o[1] = p->num_critic_obs;
  pl->h0 = h0; pl->h1 = h1; pl->h2 = h2; pl->q = H; pl->nact = p->num_actions;
  pl->ng1 = H / 32 < 4 ? H / 32 : 4; pl->ng2 = h1 / 32 < 4 ? h1 / 32 : 4;      // drain groups: 4 warps x {8, 16} columns each
  long long off = 0;
  for (int net = 0; net < 2; ++net) {
    pl->kc1[net] = (pl->o[net] + kChunkK - 1) / kChunkK;
    pl->ntiles[net] = build_schedule(pl, net);
    long long bytes = 0;
    for (int t = 0; t < pl->ntiles[net]; ++t) bytes += 128LL * pl->sched[net][t].rows;
    pl->net_bytes[net] = bytes;
    pl->net_off[net] = off;
    off += bytes;
  }
  return true;
}

long long policy_tc_workspace_bytes(const TcPlan& pl) { return pl.net_off[1] + pl.net_bytes[1] + 1024; }

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B (layout type 2), 8-row groups 1024 B apart, descriptor version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }

// A operand from tensor memory (lane = row, 32-bit column per k): D[tmem] (+)= A[tmem] x B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ weight packing
// One block per tile of the schedule.  A tile is the shared-memory image itself: row r holds 32 consecutive k as eight
// 16-byte chunks, logical chunk c at physical chunk c ^ (r & 7) (128-byte swizzle), values rounded to TF32, zero beyond K.
struct PackJobs {
  const float* w[6];        // [net*3 + layer]
  int k[6];                 // input width of that matrix
  float* dst[2];            // packed image of each net
  int ntiles[2];
  uint32_t off[2][kTcMaxTiles];      // float offset of each tile inside its net's image
  TcTile sched[2][kTcMaxTiles];
};

__global__ void __launch_bounds__(256) policy_pack_kernel(const __grid_constant__ PackJobs jobs) {
  const int net = blockIdx.y, t = blockIdx.x;
  if (t >= jobs.ntiles[net]) return;
  const TcTile T = jobs.sched[net][t];
  const float* __restrict__ W = jobs.w[net * 3 + T.layer];
  const int K = jobs.k[net * 3 + T.layer];
  float* __restrict__ dst = jobs.dst[net] + jobs.off[net][t];
  const int total = T.rows * kChunkK;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int row = e >> 5, pc = (e >> 2) & 7, w = e & 3;
    const int lc = pc ^ (row & 7);
    const int n = T.n0 + row, k = T.k0 + lc * 4 + w;
    dst[e] = (k < K) ? to_tf32(W[(size_t)n * K + k]) : 0.f;
  }
}

// ------------------------------------------------------------------ the fused forward
struct TcArgs {
  LgkPolicyParams p;
  TcPlan pl;
  const uint8_t* packed;      // workspace base (1024-byte aligned)
  long long* timeline;        // optional [16] globaltimer stamps of CTA (0,0) (lgk_policy_debug_timeline)
  int dbg_flags;              // bit 0: skip the weight copies, bit 1: skip the MMAs (profiling experiments only)
  int net0;                   // first network of the grid's y dimension (0 actor, 1 critic): LgkPolicyParams.nets
};

constexpr uint32_t kDescHi = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (1ull << 46) | (2ull << 61)) >> 32);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// ELU with one MUFU: exp(x) - 1 = 2^(x*log2e) - 1 for x <= 0 (absolute error ~1e-7, far inside the TF32 rounding that follows)
__device__ __forceinline__ float elu_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.0f;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// tcgen05.ld / st of CW (8, 16) consecutive 32-bit columns of this warp's lane quadrant
template <int CW> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[CW]);
template <int CW> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[CW]);
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <> __device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <> __device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

// biases of this warp's CW columns of a group (global memory, 16-byte aligned): loaded ahead of the accumulators
template <int CW> struct BiasR { float4 b[CW / 4]; };
template <int CW> __device__ __forceinline__ void load_bias(BiasR<CW>& B, const float* __restrict__ bias) {
#pragma unroll
  for (int q4 = 0; q4 < CW / 4; ++q4) B.b[q4] = ldg_f4(bias + 4 * q4);
}

// Activate an accumulator block of `ng` groups of 4*CW columns IN PLACE, group by group: + bias, ELU, round to TF32, store
// back to the same TMEM columns, where the next layer's MMAs read them as their A operand; group g is handed to the issuer
// (barrier bar0 + 8g) as soon as all sixteen warps have stored their CW columns of it.  `on_first` / `on_last` are the
// profiling stamps.
template <int CW, class F0, class F1, class F2>
__device__ __forceinline__ void drain_groups(uint32_t tmem_lane_base, int part, int col0, int ng, const float* __restrict__ bias,
                                             uint32_t bar0, uint32_t wait_bar, uint32_t wait_par, F0 on_full, F1 on_first, F2 on_last) {
  BiasR<CW> B;
  load_bias<CW>(B, bias + part * CW);
  mbar_wait(wait_bar, wait_par);
  tc_fence_after();
  on_full();
  for (int g = 0; g < ng; ++g) {
    const uint32_t taddr = tmem_lane_base + (uint32_t)(col0 + g * 4 * CW + part * CW);
    uint32_t v[CW];
    tmem_ld<CW>(taddr, v);
    BiasR<CW> Bn;
    if (g + 1 < ng) load_bias<CW>(Bn, bias + (g + 1) * 4 * CW + part * CW);
#pragma unroll
    for (int q4 = 0; q4 < CW / 4; ++q4) {
      v[4 * q4 + 0] = __float_as_uint(to_tf32(elu_fast(__uint_as_float(v[4 * q4 + 0]) + B.b[q4].x)));
      v[4 * q4 + 1] = __float_as_uint(to_tf32(elu_fast(__uint_as_float(v[4 * q4 + 1]) + B.b[q4].y)));
      v[4 * q4 + 2] = __float_as_uint(to_tf32(elu_fast(__uint_as_float(v[4 * q4 + 2]) + B.b[q4].z)));
      v[4 * q4 + 3] = __float_as_uint(to_tf32(elu_fast(__uint_as_float(v[4 * q4 + 3]) + B.b[q4].w)));
    }
    tmem_st<CW>(taddr, v);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar0 + 8 * g);
    if (g == 0) on_first();
    if (g + 1 < ng) B = Bn;
  }
  on_last();
}

// last layer for this thread's 32 hidden columns: acc[o] += elu(h) * W4^T[k][o], NQ float4 groups of outputs
// last layer for this thread's 32 hidden columns: acc[o] += elu(h) * W4^T[k][o] for 4*NQ outputs, two outputs per packed
// FFMA2 (acc2[j] = outputs 2j, 2j+1)
template <int NQ>
__device__ __forceinline__ void last_layer_block(const uint32_t (&v)[32], const float4 (&b3)[8], uint32_t w4t_addr,
                                                 f2_t (&acc2)[kMaxOut / 2]) {
#pragma unroll
  for (int jj = 0; jj < 32; jj += 4) {
    const float4 b = b3[jj / 4];
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float h = elu_fast(__uint_as_float(v[jj + u]) + bb[u]);
      const f2_t hh = pack2(h, h);
      const uint32_t wr = w4t_addr + (jj + u) * kMaxOut * 4;
#pragma unroll
      for (int g = 0; g < NQ; ++g) {
        f2_t w01, w23;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w01), "=l"(w23) : "r"(wr + 16 * g));
        acc2[2 * g] = fma2(hh, w01, acc2[2 * g]);
        acc2[2 * g + 1] = fma2(hh, w23, acc2[2 * g + 1]);
      }
    }
  }
}

template <bool PROF>
__global__ void __launch_bounds__(kTcThreads, 1) policy_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);     // 128-byte swizzle atoms need 1024-byte alignment
  const uint32_t s_base = smem_addr(smem);
  const uint32_t abuf = s_base + kOffA, ring = s_base + kOffRing;
  float* s_w4t = reinterpret_cast<float*>(smem + kOffA + kOffW4);       // valid once layer 1 is done with the tile
  float* s_b4 = reinterpret_cast<float*>(smem + kOffMisc);
  float* s_std = s_b4 + 16;
  const uint32_t a_w4t = abuf + kOffW4, a_partial = abuf + kOffPartial;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  const uint32_t bar_full = smem_addr(bars), bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_s = bar_empty + 8 * kStages;              // a half of layer 1 accumulated into S
  const uint32_t bar_acc = bar_s + 8;                          // layer 2 done / layer 3 done
  const uint32_t bar_wait = bar_acc + 8;                       // [kNumWaitBars] what the issuer waits for: obs chunks, activated groups
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 + kNumWaitBars);

  const LgkPolicyParams& p = a.p;
  const TcPlan& pl = a.pl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int net = a.net0 + blockIdx.y;            // 0 actor, 1 critic
  const int m0 = blockIdx.x * kTileM;
  const int N = p.num_envs;
  const int O = pl.o[net], kc1 = pl.kc1[net];
  const int h1 = pl.h1, h2 = pl.h2, H = pl.q;
  const int ntiles = pl.ntiles[net];
  const int nout = net ? 1 : pl.nact;
  const float* const* Wg = net ? p.critic_w : p.actor_w;
  const float* const* Bg = net ? p.critic_b : p.actor_b;

  // ---- one-time setup.  Programmatic dependent launch: barrier init and the TMEM allocation overlap the tail of the
  // preceding kernel on the stream (the env step that writes the observations, or the previous act()); nothing in global
  // memory is read before pdl_wait().
  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_s, 1);
    mbar_init(bar_acc, 1);
    for (int c = 0; c < kMaxChunks; ++c) mbar_init(bar_wait + 8 * (kWaitChunk0 + c), kEpiThreads);
    for (int g = 0; g < 4; ++g) { mbar_init(bar_wait + 8 * (kWaitAct + g), kEpiThreads); mbar_init(bar_wait + 8 * (kWaitAct2 + g), kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  if (tid < kMaxOut) {
    const float sd = (net == 0 && tid < nout) ? p.std[tid] : 1.f;
    s_b4[tid] = tid < nout ? Bg[3][tid] : 0.f;
    s_std[tid] = sd; s_std[16 + tid] = 1.0f / sd; s_std[32 + tid] = logf(sd);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // phase stamps of CTA (0,0): 0 setup done | 1 obs staged | 2 L1a accumulated | 3, 4 its groups activated | 5 L1b accumulated |
  // 6, 7 its groups activated | 8 layer 2 done | 9, 10 its groups activated | 11 layer 3 done | 12 last layer summed | 13 outputs written
  auto stamp = [&](int slot) {
    if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0)) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.timeline[slot] = (long long)t;
    }
  };
  stamp(0);

  if (warp < kEpiWarps) {
    // ================= staging + epilogue warps =================
    // TMEM lane quadrant = warp % 4 (hardware rule); the kColSplit warps of a quadrant split the column blocks
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;                       // TMEM lane == row of the tile
    const uint32_t tmem_lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    {   // observation tile -> A chunks (coalesced 128-byte row segments; lane = k inside the chunk).  A warp owns 8 rows and
        // issues the loads of four chunks at once (32 in flight per thread), then stores chunk by chunk: a chunk goes to the
        // tensor pipe as soon as every warp has stored its rows of it.
      const float* __restrict__ X = net ? p.critic_obs : p.obs;
      constexpr int kRowsPerWarp = kTileM / kEpiWarps, kHalfChunks = kMaxChunks / 2;
      const int r0 = warp * kRowsPerWarp;
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {          // chunks 0-3, then 4-7: the second half's loads fly under the first MMAs
        float vals[kHalfChunks][kRowsPerWarp];
#pragma unroll
        for (int c = 0; c < kHalfChunks; ++c) {
          const int kc = hc * kHalfChunks + c;
#pragma unroll
          for (int i = 0; i < kRowsPerWarp; ++i) {
            const int n = m0 + r0 + i;
            const int k = kc * kChunkK + lane;
            vals[c][i] = (kc < kc1 && n < N && k < O) ? __ldg(X + (size_t)n * O + k) : 0.f;
          }
        }
#pragma unroll
        for (int c = 0; c < kHalfChunks; ++c) {
          const int kc = hc * kHalfChunks + c;
          if (kc < kc1) {
#pragma unroll
            for (int i = 0; i < kRowsPerWarp; ++i) {
              const int r = r0 + i;
              sts_f1(abuf + kc * kChunkBytes + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4, to_tf32(vals[c][i]));
            }
          }
        }
        fence_proxy_async();                      // one generic->async proxy fence per half (it waits for the stores above)
#pragma unroll
        for (int c = 0; c < kHalfChunks; ++c)
          if (hc * kHalfChunks + c < kc1) mbar_arrive(bar_wait + 8 * (kWaitChunk0 + hc * kHalfChunks + c));
      }
    }
    stamp(1);

    // ---- layer 1, half by half: activate S in place, group by group, while the tensor pipe multiplies the groups already done
    const int ng1 = pl.ng1, ng2 = pl.ng2;
    for (int half = 0; half < 2; ++half) {
      const float* bias = Bg[0] + half * H;
      auto f0 = [&]() { stamp(2 + 3 * half); };
      auto f1 = [&]() { stamp(3 + 3 * half); };
      auto f2 = [&]() { stamp(4 + 3 * half); };
      if (H / ng1 == 64) drain_groups<16>(tmem_lane_base, part, 0, ng1, bias, bar_wait + 8 * kWaitAct, bar_s, (uint32_t)half, f0, f1, f2);
      else drain_groups<8>(tmem_lane_base, part, 0, ng1, bias, bar_wait + 8 * kWaitAct, bar_s, (uint32_t)half, f0, f1, f2);
    }
    // layer 1 is done with the observation tile: its buffer takes the transposed last layer [k][16]
    for (int i = tid; i < h2 * kMaxOut; i += kEpiThreads) {
      const int k = i >> 4, o = i & 15;
      s_w4t[i] = o < nout ? Wg[3][(size_t)o * h2 + k] : 0.f;
    }
    // ---- layer 2
    {
      auto f0 = [&]() { stamp(8); };
      auto f1 = [&]() { stamp(9); };
      auto f2 = [&]() { stamp(10); };
      if (h1 / ng2 == 64) drain_groups<16>(tmem_lane_base, part, H, ng2, Bg[1], bar_wait + 8 * kWaitAct2, bar_acc, 0u, f0, f1, f2);
      else drain_groups<8>(tmem_lane_base, part, H, ng2, Bg[1], bar_wait + 8 * kWaitAct2, bar_acc, 0u, f0, f1, f2);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // s_w4t complete
    float4 b3r[8];                                          // this warp's block of the layer-3 bias, ahead of the wait
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) b3r[q4] = part * 32 < h2 ? ldg_f4(Bg[2] + part * 32 + 4 * q4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float zq[4] = {0.f, 0.f, 0.f, 0.f};                     // this thread's action quad: the draws do not depend on the network
    if (net == 0 && 4 * part < nout && m0 + row < N) policy_draw_quad(p, m0 + row, part, zq);
    mbar_wait(bar_acc, 1);                                  // layer 3 done
    tc_fence_after();
    stamp(11);
    // ---- last layer (h2 -> nout) on FP32 FFMA: every part sums its column blocks, parts 1.. hand their partial sums
    //      to part 0 through the observation buffer, laid out [part-1][output][row]
    f2_t acc2[kMaxOut / 2];
#pragma unroll
    for (int o = 0; o < kMaxOut / 2; ++o) acc2[o] = pack2(0.f, 0.f);
    for (int cb = part; cb < h2 / 32; cb += kColSplit) {
      uint32_t v[32];
      tmem_ld32(tmem_lane_base + (uint32_t)(cb * 32), v);
      const uint32_t wa = a_w4t + cb * 32 * kMaxOut * 4;
      if (nout <= 4) last_layer_block<1>(v, b3r, wa, acc2);
      else if (nout <= 12) last_layer_block<3>(v, b3r, wa, acc2);
      else last_layer_block<4>(v, b3r, wa, acc2);
    }
    float acc[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut / 2; ++o) unpack2(acc2[o], acc[2 * o], acc[2 * o + 1]);
    tc_fence_before();
    // every part publishes its partial sums [part][output][row]; part b then owns the actions 4b..4b+3 of its rows (one
    // Philox block, two Box-Muller pairs), part 3 adds the partial log-probs in a fixed order
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o)
      if (o < nout) sts_f1(a_partial + ((part * kMaxOut + o) * kTileM + row) * 4, acc[o]);
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    stamp(12);
    const int n = m0 + row;
    const uint32_t a_logp = a_partial + kColSplit * kMaxOut * kTileM * 4;         // [4][row]
    if (net == 1) {
      if (part == 0 && n < N) {
        float t = s_b4[0];
#pragma unroll
        for (int pp = 0; pp < kColSplit; ++pp) t += lds_f1(a_partial + ((pp * kMaxOut) * kTileM + row) * 4);
        p.values[n] = t;
      }
    } else {
      if (4 * part < nout) {
        float mu[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = 4 * part + j;
          float t = s_b4[o];
#pragma unroll
          for (int pp = 0; pp < kColSplit; ++pp) t += lds_f1(a_partial + ((pp * kMaxOut + o) * kTileM + row) * 4);
          mu[j] = t;
        }
        const float lp = n < N ? policy_finish_quad(p, n, part, mu, s_std, zq) : 0.f;
        sts_f1(a_logp + (part * kTileM + row) * 4, lp);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (part == 3 && n < N) {
        float lp = 0.f;
#pragma unroll
        for (int b = 0; b < kColSplit; ++b)
          if (4 * b < nout) lp += lds_f1(a_logp + (b * kTileM + row) * 4);
        p.actions_log_prob[n] = lp;
      }
    }
    stamp(13);
  } else if (warp == kProducerWarp) {
    // ================= weight-tile producer =================
    if (lane == 0) {
      const uint8_t* src = a.packed + pl.net_off[net];
      for (int i = 0; i < ntiles; ++i) {
        const int s = i % kStages;
        const uint32_t bytes = (uint32_t)pl.sched[net][i].rows * 128u;
        mbar_wait(bar_empty + 8 * s, ((i / kStages) & 1) ^ 1);
        if (PROF && (a.dbg_flags & 1)) {
          mbar_arrive(bar_full + 8 * s);
        } else {
          mbar_expect_tx(bar_full + 8 * s, bytes);
          bulk_g2s(ring + s * kStageBytes, src, bytes, bar_full + 8 * s);
        }
        src += bytes;
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    // The whole warp walks the schedule with warp-uniform control flow and takes every MMA operand from uniform sources
    // (the tile table in the kernel-parameter constant bank, the TMEM base broadcast once): the descriptors then live in
    // uniform registers and a tcgen05.mma is a single UTCHMMA.  Operands loaded per thread (LDS) cost an R2UR + ELECT loop
    // of ~25 instructions around every MMA, which is what bounded the issue rate at ~125 cycles per MMA.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const TcTile* __restrict__ sched = pl.sched[net];
    const bool skip_mma = PROF && (a.dbg_flags & 2) != 0;
    const bool prof = PROF && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
    long long c_bar = 0, c_full = 0, c_issue = 0, c_commit = 0;
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      const TcTile T = sched[i];
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (PROF) t0 = clock64();
      if (T.wait_bar) mbar_wait(bar_wait + 8 * (T.wait_bar - 1), T.wait_par);
      if (PROF) t1 = clock64();
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (PROF) t2 = clock64();
      const uint32_t d = tmem_u + T.d_col;
      const uint32_t idesc = make_idesc(T.rows);
      const uint32_t b_lo = desc_lo(ring + stage * kStageBytes);
      const uint32_t acc0 = (T.flags & 2) ? 0u : 1u;
      const int ksteps = T.ksteps;
      if (elect_one()) {
        if (!skip_mma) {
          if (T.flags & 1) {
            const uint32_t a_t = tmem_u + T.a;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) mma_tf32_ts(d, a_t + ks * 8, desc_from_lo(b_lo + 2 * ks), idesc, ks == 0 ? acc0 : 1u);
          } else {
            const uint32_t a_lo = desc_lo(abuf + T.a * kChunkBytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) mma_tf32(d, desc_from_lo(a_lo + 2 * ks), desc_from_lo(b_lo + 2 * ks), idesc, ks == 0 ? acc0 : 1u);
          }
        }
        if (PROF) t3 = clock64();
        tc_commit(bar_empty + 8 * stage);                   // stage is free once these MMAs have read it
        if (T.commit_bar) tc_commit(T.commit_bar - 1 == kCommitS ? bar_s : bar_acc);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
      if (prof) {
        const long long t4 = clock64();
        c_bar += t1 - t0; c_full += t2 - t1; c_issue += t3 - t2; c_commit += t4 - t3;
        a.timeline[16 + i] = t2;                            // cycle stamp of the tile's start
      }
    }
    if (prof) { a.timeline[100] = c_bar; a.timeline[101] = c_full; a.timeline[102] = c_issue; a.timeline[103] = c_commit; }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
// Packed-weight cache: one record PER WORKSPACE (a process may alternate several ActorCritic instances -- the games run a
// low-level policy plus one or two high-level agents -- and each owns its workspace), keyed by the workspace pointer and
// validated against the caller's weights_version, the shapes and the weight pointers.
struct PackCache { long long version; int shape[8]; const void* w[8]; };
static std::mutex g_pack_mu;
static std::unordered_map<const void*, PackCache> g_pack;
static long long* g_timeline = nullptr;
static int g_dbg_flags = 0;
void policy_tc_set_timeline(long long* dev, int flags) { g_timeline = dev; g_dbg_flags = flags; }

int policy_tc_launch(const LgkPolicyParams* p, const TcPlan& pl, cudaStream_t st) {
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)p->workspace + 1023) & ~(uintptr_t)1023);
  PackCache want;
  want.version = p->weights_version;
  const int shape[8] = {pl.o[0], pl.o[1], pl.h0, pl.h1, pl.h2, pl.nact, 0, 0};
  memcpy(want.shape, shape, sizeof(shape));
  for (int i = 0; i < 4; ++i) { want.w[i] = p->actor_w[i]; want.w[4 + i] = p->critic_w[i]; }
  bool cached = false;
  if (p->weights_version != 0) {
    std::lock_guard<std::mutex> lk(g_pack_mu);
    auto it = g_pack.find(p->workspace);
    cached = it != g_pack.end() && it->second.version == want.version &&
             memcmp(want.shape, it->second.shape, sizeof(shape)) == 0 && memcmp(want.w, it->second.w, sizeof(want.w)) == 0;
  }
  if (!cached) {
    PackJobs jobs;
    int max_tiles = 0;
    for (int net = 0; net < 2; ++net) {
      const float* const* W = net ? p->critic_w : p->actor_w;
      const int ks[3] = {pl.o[net], pl.h0, pl.h1};
      for (int l = 0; l < 3; ++l) { jobs.w[net * 3 + l] = W[l]; jobs.k[net * 3 + l] = ks[l]; }
      jobs.dst[net] = reinterpret_cast<float*>(base + pl.net_off[net]);
      jobs.ntiles[net] = pl.ntiles[net];
      uint32_t off = 0;
      for (int t = 0; t < pl.ntiles[net]; ++t) {
        jobs.sched[net][t] = pl.sched[net][t];
        jobs.off[net][t] = off;
        off += (uint32_t)pl.sched[net][t].rows * kChunkK;
      }
      if (pl.ntiles[net] > max_tiles) max_tiles = pl.ntiles[net];
    }
    policy_pack_kernel<<<dim3(max_tiles, 2), 256, 0, st>>>(jobs);
    count_launch();
    if (int rc = check_cuda(cudaGetLastError(), "policy_pack_kernel launch")) return rc;
    if (p->weights_version != 0) {
      std::lock_guard<std::mutex> lk(g_pack_mu);
      if (g_pack.size() > 256) g_pack.clear();        // workspaces come and go with their modules: bound the table
      g_pack[p->workspace] = want;
    }
  }
  const bool prof = g_timeline != nullptr;
  auto kern = prof ? policy_tc_kernel<true> : policy_tc_kernel<false>;
  if (int rc = ensure_func_attr(reinterpret_cast<const void*>(kern), kSmemBytes, "policy_tc_kernel")) return rc;
  TcArgs args;
  args.p = *p; args.pl = pl; args.packed = base; args.timeline = g_timeline; args.dbg_flags = g_dbg_flags;
  args.net0 = p->nets == 2 ? 1 : 0;
  const int nets = (p->nets == 1 || p->nets == 2) ? 1 : 2;
  cudaError_t le = launch_chained(kern, dim3((p->num_envs + kTileM - 1) / kTileM, nets), dim3(kTcThreads), (size_t)kSmemBytes, st, args);
  if (int rc = check_cuda(le, "policy_tc_kernel launch")) return rc;
  count_launch();
  return check_cuda(cudaGetLastError(), "policy_tc_kernel launch");
}

}  // namespace lgk
