// lgk_policy_tc_plan.h -- host-side tiling plan + MMA schedule of the tcgen05 policy kernel (lgk_policy_tc.cu)
#pragma once
#include <stdint.h>
namespace lgk {

constexpr int kTcMaxTiles = 48;      // 2 x 8 (layer 1) + 2 x 4 x 2 (layer 2) + 4 x 2 (layer 3) = 40 for the 512-256-128 nets

// One weight tile (<= 256 output rows x 32 k, the unit the producer copies and the issuer multiplies) in the order the
// tensor pipe consumes it.  The same table drives the packing kernel, the TMA producer and the MMA issuer.
struct TcTile {
  uint16_t a;          // A operand: shared-memory chunk index of the observation tile (SS) / TMEM column of the first k (TS)
  uint16_t d_col;      // TMEM column of the accumulator block
  uint16_t rows;       // N of the MMAs = output rows in the tile
  uint8_t ksteps;      // MMAs of K = 8 that hold data (1..4)
  uint8_t flags;       // bit 0: A from TMEM; bit 1: the first MMA overwrites the accumulator
  uint8_t wait_bar;    // 0 none, else 1 + index of the barrier the issuer waits on before the tile
  uint8_t wait_par;    // its parity
  uint8_t commit_bar;  // 0 none, else 1 + index of the barrier committed after the tile's MMAs
  uint8_t layer;       // 0..2: which weight matrix (packing)
  uint16_t n0, k0;     // first output row / first input column of the tile (packing)
};
static_assert(sizeof(TcTile) == 16, "TcTile is 16 bytes");

struct TcPlan {
  int o[2];          // input widths: actor obs, critic obs
  int kc1[2];        // k-chunks of layer 1 per net
  int h0, h1, h2, q, nact;       // q = h0 / 2: layer 1 is computed in two column halves of q (<= 256) columns
  int ng1, ng2;      // drain groups per layer-1 half / of the layer-2 block (each 4 warps x 8 or 16 columns)
  int ntiles[2];
  long long net_bytes[2];        // packed bytes per net
  long long net_off[2];          // offset of each net's packed image in the workspace
  TcTile sched[2][kTcMaxTiles];
};

}  // namespace lgk
