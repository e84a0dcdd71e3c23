// lgk_policy_tc_plan.h -- host-side tiling plan of the tcgen05 policy kernel (lgk_policy_tc.cu)
#pragma once
namespace lgk {
struct TcPlan {
  int o[2];          // input widths: actor obs, critic obs
  int kc1[2];        // k-chunks of layer 1 per net
  int h0, h1, h2, half, nact;
  int nb1, nb2, nb3;             // rows per weight tile in L1 (per half), L2, L3
  int t1[2], t2, t3;             // tiles per phase (L1 per net; both halves)
  long long net_bytes[2];        // packed bytes per net
  long long net_off[2];          // offset of each net's packed image in the workspace
};

}  // namespace lgk
