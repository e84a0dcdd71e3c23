// TEST INFRASTRUCTURE ONLY -- exhaustive proof that the FMA-based division by a constant used in csrc/lgk_math.cuh
// (height_index2) is bit-identical to IEEE x / c for every fp32 x in [1e-20, 1e20).  gcc -O2 -ffp-contract=off -mfma divcheck.c -lm
#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline float u2f(uint32_t u){ float f; memcpy(&f,&u,4); return f; }
static inline uint32_t f2u(float f){ uint32_t u; memcpy(&u,&f,4); return u; }
int main(){
  const float cs[3] = {0.1f, 0.05f, 0.25f};
  for (int ci=0; ci<3; ++ci){
    const float c = cs[ci];
    const float r = (float)(1.0/(double)c);   // correctly rounded reciprocal of the float c
    uint64_t bad=0, n=0;
    for (uint32_t u=f2u(1e-20f); u<f2u(1e20f); ++u){
      float x=u2f(u);
      float q0 = x*r;
      float rem = fmaf(-q0, c, x);
      float q1 = fmaf(rem, r, q0);
      volatile float qe = x / c;
      if (q1 != qe) { if (bad<5) printf("c=%g x=%a q1=%a qe=%a\n", c, x, q1, qe); ++bad; }
      ++n;
    }
    printf("c=%g r=%a: %llu values, %llu mismatches\n", c, r, (unsigned long long)n, (unsigned long long)bad);
  }
  return 0;
}
