// Exhaustive check behind vt_div6 (voltools_b200/csrc/vt_common.cuh); gcc -O2 -ffp-contract=off check_div6.c -lm.
// exhaustive check: for every float x in [0, 8], is fma(fma(-6,q,x), r6, q) with q = x*r6 equal to x/6 (RN)?
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline float asf(uint32_t u){float f; memcpy(&f,&u,4); return f;}
int main(){
  const float r6 = 1.0f/6.0f;
  uint32_t hi; float top = 8.0f; memcpy(&hi,&top,4);
  unsigned long long bad=0, n=0; float maxbad = 0;
  for (uint32_t u=0; u<=hi; u++){
    float x = asf(u);
    float q = x*r6;
    float r = fmaf(-6.0f, q, x);
    float q2 = fmaf(r, r6, q);
    float d = x/6.0f;
    n++;
    if (q2 != d) { bad++; if (x > maxbad) maxbad = x; }
  }
  printf("checked %llu, mismatches %llu, largest mismatching x = %a\n", n, bad, maxbad);
  return 0;
}
