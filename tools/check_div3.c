/* Backs the claim in csrc/aggregate_nms_ms.cuh: for float32 x,
 *     q0 = x * r;  q = fma(fma(-3, q0, x), r, q0)      with r = RN(1/3) = 0x3eaaaaab
 * equals the IEEE quotient x / 3 for every finite x except -0 (and the kernel sends -0 and |x| < 1e-30 through
 * the real division anyway).
 *   check_div3           all 2^24 mantissas x both signs in 4 binades: denormals, smallest normals, [1,2), largest (seconds)
 *   check_div3 --full    every one of the 2^32 bit patterns (about 80 s)
 * Exit code 0 = the claim holds.  Build: gcc -O2 -ffp-contract=off -o check_div3 check_div3.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static int check(uint32_t bits, unsigned long long* tested, unsigned long long* bad_guarded) {
  float x;
  memcpy(&x, &bits, 4);
  if (!isfinite(x)) return 0;
  const float r = 1.0f / 3.0f;
  const float ref = x / 3.0f;
  const float q0 = x * r;
  const float q = fmaf(fmaf(-3.0f, q0, x), r, q0);
  uint32_t a, b;
  memcpy(&a, &ref, 4);
  memcpy(&b, &q, 4);
  ++*tested;
  if (a == b) return 0;
  /* the kernel's guard: tiny non-zero magnitudes and -0 take the real division */
  if (fabsf(x) < 1e-30f && bits != 0u) return 0;
  ++*bad_guarded;
  return 1;
}

int main(int argc, char** argv) {
  unsigned long long tested = 0, bad = 0;
  if (argc > 1 && strcmp(argv[1], "--full") == 0) {
    for (unsigned long long u = 0; u < (1ull << 32); ++u) check((uint32_t)u, &tested, &bad);
  } else {
    const uint32_t exps[4] = {0, 1, 127, 254};   /* rounding depends on the mantissa only, away from the range ends */
    for (int e = 0; e < 4; ++e)
      for (uint32_t m = 0; m < (1u << 23); ++m)
        for (uint32_t s = 0; s < 2; ++s) check((s << 31) | (exps[e] << 23) | m, &tested, &bad);
  }
  printf("tested %llu finite values, %llu mismatches outside the guarded range\n", tested, bad);
  return bad != 0;
}
