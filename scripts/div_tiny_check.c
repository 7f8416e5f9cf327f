// host model of the tiny-numerator division: must equal x / h bit for bit
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static inline double markstein(double x, double h, double rh) {
  double q = x * rh;
  double r = fma(-q, h, x);
  q = fma(r, rh, q);
  r = fma(-q, h, x);
  return fma(r, rh, q);
}
static inline double fix_tiny(double x, double h, double q /* = RN((x * 2^400) / h) */) {
  const double xs = x * 0x1p+400;
  const double aq = fabs(q);
  if (aq >= 0x1p-622) return q * 0x1p-400;         // the scaled-back quotient is a normal number: exact
  const double r = fma(-q, h, xs);                 // exact residual: sign tells on which side the true quotient lies
  if (r != 0.0) {                                  // (an exact q is rounded correctly by the final multiplication)
    int64_t b; memcpy(&b, &q, 8);
    // >= 2 bits lost: round to odd.  Exactly 1 bit lost: q with an odd last bit sits on a rounding boundary of the
    // coarser grid -- step off it towards the true quotient (lands on the grid: the scaling is then exact)
    const int one_bit = aq >= 0x1p-623;
    if (((b & 1) == 0) != one_bit) {
      const int up = (r > 0.0) == (h > 0.0);       // true quotient > q ?
      b += ((q > 0.0) == up) ? 1 : -1;
      memcpy(&q, &b, 8);
    }
  }
  return q * 0x1p-400;
}
static inline double div_exact2(double x, double h, double rh) {
  if (x != 0.0 && fabs(x) < 1e-280) return fix_tiny(x, h, markstein(x * 0x1p+400, h, rh));
  return markstein(x, h, rh);
}
static uint64_t s = 88172645463325252ULL;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main() {
  long bad = 0, n = 0, band = 0;
  double hs[] = {24.0, 240.0, 24 * 26.11, 2800.0, 2800.0 * 1.0371, 3131.7, 1.0 / 3.0, 7e5, 1e-2, -2800.0, 3e9};
  for (int ih = 0; ih < 11; ih++) {
    double h = hs[ih], rh = 1.0 / h;
    for (long k = 0; k < 40000000; k++) {
      uint64_t m = rnd();
      // exponent uniformly in the tiny range: biased exponent 0 .. 95 (5e-324 .. ~1e-280), random mantissa and sign
      uint64_t e = rnd() % 96;
      uint64_t bits = (m & 0x800FFFFFFFFFFFFFULL) | (e << 52);
      double x; memcpy(&x, &bits, 8);
      if (x == 0.0) continue;
      double a = div_exact2(x, h, rh), b = x / h;
      n++;
      if (memcmp(&a, &b, 8) != 0) { if (bad < 5) printf("MISMATCH x=%a h=%a got %a want %a\n", x, h, a, b); bad++; }
    }
    // also normal-range spot check of the fast path
    for (long k = 0; k < 2000000; k++) {
      uint64_t bits = rnd(); uint64_t e = 200 + rnd() % 1600; bits = (bits & 0x800FFFFFFFFFFFFFULL) | (e << 52);
      double x; memcpy(&x, &bits, 8);
      double a = div_exact2(x, h, rh), b = x / h; n++;
      if (memcmp(&a, &b, 8) != 0 && isfinite(b) && fabs(b) > 1e-290 && fabs(x) < 1e290) { if (bad < 10) printf("MISMATCH(normal) x=%a h=%a got %a want %a\n", x, h, a, b); bad++; }
    }
  }
  printf("%ld cases, %ld mismatches\n", n, bad);
  return bad != 0;
}
