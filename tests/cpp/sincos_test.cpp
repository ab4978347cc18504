// Host check of the double-precision sincos used by k_brief (orb_kernels.cuh: sincos_f64_to_f32): rounded to float
// it must equal (float)cos((double)x), (float)sin((double)x) of the host libm — what OpenCV's BRIEF steering computes —
// for float angles in [0, 2*pi].  IEEE fma/mul/add only, so the device evaluates the identical bits.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
static void sc(float xf, float& cf, float& sf) {
  const double x = (double)xf;
  const double q = rint(x * 6.36619772367581382433e-01);
  double t = fma(-q, 1.57079632679489655800e+00, x);
  t = fma(-q, 6.12323399573676603587e-17, t);
  const double z = t * t;
  double ps = 1.58962301576546568060e-10;
  ps = fma(ps, z, -2.50507477628578072866e-8);
  ps = fma(ps, z, 2.75573136213857245213e-6);
  ps = fma(ps, z, -1.98412698295895385996e-4);
  ps = fma(ps, z, 8.33333333332211858878e-3);
  ps = fma(ps, z, -1.66666666666666307295e-1);
  const double sn = fma(t * z, ps, t);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-9);
  pc = fma(pc, z, -2.75573143513906633035e-7);
  pc = fma(pc, z, 2.48015872894767294178e-5);
  pc = fma(pc, z, -1.38888888888741095749e-3);
  pc = fma(pc, z, 4.16666666666666019037e-2);
  const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
  const int iq = (int)q;
  const double s1 = (iq & 1) ? cs : sn, c1 = (iq & 1) ? sn : cs;
  sf = (float)((iq & 2) ? -s1 : s1);
  cf = (float)(((iq + 1) & 2) ? -c1 : c1);
}
int main(int argc, char** argv) {
  const uint32_t step = argc > 1 ? (uint32_t)atoi(argv[1]) : 16;
  long bad = 0, n = 0;
  for (uint32_t u = 0; u <= 0x40C90FDBu + 64; u += step) {  // 0x40C90FDB = 2*pi as float
    float x, c, s;
    memcpy(&x, &u, 4);
    sc(x, c, s);
    if (c != (float)cos((double)x) || s != (float)sin((double)x)) {
      if (bad < 5) printf("MISMATCH x=%.9g\n", x);
      bad++;
    }
    n++;
  }
  printf("checked %ld floats, mismatches %ld\n", n, bad);
  return bad != 0;
}
