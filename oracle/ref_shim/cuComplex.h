/* CPU stand-in for <cuComplex.h> (TEST INFRASTRUCTURE, authored here): the handful of complex helpers
 * the reference's kernel file uses. */
#ifndef CPB_REF_SHIM_CUCOMPLEX_H
#define CPB_REF_SHIM_CUCOMPLEX_H
struct cuDoubleComplex {
  double x, y;
};
struct cuComplex {
  float x, y;
};
static inline cuDoubleComplex make_cuDoubleComplex(double r, double i) {
  cuDoubleComplex c;
  c.x = r;
  c.y = i;
  return c;
}
static inline cuDoubleComplex cuCadd(cuDoubleComplex a, cuDoubleComplex b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
static inline cuDoubleComplex cuCmul(cuDoubleComplex a, cuDoubleComplex b) {
  return make_cuDoubleComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
#endif
