// Table of per-length kernel launchers (see axis.h, sizes.def).
#include "axis.h"

#ifndef CPB_SIZES_DEF
#define CPB_SIZES_DEF "sizes.def"
#endif

namespace cpb {

#define CPB_SIZE(N, R1, R2) const AxisKernels* axis_kernels_n##N();
#include CPB_SIZES_DEF
#undef CPB_SIZE

namespace {
typedef const AxisKernels* (*Getter)();
const Getter kGetters[] = {
#define CPB_SIZE(N, R1, R2) axis_kernels_n##N,
#include CPB_SIZES_DEF
#undef CPB_SIZE
};
constexpr int kNum = sizeof(kGetters) / sizeof(kGetters[0]);
}  // namespace

int num_axis_kernels() { return kNum; }
const AxisKernels* axis_kernels_at(int i) { return (i >= 0 && i < kNum) ? kGetters[i]() : nullptr; }
const AxisKernels* find_axis_kernels(int n) {
  for (int i = 0; i < kNum; ++i) {
    const AxisKernels* k = kGetters[i]();
    if (k->n == n) return k;
  }
  return nullptr;
}

}  // namespace cpb
