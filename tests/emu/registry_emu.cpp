// Registry for the simulator build: only the small lengths listed in tests/emu/Makefile.
#include "axis.h"

namespace cpb {
#define CPB_EMU_SIZES(X) X(16) X(20) X(24) X(30) X(32) X(36) X(40) X(48) X(60) X(64) X(72) X(128) X(192)
#define X(N) const AxisKernels* axis_kernels_n##N();
CPB_EMU_SIZES(X)
#undef X
namespace {
typedef const AxisKernels* (*Getter)();
const Getter kGetters[] = {
#define X(N) axis_kernels_n##N,
    CPB_EMU_SIZES(X)
#undef X
};
constexpr int kNum = sizeof(kGetters) / sizeof(kGetters[0]);
}  // namespace
int num_axis_kernels() { return kNum; }
const AxisKernels* axis_kernels_at(int i) { return (i >= 0 && i < kNum) ? kGetters[i]() : nullptr; }
const AxisKernels* find_axis_kernels(int n) {
  for (int i = 0; i < kNum; ++i)
    if (kGetters[i]()->n == n) return kGetters[i]();
  return nullptr;
}
}  // namespace cpb
