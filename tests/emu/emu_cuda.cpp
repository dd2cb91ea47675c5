// Fiber-based block scheduler for the CUDA functional simulator (see emu_cuda.h).
#include "emu_cuda.h"

#include <omp.h>

#include <memory>

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace emu {

namespace {
constexpr size_t kStack = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  std::unique_ptr<char[]> stack;
  bool done = false;
  int wait = 0;  // 0: runnable, 1: at __syncthreads, 2: at __syncwarp
};

struct BlockState {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  Fiber* current = nullptr;
  const std::function<void()>* body = nullptr;
  std::vector<unsigned char> smem;
};

thread_local BlockState* g_bs = nullptr;

void trampoline() {
  BlockState* bs = g_bs;
  Fiber* self = bs->current;
  (*bs->body)();
  self->done = true;
  swapcontext(&self->ctx, &bs->sched);
}
}  // namespace

void* dyn_smem() { return g_bs->smem.data(); }

static void run_block(BlockState& bs, dim3 block) {
  const unsigned nt = block.x * block.y * block.z;
  for (unsigned t = 0; t < nt; ++t) {
    Fiber& f = bs.fibers[t];
    f.done = false;
    f.wait = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.get();
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &bs.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  for (;;) {
    // run every runnable fiber until it waits at a barrier or exits
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = bs.fibers[t];
      if (f.done || f.wait != 0) continue;
      threadIdx.x = t % block.x;
      threadIdx.y = (t / block.x) % block.y;
      threadIdx.z = t / (block.x * block.y);
      bs.current = &f;
      swapcontext(&bs.sched, &f.ctx);
    }
    // warp barriers: a warp whose live fibers all wait at __syncwarp goes on
    bool released = false;
    for (unsigned w0 = 0; w0 < nt; w0 += 32) {
      const unsigned w1 = w0 + 32 < nt ? w0 + 32 : nt;
      unsigned live = 0, atwarp = 0;
      for (unsigned t = w0; t < w1; ++t) {
        if (bs.fibers[t].done) continue;
        ++live;
        if (bs.fibers[t].wait == 2) ++atwarp;
      }
      if (live != 0 && atwarp == live) {
        for (unsigned t = w0; t < w1; ++t) bs.fibers[t].wait = 0;
        released = true;
      } else if (atwarp != 0) {
        for (unsigned t = w0; t < w1; ++t)
          if (!bs.fibers[t].done && bs.fibers[t].wait == 0) {
            std::fprintf(stderr, "emu: inconsistent fiber state at __syncwarp\n");
            std::abort();
          }
      }
    }
    if (released) continue;
    // block barrier: every fiber of the block waits at __syncthreads
    unsigned ndone = 0, atblock = 0, atwarp = 0;
    for (unsigned t = 0; t < nt; ++t) {
      if (bs.fibers[t].done) ++ndone;
      else if (bs.fibers[t].wait == 1) ++atblock;
      else ++atwarp;
    }
    if (ndone == nt) break;
    if (atwarp != 0) {
      std::fprintf(stderr, "emu: deadlock (part of a warp waits at __syncwarp, the rest at __syncthreads or gone)\n");
      std::abort();
    }
    if (ndone != 0) {
      std::fprintf(stderr, "emu: divergent __syncthreads (some threads exited, others wait)\n");
      std::abort();
    }
    for (unsigned t = 0; t < nt; ++t) bs.fibers[t].wait = 0;
  }
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const long nblocks = (long)grid.x * grid.y * grid.z;
  const unsigned nt = block.x * block.y * block.z;
#pragma omp parallel
  {
    BlockState bs;
    bs.body = &body;
    bs.smem.assign(smem_bytes + 64, 0xAB);  // poison: kernels must initialise what they read
    bs.fibers.resize(nt);
    for (auto& f : bs.fibers) f.stack.reset(new char[kStack]);
    g_bs = &bs;
    blockDim = block;
    gridDim = grid;
#pragma omp for schedule(dynamic, 1)
    for (long b = 0; b < nblocks; ++b) {
      blockIdx.x = (unsigned)(b % grid.x);
      blockIdx.y = (unsigned)((b / grid.x) % grid.y);
      blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
      std::memset(bs.smem.data(), 0xAB, bs.smem.size());
      run_block(bs, block);
    }
    g_bs = nullptr;
  }
}

}  // namespace emu

void __syncthreads() {
  emu::BlockState* bs = emu::g_bs;
  bs->current->wait = 1;
  swapcontext(&bs->current->ctx, &bs->sched);
}

void __syncwarp() {
  emu::BlockState* bs = emu::g_bs;
  bs->current->wait = 2;
  swapcontext(&bs->current->ctx, &bs->sched);
}

// butterfly sum over the 32 fibers of a warp with the device's pairing order (x += shfl_xor(x, m), m = 16 .. 1)
namespace cpb {
void emu_warp_sum4(double (&q)[4]) {
  static thread_local double scratch[64][32][4];
  const unsigned t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned w = t / 32, lane = t % 32;
  if (w >= 64) {
    std::fprintf(stderr, "emu: warp_sum4 supports 64 warps per block\n");
    std::abort();
  }
  for (int m = 16; m > 0; m >>= 1) {
    for (int i = 0; i < 4; ++i) scratch[w][lane][i] = q[i];
    __syncwarp();
    double o[4];
    for (int i = 0; i < 4; ++i) o[i] = scratch[w][lane ^ m][i];
    __syncwarp();
    for (int i = 0; i < 4; ++i) q[i] += o[i];
  }
}
}  // namespace cpb
