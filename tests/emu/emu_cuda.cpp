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
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.get();
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &bs.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  bool any = true;
  while (any) {
    any = false;
    unsigned ndone = 0;
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = bs.fibers[t];
      if (f.done) {
        ++ndone;
        continue;
      }
      threadIdx.x = t % block.x;
      threadIdx.y = (t / block.x) % block.y;
      threadIdx.z = t / (block.x * block.y);
      bs.current = &f;
      swapcontext(&bs.sched, &f.ctx);
      if (!f.done) any = true; else ++ndone;
    }
    if (any && ndone != 0) {
      std::fprintf(stderr, "emu: divergent __syncthreads (some threads exited, others wait)\n");
      std::abort();
    }
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
  swapcontext(&bs->current->ctx, &bs->sched);
}
