// TEST INFRASTRUCTURE — runtime half of the CUDA execution-model emulator (see emu_cuda.h).
#include "emu_cuda.h"
#include <mutex>

namespace emu {
thread_local Block* g_blk = nullptr;
thread_local uint3_ g_blockIdx;
thread_local dim3 g_blockDim, g_gridDim;

static const int kWorkers = 8;
static Block* g_blocks[kWorkers];
static std::mutex g_launch_mu;

char* dyn_smem() { return g_blk->dyn; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  std::lock_guard<std::mutex> lk(g_launch_mu);
  const long total = (long)grid.x * grid.y * grid.z;
  if (total <= 0) return;
  std::atomic<long> next{0};
  int nw = (int)std::min<long>(kWorkers, total);
  const char* env = getenv("VAESNE_EMU_WORKERS");
  if (env) nw = std::max(1, std::min(nw, atoi(env)));
  auto work = [&](int wid) {
    if (!g_blocks[wid]) g_blocks[wid] = new Block();
    Block* b = g_blocks[wid];
    std::vector<char> dyn(smem + 64);
    b->dyn = (char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_blockDim = block; g_gridDim = grid;
    for (;;) {
      long i = next.fetch_add(1);
      if (i >= total) break;
      g_blockIdx = {(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((long)grid.x * grid.y))};
      run_block(b, body, block);
    }
  };
  if (nw == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int w = 0; w < nw; ++w) th.emplace_back(work, w);
  for (auto& t : th) t.join();
}
}  // namespace emu
