// Fiber scheduler behind approxfun.jl_b200/csrc/afb_emu.h (DEVELOPER TEST TOOL, see that header).
#include "afb_emu.h"

namespace afb_emu {
thread_local Block* g_blk = nullptr;
thread_local uint3 g_tid, g_bid;
thread_local dim3 g_bdim, g_gdim;

static void fiber_entry() {
    Block* b = g_blk;
    b->body();
    b->done[b->cur] = 1;
    swapcontext(&b->ctx[b->cur], &b->sched);
}

void yield_barrier() {
    Block* b = g_blk;
    int me = b->cur;
    swapcontext(&b->ctx[me], &b->sched);
    // resumed: restore identity
    g_tid.x = (unsigned)me;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    assert(block.y == 1 && block.z == 1 && grid.z == 1);
    const int T = (int)block.x;
    const size_t STK = 256 * 1024;
    Block blk;
    blk.ctx.resize(T);
    blk.stacks.assign(T, std::vector<char>(STK));
    blk.done.assign(T, 0);
    blk.smem.assign(smem + 64, 0);
    blk.xchg.assign(T, 0);
    blk.body = body;
    g_blk = &blk;
    g_bdim = block;
    g_gdim = grid;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            g_bid.x = bx; g_bid.y = by; g_bid.z = 0;
            std::fill(blk.done.begin(), blk.done.end(), 0);
            for (int t = 0; t < T; ++t) {
                getcontext(&blk.ctx[t]);
                blk.ctx[t].uc_stack.ss_sp = blk.stacks[t].data();
                blk.ctx[t].uc_stack.ss_size = STK;
                blk.ctx[t].uc_link = &blk.sched;
                makecontext(&blk.ctx[t], (void (*)())fiber_entry, 0);
            }
            // round-robin: each sweep runs every live fiber up to its next barrier (or its end).
            // A barrier is complete when every live fiber has yielded once, which one sweep does.
            // AFB_EMU_ORDER=reverse|shuffle runs the fibers of each sweep in another order: a missing barrier then
            // shows up as a wrong answer whichever direction the dependency points (poor man's racecheck).
            static const char* order_env = std::getenv("AFB_EMU_ORDER");
            const int order = !order_env ? 0 : (order_env[0] == 'r' ? 1 : 2);
            unsigned lcg = 12345u + bx * 7919u;
            bool live = true;
            while (live) {
                live = false;
                lcg = lcg * 1664525u + 1013904223u;
                const int rot = (int)((lcg >> 16) % (unsigned)T);
                for (int tt = 0; tt < T; ++tt) {
                    int t = tt;
                    if (order == 1) t = T - 1 - tt;
                    else if (order == 2) t = (int)(((long)tt * 37 + rot) % T);  // a permutation: T is a power of two
                    if (blk.done[t]) continue;
                    blk.cur = t;
                    g_tid.x = (unsigned)t; g_tid.y = 0; g_tid.z = 0;
                    swapcontext(&blk.sched, &blk.ctx[t]);
                    if (!blk.done[t]) live = true;
                }
            }
        }
    g_blk = nullptr;
}
}  // namespace afb_emu
