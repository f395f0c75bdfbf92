// TEST INFRASTRUCTURE ONLY -- fiber scheduler behind cuda_emul.h (see that header).
#include "cuda_emul.h"

uint3_ threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char *jdsp_emul_dyn_smem = nullptr;

namespace jdsp_emul {
State S;
int g_order = +1;
static const std::function<void()> *g_body = nullptr;
static const size_t kStack = 256 * 1024;

static void set_tid(int t) {
    S.cur = t;
    threadIdx.x = (unsigned)t % blockDim.x;
    threadIdx.y = ((unsigned)t / blockDim.x) % blockDim.y;
    threadIdx.z = (unsigned)t / (blockDim.x * blockDim.y);
}

void yield() {
    const int me = S.cur;
    swapcontext(&S.ctx[me], &S.sched);
    set_tid(me);
}

static void trampoline() {
    (*g_body)();
    const int me = S.cur;
    S.done[me] = 1;
    S.alive--;
    S.warp_alive[me >> 5]--;
    // a thread that exits while others wait at a barrier must not deadlock them
    if (S.alive > 0 && S.cta_arrived == S.alive) { S.cta_arrived = 0; S.cta_gen++; }
    const int w = me >> 5;
    if (S.warp_alive[w] > 0 && S.warp_arrived[w] == S.warp_alive[w]) { S.warp_arrived[w] = 0; S.warp_gen[w]++; }
    swapcontext(&S.ctx[me], &S.sched);
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body) {
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<char> stacks((size_t)nt * kStack);
    std::vector<unsigned char> smem(dyn_smem + 64);
    g_body = &body;
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                memset(smem.data(), 0xCD, smem.size());  // poison: uninitialised shared memory is garbage
                jdsp_emul_dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
                S.nthreads = nt; S.alive = nt; S.cta_arrived = 0; S.cta_gen = 0;
                S.order = g_order;
                S.ctx.assign(nt, ucontext_t());
                S.done.assign(nt, 0);
                for (int w = 0; w < 64; ++w) { S.warp_arrived[w] = 0; S.warp_gen[w] = 0; S.warp_alive[w] = 0; }
                for (int t = 0; t < nt; ++t) S.warp_alive[t >> 5]++;
                for (int t = 0; t < nt; ++t) {
                    getcontext(&S.ctx[t]);
                    S.ctx[t].uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
                    S.ctx[t].uc_stack.ss_size = kStack;
                    S.ctx[t].uc_link = &S.sched;
                    makecontext(&S.ctx[t], trampoline, 0);
                }
                while (S.alive > 0) {
                    for (int i = 0; i < nt; ++i) {
                        const int t = S.order > 0 ? i : nt - 1 - i;
                        if (S.done[t]) continue;
                        set_tid(t);
                        swapcontext(&S.sched, &S.ctx[t]);
                    }
                }
            }
    jdsp_emul_dyn_smem = nullptr;
}
}  // namespace jdsp_emul
