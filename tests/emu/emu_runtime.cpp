// TEST INFRASTRUCTURE (see include/cuda_runtime.h): cooperative-fiber execution of one CUDA grid on the calling thread.
#include <cuda_runtime.h>
#include <ucontext.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <unordered_map>
#include <vector>

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace emu {

uint3 g_threadIdx{0, 0, 0}, g_blockIdx{0, 0, 0};
dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;

struct Barrier {
    unsigned arrived = 0, gen = 0;
};

struct Warp {
    unsigned present = 0;   // lanes that exist in this warp
    unsigned exited = 0;
    unsigned long long slots[32];
    std::unordered_map<unsigned, Barrier> bars;   // one barrier per participation mask
};

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    bool wait_cta = false;
    unsigned wait_cta_gen = 0;
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned n = 0, n_exited = 0;
    unsigned cta_arrived = 0, cta_gen = 0;
    unsigned cur = 0;
    unsigned long long progress = 0;
    ucontext_t sched;
    std::vector<unsigned char> smem;
    const std::function<void()>* body = nullptr;
};

Cta C;
unsigned long long g_launches = 0;

// BLOBS_EMU_SEED=<n>: CTAs of a launch, warps of a CTA and lanes of a warp are visited in a seeded random order instead of
// ascending order. Results must not depend on it (atomics hand out different ranks, cells hold their records in a different
// order): a cheap check for accidental order dependence.
struct Shuffle {
    bool on = false;
    std::mt19937 rng;
    Shuffle() {
        if (const char* e = getenv("BLOBS_EMU_SEED")) { on = true; rng.seed((unsigned)atoi(e)); }
    }
    void order(std::vector<unsigned>& v, unsigned n) {
        v.resize(n);
        std::iota(v.begin(), v.end(), 0u);
        if (on) std::shuffle(v.begin(), v.end(), rng);
    }
};
Shuffle g_shuffle;

void yield() { swapcontext(&C.fibers[C.cur].ctx, &C.sched); }

void trampoline() {
    (*C.body)();
    Fiber& f = C.fibers[C.cur];
    Warp& w = C.warps[C.cur >> 5];
    f.done = true;
    w.exited |= 1u << (C.cur & 31u);
    w.slots[C.cur & 31u] = 0;
    C.n_exited++;
    C.progress++;
    // returning resumes uc_link == the scheduler
}

[[noreturn]] void deadlock(const char* what) {
    fprintf(stderr, "[emu] DEADLOCK in block %u: %s\n", g_blockIdx.x, what);
    for (unsigned t = 0; t < C.n; ++t) {
        const Fiber& f = C.fibers[t];
        if (!f.done) fprintf(stderr, "[emu]   thread %u: %s\n", t, f.wait_cta ? "at __syncthreads" : "at a warp collective");
    }
    abort();
}

void run_cta() {
    for (unsigned t = 0; t < C.n; ++t) {
        Fiber& f = C.fibers[t];
        f.done = false;
        f.wait_cta = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &C.sched;
        makecontext(&f.ctx, trampoline, 0);
    }
    const unsigned nw = (C.n + 31) / 32;
    for (unsigned w = 0; w < nw; ++w) {
        Warp& W = C.warps[w];
        const unsigned lanes = C.n - w * 32 >= 32 ? 32 : C.n - w * 32;
        W.present = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
        W.exited = 0;
        W.bars.clear();
        memset(W.slots, 0, sizeof(W.slots));
    }
    C.n_exited = 0;
    C.cta_arrived = 0;
    C.cta_gen = 0;
    std::vector<unsigned> worder, lorder;
    while (C.n_exited < C.n) {
        const unsigned long long pass_progress = C.progress;
        g_shuffle.order(worder, nw);
        for (unsigned w : worder) {
            for (;;) {   // rounds over the lanes of this warp until none of them can run
                const unsigned long long before = C.progress;
                bool any = false;
                g_shuffle.order(lorder, 32);
                for (unsigned l : lorder) {
                    if (w * 32 + l >= C.n) continue;
                    const unsigned t = w * 32 + l;
                    Fiber& f = C.fibers[t];
                    if (f.done) continue;
                    if (f.wait_cta && f.wait_cta_gen == C.cta_gen) continue;   // parked until the CTA barrier opens
                    any = true;
                    C.cur = t;
                    g_threadIdx = uint3{t, 0, 0};
                    swapcontext(&C.sched, &f.ctx);
                }
                if (!any) break;
                if (C.progress == before) deadlock("a warp made no progress in a full round (divergent collective?)");
            }
        }
        // threads that exited after others arrived can complete a CTA barrier
        if (C.cta_arrived && C.cta_arrived >= C.n - C.n_exited) {
            C.cta_arrived = 0;
            C.cta_gen++;
            C.progress++;
        }
        if (C.progress == pass_progress && C.n_exited < C.n) deadlock("no thread can run");
    }
}

}  // namespace

unsigned lane_id() { return C.cur & 31u; }
unsigned long long* warp_slots() { return C.warps[C.cur >> 5].slots; }
unsigned participants(unsigned mask) {
    const Warp& w = C.warps[C.cur >> 5];
    return mask & w.present & ~w.exited;
}
void* dynamic_smem() { return C.smem.data(); }
unsigned long long launches() { return g_launches; }

void warp_barrier(unsigned mask) {
    const unsigned me = C.cur;   // C.cur changes while this fiber is switched out
    Warp& w = C.warps[me >> 5];
    const unsigned bit = 1u << (me & 31u);
    mask |= bit;
    Barrier& b = w.bars[mask];   // references into unordered_map stay valid across inserts
    b.arrived |= bit;
    C.progress++;
    const unsigned my_gen = b.gen;
    for (;;) {
        if (b.gen != my_gen) return;
        const unsigned need = mask & w.present & ~w.exited;
        if ((b.arrived & need) == need) {
            b.arrived = 0;
            b.gen++;
            C.progress++;
            return;
        }
        yield();
    }
}

void cta_barrier() {
    const unsigned me = C.cur;
    Fiber& f = C.fibers[me];
    C.cta_arrived++;
    C.progress++;
    const unsigned my_gen = C.cta_gen;
    while (C.cta_gen == my_gen) {
        if (C.cta_arrived >= C.n - C.n_exited) {
            C.cta_arrived = 0;
            C.cta_gen++;
            C.progress++;
            break;
        }
        f.wait_cta = true;
        f.wait_cta_gen = my_gen;
        yield();
    }
    f.wait_cta = false;
}

void run_grid(unsigned grid, unsigned block, size_t dyn_smem, const std::function<void()>& body) {
    if (grid == 0 || block == 0) return;
    g_launches++;
    if (C.fibers.size() < block) {
        const size_t old = C.fibers.size();
        C.fibers.resize(block);
        for (size_t i = old; i < block; ++i) C.fibers[i].stack = static_cast<char*>(malloc(STACK_BYTES));
    }
    if (C.warps.size() < (block + 31) / 32) C.warps.resize((block + 31) / 32);
    C.n = block;
    C.body = &body;
    C.smem.assign(dyn_smem ? dyn_smem : 1, 0xCD);
    g_blockDim = dim3(block, 1, 1);
    g_gridDim = dim3(grid, 1, 1);
    std::vector<unsigned> border;
    g_shuffle.order(border, grid);
    for (unsigned bx : border) {
        g_blockIdx = uint3{bx, 0, 0};
        run_cta();
    }
}

}  // namespace emu

// ---------------------------------------------------------------------------------------------- CUDA IPC stand-in
// A block allocated with emu::ipc_alloc lives in a POSIX shared-memory segment named after the allocating process, so that
// the rank processes of the emulated strip test can map each other's receive buffers like GPUs map peer memory.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>

namespace emu {
namespace {
struct IpcSeg { std::string name; size_t bytes; bool owner; };
std::map<void*, IpcSeg> g_ipc;
std::mutex g_ipc_mu;
int g_ipc_counter = 0;
}  // namespace

cudaError_t ipc_alloc(void** p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    char name[64];
    snprintf(name, sizeof(name), "/blobs_emu_%d_%d", (int)getpid(), g_ipc_counter++);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return cudaErrorMemoryAllocation;
    if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return cudaErrorMemoryAllocation; }
    void* q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) { shm_unlink(name); return cudaErrorMemoryAllocation; }
    memset(q, 0xCD, bytes);
    g_ipc[q] = IpcSeg{name, bytes, true};
    *p = q;
    return cudaSuccess;
}

void ipc_free(void* p) {
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    auto it = g_ipc.find(p);
    if (it == g_ipc.end()) return;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    g_ipc.erase(it);
}

cudaError_t ipc_get_handle(cudaIpcMemHandle_t* h, void* p) {
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    auto it = g_ipc.find(p);
    if (it == g_ipc.end()) return cudaErrorNotSupported;
    memset(h, 0, sizeof(*h));
    snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    const unsigned long long n = it->second.bytes;
    memcpy(h->reserved + 48, &n, sizeof(n));
    return cudaSuccess;
}

cudaError_t ipc_open(void** p, const cudaIpcMemHandle_t& h) {
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    unsigned long long n = 0;
    memcpy(&n, h.reserved + 48, sizeof(n));
    char name[49];
    memcpy(name, h.reserved, 48);
    name[48] = 0;
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return cudaErrorNotSupported;
    void* q = mmap(nullptr, (size_t)n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorNotSupported;
    g_ipc[q] = IpcSeg{name, (size_t)n, false};
    *p = q;
    return cudaSuccess;
}

cudaError_t ipc_close(void* p) { ipc_free(p); return cudaSuccess; }
}  // namespace emu
