// TEST INFRASTRUCTURE ONLY -- see include/cuda_runtime.h.  Scheduler, memory, runtime-API shims and the asynchronous engines of
// the CPU emulation the kernel sources are compiled against in the GPU-less container.
#define SGM_EMU 1
#include "cuda_runtime.h"

#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cstdarg>
#include <deque>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

// ------------------------------------------------------------------------------------------------------------------------------
// context switch (x86-64 SysV): callee-saved registers + stack pointer
// ------------------------------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

uint3 threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;

static bool g_lazy = true;          // SGM_EMU_ASYNC
static int g_sms = 2;               // SGM_EMU_SMS
static bool g_verbose = false;
static bool g_shm_ipc = false;     // SGM_EMU_SHM_IPC
static int g_sched = 0;             // SGM_EMU_SCHED: 0 rr, 1 reverse, >= 2 random (seed)
static bool g_cluster_lockstep = false;   // SGM_EMU_CLUSTER=lockstep (default: each CTA of a cluster runs as far as it can in turn)
static unsigned long long g_sched_state = 1;

// ---- sticky error ------------------------------------------------------------------------------------------------------------
static int g_err = 0;               // sticky (context-destroying) error
static int g_last = 0;              // non-sticky last error
static std::string g_errmsg;

// ---- fibers ------------------------------------------------------------------------------------------------------------------
constexpr size_t STACK_BYTES = 256 * 1024;
struct CpAsyncOp { uint32_t dst; const void* src; };
struct BulkStore { void* dst; uint32_t src; uint32_t bytes; std::vector<unsigned char> snap; bool read_done = false; };
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
    std::vector<std::vector<CpAsyncOp>> cp_groups;   // committed, not yet waited for
    std::vector<CpAsyncOp> cp_open;                  // issued since the last commit
    std::vector<std::vector<BulkStore>> bulk_groups;  // committed bulk stores, oldest first
    std::vector<BulkStore> bulk_open;
    int wait_kind = 0; uint32_t wait_a = 0, wait_b = 0;   // what a blocked thread is waiting for (formatted only in a deadlock report)
    bool cl_arrived = false; unsigned cl_gen = 0;         // barrier.cluster: arrived and not yet waited; the generation arrived in
};
struct WarpEx {
    uint64_t slot[2][32];
    unsigned have[2] = {0, 0}, read[2] = {0, 0};
    unsigned gen[32];
    unsigned alive = 0;
};
struct PendingCopy { std::function<void()> run; uint32_t bytes; };
struct Mbar {
    bool inited = false;
    uint32_t count = 0, pending = 0;
    int64_t tx = 0;
    uint32_t phase = 0;
    std::deque<PendingCopy> copies;     // TMA loads that complete_tx on this barrier (lazy mode)
    int mma_commits = 0;                // tcgen05.commit arrivals owed once the MMA queue has drained up to them
};
struct MmaOp { std::function<void()> run; };
struct MmaCommit { size_t upto; uint32_t bar; };

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<WarpEx> warps;
    int nthreads = 0, alive = 0, cur = -1;
    int bar_count = 0; unsigned bar_gen = 0;
    unsigned long long progress = 0;
    unsigned char* smem = nullptr; size_t smem_bytes = 0; size_t smem_align = 16; uint32_t window = 0;
    std::unordered_map<uint32_t, Mbar> mbars;
    std::vector<uint32_t> tmem;         // [128][512]
    uint32_t tmem_alloc_cols = 0; bool tmem_permit = true;
    std::deque<MmaOp> mma_queue; size_t mma_issued = 0, mma_done = 0;
    std::deque<MmaCommit> mma_commits;
    const std::function<void()>* body = nullptr;
    const char* kname = "";
    bool aborted = false;
    unsigned bx = 0, by = 0, bz = 0;    // blockIdx of this CTA (the CTAs of a thread-block cluster run interleaved)
    int rank = 0;                       // %cluster_ctarank
    struct Cluster* cl = nullptr;
};
// a thread-block cluster: its CTAs are co-scheduled (fibers of all of them take turns), share one hardware barrier
// (barrier.cluster.arrive / wait) and can address each other's shared memory (mapa + ld/st.shared::cluster)
struct Cluster {
    std::vector<Cta*> ctas;
    int bar_arrived = 0; unsigned bar_gen = 0;
};
constexpr int MAX_CLUSTER = 16;
static Cta* g_cta = nullptr;
static void* g_main_sp = nullptr;
static std::vector<char*> g_stack_pool;
static unsigned char* g_smem_pool = nullptr;
constexpr size_t SMEM_POOL = 256 * 1024 + 4096;

static char* get_stack() {
    if (!g_stack_pool.empty()) { char* s = g_stack_pool.back(); g_stack_pool.pop_back(); return s; }
    void* p = mmap(nullptr, STACK_BYTES + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("emu: mmap stack"); abort(); }
    mprotect(p, 4096, PROT_NONE);   // guard below the stack
    return (char*)p + 4096;
}

static void fiber_exit();
static void bulk_store_write(BulkStore& b);
static void cluster_release_if_complete(struct Cluster& cl);
extern "C" void emu_trampoline() {
    (*g_cta->body)();
    fiber_exit();
}

static void switch_to_main() {
    Fiber& f = g_cta->fibers[g_cta->cur];
    emu_switch(&f.sp, g_main_sp);
}

void progress() { if (g_cta) ++g_cta->progress; }

void yield() {
    if (!g_cta) return;
    switch_to_main();
}

static void release_barrier_if_complete(Cta& c) {
    if (c.alive > 0 && c.bar_count >= c.alive) { c.bar_count = 0; ++c.bar_gen; ++c.progress; }
}

static void fiber_exit() {
    Cta& c = *g_cta;
    Fiber& f = c.fibers[c.cur];
    if (!f.cp_open.empty() || !f.cp_groups.empty()) {
        // outstanding cp.async at thread exit complete eventually on hardware; apply them now
        for (auto& g : f.cp_groups) for (auto& op : g) std::memcpy(smem_ptr(op.dst, 16, 16), op.src, 16);
        for (auto& op : f.cp_open) std::memcpy(smem_ptr(op.dst, 16, 16), op.src, 16);
        f.cp_groups.clear(); f.cp_open.clear();
    }
    for (auto& g : f.bulk_groups) for (auto& b : g) bulk_store_write(b);   // stores still in flight at thread exit complete
    for (auto& b : f.bulk_open) bulk_store_write(b);
    f.bulk_groups.clear(); f.bulk_open.clear();
    f.done = true;
    --c.alive;
    c.warps[c.cur >> 5].alive &= ~(1u << (c.cur & 31));
    ++c.progress;
    release_barrier_if_complete(c);
    if (c.cl) {
        if (f.cl_arrived && f.cl_gen == c.cl->bar_gen) --c.cl->bar_arrived;   // an arrival of this phase by a thread that is gone: the count is over live threads
        cluster_release_if_complete(*c.cl);
    }
    void* dummy;
    emu_switch(&dummy, g_main_sp);
    abort();   // never resumed
}

[[noreturn]] void fault(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    char where[256] = "";
    if (g_cta && g_cta->cur >= 0)
        snprintf(where, sizeof where, " [kernel %s, block (%u,%u,%u), thread %d]", g_cta->kname, blockIdx_.x, blockIdx_.y, blockIdx_.z, g_cta->cur);
    g_errmsg = std::string("emu fault: ") + buf + where;
    g_err = cudaErrorLaunchFailure;
    if (g_verbose || getenv("SGM_EMU_TRACE")) fprintf(stderr, "%s\n", g_errmsg.c_str());
    if (g_cta && g_cta->cur >= 0) {
        g_cta->aborted = true;
        void* dummy;
        emu_switch(&dummy, g_main_sp);
    }
    fprintf(stderr, "%s (outside a kernel)\n", g_errmsg.c_str());
    abort();
}

int linear_tid() { return g_cta ? g_cta->cur : 0; }
int lane_id() { return linear_tid() & 31; }
int warp_id() { return linear_tid() >> 5; }
long long clock() { static long long t = 0; return t += 7; }
static std::map<std::string, long long> g_notes;
void note(const char* name) { ++g_notes[name]; }

void syncthreads() {
    Cta& c = *g_cta;
    const unsigned my = c.bar_gen;
    ++c.bar_count; ++c.progress;
    release_barrier_if_complete(c);
    c.fibers[c.cur].wait_kind = 1;
    while (c.bar_gen == my) yield();
}

uint64_t shfl(unsigned mask, uint64_t v, int src) {
    Cta& c = *g_cta;
    WarpEx& w = c.warps[c.cur >> 5];
    const int lane = c.cur & 31;
    const unsigned b = w.gen[lane] & 1u;
    if (!((mask >> lane) & 1u)) fault("warp shuffle/sync executed by lane %d which is not in its mask %08x", lane, mask);
    w.slot[b][lane] = v; w.have[b] |= 1u << lane; ++c.progress;
    for (;;) {
        const unsigned need = mask & w.alive;
        if ((w.have[b] & need) == need) break;
        { Fiber& f = c.fibers[c.cur]; f.wait_kind = 2; f.wait_a = mask; f.wait_b = w.have[b]; }
        yield();
    }
    const uint64_t r = w.slot[b][src & 31];
    w.read[b] |= 1u << lane;
    const unsigned need = mask & w.alive;
    if ((w.read[b] & need) == need) { w.have[b] = 0; w.read[b] = 0; }
    ++w.gen[lane];
    return r;
}
void syncwarp(unsigned mask) { (void)shfl(mask, 0, lane_id()); }
bool elect_one() { return lane_id() == 0; }

unsigned char* dyn_smem(size_t align) {
    Cta& c = *g_cta;
    if (align > c.smem_align) {
        // the declared alignment of the extern array fixes the window base; all threads of a CTA declare the same one
        c.smem_align = align;
    }
    return c.smem;
}
static uint32_t window_base(size_t align) { return align >= 1024 ? 2048u : 1024u + (uint32_t)align; }   // aligned to `align` and no further

uint32_t smem_window(const void* p) {
    Cta& c = *g_cta;
    const unsigned char* q = (const unsigned char*)p;
    if (q < c.smem || q > c.smem + c.smem_bytes) fault("cvta.to.shared of a pointer outside the CTA's dynamic shared memory");
    return window_base(c.smem_align) + (uint32_t)(q - c.smem);
}
unsigned char* smem_ptr(uint32_t addr, size_t bytes, size_t align) {
    Cta& c = *g_cta;
    const uint32_t base = window_base(c.smem_align);
    if (addr < base || (size_t)(addr - base) + bytes > c.smem_bytes)
        fault("shared-memory access out of range: window address 0x%x + %zu, dynamic shared memory is [0x%x, 0x%zx)", addr, bytes, base, (size_t)base + c.smem_bytes);
    if (align > 1 && (addr % align) != 0) fault("misaligned shared-memory access: address 0x%x, %zu-byte alignment required", addr, align);
    return c.smem + (addr - base);
}

// ---- global memory -----------------------------------------------------------------------------------------------------------
struct Alloc { size_t bytes; void* map; size_t map_bytes; std::string shm; size_t shm_off = 0; bool foreign = false; };
static std::map<const unsigned char*, Alloc> g_allocs;
static const size_t PAGE = 4096;

static const Alloc* find_alloc(const void* p, const unsigned char** base) {
    auto it = g_allocs.upper_bound((const unsigned char*)p);
    if (it == g_allocs.begin()) return nullptr;
    --it;
    if ((const unsigned char*)p < it->first + it->second.bytes) { *base = it->first; return &it->second; }
    return nullptr;
}
void check_global(const void* p, size_t bytes, size_t align, const char* what) {
    const unsigned char* base = nullptr;
    const Alloc* a = find_alloc(p, &base);
    if (!a || (const unsigned char*)p + bytes > base + a->bytes) fault("%s: global range [%p, +%zu) is not inside a device allocation", what, p, bytes);
    if (align > 1 && ((uintptr_t)p % align) != 0) fault("%s: global address %p is not %zu-byte aligned", what, p, align);
}

static struct sigaction g_old_segv;
static void segv_handler(int sig, siginfo_t* si, void* uc) {
    if (g_cta && g_cta->cur >= 0) {
        const unsigned char* base = nullptr;
        (void)base;
        // nearest allocation, for the message
        auto it = g_allocs.upper_bound((const unsigned char*)si->si_addr);
        char near[128] = "";
        if (it != g_allocs.begin()) { --it; snprintf(near, sizeof near, " (%+lld bytes past the end of a %zu-byte allocation)", (long long)((const unsigned char*)si->si_addr - (it->first + it->second.bytes)), it->second.bytes); }
        fault("illegal address %p%s", si->si_addr, near);
    }
    // not a fault inside one of THIS library's kernels: hand it to whoever was installed before (a second emulated build loaded in
    // the same process, Python's faulthandler, ...)
    if ((g_old_segv.sa_flags & SA_SIGINFO) && g_old_segv.sa_sigaction) { g_old_segv.sa_sigaction(sig, si, uc); return; }
    if (!(g_old_segv.sa_flags & SA_SIGINFO) && g_old_segv.sa_handler != SIG_DFL && g_old_segv.sa_handler != SIG_IGN) { g_old_segv.sa_handler(sig); return; }
    signal(SIGSEGV, SIG_DFL);
    raise(SIGSEGV);
}

static void init_once() {
    static bool done = false;
    if (done) return;
    done = true;
    const char* a = getenv("SGM_EMU_ASYNC");
    g_lazy = !(a && !strcmp(a, "eager"));
    if (const char* s = getenv("SGM_EMU_SMS")) g_sms = atoi(s) > 0 ? atoi(s) : 2;
    g_verbose = getenv("SGM_EMU_VERBOSE") != nullptr;
    g_shm_ipc = getenv("SGM_EMU_SHM_IPC") != nullptr;
    if (const char* sc = getenv("SGM_EMU_SCHED")) {
        if (!strcmp(sc, "reverse")) g_sched = 1;
        else if (atoi(sc) > 0) { g_sched = 2; g_sched_state = (unsigned long long)atoi(sc) * 0x9E3779B97F4A7C15ull + 1; }
    }
    if (const char* cs = getenv("SGM_EMU_CLUSTER")) g_cluster_lockstep = !strcmp(cs, "lockstep");
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = segv_handler;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigaction(SIGSEGV, &sa, &g_old_segv);
    g_smem_pool = (unsigned char*)mmap(nullptr, SMEM_POOL * MAX_CLUSTER, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
}

// ---- launch ------------------------------------------------------------------------------------------------------------------
static std::unordered_map<const void*, int> g_dyn_limit;

// runs the CTAs of one cluster (a plain launch: a cluster of one) to completion
static void run_group(std::vector<Cta*>& grp) {
    // round-robin over the fibers of every CTA; a pass without progress while threads are still alive = deadlock
    int stalled = 0;
    auto alive = [&]() { int n = 0; for (Cta* c : grp) n += c->alive; return n; };
    auto aborted = [&]() { for (Cta* c : grp) if (c->aborted) return true; return false; };
    auto progress_sum = [&]() { unsigned long long p = 0; for (Cta* c : grp) p += c->progress; return p; };
    // one pass over the live fibers of one CTA
    auto pass = [&](Cta* cp) {
        Cta& c = *cp;
        g_cta = cp;
        blockIdx_.x = c.bx; blockIdx_.y = c.by; blockIdx_.z = c.bz;
        // SGM_EMU_SCHED: the order in which the WARPS of the CTA get their turn in a pass ("rr" = warp 0 first, "reverse", or a
        // number = seed of a fresh random order per pass).  Lanes of a warp always run in lane order: inter-warp races are real on the
        // device, intra-warp orders that converged code never produces are not explored.
        const int nwarps = (c.nthreads + 31) / 32;
        std::vector<int> order(nwarps);
        for (int w = 0; w < nwarps; ++w) order[w] = w;
        if (g_sched == 1) for (int w = 0; w < nwarps; ++w) order[w] = nwarps - 1 - w;
        else if (g_sched >= 2) for (int w = nwarps - 1; w > 0; --w) { g_sched_state = g_sched_state * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[w], order[(g_sched_state >> 33) % (unsigned)(w + 1)]); }
        for (int ti = 0; ti < c.nthreads && !aborted(); ++ti) {
            const int t = order[ti >> 5] * 32 + (ti & 31);
            if (t >= c.nthreads) continue;
            Fiber& f = c.fibers[t];
            if (f.done) continue;
            c.cur = t;
            threadIdx_.x = (unsigned)t % blockDim_.x;
            threadIdx_.y = ((unsigned)t / blockDim_.x) % blockDim_.y;
            threadIdx_.z = (unsigned)t / (blockDim_.x * blockDim_.y);
            emu_switch(&g_main_sp, f.sp);
        }
        c.cur = -1;
    };
    while (alive() > 0 && !aborted()) {
        const unsigned long long before = progress_sum();
        // CTAs of a cluster run on different SMs at unrelated speeds.  Default ("depth"): each CTA in turn runs as far as it can before
        // the next one gets the machine -- the CTA that completes a cluster barrier races ahead of its peers (reads their shared memory
        // late, overwrites its own early, exits first), the order that exposes a missing barrier.  SGM_EMU_CLUSTER=lockstep: one pass
        // per CTA in turn.  SGM_EMU_SCHED=reverse also reverses the order of the CTAs.
        for (size_t k = 0; k < grp.size(); ++k) {
            Cta* cp = grp[g_sched == 1 ? grp.size() - 1 - k : k];
            if (cp->alive == 0) continue;
            for (;;) {
                const unsigned long long p0 = cp->progress;
                pass(cp);
                if (g_cluster_lockstep || grp.size() == 1 || cp->alive == 0 || cp->progress == p0 || aborted()) break;
            }
        }
        if (progress_sum() == before) { if (++stalled >= 4) break; } else stalled = 0;
    }
    if (!aborted() && alive() > 0) {
        std::string who;
        Cta* first = nullptr;
        for (Cta* cp : grp) {
            Cta& c = *cp;
            if (c.alive == 0) continue;
            if (!first) first = cp;
            c.cur = -1;
            std::map<std::string, std::vector<int>> by_reason;
            for (int t = 0; t < c.nthreads; ++t) {
                const Fiber& f = c.fibers[t];
                if (f.done) continue;
                char why[200] = "(running)";
                if (f.wait_kind == 1) snprintf(why, sizeof why, "__syncthreads (%d of %d live threads arrived)", c.bar_count, c.alive);
                else if (f.wait_kind == 2) snprintf(why, sizeof why, "warp shuffle / __syncwarp (mask %08x, arrived %08x)", f.wait_a, f.wait_b);
                else if (f.wait_kind == 3) {
                    const Mbar& m = c.mbars[f.wait_a];
                    snprintf(why, sizeof why, "mbarrier 0x%x parity %u (phase %u, pending arrivals %u of %u, tx bytes %lld, TMA copies in flight %zu, commits owed %d)",
                             f.wait_a, f.wait_b, m.phase, m.pending, m.count, (long long)m.tx, m.copies.size(), m.mma_commits);
                }
                else if (f.wait_kind == 4) snprintf(why, sizeof why, "barrier.cluster.wait (%d threads of the cluster have arrived)", c.cl ? c.cl->bar_arrived : 0);
                by_reason[why].push_back(t);
            }
            for (auto& kv : by_reason) {
                who += "\n  " + (grp.size() > 1 ? "cluster rank " + std::to_string(c.rank) + ": " : std::string()) + std::to_string(kv.second.size()) + " thread(s) [" +
                       std::to_string(kv.second.front()) + ".." + std::to_string(kv.second.back()) + "] at " + kv.first;
            }
        }
        g_errmsg = std::string("emu fault: DEADLOCK in kernel ") + first->kname + ", block (" + std::to_string(first->bx) + "," + std::to_string(first->by) + "," +
                   std::to_string(first->bz) + "): threads still waiting: " + who;
        g_err = cudaErrorLaunchTimeout;
        for (Cta* cp : grp) cp->aborted = true;
        if (g_verbose || getenv("SGM_EMU_TRACE")) fprintf(stderr, "%s\n", g_errmsg.c_str());
    }
    for (Cta* cp : grp) cp->cur = -1;
    g_cta = nullptr;
}

static void launch_impl(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, unsigned cluster_x, const std::function<void()>& body) {
    init_once();
    if (g_err) return;
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535) {
        g_last = cudaErrorInvalidValue; g_errmsg = std::string("emu: invalid launch configuration for ") + name; return;
    }
    if (cluster_x == 0 || cluster_x > 8 || (grid.x % cluster_x) != 0) {   // portable cluster sizes only (16 needs cudaFuncAttributeNonPortableClusterSizeAllowed)
        g_last = cudaErrorInvalidValue;
        g_errmsg = std::string("emu: ") + name + ": cluster dimension " + std::to_string(cluster_x) + " must be in [1, 8] and divide gridDim.x = " + std::to_string(grid.x);
        return;
    }
    int limit = 48 * 1024;
    auto it = g_dyn_limit.find(fn);
    if (it != g_dyn_limit.end()) limit = it->second;
    if (smem > (size_t)limit) {
        g_last = cudaErrorInvalidValue;
        g_errmsg = std::string("emu: ") + name + " launched with " + std::to_string(smem) + " B of dynamic shared memory, limit " + std::to_string(limit) +
                   " (cudaFuncAttributeMaxDynamicSharedMemorySize)";
        return;
    }
    gridDim_ = grid; blockDim_ = block;
    for (unsigned bz = 0; bz < grid.z && !g_err; ++bz)
        for (unsigned by = 0; by < grid.y && !g_err; ++by)
            for (unsigned bx0 = 0; bx0 < grid.x && !g_err; bx0 += cluster_x) {
                std::vector<Cta> ctas(cluster_x);
                Cluster cl;
                std::vector<Cta*> grp;
                for (unsigned r = 0; r < cluster_x; ++r) {
                    Cta& c = ctas[r];
                    c.kname = name; c.body = &body; c.nthreads = (int)nthreads; c.alive = (int)nthreads;
                    c.bx = bx0 + r; c.by = by; c.bz = bz; c.rank = (int)r; c.cl = &cl;
                    c.fibers.resize(nthreads);
                    c.warps.resize((nthreads + 31) / 32);
                    for (size_t t = 0; t < nthreads; ++t) {
                        Fiber& f = c.fibers[t];
                        f.stack = get_stack(); f.done = false;
                        uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
                        void** sp = (void**)(top - 16);
                        sp[0] = (void*)&emu_trampoline;
                        sp -= 6;
                        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
                        f.sp = sp;
                        WarpEx& w = c.warps[t >> 5];
                        w.alive |= 1u << (t & 31); w.gen[t & 31] = 0;
                    }
                    // poison shared memory: uninitialised reads give NaNs, not zeros
                    unsigned char* pool = g_smem_pool + (size_t)r * SMEM_POOL;
                    memset(pool, 0xff, smem + 2048 < SMEM_POOL ? smem + 2048 : SMEM_POOL);
                    c.smem = pool + 1024; c.smem_bytes = smem;
                    cl.ctas.push_back(&c);
                    grp.push_back(&c);
                }
                run_group(grp);
                for (Cta& c : ctas) {
                    for (auto& f : c.fibers) g_stack_pool.push_back(f.stack);
                    if (!c.aborted && c.tmem_alloc_cols != 0 && !g_err) {
                        g_err = cudaErrorLaunchFailure; g_errmsg = std::string("emu fault: kernel ") + name + " exited without tcgen05.dealloc";
                    }
                }
            }
}
void launch(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) { launch_impl(name, fn, grid, block, smem, 1, body); }
void launch_cluster(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, unsigned cluster_x, const std::function<void()>& body) {
    launch_impl(name, fn, grid, block, smem, cluster_x, body);
}

// ---- thread-block clusters ----------------------------------------------------------------------------------------------------
unsigned cluster_ctarank() { return (unsigned)g_cta->rank; }
unsigned cluster_nctarank() { return (unsigned)g_cta->cl->ctas.size(); }
static int cluster_alive(const Cluster& cl) { int n = 0; for (const Cta* c : cl.ctas) n += c->alive; return n; }
static void cluster_release_if_complete(Cluster& cl) {
    const int need = cluster_alive(cl);
    if (need > 0 && cl.bar_arrived >= need) { cl.bar_arrived = 0; ++cl.bar_gen; for (Cta* c : cl.ctas) ++c->progress; }
}
// barrier.cluster.arrive (+ .release) and barrier.cluster.wait (+ .acquire): every thread of every CTA of the cluster arrives once
// per phase; threads that have exited count as arrived (hardware behaviour)
void cluster_arrive() {
    Cta& c = *g_cta;
    Fiber& f = c.fibers[c.cur];
    if (f.cl_arrived) fault("barrier.cluster.arrive executed twice without a barrier.cluster.wait in between");
    f.cl_arrived = true; f.cl_gen = c.cl->bar_gen;
    ++c.cl->bar_arrived; ++c.progress;
    cluster_release_if_complete(*c.cl);
}
void cluster_wait() {
    Cta& c = *g_cta;
    Fiber& f = c.fibers[c.cur];
    if (!f.cl_arrived) fault("barrier.cluster.wait without a preceding barrier.cluster.arrive");
    f.wait_kind = 4;
    while (c.cl->bar_gen == f.cl_gen) yield();
    f.cl_arrived = false;
}
// mapa.shared::cluster: the shared-window address of the same offset in CTA `rank` of the cluster (the rank rides in bits 24..)
uint32_t mapa(uint32_t addr, uint32_t rank) {
    Cta& c = *g_cta;
    if (rank >= c.cl->ctas.size()) fault("mapa: target rank %u outside the cluster of %zu CTAs", rank, c.cl->ctas.size());
    (void)smem_ptr(addr & 0xFFFFFFu, 1, 1);
    return (addr & 0xFFFFFFu) | ((rank + 1u) << 24);
}
unsigned char* dsmem_ptr(uint32_t addr, size_t bytes, size_t align) {
    Cta& c = *g_cta;
    const uint32_t tag = addr >> 24;
    if (tag == 0) return smem_ptr(addr, bytes, align);   // a shared::cta address is a valid shared::cluster address of the own CTA
    if (tag > c.cl->ctas.size()) fault("shared::cluster access with a malformed address 0x%x", addr);
    Cta& t = *c.cl->ctas[tag - 1];
    if (t.alive == 0) fault("shared::cluster access to CTA rank %u, which has already exited (its shared memory is gone): missing cluster barrier before exit", tag - 1);
    const uint32_t a = addr & 0xFFFFFFu, base = window_base(t.smem_align);
    if (a < base || (size_t)(a - base) + bytes > t.smem_bytes)
        fault("shared::cluster access out of range: rank %u window address 0x%x + %zu, dynamic shared memory is [0x%x, 0x%zx)", tag - 1, a, bytes, base, (size_t)base + t.smem_bytes);
    if (align > 1 && (a % align) != 0) fault("misaligned shared::cluster access: address 0x%x, %zu-byte alignment required", a, align);
    return t.smem + (a - base);
}

}  // namespace emu

// --------------------------------------------------------------------------------------------------------------------------------
// runtime API
// --------------------------------------------------------------------------------------------------------------------------------
using namespace emu;

struct EmuStream { int dummy; };
struct EmuEvent { std::chrono::steady_clock::time_point t; };

#define EMU_STICKY() do { init_once(); if (g_err) return g_err; } while (0)

cudaError_t cudaGetLastError() { init_once(); if (g_err) return g_err; const int e = g_last; g_last = 0; return e; }
cudaError_t cudaPeekAtLastError() { init_once(); return g_err ? g_err : g_last; }
const char* cudaGetErrorString(cudaError_t e) {
    if (e == cudaSuccess) return "no error";
    return g_errmsg.empty() ? "emu: error" : g_errmsg.c_str();
}
cudaError_t cudaGetDeviceCount(int* n) { init_once(); *n = getenv("SGM_EMU_NO_DEVICE") ? 0 : 1; return *n ? cudaSuccess : cudaErrorNoDevice; }
cudaError_t cudaSetDevice(int d) { EMU_STICKY(); return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaDeviceSynchronize() { EMU_STICKY(); return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    EMU_STICKY();
    if (a == cudaDevAttrMultiProcessorCount) *v = g_sms;
    else if (a == cudaDevAttrMaxSharedMemoryPerBlockOptin) *v = 227 * 1024;
    else return cudaErrorInvalidValue;
    return cudaSuccess;
}
cudaError_t cudaMallocImpl(void** p, size_t bytes) {
    EMU_STICKY();
    if (bytes == 0) bytes = 16;
    const size_t payload = (bytes + 15) & ~(size_t)15;
    const size_t pages = (payload + PAGE - 1) / PAGE * PAGE;
    const size_t total = pages + 2 * PAGE;
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { g_last = cudaErrorMemoryAllocation; g_errmsg = "emu: out of memory"; *p = nullptr; return cudaErrorMemoryAllocation; }
    std::string shm_name;
    if (g_shm_ipc && bytes >= (64u << 10)) {
        // SGM_EMU_SHM_IPC=1: large allocations live in POSIX shared memory so that cudaIpc* works ACROSS processes (multi-rank tests)
        static int counter = 0;
        char nm[40];
        snprintf(nm, sizeof nm, "/sgm_emu_%d_%d", (int)getpid(), counter++);
        const int fd = shm_open(nm, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)pages) == 0 && mmap((char*)m + PAGE, pages, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0) != MAP_FAILED) shm_name = nm;
        if (fd >= 0) close(fd);
        if (shm_name.empty()) { munmap(m, total); g_last = cudaErrorMemoryAllocation; g_errmsg = "emu: shared-memory allocation failed"; *p = nullptr; return cudaErrorMemoryAllocation; }
    }
    mprotect(m, PAGE, PROT_NONE);
    mprotect((char*)m + PAGE + pages, PAGE, PROT_NONE);
    unsigned char* q = (unsigned char*)m + PAGE + pages - payload;   // the end of the payload touches the guard page
    memset(q, 0xff, payload);                                         // uninitialised device memory is not zero
    Alloc al{bytes, m, total};
    al.shm = shm_name; al.shm_off = pages - payload;
    g_allocs[q] = al;
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    auto it = g_allocs.find((const unsigned char*)p);
    if (it == g_allocs.end()) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaFree of an unknown pointer"; return cudaErrorInvalidValue; }
    if (it->second.foreign) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaFree of a peer's IPC mapping (use cudaIpcCloseMemHandle)"; return cudaErrorInvalidValue; }
    if (!it->second.shm.empty()) shm_unlink(it->second.shm.c_str());
    munmap(it->second.map, it->second.map_bytes);
    g_allocs.erase(it);
    return cudaSuccess;
}
// page-locked host allocations are remembered: cudaPointerGetAttributes tells them from pageable memory
static std::map<const unsigned char*, size_t> g_pinned;
cudaError_t cudaHostAllocImpl(void** p, size_t bytes, unsigned) {
    EMU_STICKY();
    const size_t sz = (bytes + 63) & ~(size_t)63;
    *p = aligned_alloc(64, sz ? sz : 64);
    if (*p) g_pinned[(const unsigned char*)*p] = sz;
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) { g_pinned.erase((const unsigned char*)p); free(p); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    EMU_STICKY();
    a->device = 0; a->devicePointer = nullptr; a->hostPointer = nullptr; a->type = cudaMemoryTypeUnregistered;
    auto it = g_pinned.upper_bound((const unsigned char*)p);
    if (it != g_pinned.begin()) {
        --it;
        if ((const unsigned char*)p < it->first + it->second) { a->type = cudaMemoryTypeHost; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p); return cudaSuccess; }
    }
    const unsigned char* base = nullptr;
    if (find_alloc(p, &base)) { a->type = cudaMemoryTypeDevice; a->devicePointer = const_cast<void*>(p); }
    return cudaSuccess;
}

static cudaError_t checked_copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k) {
    EMU_STICKY();
    if (bytes == 0) return cudaSuccess;
    const unsigned char* b;
    const bool dst_dev = (k == cudaMemcpyHostToDevice || k == cudaMemcpyDeviceToDevice), src_dev = (k == cudaMemcpyDeviceToHost || k == cudaMemcpyDeviceToDevice);
    if (dst_dev) { const Alloc* a = find_alloc(dst, &b); if (!a || (unsigned char*)dst + bytes > b + a->bytes) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaMemcpy destination outside a device allocation"; return cudaErrorInvalidValue; } }
    if (src_dev) { const Alloc* a = find_alloc(src, &b); if (!a || (const unsigned char*)src + bytes > b + a->bytes) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaMemcpy source outside a device allocation"; return cudaErrorInvalidValue; } }
    memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k) { return checked_copy(dst, src, bytes, k); }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t) { return checked_copy(dst, src, bytes, k); }
cudaError_t cudaMemset(void* dst, int v, size_t bytes) {
    EMU_STICKY();
    const unsigned char* b; const Alloc* a = find_alloc(dst, &b);
    if (bytes && (!a || (unsigned char*)dst + bytes > b + a->bytes)) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaMemset outside a device allocation"; return cudaErrorInvalidValue; }
    memset(dst, v, bytes); return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t) { return cudaMemset(dst, v, bytes); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { EMU_STICKY(); *s = new EmuStream{0}; return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { EMU_STICKY(); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { EMU_STICKY(); *e = new EmuEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { EMU_STICKY(); e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { EMU_STICKY(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { EMU_STICKY(); return cudaSuccess; }   // every emulated stream is synchronous
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { EMU_STICKY(); *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaFuncSetAttributeImpl(const void* fn, cudaFuncAttribute a, int v) {
    EMU_STICKY();
    if (a != cudaFuncAttributeMaxDynamicSharedMemorySize) return cudaErrorInvalidValue;
    if (v < 0 || v > 227 * 1024) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaFuncSetAttribute: more than 227 KB of dynamic shared memory (" + std::to_string(v) + ")"; return cudaErrorInvalidValue; }
    g_dyn_limit[fn] = v > 48 * 1024 ? v : 48 * 1024;
    return cudaSuccess;
}
cudaError_t cudaOccupancyImpl(int* n, const void*, int threads, size_t smem) {
    EMU_STICKY();
    int by_smem = (int)((228 * 1024) / (smem + 1024)), by_thr = 2048 / (threads > 0 ? threads : 1);
    int v = by_smem < by_thr ? by_smem : by_thr;
    *n = v > 32 ? 32 : v;
    return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* sym, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    EMU_STICKY();
    if (!strcmp(sym, "cuTensorMapEncodeTiled")) { *fn = (void*)&emu::encode_tiled; if (q) *q = cudaDriverEntryPointSuccess; return cudaSuccess; }
    *fn = nullptr; if (q) *q = cudaDriverEntryPointSymbolNotFound; return cudaErrorInvalidValue;
}
// CUDA IPC stand-in.  Same process: the handle carries the pointer (one process can play several ranks).  Other process: the
// allocation must live in shared memory (SGM_EMU_SHM_IPC=1); the handle names the object and the payload's offset in it.
struct IpcRec { char name[32]; void* ptr; int pid; uint32_t pad; uint64_t off, bytes; };
static_assert(sizeof(IpcRec) <= sizeof(cudaIpcMemHandle_t), "IPC record fits the opaque handle");
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    EMU_STICKY();
    auto it = g_allocs.find((const unsigned char*)p);
    if (it == g_allocs.end()) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cudaIpcGetMemHandle of a pointer that is not the base of an allocation"; return cudaErrorInvalidValue; }
    IpcRec r; memset(&r, 0, sizeof r);
    snprintf(r.name, sizeof r.name, "%s", it->second.shm.c_str());
    r.ptr = p; r.pid = (int)getpid(); r.off = it->second.shm_off; r.bytes = it->second.bytes;
    memset(h, 0, sizeof *h); memcpy(h->reserved, &r, sizeof r);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    EMU_STICKY();
    IpcRec r; memcpy(&r, h.reserved, sizeof r);
    if (r.pid == (int)getpid()) { *p = r.ptr; return cudaSuccess; }
    if (!r.name[0]) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: IPC handle of another process whose allocation is not in shared memory (SGM_EMU_SHM_IPC=1)"; return cudaErrorInvalidValue; }
    const int fd = shm_open(r.name, O_RDWR, 0600);
    if (fd < 0) { g_last = cudaErrorInvalidValue; g_errmsg = "emu: cannot open the peer's shared-memory allocation"; return cudaErrorInvalidValue; }
    const size_t map_bytes = (r.off + r.bytes + PAGE - 1) / PAGE * PAGE;
    void* m = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { g_last = cudaErrorMemoryAllocation; g_errmsg = "emu: cannot map the peer's allocation"; return cudaErrorMemoryAllocation; }
    unsigned char* q = (unsigned char*)m + r.off;
    Alloc al{(size_t)r.bytes, m, map_bytes};
    al.foreign = true;
    g_allocs[q] = al;
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
    auto it = g_allocs.find((const unsigned char*)p);
    if (it != g_allocs.end() && it->second.foreign) { munmap(it->second.map, it->second.map_bytes); g_allocs.erase(it); }
    return cudaSuccess;
}

// test hook: clear the sticky error (a new "context")
extern "C" void emu_reset_error() { g_err = 0; g_last = 0; g_errmsg.clear(); }
extern "C" const char* emu_last_message() { return g_errmsg.c_str(); }
extern "C" long long emu_note_count(const char* name) { auto it = emu::g_notes.find(name); return it == emu::g_notes.end() ? 0 : it->second; }
extern "C" int emu_async_is_lazy() { init_once(); return g_lazy ? 1 : 0; }

// --------------------------------------------------------------------------------------------------------------------------------
// asynchronous engines
// --------------------------------------------------------------------------------------------------------------------------------
namespace emu {

// ---- cp.async
void cp_async16(uint32_t dst, const void* src) {
    check_global(src, 16, 16, "cp.async");
    (void)smem_ptr(dst, 16, 16);
    Fiber& f = g_cta->fibers[g_cta->cur];
    if (g_lazy) f.cp_open.push_back(CpAsyncOp{dst, src});
    else std::memcpy(smem_ptr(dst, 16, 16), src, 16);
}
void cp_async_commit() {
    Fiber& f = g_cta->fibers[g_cta->cur];
    f.cp_groups.push_back(std::move(f.cp_open));
    f.cp_open.clear();
}
void cp_async_wait_group(int allowed) {
    Fiber& f = g_cta->fibers[g_cta->cur];
    while ((int)f.cp_groups.size() > allowed) {
        for (auto& op : f.cp_groups.front()) std::memcpy(smem_ptr(op.dst, 16, 16), op.src, 16);
        f.cp_groups.erase(f.cp_groups.begin());
    }
}

// ---- mbarrier
static Mbar& mbar_at(uint32_t bar, bool must_be_inited = true) {
    (void)smem_ptr(bar, 8, 8);
    Mbar& m = g_cta->mbars[bar];
    if (must_be_inited && !m.inited) fault("mbarrier at 0x%x used before mbarrier.init", bar);
    return m;
}
static void mbar_check(Mbar& m) {
    if (m.pending == 0 && m.tx == 0) { m.phase ^= 1u; m.pending = m.count; progress(); }
}
static void mbar_settle(Mbar& m, uint32_t bar);
void mbar_init(uint32_t bar, uint32_t count) {
    Mbar& m = mbar_at(bar, false);
    if (count == 0) fault("mbarrier.init with count 0");
    if (m.inited) {
        // re-initialisation: whatever was still owed to the old barrier object lands on it first (a kernel that re-initialises a
        // barrier while a TMA copy to it is in flight corrupts the new phase on hardware)
        if (!m.copies.empty()) fault("mbarrier.init at 0x%x while a TMA copy that signals it is still in flight", bar);
        mbar_settle(m, bar);
    }
    m = Mbar();
    m.inited = true; m.count = count; m.pending = count;
    progress();
}
static void mbar_arrive_n(Mbar& m, uint32_t bar) {
    if (m.pending == 0) fault("mbarrier at 0x%x: more arrivals than its count (%u) in one phase", bar, m.count);
    --m.pending; progress();
    mbar_check(m);
}
void mbar_arrive(uint32_t bar) { mbar_arrive_n(mbar_at(bar), bar); }
void mbar_expect_tx(uint32_t bar, uint32_t bytes) { Mbar& m = mbar_at(bar); m.tx += bytes; progress(); mbar_check(m); }
void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { Mbar& m = mbar_at(bar); m.tx += bytes; mbar_arrive_n(m, bar); }

static void mma_drain(size_t upto);
// everything that may complete lazily on this barrier does so now (a thread is about to look at it)
static void mbar_settle(Mbar& m, uint32_t bar) {
    while (!m.copies.empty()) {
        PendingCopy pc = std::move(m.copies.front());
        m.copies.pop_front();
        pc.run();
        m.tx -= pc.bytes; progress();
        mbar_check(m);
    }
    if (m.mma_commits > 0) {
        Cta& c = *g_cta;
        // the tensor core completes MMAs and their commits IN ORDER: observing a commit owed to this barrier means every earlier
        // commit (to whatever barrier) has arrived too.  Find the last commit owed to `bar` and retire the queue up to it.
        size_t last = c.mma_commits.size();
        for (size_t i = 0; i < c.mma_commits.size(); ++i) if (c.mma_commits[i].bar == bar) last = i;
        for (size_t i = 0; i <= last && !c.mma_commits.empty(); ++i) {
            const MmaCommit mc = c.mma_commits.front();
            c.mma_commits.pop_front();
            mma_drain(mc.upto);
            Mbar& tgt = mbar_at(mc.bar);
            --tgt.mma_commits;
            mbar_arrive_n(tgt, mc.bar);
        }
    }
}
bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    Mbar& m = mbar_at(bar);
    if (m.phase != (parity & 1u)) return true;
    mbar_settle(m, bar);
    if (m.phase != (parity & 1u)) return true;
    yield();
    return g_cta->mbars[bar].phase != (parity & 1u);
}
void mbar_wait(uint32_t bar, uint32_t parity) {
    for (;;) {
        Mbar& m = mbar_at(bar);
        if (m.phase != (parity & 1u)) return;
        mbar_settle(m, bar);
        if (m.phase != (parity & 1u)) return;
        { Fiber& f = g_cta->fibers[g_cta->cur]; f.wait_kind = 3; f.wait_a = bar; f.wait_b = parity & 1u; }
        yield();
    }
}

// ---- TMA
static void tma_complete(uint32_t bar, uint32_t bytes, std::function<void()> run) {
    Mbar& m = mbar_at(bar);
    if (g_lazy) { m.copies.push_back(PendingCopy{std::move(run), bytes}); progress(); }
    else { run(); m.tx -= bytes; progress(); mbar_check(m); }
}
void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if (bytes == 0 || (bytes & 15u)) fault("cp.async.bulk: size %u is not a positive multiple of 16", bytes);
    check_global(src, bytes, 16, "cp.async.bulk (global source)");
    (void)smem_ptr(dst, bytes, 16);
    tma_complete(bar, bytes, [=]() { std::memcpy(smem_ptr(dst, bytes, 16), src, bytes); });
}
// bulk stores: lazy = shared memory is read when the issuing thread executes wait_group.read (or wait_group) for the store's
// group, global memory is written when it executes wait_group; eager = both at issue
static void bulk_store_read(BulkStore& b) {
    if (b.read_done) return;
    b.snap.assign(smem_ptr(b.src, b.bytes, 16), smem_ptr(b.src, b.bytes, 16) + b.bytes);
    b.read_done = true;
}
static void bulk_store_write(BulkStore& b) { bulk_store_read(b); std::memcpy(b.dst, b.snap.data(), b.bytes); }
void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    if (bytes == 0 || (bytes & 15u)) fault("cp.async.bulk (store): size %u is not a positive multiple of 16", bytes);
    check_global(dst, bytes, 16, "cp.async.bulk (global destination)");
    (void)smem_ptr(src, bytes, 16);
    BulkStore b; b.dst = dst; b.src = src; b.bytes = bytes;
    if (!g_lazy) bulk_store_write(b);
    else g_cta->fibers[g_cta->cur].bulk_open.push_back(std::move(b));
}
void bulk_commit_group() {
    Fiber& f = g_cta->fibers[g_cta->cur];
    f.bulk_groups.push_back(std::move(f.bulk_open));
    f.bulk_open.clear();
}
void bulk_wait_group(int allowed, bool read_only) {
    Fiber& f = g_cta->fibers[g_cta->cur];
    const int n = (int)f.bulk_groups.size();
    for (int g = 0; g < n - allowed; ++g)
        for (auto& b : f.bulk_groups[g]) { if (read_only) bulk_store_read(b); else bulk_store_write(b); }
    if (!read_only) f.bulk_groups.erase(f.bulk_groups.begin(), f.bulk_groups.begin() + (n - allowed > 0 ? n - allowed : 0));
    else {
        // groups whose reads are done stay queued until a full wait (or thread exit) performs the writes; wait_group.read counts
        // only groups that still have unread sources, so mark and skip them on later calls
    }
}

static const TensorMapRec* map_rec(const CUtensorMap* map) {
    const TensorMapRec* r = reinterpret_cast<const TensorMapRec*>(map->opaque);
    if (r->magic != 0x544d4150454d5531ull) fault("TMA with a CUtensorMap that was not produced by cuTensorMapEncodeTiled");
    return r;
}
static inline uint32_t swz(uint32_t addr, uint32_t mode) {
    switch (mode) {
        case CU_TENSOR_MAP_SWIZZLE_128B: return addr ^ (((addr >> 7) & 7u) << 4);
        case CU_TENSOR_MAP_SWIZZLE_64B: return addr ^ (((addr >> 7) & 3u) << 4);
        case CU_TENSOR_MAP_SWIZZLE_32B: return addr ^ (((addr >> 7) & 1u) << 4);
    }
    return addr;
}
// copy one box row (inner dimension) element by element with zero fill outside the tensor
static void tma_row(const TensorMapRec* r, uint32_t dst_row, const long long* c /* coordinates of the row start */, uint32_t inner) {
    bool outer_ok = true;
    size_t off = 0;
    for (uint32_t k = 1; k < r->rank; ++k) {
        if (c[k] < 0 || (unsigned long long)c[k] >= r->dim[k]) outer_ok = false;
        else off += (size_t)c[k] * r->stride[k];
    }
    for (uint32_t i = 0; i < inner; ++i) {
        const long long ci = c[0] + i;
        uint32_t v = 0;
        if (outer_ok && ci >= 0 && (unsigned long long)ci < r->dim[0]) std::memcpy(&v, r->base + off + (size_t)ci * r->elem_bytes, 4);
        const uint32_t a = swz(dst_row + i * 4u, r->swizzle);
        std::memcpy(smem_ptr(a, 4, 4), &v, 4);
    }
}
static void tma_check_common(const TensorMapRec* r, uint32_t dst, int c0, uint32_t total_bytes) {
    if (r->elem_bytes != 4) fault("emulated TMA handles 4-byte elements only");
    if (((long long)c0 * r->elem_bytes) % 16 != 0) fault("TMA: innermost start coordinate %d is not 16-byte aligned (illegal instruction on the device)", c0);
    // destination: 128-byte aligned.  The swizzle is a function of the shared-memory ADDRESS bits (16-byte chunk index ^= bits 7..9),
    // so a request may land anywhere inside a 1024-byte aligned tile (tile::gather4 fills it 512 bytes at a time) and still agree
    // with the UMMA descriptor that reads the tile -- provided the TILE base is 1024-byte aligned, which the descriptor check enforces.
    (void)smem_ptr(dst, total_bytes, 128);
}
void tma_tile(uint32_t dst, const CUtensorMap* map, const int* coord, int ncoord, uint32_t bar) {
    const TensorMapRec* r = map_rec(map);
    if ((uint32_t)ncoord != r->rank) fault("TMA tile: %d coordinates for a rank-%u tensor map", ncoord, r->rank);
    uint32_t bytes = r->elem_bytes;
    for (uint32_t k = 0; k < r->rank; ++k) bytes *= r->box[k];
    tma_check_common(r, dst, coord[0], bytes);
    long long c0[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < ncoord; ++k) c0[k] = coord[k];
    const TensorMapRec rec = *r;
    tma_complete(bar, bytes, [=]() {
        const uint32_t inner = rec.box[0];
        uint32_t rows = 1;
        for (uint32_t k = 1; k < rec.rank; ++k) rows *= rec.box[k];
        for (uint32_t row = 0; row < rows; ++row) {
            long long c[5] = {c0[0], 0, 0, 0, 0};
            uint32_t rem = row;
            for (uint32_t k = 1; k < rec.rank; ++k) { c[k] = c0[k] + (rem % rec.box[k]); rem /= rec.box[k]; }
            tma_row(&rec, dst + row * inner * 4u, c, inner);
        }
    });
}
void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    const TensorMapRec* r = map_rec(map);
    if (r->rank != 2) fault("TMA tile::gather4 needs a rank-2 tensor map");
    if (r->box[1] != 1) fault("TMA tile::gather4 needs a box of one row (box[1] = %u)", r->box[1]);
    const uint32_t inner = r->box[0], bytes = 4u * inner * 4u;
    tma_check_common(r, dst, col, bytes);
    const TensorMapRec rec = *r;
    const int rows[4] = {r0, r1, r2, r3};
    tma_complete(bar, bytes, [=]() {
        for (int k = 0; k < 4; ++k) {
            const long long c[5] = {col, rows[k], 0, 0, 0};
            tma_row(&rec, dst + (uint32_t)k * inner * 4u, c, inner);
        }
    });
}

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim, const cuuint64_t* gstride,
                      const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    init_once();
    auto bad = [&](const char* why) { g_errmsg = std::string("emu: cuTensorMapEncodeTiled: ") + why; if (g_verbose) fprintf(stderr, "%s\n", g_errmsg.c_str()); return CUDA_ERROR_INVALID_VALUE; };
    if (!map || rank < 1 || rank > 5) return bad("rank must be 1..5");
    if (dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT32 && dt != CU_TENSOR_MAP_DATA_TYPE_TFLOAT32) return bad("only float32 is modelled");
    if (il != CU_TENSOR_MAP_INTERLEAVE_NONE) return bad("interleave is not modelled");
    if (((uintptr_t)base & 15u) != 0) return bad("globalAddress must be 16-byte aligned");
    const unsigned char* ab = nullptr;
    const Alloc* al = find_alloc(base, &ab);
    if (!al) return bad("globalAddress is not a device allocation");
    TensorMapRec r;
    memset(&r, 0, sizeof r);
    r.magic = 0x544d4150454d5531ull; r.base = (const unsigned char*)base; r.rank = rank; r.elem_bytes = 4; r.swizzle = (uint32_t)sw;
    size_t extent = 0;
    for (uint32_t k = 0; k < rank; ++k) {
        if (gdim[k] == 0 || gdim[k] > (1ull << 32)) return bad("globalDim must be in [1, 2^32]");
        if (box[k] == 0 || box[k] > 256) return bad("boxDim must be in [1, 256]");
        if (estr[k] == 0 || estr[k] > 8) return bad("elementStrides must be in [1, 8]");
        if (estr[k] != 1) return bad("element strides other than 1 are not modelled");
        r.dim[k] = gdim[k]; r.box[k] = box[k];
        r.stride[k] = k == 0 ? 4 : gstride[k - 1];
        if (k > 0) {
            if (gstride[k - 1] % 16 != 0) return bad("globalStrides must be multiples of 16 bytes");
            if (gstride[k - 1] >= (1ull << 40)) return bad("globalStrides must be below 2^40");
        }
        extent += (size_t)(gdim[k] - 1) * r.stride[k];
    }
    extent += 4;
    if ((const unsigned char*)base + extent > ab + al->bytes) return bad("the tensor (globalDim x globalStrides) extends past the end of the device allocation");
    const uint32_t inner_bytes = box[0] * 4u;
    if (inner_bytes % 16 != 0) return bad("boxDim[0] x elementSize must be a multiple of 16 bytes");
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && inner_bytes > 128) return bad("SWIZZLE_128B: inner box dimension must be <= 128 bytes");
    if (sw == CU_TENSOR_MAP_SWIZZLE_64B && inner_bytes > 64) return bad("SWIZZLE_64B: inner box dimension must be <= 64 bytes");
    if (sw == CU_TENSOR_MAP_SWIZZLE_32B && inner_bytes > 32) return bad("SWIZZLE_32B: inner box dimension must be <= 32 bytes");
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && inner_bytes != 128) return bad("SWIZZLE_128B with an inner box narrower than 128 bytes is not modelled");
    static_assert(sizeof(TensorMapRec) <= sizeof(CUtensorMap), "record fits the opaque descriptor");
    memcpy(map->opaque, &r, sizeof r);
    return CUDA_SUCCESS;
}

// ---- tcgen05
static inline uint32_t& tmem_at(uint32_t lane, uint32_t col) {
    Cta& c = *g_cta;
    if (lane >= 128 || col >= 512) fault("TMEM access out of range: lane %u, column %u", lane, col);
    if (col >= c.tmem_alloc_cols) fault("TMEM column %u was not allocated (%u columns allocated)", col, c.tmem_alloc_cols);
    return c.tmem[(size_t)lane * 512 + col];
}
void tmem_alloc(uint32_t slot, uint32_t ncols) {
    Cta& c = *g_cta;
    (void)shfl(0xffffffffu, 0, 0);   // .sync.aligned: the whole warp executes it
    if (lane_id() != 0) { (void)shfl(0xffffffffu, 0, 0); return; }
    if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) fault("tcgen05.alloc: %u columns (must be a power of two in [32, 512])", ncols);
    if (!c.tmem_permit) fault("tcgen05.alloc after relinquish_alloc_permit");
    if (c.tmem_alloc_cols) fault("tcgen05.alloc twice (the emulation models one allocation per CTA)");
    c.tmem.assign(128 * 512, 0x7fc00000u);   // NaN: uninitialised accumulators are visible
    c.tmem_alloc_cols = ncols;
    const uint32_t base = 0;
    std::memcpy(smem_ptr(slot, 4, 4), &base, 4);
    progress();
    (void)shfl(0xffffffffu, 0, 0);
}
void tmem_relinquish() { g_cta->tmem_permit = false; }
void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    Cta& c = *g_cta;
    (void)shfl(0xffffffffu, 0, 0);
    if (lane_id() != 0) return;
    if (taddr != 0 || ncols != c.tmem_alloc_cols) fault("tcgen05.dealloc(%u, %u) does not match the allocation (0, %u)", taddr, ncols, c.tmem_alloc_cols);
    if (c.mma_done != c.mma_issued && !c.mma_queue.empty()) {
        // MMAs still in flight at dealloc: nobody ever observed their completion
        fault("tcgen05.dealloc with %zu tcgen05.mma still uncommitted / unobserved", c.mma_queue.size());
    }
    c.tmem_alloc_cols = 0;
}

// shared-memory matrix descriptor (sm_100): start address [0,14), LBO [16,30), SBO [32,46), version [46,48), layout type [61,64)
struct SmemDesc { uint32_t start, lbo, sbo, layout; };
static SmemDesc decode_desc(uint64_t d) {
    SmemDesc s;
    s.start = (uint32_t)(d & 0x3FFF) << 4; s.lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4; s.sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
    s.layout = (uint32_t)(d >> 61) & 7u;
    if (((d >> 46) & 3u) != 1u) fault("UMMA shared-memory descriptor: version field must be 1 on sm_100 (got %u)", (unsigned)((d >> 46) & 3u));
    return s;
}
static inline float tf32_in(uint32_t bits) { return __uint_as_float(bits & 0xFFFFE000u); }   // the tensor core reads 19 bits
// element (row, k) of a K-major operand tile, k in [0, 8): rows are 128-byte lines, 8-row groups SBO apart, 128-byte swizzle on address bits
static inline float desc_elem(const SmemDesc& s, uint32_t row, uint32_t k) {
    if (s.layout != 2u && s.layout != 0u) fault("UMMA descriptor layout type %u is not modelled (2 = SWIZZLE_128B, 0 = none)", s.layout);
    uint32_t addr;
    if (s.layout == 2u) {
        if ((s.start & ~127u & 1023u) != 0 && (s.start & 1023u) >= 128u) fault("UMMA SWIZZLE_128B operand: the tile base (0x%x minus its K offset) is not 1024-byte aligned", s.start);
        addr = s.start + (row >> 3) * s.sbo + (row & 7u) * 128u + k * 4u;
        addr = swz(addr, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
        // no swizzle, K-major: 8 x 16-byte core matrices; LBO between core matrices along K, SBO between 8-row groups
        addr = s.start + (row >> 3) * s.sbo + (k >> 2) * s.lbo + (row & 7u) * 16u + (k & 3u) * 4u;
    }
    uint32_t v;
    std::memcpy(&v, smem_ptr(addr, 4, 4), 4);
    return tf32_in(v);
}
struct IDesc { uint32_t M, N; };
static IDesc decode_idesc(uint32_t id) {
    if (((id >> 4) & 3u) != 1u) fault("UMMA instruction descriptor: D format must be F32");
    if (((id >> 7) & 7u) != 2u || ((id >> 10) & 7u) != 2u) fault("UMMA instruction descriptor: A/B format must be TF32 for kind::tf32");
    if ((id >> 15) & 1u) fault("UMMA instruction descriptor: MN-major A is not modelled");
    if ((id >> 16) & 1u) fault("UMMA instruction descriptor: MN-major B is not modelled");
    IDesc r; r.N = ((id >> 17) & 0x3Fu) << 3; r.M = ((id >> 24) & 0x1Fu) << 4;
    if (r.M != 128) fault("UMMA: M = %u (the emulation models M = 128, cta_group::1)", r.M);
    if (r.N < 16 || r.N > 256 || (r.N % 16)) fault("UMMA: N = %u is not a multiple of 16 in [16, 256]", r.N);
    return r;
}
static void mma_issue(std::function<void()> run) {
    Cta& c = *g_cta;
    ++c.mma_issued;
    if (g_lazy) c.mma_queue.push_back(MmaOp{std::move(run)});
    else { run(); ++c.mma_done; }
    progress();
}
static void mma_drain(size_t upto) {
    Cta& c = *g_cta;
    while (c.mma_done < upto && !c.mma_queue.empty()) { c.mma_queue.front().run(); c.mma_queue.pop_front(); ++c.mma_done; }
}
static void mma_body(uint32_t tmem_d, bool a_tmem, uint32_t tmem_a, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
    const IDesc id = decode_idesc(idesc);
    const SmemDesc sb = decode_desc(bdesc);
    SmemDesc sa{};
    if (!a_tmem) sa = decode_desc(adesc);
    const uint32_t dl = tmem_d >> 16, dc = tmem_d & 0xFFFFu;
    if (dl != 0) fault("UMMA: accumulator address must be at TMEM lane 0 for M = 128 (got lane %u)", dl);
    float bt[256][8];
    for (uint32_t n = 0; n < id.N; ++n) for (uint32_t k = 0; k < 8; ++k) bt[n][k] = desc_elem(sb, n, k);
    for (uint32_t m = 0; m < 128; ++m) {
        float a[8];
        for (uint32_t k = 0; k < 8; ++k) a[k] = a_tmem ? tf32_in(tmem_at(m + (tmem_a >> 16), (tmem_a & 0xFFFFu) + k)) : desc_elem(sa, m, k);
        for (uint32_t n = 0; n < id.N; ++n) {
            float s = 0.f;
            for (uint32_t k = 0; k < 8; ++k) s += a[k] * bt[n][k];
            uint32_t& d = tmem_at(m, dc + n);
            const float r = acc ? __uint_as_float(d) + s : s;
            d = __float_as_uint(r);
        }
    }
}
void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    (void)decode_idesc(idesc); (void)decode_desc(adesc); (void)decode_desc(bdesc);   // malformed descriptors fault at issue
    mma_issue([=]() { mma_body(tmem_d, false, 0, adesc, bdesc, idesc, accumulate); });
}
void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    (void)decode_idesc(idesc); (void)decode_desc(bdesc);
    mma_issue([=]() { mma_body(tmem_d, true, tmem_a, 0, bdesc, idesc, accumulate); });
}
void umma_commit(uint32_t bar) {
    Cta& c = *g_cta;
    Mbar& m = mbar_at(bar);
    if (g_lazy) { c.mma_commits.push_back(MmaCommit{c.mma_issued, bar}); ++m.mma_commits; progress(); }
    else mbar_arrive_n(m, bar);
}
static uint32_t tmem_lane_of_thread(uint32_t taddr) {
    const uint32_t lane_field = taddr >> 16;
    const uint32_t quarter = 32u * ((uint32_t)warp_id() & 3u);
    if (lane_field != quarter) fault("tcgen05.ld/st .32x32b: warp %d may only touch TMEM lanes %u..%u (address names lane %u)", warp_id(), quarter, quarter + 31, lane_field);
    return quarter + (uint32_t)lane_id();
}
void tmem_ld_32x32b(uint32_t taddr, uint32_t* regs, int n) {
    const uint32_t lane = tmem_lane_of_thread(taddr), col = taddr & 0xFFFFu;
    for (int j = 0; j < n; ++j) regs[j] = tmem_at(lane, col + j);
}
void tmem_st_32x32b(uint32_t taddr, const uint32_t* regs, int n) {
    const uint32_t lane = tmem_lane_of_thread(taddr), col = taddr & 0xFFFFu;
    for (int j = 0; j < n; ++j) tmem_at(lane, col + j) = regs[j];
}

}  // namespace emu
