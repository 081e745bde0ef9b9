// tests/host_emul/cuda_emu/emu_runtime.cpp — TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
// Fiber scheduler for emulated CUDA threads, a synchronous "stream" runtime with graph capture, guarded device
// memory, and a file-based stand-in for the NCCL calls comm.cu makes.
#include "cuda_runtime.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

// ---------------------------------------------------------------------------------------------------------
// context switch (x86-64 System V): callee-saved registers + MXCSR / x87 control word live on the fiber's stack
// ---------------------------------------------------------------------------------------------------------
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
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
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

ThreadCtx* g_cur = nullptr;

static cudaError_t g_last_error = cudaSuccess;
void set_error(cudaError_t e) { if (g_last_error == cudaSuccess) g_last_error = e; }

[[noreturn]] static void die(const char* msg)
{
    std::fprintf(stderr, "cuda_emu: fatal: %s\n", msg);
    std::abort();
}

struct Barrier {
    int expected = 0, count = 0;
    unsigned gen = 0;
};

struct Warp {
    Barrier bar;
    alignas(16) unsigned char buf[2][32][16];
    bool alive[32] = {};
};

struct Cta;
struct Fiber {
    ThreadCtx ctx;
    void* sp = nullptr;
    void* stack = nullptr;
    bool done = false;
    Barrier* wait_bar = nullptr;
    unsigned wait_gen = 0;
    Warp* warp = nullptr;
    Cta* cta = nullptr;
    int lane = 0;
    int parity = 0;
};

struct Cta {
    Barrier bar;
    std::vector<Warp> warps;
    unsigned char* smem = nullptr;
};

struct Group {   // the CTAs that run together: one CTA, or one cluster
    Barrier bar;
    std::vector<Cta> ctas;
    std::vector<Fiber> fibers;
    const std::function<void()>* body = nullptr;
};

static Group* g_group = nullptr;
static Fiber* g_fiber = nullptr;
static void* g_sched_sp = nullptr;

constexpr size_t kStackBytes = 256 * 1024;
static std::vector<void*> g_stack_pool;

static void* stack_get()
{
    if (!g_stack_pool.empty()) {
        void* s = g_stack_pool.back();
        g_stack_pool.pop_back();
        return s;
    }
    void* s = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s == MAP_FAILED) die("mmap of a fiber stack failed");
    mprotect(s, 4096, PROT_NONE);   // guard page at the low end
    return s;
}

static void yield_to_scheduler() { emu_switch(&g_fiber->sp, g_sched_sp); }

static void release(Barrier& b)
{
    b.count = 0;
    ++b.gen;
}

static void arrive_and_wait(Barrier& b)
{
    Fiber* f = g_fiber;
    if (!f) die("barrier outside a kernel");
    const unsigned g = b.gen;
    if (++b.count >= b.expected) {
        release(b);
        return;
    }
    f->wait_bar = &b;
    f->wait_gen = g;
    yield_to_scheduler();
}

static void leave(Barrier& b)
{
    --b.expected;
    if (b.expected > 0 && b.count >= b.expected) release(b);
}

static void fiber_main()
{
    Fiber* f = g_fiber;
    (*g_group->body)();
    f->done = true;
    f->warp->alive[f->lane] = false;
    leave(f->warp->bar);
    leave(f->cta->bar);
    leave(g_group->bar);
    yield_to_scheduler();
    die("a finished fiber was resumed");
}

void block_barrier() { arrive_and_wait(g_fiber->cta->bar); }
void warp_barrier() { arrive_and_wait(g_fiber->warp->bar); }
void cluster_barrier() { arrive_and_wait(g_group->bar); }

void shfl_bytes(void* value, size_t bytes, int rel)
{
    Fiber* f = g_fiber;
    if (bytes > 16) die("shuffle of more than 16 bytes");
    Warp& w = *f->warp;
    const int p = f->parity;
    std::memcpy(w.buf[p][f->lane], value, bytes);
    arrive_and_wait(w.bar);
    const int src = f->lane + rel;
    // (a source lane that has exited leaves its last value behind -- undefined on the hardware; no kernel here does
    //  that, and the lane may legitimately have exited AFTER this exchange but before this fiber resumed)
    if (src >= 0 && src < 32) std::memcpy(value, w.buf[p][src], bytes);
    f->parity = p ^ 1;
}

unsigned char* cluster_smem(unsigned rank)
{
    if (!g_group || rank >= g_group->ctas.size()) die("map_shared_rank: rank outside the cluster");
    return g_group->ctas[rank].smem;
}

// ---------------------------------------------------------------------------------------------------------
// kernel launches
// ---------------------------------------------------------------------------------------------------------
struct FuncAttr { int max_dyn_smem = 48 * 1024; bool nonportable_cluster = false; };
static std::map<const void*, FuncAttr> g_func_attr;
constexpr size_t kMaxSmemPerCta = 227 * 1024;

void func_attr(const void* fn, cudaFuncAttribute a, int v)
{
    if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) {
        if (v < 0 || (size_t)v > kMaxSmemPerCta) { set_error(cudaErrorInvalidValue); return; }
        g_func_attr[fn].max_dyn_smem = v;
    } else if (a == cudaFuncAttributeNonPortableClusterSizeAllowed) {
        g_func_attr[fn].nonportable_cluster = v != 0;
    }
}

static long long g_kernels_run = 0;

static void run_group(dim3 grid, dim3 block, size_t smem, unsigned first_cta, unsigned ncta, const std::function<void()>& body)
{
    const unsigned nthr = block.x * block.y * block.z;
    const unsigned nwarp = (nthr + 31) / 32;
    Group grp;
    grp.body = &body;
    grp.ctas.resize(ncta);
    grp.fibers.resize((size_t)ncta * nthr);
    grp.bar.expected = (int)(ncta * nthr);
    std::vector<unsigned char*> smem_bufs(ncta);
    for (unsigned c = 0; c < ncta; ++c) {
        Cta& cta = grp.ctas[c];
        cta.bar.expected = (int)nthr;
        cta.warps.resize(nwarp);
        const size_t sb = std::max<size_t>(smem, 16);
        if (posix_memalign((void**)&cta.smem, 128, sb) != 0) die("shared memory allocation failed");
        std::memset(cta.smem, 0xFF, sb);   // NaN poison: nothing may depend on stale shared memory
        const unsigned lin = first_cta + c;
        for (unsigned t = 0; t < nthr; ++t) {
            Fiber& f = grp.fibers[(size_t)c * nthr + t];
            f.ctx.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            f.ctx.bid = uint3{lin % grid.x, (lin / grid.x) % grid.y, lin / (grid.x * grid.y)};
            f.ctx.bdim = block;
            f.ctx.gdim = grid;
            f.ctx.smem = cta.smem;
            f.ctx.cta_rank = c;
            f.ctx.cluster_size = ncta;
            f.cta = &cta;
            f.warp = &cta.warps[t / 32];
            f.lane = (int)(t % 32);
            f.warp->bar.expected++;
            f.warp->alive[f.lane] = true;
            f.stack = stack_get();
            // initial frame for emu_switch: [mxcsr|fcw][r15][r14][r13][r12][rbx][rbp][return address]
            uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
            void** sp = (void**)(top - 16);          // return-address slot, 16-byte aligned => callee sees rsp % 16 == 8
            sp[0] = (void*)&fiber_main;
            for (int k = 1; k <= 6; ++k) sp[-k] = nullptr;
            unsigned csr[2] = {0x1F80u, 0x037Fu};    // default MXCSR, default x87 control word
            std::memcpy(&sp[-7], csr, 8);
            f.sp = (void*)&sp[-7];
        }
    }
    Group* saved_group = g_group;
    g_group = &grp;
    size_t live = grp.fibers.size();
    // CUDA_EMU_ORDER = forward (default) | reverse | random: the order in which runnable threads are resumed between
    // barriers.  A kernel without data races gives the same bits under every order; a missing barrier does not.
    static const int order_mode = []() {
        const char* e = getenv("CUDA_EMU_ORDER");
        return !e ? 0 : (!std::strcmp(e, "reverse") ? 1 : (!std::strcmp(e, "random") ? 2 : 0));
    }();
    static unsigned long long rng_state = 0x9E3779B97F4A7C15ULL;
    std::vector<unsigned> order(grp.fibers.size());
    for (unsigned i = 0; i < order.size(); ++i) order[i] = order_mode == 1 ? (unsigned)(order.size() - 1 - i) : i;
    while (live > 0) {
        bool progress = false;
        if (order_mode == 2)
            for (size_t i = order.size(); i > 1; --i) {   // Fisher-Yates with xorshift64
                rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
                std::swap(order[i - 1], order[rng_state % i]);
            }
        for (unsigned idx : order) {
            Fiber& f = grp.fibers[idx];
            if (f.done) continue;
            if (f.wait_bar) {
                if (f.wait_bar->gen == f.wait_gen) continue;
                f.wait_bar = nullptr;
            }
            g_fiber = &f;
            g_cur = &f.ctx;
            threadIdx = f.ctx.tid;
            blockIdx = f.ctx.bid;
            blockDim = f.ctx.bdim;
            gridDim = f.ctx.gdim;
            emu_switch(&g_sched_sp, f.sp);
            progress = true;
            if (f.done) --live;
        }
        if (!progress) die("deadlock: every live thread waits at a barrier that cannot complete (divergent barrier?)");
    }
    g_fiber = nullptr;
    g_cur = nullptr;
    g_group = saved_group;
    for (Fiber& f : grp.fibers) g_stack_pool.push_back(f.stack);
    for (Cta& c : grp.ctas) free(c.smem);
}

static void run_kernel(dim3 grid, dim3 block, size_t smem, unsigned cluster, const std::function<void()>& body)
{
    const unsigned total = grid.x * grid.y * grid.z;
    for (unsigned c = 0; c < total; c += cluster) run_group(grid, block, smem, c, cluster, body);
    ++g_kernels_run;
}

// ---------------------------------------------------------------------------------------------------------
// streams, capture, graphs
// ---------------------------------------------------------------------------------------------------------
}  // namespace emu

struct emuGraph {
    std::vector<std::function<void()>> ops;
    int forked = 0;    // streams pulled into the capture and not joined back yet
    std::vector<emuCond*> conds;                        // conditional handles created for this graph (reset at launch)
    std::vector<std::shared_ptr<emuGraph>> bodies;      // bodies of conditional nodes (shared with instantiated copies)
    std::vector<cudaGraph_t> body_ptrs;                 // phGraph_out storage
};
struct emuStream {
    emuGraph* capture = nullptr;
    bool joined_capture = false;   // this stream was pulled into another stream's capture by cudaStreamWaitEvent (fork)
    bool capture_to_graph = false; // cudaStreamBeginCaptureToGraph: the graph belongs to somebody else
};
struct emuEvent {
    std::chrono::steady_clock::time_point t;
    emuStream* stream = nullptr;    // stream of the last record
    emuGraph* capture = nullptr;    // the capture that record belonged to
};

namespace emu {

void enqueue(cudaStream_t st, std::function<void()> fn)
{
    if (st && st->capture) st->capture->ops.push_back(std::move(fn));
    else fn();
}

void submit_kernel(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                   std::function<void()> body)
{
    const unsigned nthr = block.x * block.y * block.z;
    const unsigned long long nblk = (unsigned long long)grid.x * grid.y * grid.z;
    const FuncAttr fa = g_func_attr.count(fn) ? g_func_attr[fn] : FuncAttr();
    if (nthr == 0 || nthr > 1024 || nblk == 0 || grid.y > 65535 || grid.z > 65535 || nblk > 0x7fffffffULL) {
        set_error(cudaErrorInvalidConfiguration);
        return;
    }
    if (smem > (size_t)fa.max_dyn_smem || smem > kMaxSmemPerCta) {
        std::fprintf(stderr, "cuda_emu: launch with %zu bytes of dynamic shared memory, function allows %d\n", smem, fa.max_dyn_smem);
        set_error(cudaErrorInvalidValue);
        return;
    }
    if (cluster < 1 || cluster > 16 || (cluster > 8 && !fa.nonportable_cluster) || grid.x % cluster != 0 || (cluster > 1 && (grid.y != 1 || grid.z != 1))) {
        set_error(cudaErrorInvalidConfiguration);
        return;
    }
    enqueue(st, [=]() { run_kernel(grid, block, smem, cluster, body); });
}

}  // namespace emu

// ---------------------------------------------------------------------------------------------------------
// runtime API
// ---------------------------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e)
{
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        case cudaErrorStreamCaptureUnsupported: return "operation not permitted when stream is capturing";
        default: return "unspecified launch failure";
    }
}
cudaError_t cudaGetLastError()
{
    cudaError_t e = emu::g_last_error;
    emu::g_last_error = cudaSuccess;
    return e;
}
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "cuda_emu (CPU emulation of a B200)");
    p->multiProcessorCount = 148;
    p->sharedMemPerBlockOptin = emu::kMaxSmemPerCta;
    p->totalGlobalMem = (size_t)180 << 30;
    p->major = 10;
    p->minor = 0;
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s)
{
    if (s && s->capture) return cudaErrorStreamCaptureUnsupported;
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }

// device memory: [guard page][data ... slack filled with a canary][guard page]
namespace {
struct Alloc { void* map; size_t map_bytes; size_t bytes; };
std::map<void*, Alloc> g_allocs;
constexpr size_t kPage = 4096;
constexpr unsigned char kCanary = 0xA5;
}
cudaError_t cudaMalloc(void** p, size_t bytes)
{
    const size_t data = (std::max<size_t>(bytes, 1) + kPage - 1) / kPage * kPage;
    const size_t total = data + 2 * kPage;
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
    unsigned char* base = (unsigned char*)m + kPage;
    std::memset(base, 0xFF, bytes);                      // NaN poison: device memory is not zero-initialised
    std::memset(base + bytes, kCanary, data - bytes);
    mprotect(m, kPage, PROT_NONE);
    mprotect(base + data, kPage, PROT_NONE);
    g_allocs[base] = Alloc{m, total, bytes};
    *p = base;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return cudaErrorInvalidValue;
    const Alloc a = it->second;
    const size_t data = a.map_bytes - 2 * kPage;
    const unsigned char* base = (const unsigned char*)p;
    for (size_t i = a.bytes; i < data; ++i)
        if (base[i] != kCanary) emu::die("write past the end of a device allocation detected at cudaFree");
    munmap(a.map, a.map_bytes);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes)
{
    *p = nullptr;
    return posix_memalign(p, 64, std::max<size_t>(bytes, 64)) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }

cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t s)
{
    emu::enqueue(s, [=]() { std::memset(p, value, bytes); });
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t s)
{
    emu::enqueue(s, [=]() { std::memmove(dst, src, bytes); });
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind)
{
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t s)
{
    if (width > dpitch || width > spitch) return cudaErrorInvalidValue;
    emu::enqueue(s, [=]() {
        for (size_t r = 0; r < height; ++r) std::memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    });
    return cudaSuccess;
}

cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    e->t = std::chrono::steady_clock::now();
    e->stream = s;
    e->capture = s ? s->capture : nullptr;   // inside a capture the record is a dependency edge, not a timestamp
    return cudaSuccess;
}
// Streams execute synchronously, so outside a capture a wait has nothing to do.  Inside a capture it is the fork /
// join of CUDA's cross-stream capture: waiting on an event recorded in a capturing stream pulls the waiting stream
// into that capture; when the origin stream later waits on an event recorded in the pulled-in stream, that stream
// leaves the capture again.  Work enqueued on a pulled-in stream is recorded into the same graph, in issue order.
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned)
{
    if (!s || !e) return cudaErrorInvalidValue;
    if (e->capture && !s->capture) {          // fork
        s->capture = e->capture;
        s->joined_capture = true;
        ++e->capture->forked;
    } else if (e->capture && s->capture == e->capture && e->stream && e->stream != s && e->stream->joined_capture) {   // join
        e->stream->capture = nullptr;
        e->stream->joined_capture = false;
        --e->capture->forked;
    }
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode)
{
    if (!s || s->capture) return cudaErrorInvalidValue;
    s->capture = new emuGraph();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g)
{
    if (!s || !s->capture || s->joined_capture || s->capture->forked != 0) {   // unjoined work in another stream
        if (g) *g = nullptr;
        return cudaErrorInvalidValue;
    }
    if (g) *g = s->capture;
    s->capture = nullptr;
    s->capture_to_graph = false;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long)
{
    *e = new emuGraph(*g);
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s)
{
    if (s && s->capture) return cudaErrorStreamCaptureUnsupported;
    for (emuCond* c : e->conds)
        if (c->assign_default) c->value = c->dflt;
    for (auto& op : e->ops) op();
    return cudaSuccess;
}

// ---- conditional WHILE nodes ----
cudaError_t cudaGraphConditionalHandleCreate(cudaGraphConditionalHandle* h, cudaGraph_t g, unsigned dflt, unsigned flags)
{
    if (!h || !g) return cudaErrorInvalidValue;
    emuCond* c = new emuCond();          // lives as long as the process: instantiated copies keep using it
    c->dflt = dflt;
    c->value = dflt;
    c->assign_default = (flags & cudaGraphCondAssignDefault) != 0;
    g->conds.push_back(c);
    *h = c;
    return cudaSuccess;
}
cudaError_t cudaStreamGetCaptureInfo(cudaStream_t s, cudaStreamCaptureStatus* status, unsigned long long* id, cudaGraph_t* graph,
                                     const cudaGraphNode_t** deps, size_t* ndeps)
{
    if (!s || !status) return cudaErrorInvalidValue;
    *status = s->capture ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    if (id) *id = 1;
    if (graph) *graph = s->capture;
    if (deps) *deps = nullptr;
    if (ndeps) *ndeps = 0;
    return cudaSuccess;
}
cudaError_t cudaStreamIsCapturing(cudaStream_t s, cudaStreamCaptureStatus* status)
{
    if (!status) return cudaErrorInvalidValue;
    *status = (s && s->capture) ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    return cudaSuccess;
}
cudaError_t cudaGraphAddNode(cudaGraphNode_t* node, cudaGraph_t g, const cudaGraphNode_t*, size_t, cudaGraphNodeParams* p)
{
    if (!g || !p || p->type != cudaGraphNodeTypeConditional || p->conditional.type != cudaGraphCondTypeWhile || p->conditional.size != 1)
        return cudaErrorInvalidValue;
    auto body = std::make_shared<emuGraph>();
    g->bodies.push_back(body);
    g->body_ptrs.push_back(body.get());
    emuCond* c = p->conditional.handle;
    g->ops.push_back([body, c]() {               // in issue order: everything captured before the node has run
        int guard = 0;
        while (c->value) {
            for (auto& op : body->ops) op();
            if (++guard > 1000000) break;         // a body that never clears its condition
        }
    });
    p->conditional.phGraph_out = &g->body_ptrs.back();
    if (node) *node = (cudaGraphNode_t)body.get();
    return cudaSuccess;
}
cudaError_t cudaStreamUpdateCaptureDependencies(cudaStream_t s, cudaGraphNode_t*, size_t, unsigned)
{
    return (s && s->capture) ? cudaSuccess : cudaErrorInvalidValue;   // closures replay in issue order anyway
}
cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t s, cudaGraph_t g, const cudaGraphNode_t*, const cudaGraphEdgeData*, size_t,
                                          cudaStreamCaptureMode)
{
    if (!s || s->capture || !g) return cudaErrorInvalidValue;
    s->capture = g;
    s->capture_to_graph = true;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------
// NCCL stand-in: one process per rank, messages are files in a directory named by the unique id
// ---------------------------------------------------------------------------------------------------------
namespace emu_nccl {

struct Comm {
    std::string dir;
    int rank = 0, world = 1;
    std::map<int, long long> send_seq, recv_seq;
    long long gather_seq = 0;
};

static long long g_sends = 0, g_gathers = 0;

static std::string base_dir()
{
    const char* e = getenv("MGB200_EMU_DIR");
    return e && *e ? e : "/tmp";
}

int GetUniqueId(void* id128)
{
    std::memset(id128, 0, 128);
    static int counter = 0;
    const auto now = std::chrono::steady_clock::now().time_since_epoch().count();
    std::snprintf((char*)id128, 128, "mgbemu_%d_%d_%llx", (int)getpid(), counter++, (unsigned long long)now);
    return 0;
}

int CommInitRank(Comm** c, int world, const void* id128, int rank)
{
    char name[129];
    std::memcpy(name, id128, 128);
    name[128] = 0;
    if (std::strncmp(name, "mgbemu_", 7) != 0) return 1;
    Comm* x = new Comm();
    x->dir = base_dir() + "/" + name;
    x->rank = rank;
    x->world = world;
    mkdir(x->dir.c_str(), 0700);   // EEXIST is fine
    *c = x;
    return 0;
}

int CommDestroy(Comm* c)
{
    if (c) {
        rmdir(c->dir.c_str());   // succeeds for the last rank out (directory empty)
        delete c;
    }
    return 0;
}

static int put(const std::string& path, const void* buf, size_t bytes)
{
    const std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) {
        // a rank that finished early removes the (momentarily empty) directory in CommDestroy while slower ranks still
        // exchange messages among themselves: re-create it
        mkdir(path.substr(0, path.rfind('/')).c_str(), 0700);
        f = std::fopen(tmp.c_str(), "wb");
    }
    if (!f) return 2;
    const size_t w = bytes ? std::fwrite(buf, 1, bytes, f) : 0;
    std::fclose(f);
    if (w != bytes) return 2;
    return std::rename(tmp.c_str(), path.c_str()) == 0 ? 0 : 2;
}

static int get(const std::string& path, void* buf, size_t bytes)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (f) {
            const size_t r = bytes ? std::fread(buf, 1, bytes, f) : 0;
            std::fclose(f);
            unlink(path.c_str());
            return r == bytes ? 0 : 3;
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) return 4;
        usleep(200);
    }
}

int Send(const void* buf, size_t bytes, int peer, Comm* c, cudaStream_t s)
{
    emu::enqueue(s, [=]() {
        ++g_sends;
        const long long q = c->send_seq[peer]++;
        if (put(c->dir + "/p2p_" + std::to_string(c->rank) + "_" + std::to_string(peer) + "_" + std::to_string(q), buf, bytes))
            emu::set_error(cudaErrorLaunchFailure);
    });
    return 0;
}

int Recv(void* buf, size_t bytes, int peer, Comm* c, cudaStream_t s)
{
    emu::enqueue(s, [=]() {
        const long long q = c->recv_seq[peer]++;
        if (get(c->dir + "/p2p_" + std::to_string(peer) + "_" + std::to_string(c->rank) + "_" + std::to_string(q), buf, bytes))
            emu::set_error(cudaErrorLaunchFailure);
    });
    return 0;
}

int AllGather(const void* send, void* recv, size_t bytes, Comm* c, cudaStream_t s)
{
    emu::enqueue(s, [=]() {
        ++g_gathers;
        const long long q = c->gather_seq++;
        std::vector<unsigned char> mine((const unsigned char*)send, (const unsigned char*)send + bytes);   // in-place safe
        for (int d = 0; d < c->world; ++d)
            if (d != c->rank &&
                put(c->dir + "/ag_" + std::to_string(q) + "_" + std::to_string(c->rank) + "_" + std::to_string(d), mine.data(), bytes))
                emu::set_error(cudaErrorLaunchFailure);
        std::memcpy((char*)recv + (size_t)c->rank * bytes, mine.data(), bytes);
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank &&
                get(c->dir + "/ag_" + std::to_string(q) + "_" + std::to_string(r) + "_" + std::to_string(c->rank),
                    (char*)recv + (size_t)r * bytes, bytes))
                emu::set_error(cudaErrorLaunchFailure);
    });
    return 0;
}

const char* GetErrorString(int r)
{
    switch (r) {
        case 0: return "ok";
        case 1: return "bad unique id";
        default: return "emulated NCCL failure";
    }
}

}  // namespace emu_nccl

extern "C" long long cuda_emu_nccl_sends() { return emu_nccl::g_sends; }
extern "C" long long cuda_emu_nccl_allgathers() { return emu_nccl::g_gathers; }
extern "C" long long cuda_emu_kernels_run() { return emu::g_kernels_run; }
extern "C" int cuda_emu_live_allocations() { return (int)g_allocs.size(); }
