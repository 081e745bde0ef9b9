// tests/host_emul/cuda_emu/cuda_runtime.h — TEST INFRASTRUCTURE, never part of the product.
//
// A small CUDA-on-CPU emulation, just large enough to compile multigrid_nikhil_c-_b200/csrc/*.cu with g++ and
// run the WHOLE library (host orchestration + every kernel body, thread by thread) in the CPU test-suite:
//   * each CUDA thread is a fiber; __syncthreads / warp shuffles / cluster.sync are cooperative barriers
//   * kernels of one launch run CTA after CTA (cluster launches: all CTAs of the cluster together)
//   * streams are synchronous; stream capture records closures, cudaGraphLaunch replays them
//   * dynamic shared memory is poisoned with NaNs; device allocations sit between guard pages
//   * launch limits that cannot be checked without a GPU are enforced: dynamic shared memory above 48 KB
//     needs cudaFuncAttributeMaxDynamicSharedMemorySize, <= 227 KB per CTA, <= 1024 threads per CTA,
//     cluster size <= 8 (16 with cudaFuncAttributeNonPortableClusterSizeAllowed)
// tests/host_emul/build_emu.py rewrites `kernel<<<g, b, s, st>>>(args)` into emu::launch(...) and
// `extern __shared__ ... name[]` into a pointer to the CTA's buffer, then compiles with
// -ffp-contract=off (the product is built with --fmad=false), so the arithmetic is the product's.
// What this cannot show: anything about speed, occupancy, register pressure, PTX (cp.async is executed as a
// synchronous copy) or memory-model races between threads that run concurrently on a GPU.
#pragma once
#define MGB_EMU 1

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

// ---- qualifiers ----------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define EMU_NOINLINE __attribute__((noinline))   // build_emu.py rewrites __noinline__ (libstdc++ uses that token)
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static                          // static __shared__ arrays: CTAs of a launch run one after the other

// ---- vector / index types ------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

using std::max;
using std::min;

// ---- runtime types -------------------------------------------------------------------------------------
enum cudaError_t {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidConfiguration = 9,
    cudaErrorLaunchFailure = 719,
    cudaErrorStreamCaptureUnsupported = 900,
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributeNonPortableClusterSizeAllowed = 12 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4, cudaLaunchAttributeProgrammaticStreamSerialization = 6 };

struct emuStream;
struct emuEvent;
struct emuGraph;
typedef emuStream* cudaStream_t;
typedef emuEvent* cudaEvent_t;
typedef emuGraph* cudaGraph_t;
typedef emuGraph* cudaGraphExec_t;

struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
    size_t totalGlobalMem;
    int major, minor;
};
struct cudaLaunchAttributeValue { struct { unsigned x, y, z; } clusterDim; int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};

// ---- the per-thread context (what threadIdx & co. read) --------------------------------------------------
namespace emu {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    unsigned char* smem;     // this CTA's dynamic shared memory
    unsigned cta_rank;       // rank of the CTA in its cluster
    unsigned cluster_size;
};
extern ThreadCtx* g_cur;     // the running fiber's context
inline unsigned char* dyn_smem() { return g_cur->smem; }

void block_barrier();
void warp_barrier();
void cluster_barrier();
void shfl_bytes(void* value, size_t bytes, int rel);   // value of lane (own lane + rel); outside 0..31: keep own value
unsigned char* cluster_smem(unsigned rank);

// every stream operation goes through here: executed now, or recorded when the stream is capturing
void enqueue(cudaStream_t st, std::function<void()> fn);
void submit_kernel(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                   std::function<void()> body);
void func_attr(const void* fn, cudaFuncAttribute a, int v);
void set_error(cudaError_t e);

template <typename... P, typename... A>
inline void launch(void (*k)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args)
{
    std::tuple<std::decay_t<P>...> t(std::forward<A>(args)...);
    submit_kernel((const void*)k, grid, block, smem, st, 1, [k, t]() { std::apply(k, t); });
}
}  // namespace emu

// plain globals, reloaded by the scheduler every time a fiber is resumed
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- device intrinsics ---------------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int = 32)
{
    emu::shfl_bytes(&v, sizeof(T), -(int)delta);
    return v;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32)
{
    emu::shfl_bytes(&v, sizeof(T), (int)delta);
    return v;
}
// fibers of one emulated device run one at a time (cooperative scheduling), so a plain read-modify-write is atomic
template <typename T>
inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }
inline unsigned __float_as_uint(float f) { unsigned r; std::memcpy(&r, &f, 4); return r; }
inline float __uint_as_float(unsigned u) { float r; std::memcpy(&r, &u, 4); return r; }

// ---- runtime API ---------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaMalloc(void** p, size_t bytes);
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMallocHost((void**)p, bytes); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t s);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode m);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s);

// ---- conditional graph nodes (CUDA 12.4): enough for a WHILE node whose body is populated by capture-to-graph ----
struct emuCond { unsigned value = 0, dflt = 0; bool assign_default = false; };
typedef emuCond* cudaGraphConditionalHandle;
typedef void* cudaGraphNode_t;
struct cudaGraphEdgeData { int unused; };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1, cudaStreamCaptureStatusInvalidated = 2 };
enum cudaGraphNodeType { cudaGraphNodeTypeKernel = 0, cudaGraphNodeTypeConditional = 13 };
enum cudaGraphConditionalNodeType { cudaGraphCondTypeIf = 0, cudaGraphCondTypeWhile = 1 };
enum { cudaGraphCondAssignDefault = 1 };
enum { cudaStreamAddCaptureDependencies = 0, cudaStreamSetCaptureDependencies = 1 };
struct cudaConditionalNodeParams {
    cudaGraphConditionalHandle handle;
    cudaGraphConditionalNodeType type;
    unsigned size;
    cudaGraph_t* phGraph_out;
};
struct cudaGraphNodeParams {
    cudaGraphNodeType type;
    cudaConditionalNodeParams conditional;
};
cudaError_t cudaGraphConditionalHandleCreate(cudaGraphConditionalHandle* h, cudaGraph_t g, unsigned defaultLaunchValue = 0, unsigned flags = 0);
cudaError_t cudaStreamGetCaptureInfo(cudaStream_t s, cudaStreamCaptureStatus* status, unsigned long long* id = nullptr,
                                     cudaGraph_t* graph = nullptr, const cudaGraphNode_t** deps = nullptr, size_t* ndeps = nullptr);
cudaError_t cudaStreamIsCapturing(cudaStream_t s, cudaStreamCaptureStatus* status);
cudaError_t cudaGraphAddNode(cudaGraphNode_t* node, cudaGraph_t g, const cudaGraphNode_t* deps, size_t ndeps, cudaGraphNodeParams* p);
cudaError_t cudaStreamUpdateCaptureDependencies(cudaStream_t s, cudaGraphNode_t* deps, size_t ndeps, unsigned flags);
cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t s, cudaGraph_t g, const cudaGraphNode_t* deps, const cudaGraphEdgeData* data,
                                          size_t ndeps, cudaStreamCaptureMode m);
inline void cudaGraphSetConditional(cudaGraphConditionalHandle h, unsigned value) { h->value = value; }   // "device" side

template <typename F>
inline cudaError_t cudaFuncSetAttribute(F* fn, cudaFuncAttribute a, int v)
{
    emu::func_attr((const void*)fn, a, v);
    return cudaSuccess;
}

template <typename... P, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*k)(P...), A&&... args)
{
    unsigned cluster = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) {
            if (cfg->attrs[i].val.clusterDim.y != 1 || cfg->attrs[i].val.clusterDim.z != 1) return cudaErrorInvalidConfiguration;
            cluster = cfg->attrs[i].val.clusterDim.x;
        }
    std::tuple<std::decay_t<P>...> t(std::forward<A>(args)...);
    emu::submit_kernel((const void*)k, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, cluster,
                       [k, t]() { std::apply(k, t); });
    return cudaGetLastError();
}

// ---- NCCL stand-in (comm.cu binds these instead of dlopen("libnccl.so.2") in the emulation build) ---------
// ranks are separate processes; messages are files under $MGB200_EMU_DIR (default /tmp)
namespace emu_nccl {
struct Comm;
int GetUniqueId(void* id128);
int CommInitRank(Comm** c, int world, const void* id128, int rank);
int CommDestroy(Comm* c);
int Send(const void* buf, size_t bytes, int peer, Comm* c, cudaStream_t s);
int Recv(void* buf, size_t bytes, int peer, Comm* c, cudaStream_t s);
int AllGather(const void* send, void* recv, size_t bytes, Comm* c, cudaStream_t s);
const char* GetErrorString(int r);
}  // namespace emu_nccl
