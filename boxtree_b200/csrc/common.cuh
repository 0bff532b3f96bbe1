// Shared helpers for the boxtree_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define BT_OK 0
#define BT_ERR_BAD_ARG 10001
#define BT_ERR_UNSUPPORTED 10002

#define BT_CHECK(call)                                   \
    do {                                                 \
        cudaError_t bt_e_ = (call);                      \
        if (bt_e_ != cudaSuccess) return (int)bt_e_;     \
    } while (0)
#define BT_LAUNCH_CHECK()                                \
    do {                                                 \
        ++bt::g_launch_count;                            \
        BT_CHECK(cudaGetLastError());                    \
    } while (0)
#define BT_TRY(expr)                                     \
    do {                                                 \
        int bt_r_ = (expr);                              \
        if (bt_r_ != 0) return bt_r_;                    \
    } while (0)

namespace bt {

// ---- instrumentation (bench.py): kernel launch counter and optional per-scope
// CUDA-event timing on the launching stream.  Defined in tree_build.cu.
extern long long g_launch_count;
extern int g_prof_enabled;
int prof_begin(const char* name, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);

struct ProfScope {
    int slot; cudaStream_t s;
    ProfScope(const char* name, cudaStream_t st) : slot(-1), s(st)
    { if (g_prof_enabled) slot = prof_begin(name, st); }
    ~ProfScope() { if (slot >= 0) prof_end(slot, s); }
};
#define BT_PROF(name, stream) bt::ProfScope bt_prof_scope_(name, stream)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// Stream-ordered scratch memory.  The default pool's release threshold is raised once so
// that scratch survives the host synchronisations of the level loop instead of being
// returned to the OS (and re-mapped) at every sync.
static inline cudaError_t temp_alloc(void** p, size_t bytes, cudaStream_t s)
{
    // per device; bounded (2 GiB) so that scratch does not starve torch's caching allocator
    static unsigned long long configured_devices = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && !((configured_devices >> (dev & 63)) & 1ull)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            // scratch kept across synchronisations: 8 GiB unless BT_MEMPOOL_KEEP_GB says otherwise
            // (0 = everything)
            unsigned long long thr = 8ull << 30;
            if (const char* e = getenv("BT_MEMPOOL_KEEP_GB")) {
                const long long gb = atoll(e);
                thr = gb > 0 ? (unsigned long long)gb << 30 : ~0ull;
            }
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        configured_devices |= 1ull << (dev & 63);
    }
    return cudaMallocAsync(p, bytes ? bytes : 16, s);
}

// Grid for a grid-stride loop over n items: enough blocks to cover n, capped at
// a multiple of the SM count so the tail wave stays balanced.
static inline int grid_for(int64_t n, int block, int blocks_per_sm = 8)
{
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Grid for a grid-stride loop run as ONE resident wave: the blocks the kernel can keep resident
// (occupancy of this kernel at this block size x SM count), or fewer when n is small.  A fixed
// "8 blocks per SM" leaves a partial second wave behind whenever registers allow fewer
// (46 registers x 256 threads: 5 resident of 8 launched = 1.6 waves, i.e. 2 rounds of work for 1.6).
// Used where the items cost about the same (heavy-row expansion and extraction: 1.73 -> 1.55 ms,
// 0.93 -> 0.57 ms with the compacting extract); the row walks, whose rows differ widely, measured
// faster with a second wave of blocks.
template <class K>
static inline int grid_resident(K kernel, int64_t n, int block, size_t dyn_smem = 0)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem) != cudaSuccess
        || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 4;
    }
    return grid_for(n, block, per_sm);
}

template <typename T> struct CoordTraits;
template <> struct CoordTraits<float> {
    __host__ __device__ static float maxval() { return 3.402823466e+38f; }
    __host__ __device__ static float eps() { return 1.1920928955078125e-07f; }
};
template <> struct CoordTraits<double> {
    __host__ __device__ static double maxval() { return 1.7976931348623157e+308; }
    __host__ __device__ static double eps() { return 2.220446049250313e-16; }
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// streaming (read-once) loads / stores that bypass L1 allocation
__device__ __forceinline__ unsigned long long ld_stream_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_stream_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

}  // namespace bt
