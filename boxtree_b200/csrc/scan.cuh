// Single-pass exclusive prefix sum with decoupled look-back (chained scan).
//
// Replaces every pyopencl GenericScanKernel the reference uses on this path
// (tree_build_kernels.py:1555,1578,1697,1724,1770) and the count->starts scan
// inside ListOfListsBuilder (traversal.py:1854,1950).  The input is a functor
// evaluated exactly once per element (it may have side effects, like the
// reference's input_expr), the output functor receives (i, exclusive prefix).
//
// Tiles take a dynamic ticket so a tile never waits on one that is not yet
// resident.  Tile descriptors pack {status:2, value:62} into one 64-bit word,
// so publishing a descriptor is a single store and needs no fence.
#pragma once
#include "common.cuh"

namespace bt {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

constexpr unsigned long long kScanInvalid = 0ull;
constexpr unsigned long long kScanAggregate = 1ull << 62;
constexpr unsigned long long kScanPrefix = 2ull << 62;
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Look back over the descriptors of tiles [0, tile) and return their total.
// Called by one full warp.  desc[t] must have been zero-initialised.
__device__ __forceinline__ long long lookback_exclusive(unsigned long long* desc, int tile)
{
    const int lane = threadIdx.x & 31;
    long long excl = 0;
    int look = tile - 1;
    while (true) {
        const int idx = look - lane;
        unsigned long long s = (idx >= 0) ? ld_desc(desc + idx) : kScanPrefix;
        while (__any_sync(0xffffffffu, (s >> 62) == 0)) {
            if ((s >> 62) == 0) s = ld_desc(desc + idx);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const int first = pm ? (__ffs(pm) - 1) : 32;
        long long v = (lane <= first) ? (long long)(s & kScanValueMask) : 0;
        excl += warp_sum_ll(v);
        if (pm) break;
        look -= 32;
    }
    return excl;
}

template <class InOp, class OutOp>
__global__ void __launch_bounds__(kScanBlock)
scan_kernel(int64_t n_upper, const int* __restrict__ n_dev, InOp in, OutOp out,
            unsigned long long* __restrict__ desc, unsigned* __restrict__ ticket)
{
    __shared__ int s_vals[kScanTile];
    __shared__ long long s_warp[kScanBlock / 32];
    __shared__ long long s_tile_prefix;
    __shared__ int s_tile;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    int64_t n = n_upper;
    if (n_dev) { int64_t nd = *n_dev; n = nd < n ? nd : n; }
    const int64_t base = (int64_t)tile * kScanTile;
    if (base >= n) {
        if (tile == 0 && t == 0) out.total(0);
        return;
    }

#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k * kScanBlock + t;
        s_vals[k * kScanBlock + t] = (i < n) ? in(i) : 0;
    }
    __syncthreads();

    int v[kScanItems];
    {
        const int4* p = reinterpret_cast<const int4*>(s_vals + t * kScanItems);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q) {
            int4 x = p[q];
            v[4 * q + 0] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    }
    long long tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];

    // inclusive warp scan of per-thread sums
    long long inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();

    if (warp == 0) {
        long long w = (lane < kScanBlock / 32) ? s_warp[lane] : 0;
        long long winc = w;
#pragma unroll
        for (int o = 1; o < kScanBlock / 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        if (lane < kScanBlock / 32) s_warp[lane] = winc - w;   // exclusive warp offsets
        const long long agg = __shfl_sync(0xffffffffu, winc, kScanBlock / 32 - 1);
        if (lane == 0)
            st_desc(desc + tile, (tile == 0 ? kScanPrefix : kScanAggregate)
                                     | ((unsigned long long)agg & kScanValueMask));
        long long excl = 0;
        if (tile > 0) {
            excl = lookback_exclusive(desc, tile);
            if (lane == 0)
                st_desc(desc + tile, kScanPrefix | ((unsigned long long)(excl + agg) & kScanValueMask));
        }
        if (lane == 0) {
            s_tile_prefix = excl;
            if (base + kScanTile >= n) out.total(excl + agg);
        }
    }
    __syncthreads();

    long long running = s_tile_prefix + s_warp[warp] + (inc - tsum);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)t * kScanItems + k;
        if (i < n) out(i, running);
        running += v[k];
    }
}

// Exclusive scan of in(i), i in [0, n) where n = min(n_upper, *n_dev).
template <class InOp, class OutOp>
static int scan_exclusive(int64_t n_upper, const int* n_dev, InOp in, OutOp out, cudaStream_t stream)
{
    int64_t ntiles = (n_upper + kScanTile - 1) / kScanTile;
    if (ntiles < 1) ntiles = 1;
    unsigned long long* tmp = nullptr;
    const size_t bytes = (size_t)(ntiles + 1) * sizeof(unsigned long long);
    BT_CHECK(bt::temp_alloc((void**)&tmp, bytes, stream));
    BT_CHECK(cudaMemsetAsync(tmp, 0, bytes, stream));
    unsigned* ticket = reinterpret_cast<unsigned*>(tmp + ntiles);
    scan_kernel<InOp, OutOp><<<(unsigned)ntiles, kScanBlock, 0, stream>>>(
        n_upper, n_dev, in, out, tmp, ticket);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(tmp, stream));
    return BT_OK;
}

}  // namespace bt
