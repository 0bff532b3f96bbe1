// One-sweep LSD radix sort of (u64 key, u32 value) pairs, stable.
//
// This is the "device radix sort of particles" of the north star: it replaces
// the reference's per-level stable partition (morton_scan + renumber_particles,
// tree_build_kernels.py:247-508, 717-819, run once per level over 64-byte
// structs) by ceil(bits/8) passes over 12-byte pairs.
//
//  * one upfront kernel builds the digit histograms of every pass,
//  * each pass is ONE kernel: a tile (4096 pairs) ranks its keys with
//    warp-level match_any multisplit, publishes its 256 digit counts and
//    resolves its global offsets by decoupled look-back over earlier tiles,
//    then scatters through shared memory so global stores are coalesced runs.
//  * tiles take a dynamic ticket (forward progress), descriptors are
//    {status:2,value:30} words.
#pragma once
#include "common.cuh"

namespace bt {

constexpr int kRsBits = 8;
constexpr int kRsRadix = 1 << kRsBits;
constexpr int kRsThreads = 256;            // == kRsRadix: thread d owns digit d
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;
constexpr int kRsMaxPasses = 8;

constexpr unsigned kRsAgg = 1u << 30;
constexpr unsigned kRsPrefix = 2u << 30;
constexpr unsigned kRsValMask = (1u << 30) - 1;

struct RsSmem {
    unsigned long long keys[kRsTile];
    unsigned vals[kRsTile];
    unsigned whist[kRsWarps][kRsRadix];
    unsigned tile_base[kRsRadix];
    long long gbase[kRsRadix];
    unsigned scan_tmp[kRsWarps];
    int tile;
    unsigned long long mbar;       // mbarrier of the TMA bulk load of the tile's keys
};

// 1-D TMA: cp.async.bulk global -> shared, completion counted in bytes on an mbarrier
#ifndef BT_RS_TMA
#define BT_RS_TMA 1
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p)
{ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes,
                                            unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BT_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BT_MBAR_DONE;\n"
        "bra BT_MBAR_WAIT;\n"
        "BT_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan over the 256 threads of the block
__device__ __forceinline__ unsigned block_excl_scan_256(unsigned v, unsigned* tmp /*[8]*/)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();          // protect tmp reuse
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    unsigned woff = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) woff += (w < warp) ? tmp[w] : 0u;
    return woff + inc - v;
}

// histograms of all passes in one read of the keys
static __global__ void __launch_bounds__(256)
rs_histogram_kernel(const unsigned long long* __restrict__ keys, int64_t n, int begin_bit,
                    int end_bit, int npasses, unsigned* __restrict__ ghist)
{
    __shared__ unsigned sh[kRsMaxPasses][kRsRadix];
    for (int i = threadIdx.x; i < kRsMaxPasses * kRsRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long k = ld_stream_u64(keys + i);
        for (int p = 0; p < npasses; ++p) {
            const int shift = begin_bit + p * kRsBits;
            const int bits = min(kRsBits, end_bit - shift);
            const unsigned dg = (unsigned)(k >> shift) & ((1u << bits) - 1u);
            atomicAdd(&sh[p][dg], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npasses * kRsRadix; i += blockDim.x) {
        const unsigned c = (&sh[0][0])[i];
        if (c) atomicAdd(ghist + i, c);
    }
}

// 3 resident tiles per SM (<= 85 registers, 3 x 59 KB shared memory): measured 1.37 -> 1.00 ms
// for the 8 passes over 1e7 particles against the unconstrained 104-register build
#ifndef BT_RS_MIN_BLOCKS
#define BT_RS_MIN_BLOCKS 3
#endif
template <bool kIdentityVals, bool kKeysOnly = false>
__global__ void __launch_bounds__(kRsThreads, BT_RS_MIN_BLOCKS)
rs_onesweep_kernel(const unsigned long long* __restrict__ kin, unsigned long long* __restrict__ kout,
                   const unsigned* __restrict__ vin, unsigned* __restrict__ vout, int64_t n,
                   int shift, unsigned mask, const unsigned* __restrict__ ghist_pass,
                   unsigned* __restrict__ desc, unsigned* __restrict__ ticket)
{
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    if (t == 0) {
        sm.tile = (int)atomicAdd(ticket, 1u);
        if (BT_RS_TMA) mbar_init(&sm.mbar, 1);
    }
    for (int i = t; i < kRsWarps * kRsRadix; i += kRsThreads) (&sm.whist[0][0])[i] = 0;
    __syncthreads();
    const int tile = sm.tile;
    const int64_t base = (int64_t)tile * kRsTile;
    if (base >= n) return;
    const int cnt = (int)((n - base < kRsTile) ? (n - base) : kRsTile);

    // ---- load keys, warp-striped: warp w owns [w*512, (w+1)*512) of the tile.  A full tile whose
    //      32 KB of keys are 16-byte aligned arrives by ONE TMA bulk copy into sm.keys (which is
    //      only overwritten by the scatter after the ranking barrier); other tiles use plain loads
    unsigned long long key[kRsItems];
    unsigned short rank[kRsItems];
    const int wbase = warp * (32 * kRsItems);
    const bool by_tma = BT_RS_TMA && cnt == kRsTile &&
                        ((reinterpret_cast<uintptr_t>(kin + base) & 15u) == 0);
    if (by_tma) {
        if (t == 0) tma_load_1d(sm.keys, kin + base, kRsTile * (unsigned)sizeof(unsigned long long), &sm.mbar);
        mbar_wait(&sm.mbar, 0);
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) key[j] = sm.keys[wbase + j * 32 + lane];
    } else {
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            const int idx = wbase + j * 32 + lane;
            key[j] = (idx < cnt) ? ld_stream_u64(kin + base + idx) : ~0ull;
        }
    }

    // ---- rank inside the warp, round by round (stable)
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const unsigned dg = (unsigned)(key[j] >> shift) & mask;
        const unsigned peers = __match_any_sync(0xffffffffu, dg);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) {
            old = sm.whist[warp][dg];
            sm.whist[warp][dg] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (unsigned short)(old + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // ---- thread d: offsets of each warp inside digit d, tile count of digit d
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
        const unsigned c = sm.whist[w][t];
        sm.whist[w][t] = tot;
        tot += c;
    }

    // ---- publish + decoupled look-back for digit t
    unsigned* my_desc = desc + (size_t)tile * kRsRadix + t;
    st_relaxed_u32(my_desc, (tile == 0 ? kRsPrefix : kRsAgg) | tot);
    unsigned excl = 0;
    if (tile > 0) {
        for (int lt = tile - 1; lt >= 0; --lt) {
            const unsigned* pd = desc + (size_t)lt * kRsRadix + t;
            unsigned s;
            do { s = ld_relaxed_u32(pd); } while ((s >> 30) == 0);
            excl += s & kRsValMask;
            if ((s >> 30) == 2) break;
        }
        st_relaxed_u32(my_desc, kRsPrefix | ((excl + tot) & kRsValMask));
    }

    // ---- digit bases: inside the tile and globally
    const unsigned gcount = ghist_pass[t];
    const unsigned gexcl = block_excl_scan_256(gcount, sm.scan_tmp);
    const unsigned tbase = block_excl_scan_256(tot, sm.scan_tmp);
    sm.tile_base[t] = tbase;
    sm.gbase[t] = (long long)gexcl + (long long)excl - (long long)tbase;
    __syncthreads();

    // ---- scatter keys into tile-sorted order in shared memory
    unsigned short pos[kRsItems];
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const unsigned dg = (unsigned)(key[j] >> shift) & mask;
        pos[j] = (unsigned short)(sm.tile_base[dg] + sm.whist[warp][dg] + rank[j]);
        sm.keys[pos[j]] = key[j];
    }
    // values ride along
    if (!kKeysOnly) {
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            const int idx = wbase + j * 32 + lane;
            unsigned v = 0;
            if (idx < cnt) v = kIdentityVals ? (unsigned)(base + idx) : ld_stream_u32(vin + base + idx);
            sm.vals[pos[j]] = v;
        }
    }
    __syncthreads();

    // ---- coalesced runs to global memory
    for (int i = t; i < cnt; i += kRsThreads) {
        const unsigned long long k = sm.keys[i];
        const unsigned dg = (unsigned)(k >> shift) & mask;
        const long long g = sm.gbase[dg] + i;
        kout[g] = k;
        if (!kKeysOnly) vout[g] = sm.vals[i];
    }
}

// Sort pairs by key bits [begin_bit, end_bit).  Buffers ping-pong; returns in
// *result_in_alt whether the sorted data ended in (keys_alt, vals_alt).
// vals may be generated as the identity permutation (identity_vals, vals == nullptr on input);
// vals == vals_alt == nullptr without identity_vals sorts the keys alone.
static int radix_sort_pairs(int64_t n, unsigned long long* keys, unsigned long long* keys_alt,
                            unsigned* vals, unsigned* vals_alt, int identity_vals,
                            int begin_bit, int end_bit, int* result_in_alt, cudaStream_t stream,
                            const char* pass_scope = "rs_onesweep_pass")
{
    *result_in_alt = 0;
    if (n <= 0 || end_bit <= begin_bit) {
        if (identity_vals) return BT_ERR_BAD_ARG;   // caller must fill ids itself
        return BT_OK;
    }
    if (n >= (1ll << 30)) return BT_ERR_UNSUPPORTED;
    const int npasses = (end_bit - begin_bit + kRsBits - 1) / kRsBits;
    if (npasses > kRsMaxPasses) return BT_ERR_BAD_ARG;
    const int64_t ntiles = (n + kRsTile - 1) / kRsTile;

    const size_t hist_bytes = (size_t)npasses * kRsRadix * sizeof(unsigned);
    const size_t desc_bytes = (size_t)ntiles * kRsRadix * sizeof(unsigned) + 16;
    unsigned char* tmp = nullptr;
    BT_CHECK(bt::temp_alloc((void**)&tmp, hist_bytes + desc_bytes, stream));
    unsigned* ghist = reinterpret_cast<unsigned*>(tmp);
    unsigned* desc = reinterpret_cast<unsigned*>(tmp + hist_bytes);
    unsigned* ticket = desc + (size_t)ntiles * kRsRadix;
    BT_CHECK(cudaMemsetAsync(ghist, 0, hist_bytes, stream));
    {
    BT_PROF("rs_histogram", stream);
    rs_histogram_kernel<<<grid_for(n, 256, 4), 256, 0, stream>>>(keys, n, begin_bit, end_bit,
                                                                 npasses, ghist);
    BT_LAUNCH_CHECK();
    }

    // the attribute is per device
    static unsigned long long attr_set_devices = 0;
    int dev = 0;
    BT_CHECK(cudaGetDevice(&dev));
    if (!((attr_set_devices >> (dev & 63)) & 1ull)) {
        BT_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        BT_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        BT_CHECK(cudaFuncSetAttribute((rs_onesweep_kernel<false, true>),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        attr_set_devices |= 1ull << (dev & 63);
    }

    unsigned long long* kin = keys; unsigned long long* kout = keys_alt;
    unsigned* vin = vals; unsigned* vout = vals_alt;
    for (int p = 0; p < npasses; ++p) {
        const int shift = begin_bit + p * kRsBits;
        const int bits = (end_bit - shift < kRsBits) ? (end_bit - shift) : kRsBits;
        const unsigned mask = (1u << bits) - 1u;
        BT_CHECK(cudaMemsetAsync(desc, 0, desc_bytes, stream));
        BT_PROF(pass_scope, stream);
        if (!identity_vals && vals == nullptr)      // keys only
            rs_onesweep_kernel<false, true><<<(unsigned)ntiles, kRsThreads, sizeof(RsSmem), stream>>>(
                kin, kout, nullptr, nullptr, n, shift, mask, ghist + p * kRsRadix, desc, ticket);
        else if (p == 0 && identity_vals)
            rs_onesweep_kernel<true><<<(unsigned)ntiles, kRsThreads, sizeof(RsSmem), stream>>>(
                kin, kout, vin, vout, n, shift, mask, ghist + p * kRsRadix, desc, ticket);
        else
            rs_onesweep_kernel<false><<<(unsigned)ntiles, kRsThreads, sizeof(RsSmem), stream>>>(
                kin, kout, vin, vout, n, shift, mask, ghist + p * kRsRadix, desc, ticket);
        BT_LAUNCH_CHECK();
        unsigned long long* tk = kin; kin = kout; kout = tk;
        unsigned* tv = vin; vin = vout; vout = tv;
        *result_in_alt ^= 1;
    }
    BT_CHECK(cudaFreeAsync(tmp, stream));
    return BT_OK;
}

}  // namespace bt
