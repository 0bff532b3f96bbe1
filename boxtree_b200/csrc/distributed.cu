// Distributed setup kernels: DFS box order for the work
// partition, box masks, local-particle compaction, box -> user-rank CSR and the local
// target flags.  They restate the OpenCL kernels of boxtree/distributed/partition.py and
// boxtree/distributed/local_tree.py (cited per kernel).
#include "common.cuh"
#include "scan.cuh"
#include "../../include/boxtree_b200.h"

namespace bt {

// ---- c1: get_box_ids_dfs_order, partition.py:38-57 -----------------------------------
// The reference pops an explicit stack after pushing children in Morton order, i.e. it
// visits the HIGHEST Morton child first: pre-order with reversed child order.
template <int DIM>
__global__ void dist_subtree_size_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                         const int* __restrict__ child_ids, int* __restrict__ size)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        int sz = 1;
#pragma unroll
        for (int m = 0; m < NB; ++m) { const int c = child_ids[m * aligned + b]; if (c > 0) sz += size[c]; }
        size[b] = sz;
    }
}
template <int DIM>
__global__ void dist_dfs_order_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                      const int* __restrict__ child_ids, const int* __restrict__ size,
                                      int* __restrict__ rank, int* __restrict__ dfs_order)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        const int my = (b == 0) ? 0 : rank[b];
        if (b == 0) rank[0] = 0;
        dfs_order[my] = b;
        int r = my + 1;
#pragma unroll
        for (int m = NB - 1; m >= 0; --m) {
            const int c = child_ids[m * aligned + b];
            if (c > 0) { rank[c] = r; r += size[c]; }
        }
    }
}

// ---- c2: box masks, partition.py:124-357 ------------------------------------------------
__global__ void dist_mask_from_list_kernel(int n, const int* __restrict__ list, signed char* __restrict__ mask)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mask[list[i]] = 1;
}

// ancestors of the responsible boxes: the fixpoint of add_parent_boxes (partition.py:164-194)
// equals "every strict ancestor of a responsible box"
__global__ void dist_ancestor_mask_kernel(int nboxes, const signed char* __restrict__ responsible,
                                          const int* __restrict__ parent_ids, signed char* __restrict__ anc)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        if (!responsible[b]) continue;
        int cur = b;
        while (cur != 0) {
            cur = parent_ids[cur];
            if (anc[cur]) break;          // everything above is already marked (or being marked)
            anc[cur] = 1;
        }
    }
}

// add_interaction_list_boxes, partition.py:135-162 (mask_b, optional, is OR-ed in).  Rows with
// ~1e6 entries exist, so the work is cut by ENTRIES: every warp takes a span of kSpan consecutive
// list entries, finds the row of its first entry by one binary search in `starts` and then walks
// the rows of its span, 32 coalesced entries at a time.
constexpr int kMaskSpan = 2048;
__global__ void __launch_bounds__(256)
dist_add_list_boxes_kernel(int nrows, const int* __restrict__ box_list,
                           const signed char* __restrict__ mask_a, const signed char* __restrict__ mask_b,
                           const int* __restrict__ starts, const int* __restrict__ lists,
                           signed char* __restrict__ out_mask)
{
    const int nentries = starts[nrows];
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sp = w0 * kMaskSpan; sp < nentries; sp += nw * kMaskSpan) {
        const int first = (int)sp, last = (int)((sp + kMaskSpan < nentries) ? sp + kMaskSpan : nentries);
        int lo = 0, hi = nrows;                // last row with starts[row] <= first
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= first) lo = mid; else hi = mid; }
        int row = lo, k = first;
        while (k < last) {
            int rend = starts[row + 1];
            while (rend <= k) { ++row; rend = starts[row + 1]; }       // skip empty rows
            const int e = rend < last ? rend : last;
            const int box = box_list[row];
            if (mask_a[box] || (mask_b && mask_b[box]))
                for (int q = k + lane; q < e; q += 32) out_mask[lists[q]] = 1;
            k = e;
        }
    }
}

// every entry of a list marks its box (all rows of the list are rows the mask reads)
__global__ void __launch_bounds__(256)
dist_mark_list_boxes_kernel(int64_t n, const int* __restrict__ lists, signed char* __restrict__ out_mask)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t head = ((16 - ((uintptr_t)lists & 15)) & 15) / 4;      // entries before 16-byte alignment
    const int64_t h = head < n ? head : n;
    if (tid < h) out_mask[lists[tid]] = 1;
    const int4* v = reinterpret_cast<const int4*>(lists + h);
    const int64_t nv = (n - h) / 4;
    for (int64_t i = tid; i < nv; i += stride) {
        const int4 x = v[i];
        out_mask[x.x] = 1; out_mask[x.y] = 1; out_mask[x.z] = 1; out_mask[x.w] = 1;
    }
    const int64_t tail = h + nv * 4;
    if (tid < n - tail) out_mask[lists[tail + tid]] = 1;
}

// distributed build: target bits that the rank's local flags lack although the global flags
// have them, on the rows the masks read (responsible boxes and their ancestors) -- the rows of
// the global traversal that the local traversal does not contain (partition.py:197-297 reads
// them); source bits are kept.  any_out[0] != 0 if any such row exists.
__global__ void dist_corner_flags_kernel(int nboxes, const unsigned char* __restrict__ gflags,
                                         const unsigned char* __restrict__ lflags,
                                         const signed char* __restrict__ mask_a,
                                         const signed char* __restrict__ mask_b,
                                         unsigned char* __restrict__ out_flags, int* __restrict__ any_out)
{
    const unsigned char tbits = BT_BOX_IS_TARGET_BOX | BT_BOX_HAS_TARGET_CHILD_BOXES;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const unsigned char g = gflags[b], l = lflags[b];
        unsigned char keep = 0;
        if (mask_a[b] || mask_b[b]) keep = (unsigned char)(g & ~l & tbits);
        out_flags[b] = (unsigned char)((g & ~tbits) | keep);
        if (keep) *any_out = 1;
    }
}

// ---- c3: local particles, local_tree.py:70-151, 198-284 -------------------------------
__global__ void dist_particle_mask_kernel(int nboxes, const signed char* __restrict__ box_mask,
                                          const int* __restrict__ starts, const int* __restrict__ counts_nonchild,
                                          int* __restrict__ particle_mask)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int b = w; b < nboxes; b += nw) {
        if (!box_mask[b]) continue;
        const int s = starts[b], e = s + counts_nonchild[b];
        for (int p = s + lane; p < e; p += 32) particle_mask[p] = 1;
    }
}

struct MaskScanIn {
    const int* mask;
    __device__ int operator()(int64_t i) const { return mask[i]; }
};
struct MaskScanOut {   // mask_scan_kernel, local_tree.py:94-107: scan[i + 1] = inclusive sum
    int* g2l; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { g2l[i] = (int)excl; }
    __device__ void total(long long t) const { g2l[n] = (int)t; }
};

// fetch_local_particles, local_tree.py:124-151 (+ the global index of every local particle)
template <typename T>
__global__ void __launch_bounds__(256)
dist_fetch_kernel(int dim, int64_t n, const int* __restrict__ mask, const int* __restrict__ g2l,
                  const T* p0, const T* p1, const T* p2, const T* __restrict__ radii, T* o0, T* o1, T* o2,
                  T* __restrict__ oradii, long long* __restrict__ idx_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!mask[i]) continue;
        const int des = g2l[i];
        o0[des] = p0[i];
        if (dim > 1) o1[des] = p1[i];
        if (dim > 2) o2[des] = p2[i];
        if (radii) oradii[des] = radii[i];
        idx_out[des] = i;
    }
}

__global__ void dist_local_lists_kernel(int nboxes, const signed char* __restrict__ box_mask,
                                        const int* __restrict__ g2l, const int* __restrict__ starts,
                                        const int* __restrict__ counts_nonchild,
                                        const int* __restrict__ counts_cumul, int* __restrict__ lstarts,
                                        int* __restrict__ lnonchild, int* __restrict__ lcumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int s = starts[b];
        lstarts[b] = g2l[s];
        lnonchild[b] = box_mask[b] ? counts_nonchild[b] : 0;
        lcumul[b] = g2l[s + counts_cumul[b]] - g2l[s];
    }
}

// modify_target_flags, local_tree.py:163-185
__global__ void dist_modify_target_flags_kernel(int nboxes, const int* __restrict__ tgt_nonchild,
                                                const int* __restrict__ tgt_cumul,
                                                unsigned char* __restrict__ flags)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        unsigned char f = flags[b];
        f &= (unsigned char)~BT_BOX_IS_TARGET_BOX;
        f &= (unsigned char)~BT_BOX_HAS_TARGET_CHILD_BOXES;
        if (tgt_nonchild[b]) f |= BT_BOX_IS_TARGET_BOX;
        if (tgt_nonchild[b] < tgt_cumul[b]) f |= BT_BOX_HAS_TARGET_CHILD_BOXES;
        flags[b] = f;
    }
}

// flags with the target bits kept only for boxes of mask_a | mask_b (sharded setup: the
// rows of the partial traversal); need_mask = mask_a | mask_b
__global__ void dist_restrict_target_flags_kernel(int nboxes, const unsigned char* __restrict__ flags,
                                                  const signed char* __restrict__ mask_a,
                                                  const signed char* __restrict__ mask_b,
                                                  unsigned char* __restrict__ out_flags,
                                                  signed char* __restrict__ need_mask)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const bool need = mask_a[b] || mask_b[b];
        unsigned char f = flags[b];
        if (!need) f &= (unsigned char)~(BT_BOX_IS_TARGET_BOX | BT_BOX_HAS_TARGET_CHILD_BOXES);
        out_flags[b] = f;
        need_mask[b] = need ? 1 : 0;
    }
}

// MaskCompressorKernel, 2-D case (tools.py:647-740): masks[nranks][nboxes] -> per box the
// ascending list of ranks whose mask is set
struct RankCountIn {
    const signed char* masks; int nranks; int64_t nboxes; int bitsel;
    __device__ int operator()(int64_t b) const
    {
        int c = 0;
        for (int r = 0; r < nranks; ++r) c += (masks[(int64_t)r * nboxes + b] & bitsel) ? 1 : 0;
        return c;
    }
};
struct RankCountOut {
    int* starts; int64_t n; long long* total_out;
    __device__ void operator()(int64_t b, long long excl) const { starts[b] = (int)excl; }
    __device__ void total(long long t) const { starts[n] = (int)t; *total_out = t; }
};
__global__ void dist_rank_fill_kernel(int nboxes, int nranks, int bitsel, const signed char* __restrict__ masks,
                                      const int* __restrict__ starts, int* __restrict__ lists)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        int k = starts[b];
        for (int r = 0; r < nranks; ++r) if (masks[(int64_t)r * nboxes + b] & bitsel) lists[k++] = r;
    }
}

template <typename T>
static int fetch_impl(int dim, int64_t n, const int* mask, const int* g2l, void* const* parts,
                      const void* radii, void* const* outs, void* out_radii, long long* idx_out,
                      cudaStream_t s)
{
    if (n <= 0) return BT_OK;
    dist_fetch_kernel<T><<<grid_for(n, 256, 8), 256, 0, s>>>(
        dim, n, mask, g2l, (const T*)parts[0], dim > 1 ? (const T*)parts[1] : nullptr,
        dim > 2 ? (const T*)parts[2] : nullptr, (const T*)radii, (T*)outs[0],
        dim > 1 ? (T*)outs[1] : nullptr, dim > 2 ? (T*)outs[2] : nullptr, (T*)out_radii, idx_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}


// ---- (e): particle exchange of the distributed build ---------------------------------------
// After the distributed tree build every rank holds the GLOBAL box arrays and its OWN input
// particles in tree order.  A rank's local tree (local_tree.py:198-284) needs the sources of
// its point-source boxes and the targets of its responsible boxes, wherever they live: each
// owner packs, per destination rank, one record per needed particle
//     [coords (dim) | radius (optional) | box id (i32) | index in the owner's own range (i32)]
// the records travel in ONE all_to_all, and the receiver scatters them to
//     local_start[box] + (own particles of the box held by lower ranks) + index,
// the particle's place in the global tree order restricted to the rank's boxes (the order
// construct_local_particles_and_lists produces; inside a box's own range the global order is
// rank-major because global particle ids are).  The per-(sender, box) counts that the middle
// term needs are counted from the records themselves.

// dest_bits[b] = OR over ranks r with (masks[r][b] & bitsel) != 0 of (1 << r)
__global__ void dist_mask_bits_kernel(int nboxes, int nranks, int bitsel, const signed char* __restrict__ masks,
                                      unsigned* __restrict__ bits)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        unsigned v = 0;
        for (int r = 0; r < nranks; ++r) if (masks[(int64_t)r * nboxes + b] & bitsel) v |= 1u << r;
        bits[b] = v;
    }
}

constexpr int kMaxRanks = 32;
struct DestBase { long long v[kMaxRanks + 1]; };

// box id of every local particle (the boxes' own ranges tile the local array); 8 lanes per box,
// boxes with many own particles by the whole grid afterwards
constexpr int kPboxBig = 512;
__global__ void __launch_bounds__(256)
dist_particle_box_kernel(int nboxes, const int* __restrict__ lstart, const int* __restrict__ lown,
                         int* __restrict__ pbox, int* __restrict__ big_list, int* __restrict__ nbig)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    const int ng = (gridDim.x * blockDim.x) >> 3;
    for (int b = g; b < nboxes; b += ng) {
        const int n = lown[b];
        if (n > kPboxBig) { if (gl == 0) big_list[atomicAdd(nbig, 1)] = b; continue; }
        const int s = lstart[b];
        for (int p = gl; p < n; p += 8) pbox[s + p] = b;
    }
}
__global__ void __launch_bounds__(256)
dist_particle_box_big_kernel(const int* __restrict__ lstart, const int* __restrict__ lown,
                             int* __restrict__ pbox, const int* __restrict__ big_list,
                             const int* __restrict__ nbig)
{
    const int n = *nbig;
    for (int e = 0; e < n; ++e) {
        const int b = big_list[e], s = lstart[b], c = lown[b];
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < c; p += gridDim.x * blockDim.x) pbox[s + p] = b;
    }
}

// Multisplit of the local particles by destination rank (a particle may have several):
// tiles of kPackTile consecutive particles; pass 1 counts the records per (tile, destination),
// one scan over the destination-major counts gives every tile's record offset in every
// destination's chunk, pass 2 writes the records -- in tree order inside each chunk.
constexpr int kPackBlock = 256;
constexpr int kPackIters = 4;
constexpr int kPackTile = kPackBlock * kPackIters;

__global__ void __launch_bounds__(kPackBlock)
dist_pack_count_kernel(int64_t n, int nranks, int64_t ntiles, const int* __restrict__ pbox,
                       const unsigned* __restrict__ dest_bits, int* __restrict__ tile_counts /*[nranks][ntiles]*/)
{
    __shared__ int sc[kMaxRanks];
    const int lane = threadIdx.x & 31;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < kMaxRanks) sc[threadIdx.x] = 0;
        __syncthreads();
        int mine = 0;                 // lane d of every warp accumulates destination d
        for (int it = 0; it < kPackIters; ++it) {
            const int64_t p = tile * kPackTile + it * kPackBlock + threadIdx.x;
            const unsigned bits = (p < n) ? dest_bits[pbox[p]] : 0u;
            for (int d = 0; d < nranks; ++d) {
                const unsigned m = __ballot_sync(0xffffffffu, (bits >> d) & 1u);
                if (lane == d) mine += __popc(m);
            }
        }
        if (lane < nranks && mine) atomicAdd(&sc[lane], mine);
        __syncthreads();
        if (threadIdx.x < nranks) tile_counts[(int64_t)threadIdx.x * ntiles + tile] = sc[threadIdx.x];
        __syncthreads();
    }
}

struct TileScanIn {
    const int* counts;
    __device__ int operator()(int64_t i) const { return counts[i]; }
};
struct TileScanOut {
    long long* offs; long long* dest_offsets; int64_t ntiles; int nranks;
    __device__ void operator()(int64_t i, long long excl) const
    {
        offs[i] = excl;
        if (i % ntiles == 0) dest_offsets[i / ntiles] = excl;
    }
    __device__ void total(long long t) const { dest_offsets[nranks] = t; }
};

template <typename T>
__global__ void __launch_bounds__(kPackBlock)
dist_pack_kernel(int64_t n, int nranks, int64_t ntiles, int dim, int recbytes, const int* __restrict__ pbox,
                 const unsigned* __restrict__ dest_bits, const long long* __restrict__ tile_offs,
                 const int* __restrict__ lstart, const T* c0, const T* c1, const T* c2,
                 const T* __restrict__ radii, unsigned char* __restrict__ sendbuf)
{
    __shared__ int swarp[kPackBlock / 32][kMaxRanks];     // per-warp counts of this iteration
    __shared__ long long srun[kMaxRanks];                 // next record of the tile per destination
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < nranks) srun[threadIdx.x] = tile_offs[(int64_t)threadIdx.x * ntiles + tile];
        for (int it = 0; it < kPackIters; ++it) {
            const int64_t p = tile * kPackTile + it * kPackBlock + threadIdx.x;
            int b = 0, k = 0;
            unsigned bits = 0;
            T v0 = 0, v1 = 0, v2 = 0, vr = 0;
            if (p < n) {
                b = pbox[p]; bits = dest_bits[b];
                if (bits) {
                    k = (int)(p - lstart[b]);
                    v0 = c0[p];
                    if (dim > 1) v1 = c1[p];
                    if (dim > 2) v2 = c2[p];
                    if (radii) vr = radii[p];
                }
            }
            for (int d = 0; d < nranks; ++d) {
                const unsigned m = __ballot_sync(0xffffffffu, (bits >> d) & 1u);
                if (lane == 0) swarp[warp][d] = __popc(m);
            }
            __syncthreads();
            for (int d = 0; d < nranks; ++d) {
                const unsigned m = __ballot_sync(0xffffffffu, (bits >> d) & 1u);
                if (!((bits >> d) & 1u)) continue;
                long long pos = srun[d] + __popc(m & lt);
                for (int w = 0; w < warp; ++w) pos += swarp[w][d];
                unsigned char* rec = sendbuf + pos * recbytes;
                T* c = reinterpret_cast<T*>(rec);
                c[0] = v0;
                if (dim > 1) c[1] = v1;
                if (dim > 2) c[2] = v2;
                int q = dim;
                if (radii) c[q++] = vr;
                int* tail = reinterpret_cast<int*>(c + q);
                tail[0] = b;
                tail[1] = k;
            }
            __syncthreads();
            if (threadIdx.x < nranks) {
                int t = 0;
                for (int w = 0; w < kPackBlock / 32; ++w) t += swarp[w][threadIdx.x];
                srun[threadIdx.x] += t;
            }
            __syncthreads();
        }
    }
}

// receiver, step 1: records per (sender, box of my mask); compact[b] = index of box b among
// the boxes of the mask.  The records of one sender arrive in tree order with the indices of a
// box's own range ascending, so the last record of every (sender, box) run carries the count.
template <typename T>
__global__ void __launch_bounds__(256)
dist_unpack_count_kernel(int64_t nrec, int nranks, int nfields, int recbytes, DestBase chunk,
                         const unsigned char* __restrict__ recv, const int* __restrict__ compact,
                         int nmasked, int* __restrict__ cnt /*[nranks][nmasked]*/)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nrec;
         k += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < nranks && chunk.v[s + 1] <= k) ++s;
        const int* tail = reinterpret_cast<const int*>(recv + k * recbytes + (size_t)nfields * sizeof(T));
        bool last = (k + 1 == chunk.v[s + 1]);
        if (!last) {
            const int* nxt = reinterpret_cast<const int*>(recv + (k + 1) * recbytes + (size_t)nfields * sizeof(T));
            last = nxt[0] != tail[0];
        }
        if (last) cnt[(int64_t)s * nmasked + compact[tail[0]]] = tail[1] + 1;
    }
}
// step 2: exclusive prefix over the senders, per box
__global__ void dist_unpack_prefix_kernel(int nranks, int nmasked, int* __restrict__ cnt)
{
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < nmasked; m += gridDim.x * blockDim.x) {
        int run = 0;
        for (int s = 0; s < nranks; ++s) {
            const int v = cnt[(int64_t)s * nmasked + m];
            cnt[(int64_t)s * nmasked + m] = run;
            run += v;
        }
    }
}
// step 3: scatter
template <typename T>
__global__ void __launch_bounds__(256)
dist_unpack_kernel(int64_t nrec, int nranks, int dim, int recbytes, int has_radii, DestBase chunk,
                   const unsigned char* __restrict__ recv, const int* __restrict__ compact, int nmasked,
                   const int* __restrict__ cnt, const int* __restrict__ dst_start,
                   const int* __restrict__ gstart, T* o0, T* o1, T* o2, T* __restrict__ oradii,
                   long long* __restrict__ idx_out)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nrec;
         k += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < nranks && chunk.v[s + 1] <= k) ++s;
        const unsigned char* rec = recv + k * recbytes;
        const T* c = reinterpret_cast<const T*>(rec);
        const int q = dim + (has_radii ? 1 : 0);
        const int* tail = reinterpret_cast<const int*>(c + q);
        const int b = tail[0];
        const int rel = cnt[(int64_t)s * nmasked + compact[b]] + tail[1];
        const int64_t pos = (int64_t)dst_start[b] + rel;
        o0[pos] = c[0];
        if (dim > 1) o1[pos] = c[1];
        if (dim > 2) o2[pos] = c[2];
        if (has_radii) oradii[pos] = c[dim];
        idx_out[pos] = (long long)gstart[b] + rel;
    }
}

// compact index of the boxes of a mask (box-id order)
struct CompactIn {
    const signed char* mask;
    __device__ int operator()(int64_t i) const { return mask[i] ? 1 : 0; }
};
struct CompactOut {
    int* compact; int* total_out;
    __device__ void operator()(int64_t i, long long excl) const { compact[i] = (int)excl; }
    __device__ void total(long long t) const { *total_out = (int)t; }
};

// ranges of the rank's local particle arrays (local_tree.py:249-284) without a global particle
// array: the global tree order is the boxes' pre-order (own particles, then the children in
// Morton order), so the local start of a box is the prefix, in pre-order, of the own counts of
// the masked boxes before it
struct OwnScanIn {
    const signed char* mask; const int* own; const int* order;
    __device__ int operator()(int64_t i) const { const int b = order[i]; return mask[b] ? own[b] : 0; }
};
struct OwnScanOut {
    int* prefix; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { prefix[i] = (int)excl; }
    __device__ void total(long long t) const { prefix[n] = (int)t; }
};
__global__ void dist_local_ranges_kernel(int nboxes, const signed char* __restrict__ mask,
                                         const int* __restrict__ own, const int* __restrict__ rank,
                                         const int* __restrict__ subtree, const int* __restrict__ prefix,
                                         int* __restrict__ lstarts, int* __restrict__ lnonchild,
                                         int* __restrict__ lcumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int r = rank[b];
        lstarts[b] = prefix[r];
        lnonchild[b] = mask[b] ? own[b] : 0;
        lcumul[b] = prefix[r + subtree[b]] - prefix[r];
    }
}

static int pack_count_impl(int nranks, int nboxes, int64_t n, const unsigned* dest_bits, const int* lstart,
                           const int* lown, int* pbox, int* tile_counts, long long* tile_offs,
                           long long* dest_offsets, cudaStream_t s)
{
    const int64_t ntiles = (n + kPackTile - 1) / kPackTile;
    if (n <= 0) return (int)cudaMemsetAsync(dest_offsets, 0, sizeof(long long) * (nranks + 1), s);
    int* big = nullptr;
    BT_CHECK(temp_alloc((void**)&big, sizeof(int) * ((size_t)n / kPboxBig + 2), s));
    BT_CHECK(cudaMemsetAsync(big, 0, sizeof(int), s));
    dist_particle_box_kernel<<<grid_for((int64_t)nboxes * 8, 256, 8), 256, 0, s>>>(
        nboxes, lstart, lown, pbox, big + 1, big);
    BT_LAUNCH_CHECK();
    dist_particle_box_big_kernel<<<kNumSMs * 4, 256, 0, s>>>(lstart, lown, pbox, big + 1, big);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(big, s));
    dist_pack_count_kernel<<<(unsigned)(ntiles < kNumSMs * 8 ? ntiles : kNumSMs * 8), kPackBlock, 0, s>>>(
        n, nranks, ntiles, pbox, dest_bits, tile_counts);
    BT_LAUNCH_CHECK();
    TileScanIn in{tile_counts};
    TileScanOut out{tile_offs, dest_offsets, ntiles, nranks};
    return scan_exclusive((int64_t)nranks * ntiles, nullptr, in, out, s);
}

template <typename T>
static int pack_impl(int nranks, int dim, int64_t n, const int* pbox, const unsigned* dest_bits,
                     const long long* tile_offs, void* const* parts, const void* radii,
                     const int* lstart, void* sendbuf, cudaStream_t s)
{
    if (n <= 0) return BT_OK;
    const int recbytes = (int)sizeof(T) * (dim + (radii ? 1 : 0)) + 8;
    const int64_t ntiles = (n + kPackTile - 1) / kPackTile;
    dist_pack_kernel<T><<<(unsigned)(ntiles < kNumSMs * 8 ? ntiles : kNumSMs * 8), kPackBlock, 0, s>>>(
        n, nranks, ntiles, dim, recbytes, pbox, dest_bits, tile_offs, lstart, (const T*)parts[0],
        dim > 1 ? (const T*)parts[1] : nullptr, dim > 2 ? (const T*)parts[2] : nullptr,
        (const T*)radii, (unsigned char*)sendbuf);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T>
static int unpack_impl(int nranks, int dim, int64_t nrec, int has_radii, const void* recvbuf,
                       const long long* chunk_host, const int* compact, int nmasked, int* cnt,
                       const int* dst_start, const int* gstart, void* const* outs, void* oradii,
                       long long* idx_out, cudaStream_t s)
{
    const int nfields = dim + (has_radii ? 1 : 0);
    const int recbytes = (int)sizeof(T) * nfields + 8;
    DestBase chunk;
    for (int d = 0; d <= kMaxRanks; ++d) chunk.v[d] = d <= nranks ? chunk_host[d] : 0;
    BT_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nranks * (nmasked > 0 ? nmasked : 1), s));
    if (nrec <= 0) return BT_OK;
    const int grid = grid_for(nrec, 256, 8);
    dist_unpack_count_kernel<T><<<grid, 256, 0, s>>>(nrec, nranks, nfields, recbytes, chunk,
                                                     (const unsigned char*)recvbuf, compact, nmasked, cnt);
    BT_LAUNCH_CHECK();
    dist_unpack_prefix_kernel<<<grid_for(nmasked, 256), 256, 0, s>>>(nranks, nmasked, cnt);
    BT_LAUNCH_CHECK();
    dist_unpack_kernel<T><<<grid, 256, 0, s>>>(
        nrec, nranks, dim, recbytes, has_radii, chunk, (const unsigned char*)recvbuf, compact, nmasked,
        cnt, dst_start, gstart, (T*)outs[0], dim > 1 ? (T*)outs[1] : nullptr,
        dim > 2 ? (T*)outs[2] : nullptr, (T*)oradii, idx_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // namespace bt

extern "C" {

int bt_dist_dfs_order(int dim, int nboxes, int aligned_nboxes, int nlevels,
                      const int32_t* level_start_box_nrs, const int32_t* box_child_ids,
                      int32_t* subtree_size, int32_t* rank_tmp, int32_t* dfs_order, void* stream)
{
    BT_PROF("bt_dist_dfs_order", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    const int grid = bt::grid_for(nboxes, 256, 4);
    for (int lev = nlevels - 1; lev >= 0; --lev) {
        if (dim == 1) bt::dist_subtree_size_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else if (dim == 2) bt::dist_subtree_size_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else bt::dist_subtree_size_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        BT_LAUNCH_CHECK();
    }
    for (int lev = 0; lev < nlevels; ++lev) {
        if (dim == 1) bt::dist_dfs_order_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        else if (dim == 2) bt::dist_dfs_order_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        else bt::dist_dfs_order_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

// ---- work partition for the default cost 1 + own sources + own targets (partition.py:81-116) --
// The reference accumulates the costs in depth-first order and cuts where the running sum first
// exceeds (k + 1) * total / size.  With integer costs the running sums are exact in int64 as in
// float64 (any summation order), so one look-back scan of the gathered costs + a binary search
// per cut give the reference's cut positions; thresholds in the reference's float64 expression.
namespace bt {
struct PartCostIn {
    const int* order; const int* nsrc; const int* ntgt;
    __device__ int operator()(int64_t i) const { const int b = order[i]; return 1 + nsrc[b] + ntgt[b]; }
};
struct PartPrefixOut {
    long long* excl; long long* tot;
    __device__ void operator()(int64_t i, long long e) const { excl[i] = e; }
    __device__ void total(long long t) const { *tot = t; }
};
__global__ void dist_partition_cuts_kernel(int nboxes, int nranks, const long long* __restrict__ excl,
                                           const long long* __restrict__ tot, long long* __restrict__ cuts)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;       // cut between segments k-1 and k
    if (k >= nranks) return;
    const double total = (double)*tot;
    const double thr = (double)k * total / (double)nranks;
    // first i with (inclusive running sum of i) > thr; inclusive(i) = excl[i + 1], inclusive(n-1) = total
    int lo = 0, hi = nboxes;                                        // answer in [lo, hi]
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        const double inc = (double)(mid + 1 < nboxes ? excl[mid + 1] : *tot);
        if (inc > thr) hi = mid; else lo = mid + 1;
    }
    cuts[k - 1] = lo;
    if (k == 1) cuts[nranks - 1] = *tot;
}
}  // namespace bt

int bt_dist_partition_cuts(int nboxes, int nranks, const int32_t* dfs_order,
                           const int32_t* box_source_counts_nonchild,
                           const int32_t* box_target_counts_nonchild, int64_t* cuts, void* stream)
{
    BT_PROF("bt_dist_partition_cuts", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0 || nranks < 1) return BT_ERR_BAD_ARG;
    long long* tmp = nullptr;
    BT_CHECK(bt::temp_alloc((void**)&tmp, sizeof(long long) * ((size_t)nboxes + 1), s));
    bt::PartCostIn in{dfs_order, box_source_counts_nonchild, box_target_counts_nonchild};
    bt::PartPrefixOut out{tmp, tmp + nboxes};
    BT_TRY(bt::scan_exclusive(nboxes, nullptr, in, out, s));
    bt::dist_partition_cuts_kernel<<<(nranks + 63) / 64, 64, 0, s>>>(nboxes, nranks, tmp, tmp + nboxes,
                                                                     (long long*)cuts);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(tmp, s));
    return BT_OK;
}

int bt_dist_mask_from_list(int n, const int32_t* list, int8_t* mask, void* stream)
{
    BT_PROF("bt_dist_mask_from_list", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::dist_mask_from_list_kernel<<<bt::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, list, (signed char*)mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_ancestor_mask(int nboxes, const int8_t* responsible, const int32_t* box_parent_ids,
                          int8_t* ancestor_mask, void* stream)
{
    BT_PROF("bt_dist_ancestor_mask", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_ancestor_mask_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)responsible, box_parent_ids, (signed char*)ancestor_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_add_list_boxes(int nrows, const int32_t* box_list, const int8_t* mask_a, const int8_t* mask_b,
                           const int32_t* starts, const int32_t* lists, int8_t* out_mask, void* stream)
{
    BT_PROF("bt_dist_add_list_boxes", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    bt::dist_add_list_boxes_kernel<<<bt::kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(
        nrows, box_list, (const signed char*)mask_a, (const signed char*)mask_b, starts, lists,
        (signed char*)out_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_particle_mask(int nboxes, const int8_t* box_mask, const int32_t* starts,
                          const int32_t* counts_nonchild, int32_t* particle_mask, void* stream)
{
    BT_PROF("bt_dist_particle_mask", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_particle_mask_kernel<<<bt::grid_for((int64_t)nboxes * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)box_mask, starts, counts_nonchild, particle_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mask_scan(int64_t n, const int32_t* particle_mask, int32_t* global_to_local, void* stream)
{
    BT_PROF("bt_dist_mask_scan", (cudaStream_t)stream);
    bt::MaskScanIn in{particle_mask};
    bt::MaskScanOut out{global_to_local, n};
    return bt::scan_exclusive(n, nullptr, in, out, (cudaStream_t)stream);
}

int bt_dist_fetch_local_particles(int dtype, int dim, int64_t n, const int32_t* particle_mask,
                                  const int32_t* global_to_local /*[n+1]*/, void* const* particles,
                                  const void* radii, void* const* local_particles, void* local_radii,
                                  int64_t* particle_idx, void* stream)
{
    BT_PROF("bt_dist_fetch_local_particles", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == BT_F32)
        return bt::fetch_impl<float>(dim, n, particle_mask, global_to_local, particles, radii,
                                     local_particles, local_radii, (long long*)particle_idx, s);
    if (dtype == BT_F64)
        return bt::fetch_impl<double>(dim, n, particle_mask, global_to_local, particles, radii,
                                      local_particles, local_radii, (long long*)particle_idx, s);
    return BT_ERR_BAD_ARG;
}

int bt_dist_local_lists(int nboxes, const int8_t* box_mask, const int32_t* global_to_local,
                        const int32_t* starts, const int32_t* counts_nonchild, const int32_t* counts_cumul,
                        int32_t* local_starts, int32_t* local_nonchild, int32_t* local_cumul, void* stream)
{
    BT_PROF("bt_dist_local_lists", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_local_lists_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)box_mask, global_to_local, starts, counts_nonchild, counts_cumul,
        local_starts, local_nonchild, local_cumul);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_modify_target_flags(int nboxes, const int32_t* tgt_nonchild, const int32_t* tgt_cumul,
                                uint8_t* box_flags, void* stream)
{
    BT_PROF("bt_dist_modify_target_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_modify_target_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, tgt_nonchild, tgt_cumul, box_flags);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_restrict_target_flags(int nboxes, const uint8_t* box_flags, const int8_t* mask_a,
                                  const int8_t* mask_b, uint8_t* out_flags, int8_t* need_mask, void* stream)
{
    BT_PROF("bt_dist_restrict_target_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_restrict_target_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, box_flags, (const signed char*)mask_a, (const signed char*)mask_b, out_flags,
        (signed char*)need_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mark_list_boxes(int64_t nentries, const int32_t* lists, int8_t* out_mask, void* stream)
{
    BT_PROF("bt_dist_mark_list_boxes", (cudaStream_t)stream);
    if (nentries <= 0) return BT_OK;
    bt::dist_mark_list_boxes_kernel<<<bt::grid_for(nentries / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nentries, lists, (signed char*)out_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_corner_flags(int nboxes, const uint8_t* global_flags, const uint8_t* local_flags,
                         const int8_t* mask_a, const int8_t* mask_b, uint8_t* out_flags, int32_t* any_out,
                         void* stream)
{
    BT_PROF("bt_dist_corner_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_corner_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, global_flags, local_flags, (const signed char*)mask_a, (const signed char*)mask_b,
        out_flags, any_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mask_bits(int nboxes, int nranks, int bitsel, const int8_t* masks_all_ranks,
                      uint32_t* dest_bits, void* stream)
{
    BT_PROF("bt_dist_mask_bits", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    if (nranks > bt::kMaxRanks) return BT_ERR_UNSUPPORTED;
    bt::dist_mask_bits_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, nranks, bitsel, (const signed char*)masks_all_ranks, dest_bits);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_pack_ntiles(int64_t n) { return (int)((n + bt::kPackTile - 1) / bt::kPackTile); }

int bt_dist_pack_count(int nranks, int nboxes, int64_t n, const uint32_t* dest_bits,
                       const int32_t* local_start, const int32_t* local_own, int32_t* particle_box,
                       int32_t* tile_counts, int64_t* tile_offsets, int64_t* dest_offsets, void* stream)
{
    BT_PROF("bt_dist_pack_count", (cudaStream_t)stream);
    if (nranks > bt::kMaxRanks) return BT_ERR_UNSUPPORTED;
    return bt::pack_count_impl(nranks, nboxes, n, dest_bits, local_start, local_own, particle_box,
                               tile_counts, (long long*)tile_offsets, (long long*)dest_offsets,
                               (cudaStream_t)stream);
}

int bt_dist_pack_records(int dtype, int nranks, int dim, int64_t n, const int32_t* particle_box,
                         const uint32_t* dest_bits, const int64_t* tile_offsets, void* const* particles,
                         const void* radii, const int32_t* local_start, void* sendbuf, void* stream)
{
    BT_PROF("bt_dist_pack_records", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nranks > bt::kMaxRanks) return BT_ERR_UNSUPPORTED;
    if (dtype == BT_F32)
        return bt::pack_impl<float>(nranks, dim, n, particle_box, dest_bits, (const long long*)tile_offsets,
                                    particles, radii, local_start, sendbuf, s);
    if (dtype == BT_F64)
        return bt::pack_impl<double>(nranks, dim, n, particle_box, dest_bits, (const long long*)tile_offsets,
                                     particles, radii, local_start, sendbuf, s);
    return BT_ERR_BAD_ARG;
}

int bt_dist_compact_index(int nboxes, const int8_t* box_mask, int32_t* compact, int32_t* nmasked_dev,
                          void* stream)
{
    BT_PROF("bt_dist_compact_index", (cudaStream_t)stream);
    bt::CompactIn in{(const signed char*)box_mask};
    bt::CompactOut out{compact, nmasked_dev};
    return bt::scan_exclusive(nboxes, nullptr, in, out, (cudaStream_t)stream);
}

int bt_dist_unpack_records(int dtype, int nranks, int dim, int64_t nrec, int has_radii, const void* recvbuf,
                           const int64_t* chunk_offsets_host, const int32_t* compact, int nmasked,
                           int32_t* count_tmp, const int32_t* dst_start, const int32_t* box_global_start,
                           void* const* local_particles, void* local_radii, int64_t* particle_idx,
                           void* stream)
{
    BT_PROF("bt_dist_unpack_records", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nranks > bt::kMaxRanks) return BT_ERR_UNSUPPORTED;
    if (dtype == BT_F32)
        return bt::unpack_impl<float>(nranks, dim, nrec, has_radii, recvbuf, (const long long*)chunk_offsets_host,
                                      compact, nmasked, count_tmp, dst_start, box_global_start,
                                      local_particles, local_radii, (long long*)particle_idx, s);
    if (dtype == BT_F64)
        return bt::unpack_impl<double>(nranks, dim, nrec, has_radii, recvbuf, (const long long*)chunk_offsets_host,
                                       compact, nmasked, count_tmp, dst_start, box_global_start,
                                       local_particles, local_radii, (long long*)particle_idx, s);
    return BT_ERR_BAD_ARG;
}

int bt_dist_local_ranges(int nboxes, const int8_t* box_mask, const int32_t* own_counts,
                         const int32_t* preorder_rank, const int32_t* preorder_boxes,
                         const int32_t* subtree_size, int32_t* prefix_tmp /*[nboxes+1]*/,
                         int32_t* local_starts, int32_t* local_nonchild, int32_t* local_cumul, void* stream)
{
    BT_PROF("bt_dist_local_ranges", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    bt::OwnScanIn in{(const signed char*)box_mask, own_counts, preorder_boxes};
    bt::OwnScanOut out{prefix_tmp, nboxes};
    BT_TRY(bt::scan_exclusive(nboxes, nullptr, in, out, s));
    bt::dist_local_ranges_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
        nboxes, (const signed char*)box_mask, own_counts, preorder_rank, subtree_size, prefix_tmp,
        local_starts, local_nonchild, local_cumul);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_box_to_user_rank_bits(int phase, int nboxes, int nranks, int bitsel, const int8_t* masks_all_ranks,
                                  int32_t* starts, int32_t* lists, int64_t* total_dev, void* stream)
{
    BT_PROF("bt_dist_box_to_user_rank", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::RankCountIn in{(const signed char*)masks_all_ranks, nranks, nboxes, bitsel};
        bt::RankCountOut out{starts, nboxes, (long long*)total_dev};
        return bt::scan_exclusive(nboxes, nullptr, in, out, s);
    }
    if (nboxes <= 0) return BT_OK;
    bt::dist_rank_fill_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
        nboxes, nranks, bitsel, (const signed char*)masks_all_ranks, starts, lists);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_box_to_user_rank(int phase, int nboxes, int nranks, const int8_t* masks_all_ranks,
                             int32_t* starts, int32_t* lists, int64_t* total_dev, void* stream)
{
    return bt_dist_box_to_user_rank_bits(phase, nboxes, nranks, 0xff, masks_all_ranks, starts, lists,
                                         total_dev, stream);
}

}  // extern "C"
