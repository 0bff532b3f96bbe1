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

// add_interaction_list_boxes, partition.py:135-162 (mask_b, optional, is OR-ed in).  One
// thread per list ENTRY (its row is found by binary search in `starts`), so rows with ~1e6
// entries do not serialise on one thread.
__global__ void __launch_bounds__(256)
dist_add_list_boxes_kernel(int nrows, const int* __restrict__ box_list,
                           const signed char* __restrict__ mask_a, const signed char* __restrict__ mask_b,
                           const int* __restrict__ starts, const int* __restrict__ lists,
                           signed char* __restrict__ out_mask)
{
    const int nentries = starts[nrows];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nentries; k += gridDim.x * blockDim.x) {
        int lo = 0, hi = nrows;                // last row with starts[row] <= k
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= k) lo = mid; else hi = mid; }
        const int box = box_list[lo];
        if (mask_a[box] || (mask_b && mask_b[box])) out_mask[lists[k]] = 1;
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
    const signed char* masks; int nranks; int64_t nboxes;
    __device__ int operator()(int64_t b) const
    {
        int c = 0;
        for (int r = 0; r < nranks; ++r) c += masks[(int64_t)r * nboxes + b] ? 1 : 0;
        return c;
    }
};
struct RankCountOut {
    int* starts; int64_t n; long long* total_out;
    __device__ void operator()(int64_t b, long long excl) const { starts[b] = (int)excl; }
    __device__ void total(long long t) const { starts[n] = (int)t; *total_out = t; }
};
__global__ void dist_rank_fill_kernel(int nboxes, int nranks, const signed char* __restrict__ masks,
                                      const int* __restrict__ starts, int* __restrict__ lists)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        int k = starts[b];
        for (int r = 0; r < nranks; ++r) if (masks[(int64_t)r * nboxes + b]) lists[k++] = r;
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
    bt::dist_add_list_boxes_kernel<<<bt::kNumSMs * 16, 256, 0, (cudaStream_t)stream>>>(
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

int bt_dist_box_to_user_rank(int phase, int nboxes, int nranks, const int8_t* masks_all_ranks,
                             int32_t* starts, int32_t* lists, int64_t* total_dev, void* stream)
{
    BT_PROF("bt_dist_box_to_user_rank", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::RankCountIn in{(const signed char*)masks_all_ranks, nranks, nboxes};
        bt::RankCountOut out{starts, nboxes, (long long*)total_dev};
        return bt::scan_exclusive(nboxes, nullptr, in, out, s);
    }
    if (nboxes <= 0) return BT_OK;
    bt::dist_rank_fill_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
        nboxes, nranks, (const signed char*)masks_all_ranks, starts, lists);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // extern "C"
