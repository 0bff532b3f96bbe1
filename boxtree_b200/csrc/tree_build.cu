// Tree build kernels for sm_100a: bounding box, Morton keys (+ per-particle stop
// level for particles with extent), one-sweep radix sort, the per-level
// split-box loop, level restriction, pruning / renumbering, source/target
// split, coordinate permutation, box flags and box particle-extents.
//
// Design: particles are sorted ONCE by a full-depth key
//     K = [Morton digits of levels 1..min(stop,D), zero padded] << 6 | stop
// after which every box of every level is a contiguous range of the sorted
// order.  The level loop then only touches per-box data: child ranges are
// found by binary search on the sorted keys.  Box ids live in a creation-order
// "pool"; the final level-major numbering and the pruning of empty boxes are
// applied at the end.  Every float expression below restates the reference's
// OpenCL expression (cited per function) and is compiled with -fmad=false.
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"
#include "../../include/boxtree_b200.h"

#include <string.h>
#include <stdio.h>
#include <vector>
#include <string>

namespace bt {

// ---------------------------------------------------------------------------
// instrumentation
// ---------------------------------------------------------------------------
long long g_launch_count = 0;
int g_prof_enabled = 0;
struct ProfRec { std::string name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;

int prof_begin(const char* name, cudaStream_t s)
{
    ProfRec r; r.name = name;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int slot, cudaStream_t s) { cudaEventRecord(g_prof[slot].b, s); }

// ---------------------------------------------------------------------------
// key layout
// ---------------------------------------------------------------------------
constexpr int kStopBits = 6;
constexpr unsigned kStopNever = 63;

__host__ __device__ __forceinline__ int key_shift(int dim, int D, int level)
{   // bit position of the Morton digit of `level` (1..D) inside K
    return kStopBits + (D - level) * dim;
}

template <typename T, int DIM>
struct Particles {          // virtual concatenation sources ++ targets
    const T* src[DIM];
    const T* tgt[DIM];
    const T* src_radii;
    const T* tgt_radii;
    int nsrc;
    int n;
    __device__ __forceinline__ T coord(int a, int i) const
    { return i < nsrc ? src[a][i] : tgt[a][i - nsrc]; }
    __device__ __forceinline__ T radius(int i) const
    {
        if (i < nsrc) return src_radii ? src_radii[i] : (T)0;
        return tgt_radii ? tgt_radii[i - nsrc] : (T)0;
    }
};

// ---------------------------------------------------------------------------
// a1: bounding box -- boxtree/bounding_box.py:54-122
// ---------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
bbox_partial_kernel(Particles<T, DIM> P, int have_radii, T* __restrict__ partial /*[grid][2*DIM]*/)
{
    T mn[DIM], mx[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { mn[a] = CoordTraits<T>::maxval(); mx[a] = -CoordTraits<T>::maxval(); }
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const T r = have_radii ? P.radius(i) : (T)0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const T c = P.coord(a, i);
            const T lo = c - r, hi = c + r;
            mn[a] = (lo < mn[a]) ? lo : mn[a];
            mx[a] = (mx[a] < hi) ? hi : mx[a];
        }
    }
    __shared__ T s[8][2 * DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T lo = __shfl_xor_sync(0xffffffffu, mn[a], o);
            const T hi = __shfl_xor_sync(0xffffffffu, mx[a], o);
            mn[a] = (lo < mn[a]) ? lo : mn[a];
            mx[a] = (mx[a] < hi) ? hi : mx[a];
        }
        if (lane == 0) { s[warp][2 * a] = mn[a]; s[warp][2 * a + 1] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * DIM) {
        const int k = threadIdx.x;
        T v = s[0][k];
        for (int w = 1; w < 8; ++w) {
            const T o = s[w][k];
            v = (k & 1) ? ((v < o) ? o : v) : ((o < v) ? o : v);
        }
        partial[blockIdx.x * 2 * DIM + k] = v;
    }
}

// one block of 2 * DIM warps: warp k reduces component k of the partial results (its lanes
// stride over the blocks' rows, so the loads are independent, not a 592-long dependent chain)
template <typename T, int DIM>
__global__ void bbox_final_kernel(const T* __restrict__ partial, int nparts, T* __restrict__ out)
{
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= 2 * DIM) return;
    const bool is_max = k & 1;
    T v = partial[k];
    for (int p = 1 + lane; p < nparts; p += 32) {
        const T o = partial[p * 2 * DIM + k];
        v = is_max ? ((v < o) ? o : v) : ((o < v) ? o : v);
    }
#pragma unroll
    for (int o_ = 16; o_ > 0; o_ >>= 1) {
        const T o = __shfl_xor_sync(0xffffffffu, v, o_);
        v = is_max ? ((v < o) ? o : v) : ((o < v) ? o : v);
    }
    if (lane == 0) out[k] = v;     // layout: min_x, max_x, min_y, max_y, ...
}

// ---------------------------------------------------------------------------
// a3 (digit part): Morton key + stop level
// restates scan_t_from_particle, boxtree/tree_build_kernels.py:308-470
// ---------------------------------------------------------------------------
template <typename T> struct BBox { T mn[3]; T mx[3]; };

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
make_keys_kernel(Particles<T, DIM> P, BBox<T> bb, int extent_norm, T stick_out_factor, int D, int Dh,
                 int points_never_stop, double so_minext, double skip_slack,
                 unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ keys_lo, T* __restrict__ records)
{
    // D levels in total; `keys` resolves the first Dh of them, `keys_lo` (two-word keys, only for
    // trees deeper than one word holds) levels Dh+1..D.  One word: Dh == D, keys_lo == nullptr.
    // records (optional): the particle's coordinates and radius side by side (4 values), so
    // that the later gather into tree order (bt_permute) reads ONE aligned 16/32-byte record per
    // particle instead of one 32-byte sector per coordinate
    constexpr int RW = 4;
    const int stride = gridDim.x * blockDim.x;
    const T one_half = ((T)1) / 2;
    const T box_radius_factor = (T)((1. + (double)(extent_norm ? stick_out_factor : (T)0)) * (double)one_half);
    T gmin[DIM], gext[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { gmin[a] = bb.mn[a]; gext[a] = bb.mx[a] - gmin[a]; }
    const T scaleD = (T)(1ull << D);

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        T pos[DIM];
        unsigned long long q[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            pos[a] = P.coord(a, i);
            // (unsigned)(((x - min) / extent) * 2^(1+level)) for every level at once:
            // multiplying by a power of two is exact, so the level-(k) bits are q >> (D-k)
            q[a] = (unsigned long long)(((pos[a] - gmin[a]) / gext[a]) * scaleD);
        }
        unsigned stop = kStopNever;
        const T radius = (extent_norm || records) ? P.radius(i) : (T)0;
        if (records) {
            T rec[RW] = {0, 0, 0, 0};
#pragma unroll
            for (int a = 0; a < DIM; ++a) rec[a] = pos[a];
            rec[3] = radius;
            if (sizeof(T) == 8) {
                double2* dst = reinterpret_cast<double2*>(records + (size_t)i * RW);
                dst[0] = make_double2((double)rec[0], (double)rec[1]);
                dst[1] = make_double2((double)rec[2], (double)rec[3]);
            } else {
                *reinterpret_cast<float4*>(records + (size_t)i * RW) =
                    make_float4((float)rec[0], (float)rec[1], (float)rec[2], (float)rec[3]);
            }
        }
        // A particle without extent lies inside its box at every level; with a stick-out factor
        // that dwarfs the rounding errors of the tests below (make_keys_impl decides) none of them
        // can fire for it, so the level loop is skipped with the identical result.
        if (extent_norm && !(points_never_stop && radius == (T)0)) {
            // The same argument with the radius in it: at loop level lev the tests compare against
            // the box inflated by stick_out * min_ext * 2^-(2+lev); while that exceeds
            // radius + (the rounding bound, skip_slack = 64 eps M) none of them can fire, so the
            // walk starts at the first level where one could (a target of radius 2^-10 in a
            // 13-level key is tested on 4 levels instead of 13).  so_minext = 0 switches it off.
            int lev0 = 0;
            if (so_minext > 0) {
                const double R = so_minext / ((double)radius + skip_slack);
                if (R > 8.0) { const int l = ilogb(R) - 2; lev0 = l < D ? l : D; }
            }
            for (int lev = lev0; lev < D; ++lev) {    // lev = level of the box the particle sits in
                const T size_factor = ((T)1) / ((T)(1u << (1 + lev)));
                bool st = false;
                T center[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const unsigned bits = (unsigned)(q[a] >> (D - 1 - lev));
                    center[a] = gmin[a] + gext[a] * ((T)bits + one_half) * size_factor;
                }
                if (extent_norm == 1) {
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const T so_rad = box_radius_factor * gext[a] * size_factor;
                        st = st || (pos[a] + radius >= center[a] + so_rad);
                        st = st || (pos[a] - radius < center[a] - so_rad);
                    }
                } else {
                    const T so_rad = box_radius_factor * gext[0] * size_factor;
                    T acc = 0;
#pragma unroll
                    for (int a = 0; a < DIM; ++a)
                        acc = acc + (pos[a] - center[a]) * (pos[a] - center[a]);
                    const T dist = sqrt(acc) + radius;
                    st = (dist * dist >= DIM * so_rad * so_rad);
                }
                if (st) { stop = (unsigned)lev; break; }
            }
        }
        unsigned long long digits = 0, digits_lo = 0;
        const int maxlev = (stop < (unsigned)D) ? (int)stop : D;
        for (int lev = 1; lev <= maxlev; ++lev) {
            unsigned dg = 0;
#pragma unroll
            for (int a = 0; a < DIM; ++a)
                dg |= (unsigned)((q[a] >> (D - lev)) & 1ull) << (DIM - 1 - a);
            if (lev <= Dh) digits |= (unsigned long long)dg << ((Dh - lev) * DIM);
            else digits_lo |= (unsigned long long)dg << ((D - lev) * DIM);
        }
        // a particle that stops below the first word's levels does not stop inside them
        keys[i] = (digits << kStopBits) | ((stop <= (unsigned)Dh) ? stop : kStopNever);
        if (keys_lo) keys_lo[i] = (digits_lo << kStopBits) | stop;
    }
}

// ---------------------------------------------------------------------------
// box pool
// ---------------------------------------------------------------------------
template <typename T, int DIM>
struct Pool {
    // start/count/nn: the box's range in THIS rank's sorted particle array; gstart/gcount/gnn:
    // the same over all ranks (alias the local arrays on one GPU).  Split decisions, pruning
    // and the output box arrays use the global values, child ranges the local ones.
    int* start; int* count; unsigned char* level; int* parent; int* child0;
    unsigned char* has_children; unsigned char* force_split; int* nn;
    int* gstart; int* gcount; int* gnn;
    int* xch;           // distributed build: (lower bound, count, nonchild) of the new children
    T* center[DIM];
};

template <typename T, int DIM>
static Pool<T, DIM> make_pool(const bt_pool* p)
{
    Pool<T, DIM> r;
    r.start = p->start; r.count = p->count; r.level = p->level; r.parent = p->parent;
    r.child0 = p->child0; r.has_children = p->has_children; r.force_split = p->force_split;
    r.nn = p->nonchild;
    r.gstart = p->gstart ? p->gstart : p->start;
    r.gcount = p->gcount ? p->gcount : p->count;
    r.gnn = p->gnonchild ? p->gnonchild : p->nonchild;
    r.xch = p->xch;
    for (int a = 0; a < DIM; ++a) r.center[a] = (T*)p->center[a];
    return r;
}

__device__ __forceinline__ int clamp_weight(long long w)
{ return w > 2147483647ll ? 2147483647 : (int)w; }

// number of keys in [lo, hi) that compare <= kc  (upper bound)
__device__ __forceinline__ int upper_bound_key(const unsigned long long* __restrict__ keys, int lo, int hi,
                                               unsigned long long kc)
{
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (keys[mid] <= kc) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// root box (tree_build.py:586-618): box 0 holds all particles, is its own parent
template <typename T, int DIM>
__global__ void pool_init_kernel(Pool<T, DIM> pool, const unsigned long long* __restrict__ keys, int n,
                                 int have_ext, BBox<T> center, int* __restrict__ ctl)
{
    if (threadIdx.x || blockIdx.x) return;
    int nn = 0;
    if (have_ext && n > 0) nn = upper_bound_key(keys, 0, n, 0ull /* empty prefix, stop level 0 */);
    pool.start[0] = 0; pool.count[0] = n; pool.level[0] = 0; pool.parent[0] = 0; pool.child0[0] = 0;
    pool.has_children[0] = 0; pool.force_split[0] = 0; pool.nn[0] = nn;
    // distributed build: the host overwrites the global root entries with the all-reduced sums
    pool.gstart[0] = 0; pool.gcount[0] = n; pool.gnn[0] = nn;
    for (int a = 0; a < DIM; ++a) pool.center[a][0] = center.mn[a];
    for (int i = 0; i < BT_CTL_SIZE; ++i) ctl[i] = 0;
    ctl[BT_CTL_NBOXES] = 1;
}

// split decision -- restates count_new_boxes_needed,
// boxtree/tree_build_kernels.py:535-614, for one pool box
template <typename T, int DIM>
struct DecideIn {
    Pool<T, DIM> pool;
    const long long* wprefix;      // nullptr: every particle has weight 1
    int* ctl;
    unsigned char* flag;
    int lo, level, maxw, adaptive, level_restrict;
    __device__ int operator()(int64_t i) const
    {
        const int b = lo + (int)i;
        bool split = false, regular = false;
        if (pool.level[b] + 1 == level) {
            const int a = pool.gstart[b], c = pool.gcount[b], nn = pool.gnn[b];
            int wdesc;
            if (wprefix) wdesc = (c > 0) ? clamp_weight(wprefix[a + c] - wprefix[a + nn]) : 0;
            else wdesc = c - nn;
            regular = adaptive ? (wdesc > maxw) : true;
            split = regular;
        }
        if (level_restrict && pool.force_split[b]) split = true;
        if (split) pool.has_children[b] = 1;
        if (regular) atomicAdd(ctl + BT_CTL_NSPLIT_REGULAR, 1);
        flag[b] = split ? 1 : 0;
        return split ? 1 : 0;
    }
};

struct DecideOut {
    const unsigned char* flag; int* split_list; int* ctl; int lo;
    __device__ void operator()(int64_t i, long long excl) const
    { if (flag[lo + i]) split_list[excl] = lo + (int)i; }
    __device__ void total(long long t) const { ctl[BT_CTL_NSPLIT] = (int)t; }
};

// box splitter -- restates boxtree/tree_build_kernels.py:646-711 on the pool
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
create_children_kernel(Pool<T, DIM> pool, const unsigned long long* __restrict__ keys_hi,
                       const unsigned long long* __restrict__ keys_lo, int Dlo,
                       const long long* __restrict__ wprefix, const int* __restrict__ split_list,
                       int* __restrict__ ctl, int capacity, int D, int have_ext, int maxw,
                       int skip_if_no_regular, T root_extent)
{
    constexpr int NB = 1 << DIM;
    const int nsplit = ctl[BT_CTL_NSPLIT];
    const int base = ctl[BT_CTL_NBOXES];
    if (skip_if_no_regular && ctl[BT_CTL_NSPLIT_REGULAR] == 0) return;
    if ((long long)base + (long long)NB * nsplit > capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl[BT_CTL_OVERFLOW] = 1;
        return;
    }
    const long long total = (long long)nsplit * NB;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         tid < ((total + 31) & ~31ll); tid += stride) {
        const bool active = tid < total;
        const int r = (int)(tid / NB), m = (int)(tid % NB);
        int lb = 0, hi = 0, b = 0, lev = 0;
        // the key word that resolves the new level (two-word keys: levels above D are in keys_lo)
        const unsigned long long* keys = keys_hi;
        int Dw = D, lev_off = 0;
        if (active) {
            b = split_list[r];
            lev = pool.level[b];
            if (keys_lo && lev + 1 > D) { keys = keys_lo; Dw = Dlo; lev_off = D; }
            const int a = pool.start[b], c = pool.count[b];
            const int lo = a + pool.nn[b];
            hi = a + c;
            // first index in [lo, hi) whose level-(lev+1) digit is >= m
            const int sh = key_shift(DIM, Dw, lev + 1 - lev_off);
            int l = lo, h = hi;
            if (m == 0 || c == 0) h = l;
            while (l < h) {
                const int mid = l + ((h - l) >> 1);
                if ((int)((keys[mid] >> sh) & (NB - 1)) >= m) h = mid; else l = mid + 1;
            }
            lb = l;
        }
        int lb_next = __shfl_down_sync(0xffffffffu, lb, 1);
        if (!active) continue;
        if (m == NB - 1) lb_next = hi;
        const int cnt = lb_next - lb;
        const int child = base + r * NB + m;
        const int new_level = lev + 1;
        pool.parent[child] = b;
        pool.level[child] = (unsigned char)new_level;
        pool.count[child] = cnt;
        // distributed build: the local lower bound is kept for empty ranges too (the global
        // start of a box is the sum of the ranks' lower bounds)
        pool.start[child] = (cnt > 0 || pool.xch) ? lb : 0;
        pool.child0[child] = 0;
        pool.has_children[child] = 0;
        pool.force_split[child] = 0;
        int nn = 0;
        if (have_ext && cnt > 0) {
            const int shc = key_shift(DIM, Dw, new_level - lev_off);
            const unsigned long long kc = ((keys[lb] >> shc) << shc) | (unsigned long long)new_level;
            nn = upper_bound_key(keys, lb, lb_next, kc) - lb;
        }
        pool.nn[child] = nn;
        if (pool.xch) {         // sums over ranks follow (children_finish_kernel)
            int* x = pool.xch + 2ll * (r * NB + m);
            x[0] = cnt; x[1] = nn;
        } else {
            const int w = wprefix ? ((cnt > 0) ? clamp_weight(wprefix[lb_next] - wprefix[lb]) : 0) : cnt;
            if (w > maxw) ctl[BT_CTL_OVERSIZE] = 1;
        }
        // centre chain: parent +- root_extent / 2^(1+new_level)   (:698-705)
        const T radius = (root_extent * 1 / (T)(1 << (1 + new_level)));
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const bool has_bit = (m >> (DIM - 1 - a)) & 1;
            const T pc = pool.center[a][b];
            pool.center[a][child] = has_bit ? pc + radius : pc - radius;
        }
        if (m == 0) pool.child0[b] = base + r * NB;
    }
}

// distributed build: the all-reduced (count, nonchild) of the new children become their global
// ranges -- a child starts where the parent's descendants start plus its lower siblings' counts,
// like the splitter's local ranges; the oversize test (:690-696) runs on the global count
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
children_finish_kernel(Pool<T, DIM> pool, const int* __restrict__ split_list, int* __restrict__ ctl, int maxw,
                       int skip_if_no_regular)
{
    constexpr int NB = 1 << DIM;
    if (ctl[BT_CTL_OVERFLOW]) return;
    if (skip_if_no_regular && ctl[BT_CTL_NSPLIT_REGULAR] == 0) return;
    const int total = ctl[BT_CTL_NSPLIT] * NB, base = ctl[BT_CTL_NBOXES];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int r = k / NB, m = k % NB;
        const int* x = pool.xch + 2ll * (r * NB);
        const int b = split_list[r];
        int start = pool.gstart[b] + pool.gnn[b];
        for (int q = 0; q < m; ++q) start += x[2 * q];
        const int cnt = x[2 * m];
        pool.gstart[base + k] = (cnt > 0) ? start : 0;
        pool.gcount[base + k] = cnt;
        pool.gnn[base + k] = x[2 * m + 1];
        if (cnt > maxw) ctl[BT_CTL_OVERSIZE] = 1;
    }
}

template <int DIM>
__global__ void commit_level_kernel(int* ctl, int skip_if_no_regular)
{
    if (threadIdx.x || blockIdx.x) return;
    if (ctl[BT_CTL_OVERFLOW]) return;
    if (skip_if_no_regular && ctl[BT_CTL_NSPLIT_REGULAR] == 0) return;
    ctl[BT_CTL_NBOXES] = ctl[BT_CTL_NBOXES] + (1 << DIM) * ctl[BT_CTL_NSPLIT];
    ctl[BT_CTL_COMMITTED] = 1;
}

// adjacency predicate -- boxtree/traversal.py:279-318
template <typename T, int DIM>
__device__ __forceinline__ bool adjacent_nbhd(T root_extent, const T* tc, int tl, T nbhd, const T* sc, int sl)
{
    const T target_rad = (root_extent * 1 / (T)(1 << (tl + 1)));
    const T source_rad = (root_extent * 1 / (T)(1 << (sl + 1)));
    const T rad_sum = ((2 * (nbhd - 1) + 1) * target_rad + source_rad);
    const T slack = rad_sum + fmin(target_rad, source_rad);
    T l_inf_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) l_inf_dist = fmax(l_inf_dist, fabs(tc[a] - sc[a]));
    return l_inf_dist <= slack;
}

// level restriction -- restates boxtree/tree_build_kernels.py:825-915 and the
// sweep control of boxtree/tree_build.py:1163-1200 (device-side break chain)
template <typename T, int DIM>
__global__ void __launch_bounds__(128)
level_restrict_kernel(Pool<T, DIM> pool, int* __restrict__ ctl, int level, int first, T root_extent)
{
    constexpr int NB = 1 << DIM;
    if (ctl[BT_CTL_OVERFLOW] || !ctl[BT_CTL_COMMITTED]) return;
    if (!first && ctl[BT_CTL_LR_FOUND + level + 1] == 0) return;
    const int nboxes = ctl[BT_CTL_NBOXES];
    const int stride = gridDim.x * blockDim.x;
    for (int box_id = blockIdx.x * blockDim.x + threadIdx.x; box_id < nboxes; box_id += stride) {
        if (pool.level[box_id] != level || pool.has_children[box_id]) continue;
        T bc[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) bc[a] = pool.center[a][box_id];
        int stack_box[40]; signed char stack_mnr[40];
        int ssize = 0, wparent = 0, wmnr = 0;
        bool cont = true;
        while (cont) {
            const int c0 = pool.child0[wparent];
            const int child = c0 ? c0 + wmnr : 0;
            bool pushed = false;
            if (child) {
                const int child_level = ssize + 1;
                bool adj = false;
                if (child != box_id) {
                    T cc[DIM];
#pragma unroll
                    for (int a = 0; a < DIM; ++a) cc[a] = pool.center[a][child];
                    adj = adjacent_nbhd<T, DIM>(root_extent, cc, child_level, (T)1, bc, level);
                }
                if (adj) {
                    if (pool.has_children[child]) {
                        if (child_level <= 1 + level) {
                            stack_box[ssize] = wparent; stack_mnr[ssize] = (signed char)wmnr; ++ssize;
                            wparent = child; wmnr = 0; pushed = true;
                        }
                    } else if (child_level == 2 + level ||
                               (child_level == 1 + level && pool.force_split[child])) {
                        pool.force_split[box_id] = 1;
                        ctl[BT_CTL_LR_FOUND + level] = 1;
                        cont = false;
                    }
                }
            }
            if (pushed) continue;
            while (true) {          // walk_advance, traversal.py:115-143
                ++wmnr;
                if (wmnr < NB) break;
                cont = cont && (ssize > 0);
                if (ssize > 0) { --ssize; wparent = stack_box[ssize]; wmnr = stack_mnr[ssize]; }
                else break;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// finalize: numbering, pruning, output box arrays
// ---------------------------------------------------------------------------
struct KeepIn {
    const int* order; const int* count; int skip_prune;
    __device__ int operator()(int64_t i) const
    {
        const int b = order ? order[i] : (int)i;
        return (skip_prune || count[b] != 0) ? 1 : 0;
    }
};
struct KeepOut {
    const int* order; const int* count; const unsigned char* level; int skip_prune;
    int* map_old2new; int* src_of_new; int* level_start; int* ctl;
    __device__ void operator()(int64_t i, long long excl) const
    {
        const int b = order ? order[i] : (int)i;
        const bool keep = skip_prune || count[b] != 0;
        map_old2new[b] = keep ? (int)excl : 0;       // pruned boxes map to 0 (tree_build.py:1336-1340)
        if (keep) src_of_new[excl] = b;
        const int lev = level[b];
        const int prev = (i == 0) ? -1 : (int)level[order ? order[i - 1] : (int)(i - 1)];
        for (int l = prev + 1; l <= lev; ++l) level_start[l] = (int)excl;
    }
    __device__ void total(long long t) const { ctl[BT_CTL_NBOXES_FINAL] = (int)t; }
};

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
gather_boxes_kernel(Pool<T, DIM> pool, const int* __restrict__ src_of_new,
                    const int* __restrict__ map_old2new, int nfinal, int aligned, int have_ext,
                    int* __restrict__ out_start, int* __restrict__ out_count,
                    int* __restrict__ out_nn, unsigned char* __restrict__ out_level,
                    int* __restrict__ out_parent, int* __restrict__ out_child /*[NB, aligned]*/,
                    T* __restrict__ out_center /*[DIM, aligned]*/,
                    unsigned char* __restrict__ out_has_children,
                    unsigned char* __restrict__ out_real_children,
                    int* __restrict__ out_lstart, int* __restrict__ out_lcount, int* __restrict__ out_lnn)
{
    constexpr int NB = 1 << DIM;
    const int stride = gridDim.x * blockDim.x;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nfinal; j += stride) {
        const int b = src_of_new[j];
        out_start[j] = pool.gstart[b];
        out_count[j] = pool.gcount[b];
        out_nn[j] = have_ext ? pool.gnn[b] : 0;
        if (out_lstart) {       // distributed build: the box's range in this rank's particles
            out_lstart[j] = pool.start[b];
            out_lcount[j] = pool.count[b];
            out_lnn[j] = have_ext ? pool.nn[b] : 0;
        }
        out_level[j] = pool.level[b];
        out_parent[j] = map_old2new[pool.parent[b]];
        const int c0 = pool.child0[b];
#pragma unroll
        for (int m = 0; m < NB; ++m) out_child[m * aligned + j] = c0 ? map_old2new[c0 + m] : 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) out_center[a * aligned + j] = pool.center[a][b];
        out_has_children[j] = pool.has_children[b];
        out_real_children[j] = c0 ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// leaf order fix-up: inside every box that was never partitioned the reference
// keeps particles in ascending user id (stable counting partition starting
// from arange ids, tree_build_kernels.py:766-798, tree_build.py:395)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_bitonic_sort(unsigned v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            const unsigned mn = v < o ? v : o, mx = v < o ? o : v;
            v = (lower == up) ? mn : mx;
        }
    }
    return v;
}

constexpr int kFixBlock = 256;
constexpr int kFixSmemCap = 4096;

// every lane looks at one box; the boxes that need work (never partitioned, >= 2 particles) are
// then sorted one after the other by the whole warp (count <= 32); larger boxes are sorted by a
// block in shared memory (bitonic), boxes above kFixSmemCap are listed for the host.
__global__ void __launch_bounds__(kFixBlock)
leaf_fixup_kernel(const int* __restrict__ box_start, const int* __restrict__ box_count,
                  const unsigned char* __restrict__ real_children, int nboxes,
                  unsigned* __restrict__ ids, int* __restrict__ big_list, int* __restrict__ ctl,
                  int big_cap)
{
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b0 = wglobal * 32; b0 < nboxes; b0 += nwarps * 32) {
        const int b = b0 + lane;
        int c = 0, s = 0;
        if (b < nboxes && !real_children[b]) { c = box_count[b]; s = box_start[b]; }
        if (c > 32) {
            const int k = atomicAdd(ctl + BT_CTL_NBIG, 1);
            if (k < big_cap) big_list[k] = b;
        }
        unsigned todo = __ballot_sync(0xffffffffu, c >= 2 && c <= 32);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int cc = __shfl_sync(0xffffffffu, c, src), ss = __shfl_sync(0xffffffffu, s, src);
            unsigned v = (lane < cc) ? ids[ss + lane] : 0xffffffffu;
            v = warp_bitonic_sort(v);
            if (lane < cc) ids[ss + lane] = v;
        }
    }
}

__global__ void __launch_bounds__(kFixBlock)
leaf_fixup_big_kernel(const int* __restrict__ box_start, const int* __restrict__ box_count,
                      const int* __restrict__ big_list, int* __restrict__ ctl, int big_cap,
                      unsigned* __restrict__ ids, int* __restrict__ huge_list)
{
    __shared__ unsigned s[kFixSmemCap];
    int nbig = ctl[BT_CTL_NBIG];
    if (nbig > big_cap) nbig = big_cap;
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
        const int b = big_list[q];
        const int c = box_count[b], st = box_start[b];
        if (c > kFixSmemCap) {
            if (threadIdx.x == 0) { const int k = atomicAdd(ctl + BT_CTL_NHUGE, 1); if (k < big_cap) huge_list[k] = b; }
            continue;
        }
        int p2 = 64; while (p2 < c) p2 <<= 1;
        for (int i = threadIdx.x; i < p2; i += blockDim.x) s[i] = (i < c) ? ids[st + i] : 0xffffffffu;
        __syncthreads();
        for (int k = 2; k <= p2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned a = s[i], bb = s[ixj];
                        const bool up = ((i & k) == 0);
                        if ((a > bb) == up) { s[i] = bb; s[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < c; i += blockDim.x) ids[st + i] = s[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// a11: source/target split -- boxtree/tree_build_kernels.py:1770-1782, 1013-1164
// ---------------------------------------------------------------------------
struct SourceIn {
    const unsigned* ids; unsigned nsources;
    __device__ int operator()(int64_t i) const { return ids[i] < nsources ? 1 : 0; }
};
struct SourceOut {
    const unsigned* ids; unsigned nsources;
    int* source_numbers;    // [n+1]
    int* user_source_ids; int* srcntgt_target_ids; int* sorted_target_ids; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const
    {
        source_numbers[i] = (int)excl;
        const unsigned id = ids[i];
        if (id < nsources) user_source_ids[excl] = (int)id;
        else {
            const int tnr = (int)(i - excl);
            srcntgt_target_ids[tnr] = (int)id;
            sorted_target_ids[id - nsources] = tnr;
        }
    }
    __device__ void total(long long t) const { source_numbers[n] = (int)t; }
};

__global__ void reverse_index_kernel(const unsigned* __restrict__ ids, int n, int* __restrict__ out)
{   // tools.py:81-109 (reverse_index_array)
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[ids[i]] = i;
}

// a12: permute -- boxtree/tree_build_kernels.py:1170-1186 and cl_array.take
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
permute_kernel(Particles<T, DIM> P, const T* __restrict__ records, const int* __restrict__ from_ids, int n,
               int want_radii, T* o0, T* o1, T* o2, T* out_radii)
{
    T* outs[3] = {o0, o1, o2};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int f = from_ids[i];
        if (records) {      // one aligned record per particle (make_keys_kernel)
            T rec[4];
            if (sizeof(T) == 8) {
                const double2* src = reinterpret_cast<const double2*>(records + (size_t)f * 4);
                const double2 a = src[0], b = src[1];
                rec[0] = (T)a.x; rec[1] = (T)a.y; rec[2] = (T)b.x; rec[3] = (T)b.y;
            } else {
                const float4 a = *reinterpret_cast<const float4*>(records + (size_t)f * 4);
                rec[0] = (T)a.x; rec[1] = (T)a.y; rec[2] = (T)a.z; rec[3] = (T)a.w;
            }
#pragma unroll
            for (int a = 0; a < DIM; ++a) outs[a][i] = rec[a];
            if (want_radii) out_radii[i] = rec[3];
        } else {
#pragma unroll
            for (int a = 0; a < DIM; ++a) outs[a][i] = P.coord(a, f);
            if (want_radii) out_radii[i] = P.radius(f);
        }
    }
}

// ---------------------------------------------------------------------------
// a11 (box part) + a13: per-box source/target ranges and flags
// restates boxtree/tree_build_kernels.py:1062-1147 (as prefix differences)
// and :1192-1305 (box_info)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
box_info_kernel(int nboxes, int sources_are_targets, int have_ext,
                const int* __restrict__ box_start, const int* __restrict__ box_count,
                const int* __restrict__ box_nn, const unsigned char* __restrict__ has_children,
                const int* __restrict__ source_numbers /*[n+1] or null*/,
                int* __restrict__ src_starts, int* __restrict__ src_nonchild, int* __restrict__ src_cumul,
                int* __restrict__ tgt_starts, int* __restrict__ tgt_nonchild, int* __restrict__ tgt_cumul,
                unsigned char* __restrict__ box_flags)
{
    const int stride = gridDim.x * blockDim.x;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += stride) {
        const int st = box_start[b], c = box_count[b];
        const int nn = have_ext ? box_nn[b] : 0;
        unsigned char fl = 0;
        if (sources_are_targets) {
            // src_* and tgt_* alias (tree_build.py:1469-1474)
            int nonchild;
            if (has_children[b]) {
                // tree_build_kernels.py:1256 sets both child-box bits for every non-leaf
                fl |= BT_BOX_HAS_SOURCE_CHILD_BOXES | BT_BOX_HAS_TARGET_CHILD_BOXES;
                nonchild = 0;
            } else {
                if (c) fl |= BT_BOX_IS_SOURCE_BOX | BT_BOX_IS_TARGET_BOX;
                nonchild = c;
            }
            src_starts[b] = st; src_cumul[b] = c; src_nonchild[b] = nonchild;
        } else {
            int s_st = 0, t_st = 0, s_cu = 0, t_cu = 0, s_nc = 0, t_nc = 0;
            if (c > 0) {
                const int s0 = source_numbers[st];
                s_st = s0; t_st = st - s0;
                s_cu = source_numbers[st + c] - s0; t_cu = c - s_cu;
                if (have_ext && nn > 0) { s_nc = source_numbers[st + nn] - s0; t_nc = nn - s_nc; }
            }
            if (has_children[b]) {
                // tree_build_kernels.py:1256 sets both child-box bits for every non-leaf
                fl |= BT_BOX_HAS_SOURCE_CHILD_BOXES | BT_BOX_HAS_TARGET_CHILD_BOXES;
                if (s_nc) fl |= BT_BOX_IS_SOURCE_BOX;
                if (t_nc) fl |= BT_BOX_IS_TARGET_BOX;
            } else {
                if (s_cu) fl |= BT_BOX_IS_SOURCE_BOX;
                if (c - s_cu) fl |= BT_BOX_IS_TARGET_BOX;
                s_nc = s_cu; t_nc = c - s_cu;
            }
            src_starts[b] = s_st; src_cumul[b] = s_cu; src_nonchild[b] = s_nc;
            tgt_starts[b] = t_st; tgt_cumul[b] = t_cu; tgt_nonchild[b] = t_nc;
        }
        box_flags[b] = fl;
    }
}

// Distributed build, part 1: the source counts of the box's range in THIS rank's particles
// (lower bounds kept for empty ranges so that the sums over ranks are the global values).
// out3 = [3][nboxes]: sources before the box, in the box (cumulative), stopping in the box.
__global__ void __launch_bounds__(256)
box_info_local_kernel(int nboxes, int have_ext, const int* __restrict__ lstart,
                      const int* __restrict__ lcount, const int* __restrict__ lnn,
                      const unsigned char* __restrict__ has_children,
                      const int* __restrict__ source_numbers, int* __restrict__ out3,
                      int* __restrict__ src_starts, int* __restrict__ src_nonchild, int* __restrict__ src_cumul,
                      int* __restrict__ tgt_starts, int* __restrict__ tgt_nonchild, int* __restrict__ tgt_cumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int st = lstart[b], c = lcount[b];
        const int nn = have_ext ? lnn[b] : 0;
        const int s0 = source_numbers[st];
        const int s_cu = source_numbers[st + c] - s0;
        int s_nc = (have_ext && nn > 0) ? source_numbers[st + nn] - s0 : 0;
        int t_nc = nn - s_nc;
        if (!has_children[b]) { s_nc = s_cu; t_nc = c - s_cu; }
        out3[b] = s0; out3[nboxes + b] = s_cu; out3[2 * nboxes + b] = s_nc;
        src_starts[b] = s0; src_cumul[b] = s_cu; src_nonchild[b] = s_nc;
        tgt_starts[b] = st - s0; tgt_cumul[b] = c - s_cu; tgt_nonchild[b] = t_nc;
    }
}

// Distributed build, part 2: the global tree's per-box ranges and flags from the global
// srcntgt ranges and the all-reduced source counts (same case analysis as box_info_kernel)
__global__ void __launch_bounds__(256)
box_info_global_kernel(int nboxes, int have_ext, const int* __restrict__ gstart,
                       const int* __restrict__ gcount, const int* __restrict__ gnn,
                       const unsigned char* __restrict__ has_children, const int* __restrict__ src3,
                       int* __restrict__ src_starts, int* __restrict__ src_nonchild, int* __restrict__ src_cumul,
                       int* __restrict__ tgt_starts, int* __restrict__ tgt_nonchild, int* __restrict__ tgt_cumul,
                       unsigned char* __restrict__ box_flags)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int st = gstart[b], c = gcount[b];
        const int nn = have_ext ? gnn[b] : 0;
        int s_st = 0, t_st = 0, s_cu = 0, t_cu = 0, s_nc = 0, t_nc = 0;
        unsigned char fl = 0;
        if (c > 0) {
            s_st = src3[b]; t_st = st - s_st;
            s_cu = src3[nboxes + b]; t_cu = c - s_cu;
            s_nc = src3[2 * nboxes + b];
            t_nc = (has_children[b] ? nn : c) - s_nc;
        }
        if (has_children[b]) {
            fl |= BT_BOX_HAS_SOURCE_CHILD_BOXES | BT_BOX_HAS_TARGET_CHILD_BOXES;
            if (s_nc) fl |= BT_BOX_IS_SOURCE_BOX;
            if (t_nc) fl |= BT_BOX_IS_TARGET_BOX;
        } else {
            if (s_cu) fl |= BT_BOX_IS_SOURCE_BOX;
            if (c - s_cu) fl |= BT_BOX_IS_TARGET_BOX;
        }
        src_starts[b] = s_st; src_cumul[b] = s_cu; src_nonchild[b] = s_nc;
        tgt_starts[b] = t_st; tgt_cumul[b] = t_cu; tgt_nonchild[b] = t_nc;
        box_flags[b] = fl;
    }
}

// ---------------------------------------------------------------------------
// a14: box particle extents -- boxtree/tree_build_kernels.py:1311-1399,
// launched per level bottom-up like boxtree/tree_build.py:1751-1802
// ---------------------------------------------------------------------------
// Phase A (all boxes at once): min/max over the box's own particles, seeded with the box
// centre.  min/max are exact and order-free, so the parallel reduction returns the same bits
// as the reference's serial loop (:1345-1368).  Four boxes per warp (8 lanes each: a leaf
// holds at most a few dozen particles); a box with many own particles (upper-level boxes of
// a tree with extents) is taken by the whole warp.
constexpr int kExtBig = 64, kExtHuge = 256;

// kExtGroup lanes per box: 8 for a tree whose leaves hold a few dozen particles, 1 when most
// boxes hold at most a couple (a rank's share of the particles in a distributed build)
template <typename T, int DIM, int kExtGroup>
__global__ void __launch_bounds__(256)
box_extents_own_kernel(int nboxes, int aligned, const T* __restrict__ box_centers,
                       const int* __restrict__ pstarts, const int* __restrict__ pcounts,
                       const T* p0, const T* p1, const T* p2, const T* __restrict__ radii,
                       T* __restrict__ bb_min, T* __restrict__ bb_max, int* __restrict__ huge_list,
                       int* __restrict__ huge_count)
{
    constexpr int GPW = 32 / kExtGroup;
    const T* parts[3] = {p0, p1, p2};
    const int lane = threadIdx.x & 31, g = lane / kExtGroup, gl = lane % kExtGroup;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = wglobal * GPW; base < nboxes; base += nwarps * GPW) {
        const int ibox = base + g;
        const bool valid = ibox < nboxes;
        T mn[DIM], mx[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) mn[a] = mx[a] = valid ? box_centers[a * aligned + ibox] : (T)0;
        const int s = valid ? pstarts[ibox] : 0, e = valid ? s + pcounts[ibox] : 0;
        // boxes with thousands of own particles go to a list for box_extents_huge_kernel
        const bool huge = (e - s) > kExtHuge;
        if (huge && gl == 0) huge_list[atomicAdd(huge_count, 1)] = ibox;
        const bool big = !huge && (e - s) > (kExtGroup == 1 ? 8 : kExtBig);
        if (!big && !huge) {        // (a huge box is box_extents_huge_kernel's: do not walk it here)
            for (int ip = s + gl; ip < e; ip += kExtGroup) {
                const T rad = radii ? radii[ip] : (T)0;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const T c = parts[a][ip];
                    const T lo = c - rad, hi = c + rad;
                    mn[a] = (lo < mn[a]) ? lo : mn[a];
                    mx[a] = (mx[a] < hi) ? hi : mx[a];
                }
            }
        }
        // boxes with many own particles: all 32 lanes, one such box after the other
        unsigned bigm = __ballot_sync(0xffffffffu, big && gl == 0);
        while (bigm) {
            const int src = __ffs(bigm) - 1;
            bigm &= bigm - 1;
            const int bs = __shfl_sync(0xffffffffu, s, src), be = __shfl_sync(0xffffffffu, e, src);
            T wmn[DIM], wmx[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                wmn[a] = __shfl_sync(0xffffffffu, mn[a], src); wmx[a] = __shfl_sync(0xffffffffu, mx[a], src);
            }
            for (int ip = bs + lane; ip < be; ip += 32) {
                const T rad = radii ? radii[ip] : (T)0;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const T c = parts[a][ip];
                    const T lo = c - rad, hi = c + rad;
                    wmn[a] = (lo < wmn[a]) ? lo : wmn[a];
                    wmx[a] = (wmx[a] < hi) ? hi : wmx[a];
                }
            }
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
#pragma unroll
                for (int o = 16; o >= (kExtGroup > 1 ? kExtGroup : 1); o >>= 1) {   // across groups; the rest below
                    const T lo = __shfl_xor_sync(0xffffffffu, wmn[a], o);
                    const T hi = __shfl_xor_sync(0xffffffffu, wmx[a], o);
                    wmn[a] = (lo < wmn[a]) ? lo : wmn[a];
                    wmx[a] = (wmx[a] < hi) ? hi : wmx[a];
                }
                if (lane / kExtGroup == src / kExtGroup) { mn[a] = wmn[a]; mx[a] = wmx[a]; }
            }
        }
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
#pragma unroll
            for (int o = kExtGroup / 2; o > 0; o >>= 1) {
                const T lo = __shfl_xor_sync(0xffffffffu, mn[a], o);
                const T hi = __shfl_xor_sync(0xffffffffu, mx[a], o);
                mn[a] = (lo < mn[a]) ? lo : mn[a];
                mx[a] = (mx[a] < hi) ? hi : mx[a];
            }
            if (valid && !huge && gl == 0) { bb_min[a * aligned + ibox] = mn[a]; bb_max[a * aligned + ibox] = mx[a]; }
        }
        (void)GPW;
    }
}

// boxes with more than kExtHuge own particles: one block per box
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
box_extents_huge_kernel(int aligned, const T* __restrict__ box_centers, const int* __restrict__ pstarts,
                        const int* __restrict__ pcounts, const T* p0, const T* p1, const T* p2,
                        const T* __restrict__ radii, T* __restrict__ bb_min, T* __restrict__ bb_max,
                        const int* __restrict__ huge_list, const int* __restrict__ huge_count)
{
    const T* parts[3] = {p0, p1, p2};
    __shared__ T smn[8][DIM], smx[8][DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nhuge = *huge_count;
    for (int i = blockIdx.x; i < nhuge; i += gridDim.x) {
        const int ibox = huge_list[i];
        T mn[DIM], mx[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) mn[a] = mx[a] = box_centers[a * aligned + ibox];
        const int s = pstarts[ibox], e = s + pcounts[ibox];
        for (int ip = s + threadIdx.x; ip < e; ip += blockDim.x) {
            const T rad = radii ? radii[ip] : (T)0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const T c = parts[a][ip];
                const T lo = c - rad, hi = c + rad;
                mn[a] = (lo < mn[a]) ? lo : mn[a];
                mx[a] = (mx[a] < hi) ? hi : mx[a];
            }
        }
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T lo = __shfl_xor_sync(0xffffffffu, mn[a], o);
                const T hi = __shfl_xor_sync(0xffffffffu, mx[a], o);
                mn[a] = (lo < mn[a]) ? lo : mn[a];
                mx[a] = (mx[a] < hi) ? hi : mx[a];
            }
            if (lane == 0) { smn[warp][a] = mn[a]; smx[warp][a] = mx[a]; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                T lo = smn[0][a], hi = smx[0][a];
                for (int w = 1; w < 8; ++w) {
                    lo = (smn[w][a] < lo) ? smn[w][a] : lo;
                    hi = (hi < smx[w][a]) ? smx[w][a] : hi;
                }
                bb_min[a * aligned + ibox] = lo; bb_max[a * aligned + ibox] = hi;
            }
        }
        __syncthreads();
    }
}

// Phase B (one launch per level, bottom-up): merge the children's boxes (:1370-1389)
template <typename T, int DIM>
__global__ void __launch_bounds__(128)
box_extents_merge_kernel(int start, int stop, int aligned, const int* __restrict__ box_child_ids,
                         T* __restrict__ bb_min, T* __restrict__ bb_max)
{
    constexpr int NB = 1 << DIM;
    const int stride = gridDim.x * blockDim.x;
    for (int ibox = start + blockIdx.x * blockDim.x + threadIdx.x; ibox < stop; ibox += stride) {
        T mn[DIM], mx[DIM];
        bool any = false;
#pragma unroll
        for (int a = 0; a < DIM; ++a) { mn[a] = bb_min[a * aligned + ibox]; mx[a] = bb_max[a * aligned + ibox]; }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const int child = box_child_ids[m * aligned + ibox];
            if (child == 0) continue;
            any = true;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const T cmn = bb_min[a * aligned + child], cmx = bb_max[a * aligned + child];
                mn[a] = (cmn < mn[a]) ? cmn : mn[a];
                mx[a] = (mx[a] < cmx) ? cmx : mx[a];
            }
        }
        if (any) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) { bb_min[a * aligned + ibox] = mn[a]; bb_max[a * aligned + ibox] = mx[a]; }
        }
    }
}

// weights in sorted order -> exclusive prefix (int64), wprefix[n] = total
struct WeightIn {
    const unsigned* ids; const int* weights;
    __device__ int operator()(int64_t i) const { return weights[ids[i]]; }
};
struct WeightOut {
    long long* wprefix; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { wprefix[i] = excl; }
    __device__ void total(long long t) const { wprefix[n] = t; }
};

__global__ void widen_ids_kernel(const unsigned* __restrict__ ids, int n, unsigned long long* __restrict__ keys)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) keys[i] = ids[i];
}
__global__ void narrow_ids_kernel(const unsigned long long* __restrict__ keys, int n, unsigned* __restrict__ ids)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ids[i] = (unsigned)keys[i];
}

// keys for the stable by-level grouping of the pool (level-restricted trees)
__global__ void level_keys_kernel(const unsigned char* __restrict__ level, int n,
                                  unsigned long long* __restrict__ keys)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) keys[i] = level[i];
}

// ---------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------
template <typename T, int DIM>
static Particles<T, DIM> make_particles(const bt_particles* p)
{
    Particles<T, DIM> P;
    for (int a = 0; a < DIM; ++a) {
        P.src[a] = (const T*)p->sources[a];
        P.tgt[a] = (const T*)p->targets[a];
    }
    P.src_radii = (const T*)p->source_radii;
    P.tgt_radii = (const T*)p->target_radii;
    P.nsrc = (int)p->nsources;
    P.n = (int)(p->nsources + p->ntargets);
    return P;
}

template <typename T, int DIM>
static int bbox_impl(const bt_particles* p, void* out, cudaStream_t s)
{
    Particles<T, DIM> P = make_particles<T, DIM>(p);
    const int have_radii = (p->source_radii || p->target_radii) ? 1 : 0;
    const int grid = grid_for(P.n, 256, 4);
    T* partial = nullptr;
    BT_CHECK(bt::temp_alloc((void**)&partial, sizeof(T) * 2 * DIM * grid, s));
    bbox_partial_kernel<T, DIM><<<grid, 256, 0, s>>>(P, have_radii, partial);
    BT_LAUNCH_CHECK();
    bbox_final_kernel<T, DIM><<<1, 32 * 2 * DIM, 0, s>>>(partial, grid, (T*)out);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(partial, s));
    return BT_OK;
}

template <typename T, int DIM>
static int make_keys_impl(const bt_particles* p, const double* bmin, const double* bmax,
                          int extent_norm, double stick_out, int D, int Dh, unsigned long long* keys,
                          unsigned long long* keys_lo, void* records, cudaStream_t s)
{
    Particles<T, DIM> P = make_particles<T, DIM>(p);
    BBox<T> bb;
    for (int a = 0; a < 3; ++a) { bb.mn[a] = (a < DIM) ? (T)bmin[a] : (T)0; bb.mx[a] = (a < DIM) ? (T)bmax[a] : (T)1; }
    if (P.n == 0) return BT_OK;
    // May the stop-level loop be skipped for particles of radius 0?  Such a particle is inside its
    // level-k box up to the rounding of (x - min) / extent (2 roundings) and of the box centre
    // (3 roundings), together < 8 eps * M with M = max(|min|, |max|, extent); the tests compare
    // against the box inflated by stick_out * (half the box size) (linf per axis; l2 the same up to
    // a factor (1 + O(eps))).  Skipping is exact while that margin, at the finest level D, exceeds
    // the error bound; the check asks for 64 eps M, 8x the bound.  fp32 builds seldom qualify.
    int points_never_stop = 0;
    double so_minext = 0, skip_slack = 0;
    if (extent_norm && stick_out > 0 && !(getenv("BT_KEYS_NO_SKIP"))) {
        double min_ext = 1e300, M = 0;
        for (int a = 0; a < DIM; ++a) {
            const double ext = (double)bb.mx[a] - (double)bb.mn[a];
            min_ext = ext < min_ext ? ext : min_ext;
            M = fmax(M, fmax(fabs((double)bb.mn[a]), fmax(fabs((double)bb.mx[a]), ext)));
        }
        const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
        const double margin = (double)(T)stick_out * min_ext / (double)(1ull << (D + 1));
        points_never_stop = (min_ext > 0 && margin > 64.0 * eps * M) ? 1 : 0;
        if (min_ext > 0) { so_minext = (double)(T)stick_out * min_ext; skip_slack = 64.0 * eps * M; }
    }
    make_keys_kernel<T, DIM><<<grid_for(P.n, 256, 8), 256, 0, s>>>(P, bb, extent_norm, (T)stick_out, D, Dh,
                                                                  points_never_stop, so_minext, skip_slack,
                                                                  keys, keys_lo, (T*)records);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T, int DIM>
static int pool_init_impl(const bt_pool* pool, int64_t n, int have_ext, const unsigned long long* keys,
                          const double* root_center, int* ctl, cudaStream_t s)
{
    Pool<T, DIM> P = make_pool<T, DIM>(pool);
    BBox<T> c;
    for (int a = 0; a < 3; ++a) { c.mn[a] = (a < DIM) ? (T)root_center[a] : (T)0; c.mx[a] = 0; }
    pool_init_kernel<T, DIM><<<1, 32, 0, s>>>(P, keys, (int)n, have_ext, c, ctl);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T, int DIM>
static int level_step_impl(const bt_pool* pool, const unsigned long long* keys,
                           const unsigned long long* keys_lo, int depth_lo, const long long* wprefix,
                           int* ctl, int* split_list, unsigned char* flag, int lo, int nboxes_host,
                           int level, int max_key_level, int maxw, int adaptive, int level_restrict,
                           int have_ext, int skip_if_no_regular, double root_extent, int phases,
                           cudaStream_t s)
{
    Pool<T, DIM> P = make_pool<T, DIM>(pool);
    const bool decide = phases & BT_STEP_DECIDE, create = phases & BT_STEP_CREATE,
               commit = phases & BT_STEP_COMMIT;
    // per-iteration control slots: NSPLIT, OVERSIZE, OVERFLOW, NSPLIT_REGULAR, COMMITTED
    if (decide) BT_CHECK(cudaMemsetAsync(ctl + BT_CTL_NSPLIT, 0, sizeof(int) * 5, s));
    else if (create) BT_CHECK(cudaMemsetAsync(ctl + BT_CTL_OVERFLOW, 0, sizeof(int), s));
    if (decide) {
        DecideIn<T, DIM> in{P, wprefix, ctl, flag, lo, level, maxw, adaptive, level_restrict};
        DecideOut out{flag, split_list, ctl, lo};
        BT_TRY(scan_exclusive((int64_t)(nboxes_host - lo), nullptr, in, out, s));
    }
    // at most every box in [lo, nboxes) splits
    const int64_t max_threads = (int64_t)(nboxes_host - lo) * (1 << DIM);
    if (create) {
        create_children_kernel<T, DIM><<<grid_for(max_threads, 256, 8), 256, 0, s>>>(
            P, keys, keys_lo, depth_lo, wprefix, split_list, ctl, pool->capacity, max_key_level, have_ext,
            maxw, skip_if_no_regular, (T)root_extent);
        BT_LAUNCH_CHECK();
    }
    if (commit) {
        if (P.xch) {
            children_finish_kernel<T, DIM><<<grid_for(max_threads, 256, 8), 256, 0, s>>>(
                P, split_list, ctl, maxw, skip_if_no_regular);
            BT_LAUNCH_CHECK();
        }
        commit_level_kernel<DIM><<<1, 32, 0, s>>>(ctl, skip_if_no_regular);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

template <typename T, int DIM>
static int level_restrict_impl(const bt_pool* pool, int* ctl, int built_level, int nboxes_upper,
                               double root_extent, cudaStream_t s)
{
    Pool<T, DIM> P = make_pool<T, DIM>(pool);
    BT_CHECK(cudaMemsetAsync(pool->force_split, 0, (size_t)pool->capacity, s));
    BT_CHECK(cudaMemsetAsync(ctl + BT_CTL_LR_FOUND, 0, sizeof(int) * BT_CTL_LR_SLOTS, s));
    bool first = true;
    for (int u = built_level - 2; u >= 1; --u) {
        level_restrict_kernel<T, DIM><<<grid_for(nboxes_upper, 128, 8), 128, 0, s>>>(
            P, ctl, u, first ? 1 : 0, (T)root_extent);
        BT_LAUNCH_CHECK();
        first = false;
    }
    return BT_OK;
}

template <typename T, int DIM>
static int finalize_boxes_impl(const bt_pool* pool, int nboxes, int level_restrict, int skip_prune,
                               int have_ext, int* ctl, int* map_old2new, int* src_of_new,
                               int* level_start, int phase, int nfinal, int aligned,
                               const bt_box_out* o, cudaStream_t s)
{
    Pool<T, DIM> P = make_pool<T, DIM>(pool);
    if (phase == 0) {
        int* order = nullptr;
        unsigned long long *k0 = nullptr, *k1 = nullptr; unsigned *v0 = nullptr, *v1 = nullptr;
        if (level_restrict) {
            // pool order -> level-major, stable: creation order inside a level
            BT_CHECK(bt::temp_alloc((void**)&k0, sizeof(unsigned long long) * nboxes * 2, s));
            BT_CHECK(bt::temp_alloc((void**)&v0, sizeof(unsigned) * nboxes * 2, s));
            k1 = k0 + nboxes; v1 = v0 + nboxes;
            level_keys_kernel<<<grid_for(nboxes, 256), 256, 0, s>>>(pool->level, nboxes, k0);
            BT_LAUNCH_CHECK();
            int in_alt = 0;
            BT_TRY(radix_sort_pairs(nboxes, k0, k1, v0, v1, 1, 0, 8, &in_alt, s));
            order = (int*)(in_alt ? v1 : v0);
        }
        // level_start: slots above the deepest level keep the final count (host fills)
        KeepIn in{order, P.gcount, skip_prune};
        KeepOut out{order, P.gcount, pool->level, skip_prune, map_old2new, src_of_new, level_start, ctl};
        BT_TRY(scan_exclusive(nboxes, nullptr, in, out, s));
        if (k0) { BT_CHECK(cudaFreeAsync(k0, s)); BT_CHECK(cudaFreeAsync(v0, s)); }
        return BT_OK;
    }
    gather_boxes_kernel<T, DIM><<<grid_for(nfinal, 256), 256, 0, s>>>(
        P, src_of_new, map_old2new, nfinal, aligned, have_ext, o->box_start, o->box_count, o->box_nonchild,
        o->box_levels, o->box_parent_ids, o->box_child_ids, (T*)o->box_centers, o->has_children,
        o->real_children, o->local_start, o->local_count, o->local_nonchild);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T, int DIM>
static int permute_impl(const bt_particles* p, const void* records, const int* from_ids, int n,
                        void* const* outs, void* out_radii, cudaStream_t s)
{
    if (n == 0) return BT_OK;
    Particles<T, DIM> P = make_particles<T, DIM>(p);
    permute_kernel<T, DIM><<<grid_for(n, 256, 8), 256, 0, s>>>(
        P, (const T*)records, from_ids, n, out_radii ? 1 : 0, (T*)outs[0], DIM > 1 ? (T*)outs[1] : nullptr,
        DIM > 2 ? (T*)outs[2] : nullptr, (T*)out_radii);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T, int DIM>
static int box_extents_impl(int nboxes, int aligned, int nlevels, const int* level_start_host,
                            const int* child_ids, const void* centers, const int* pstarts,
                            const int* pcounts, void* const* parts, const void* radii, void* bmin,
                            void* bmax, int phases, cudaStream_t s)
{
    if (nboxes <= 0) return BT_OK;
    if (phases & 1) {
    // list of the boxes with more than kExtHuge own particles ([0] = count)
    int* huge = nullptr;
    const size_t huge_cap = (size_t)nboxes;
    BT_CHECK(temp_alloc((void**)&huge, sizeof(int) * (huge_cap + 1), s));
    BT_CHECK(cudaMemsetAsync(huge, 0, sizeof(int), s));
    if (phases & 4)     // sparse boxes: one lane per box
        box_extents_own_kernel<T, DIM, 1><<<grid_for((int64_t)nboxes, 256, 8), 256, 0, s>>>(
            nboxes, aligned, (const T*)centers, pstarts, pcounts, (const T*)parts[0],
            DIM > 1 ? (const T*)parts[1] : nullptr, DIM > 2 ? (const T*)parts[2] : nullptr,
            (const T*)radii, (T*)bmin, (T*)bmax, huge + 1, huge);
    else
        box_extents_own_kernel<T, DIM, 8><<<grid_for((int64_t)nboxes * 8, 256, 8), 256, 0, s>>>(
            nboxes, aligned, (const T*)centers, pstarts, pcounts, (const T*)parts[0],
            DIM > 1 ? (const T*)parts[1] : nullptr, DIM > 2 ? (const T*)parts[2] : nullptr,
            (const T*)radii, (T*)bmin, (T*)bmax, huge + 1, huge);
    BT_LAUNCH_CHECK();
    box_extents_huge_kernel<T, DIM><<<(unsigned)(huge_cap < 4 * kNumSMs ? huge_cap : 4 * kNumSMs), 256, 0, s>>>(
        aligned, (const T*)centers, pstarts, pcounts, (const T*)parts[0],
        DIM > 1 ? (const T*)parts[1] : nullptr, DIM > 2 ? (const T*)parts[2] : nullptr,
        (const T*)radii, (T*)bmin, (T*)bmax, huge + 1, huge);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(huge, s));
    }
    if (!(phases & 2)) return BT_OK;
    for (int lev = nlevels - 2; lev >= 0; --lev) {       // the deepest level has no children
        const int start = level_start_host[lev], stop = level_start_host[lev + 1];
        if (stop <= start) continue;
        box_extents_merge_kernel<T, DIM><<<grid_for(stop - start, 128, 8), 128, 0, s>>>(
            start, stop, aligned, child_ids, (T*)bmin, (T*)bmax);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

}  // namespace bt

#define BT_DISPATCH(dtype, dim, FN, ...)                                             \
    do {                                                                             \
        if ((dtype) == BT_F32) {                                                     \
            if ((dim) == 1) return bt::FN<float, 1>(__VA_ARGS__);                     \
            if ((dim) == 2) return bt::FN<float, 2>(__VA_ARGS__);                     \
            if ((dim) == 3) return bt::FN<float, 3>(__VA_ARGS__);                     \
        } else if ((dtype) == BT_F64) {                                              \
            if ((dim) == 1) return bt::FN<double, 1>(__VA_ARGS__);                    \
            if ((dim) == 2) return bt::FN<double, 2>(__VA_ARGS__);                    \
            if ((dim) == 3) return bt::FN<double, 3>(__VA_ARGS__);                    \
        }                                                                            \
        return BT_ERR_BAD_ARG;                                                       \
    } while (0)

extern "C" {

long long bt_launch_count(void) { return bt::g_launch_count; }
void bt_prof_enable(int on) { bt::g_prof_enabled = on; }
void bt_prof_reset(void)
{
    for (auto& r : bt::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    bt::g_prof.clear();
}
// Writes "name\tcalls\ttotal_ms\n" lines (scopes aggregated by name) into buf; the
// caller must have synchronised the stream.  Returns the number of bytes needed.
int bt_prof_report(char* buf, int len)
{
    std::vector<std::string> names; std::vector<double> ms; std::vector<int> calls;
    for (auto& r : bt::g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
        size_t k = 0;
        for (; k < names.size(); ++k) if (names[k] == r.name) break;
        if (k == names.size()) { names.push_back(r.name); ms.push_back(0); calls.push_back(0); }
        ms[k] += t; calls[k] += 1;
    }
    std::string out;
    char line[256];
    for (size_t k = 0; k < names.size(); ++k) {
        snprintf(line, sizeof line, "%s\t%d\t%.6f\n", names[k].c_str(), calls[k], ms[k]);
        out += line;
    }
    if (buf && len > 0) { strncpy(buf, out.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
    return (int)out.size() + 1;
}

int bt_max_key_level(int dim)
{
    if (dim == 1) return 31;
    if (dim == 2) return 28;
    if (dim == 3) return 19;
    return 0;
}

int bt_bounding_box(int dtype, int dim, const bt_particles* p, void* out_minmax, void* stream)
{
    BT_PROF("bt_bounding_box", (cudaStream_t)stream); BT_DISPATCH(dtype, dim, bbox_impl, p, out_minmax, (cudaStream_t)stream); }

static int key_depth(int dim, int depth)
{   // levels resolved by the sort key: all the 64-bit key holds unless the caller asks for fewer
    const int dmax = bt_max_key_level(dim);
    return (depth > 0 && depth < dmax) ? depth : dmax;
}

int bt_max_tree_level(int dim) { (void)dim; return 31; }

int bt_make_keys(int dtype, int dim, const bt_particles* p, const double* bbox_min, const double* bbox_max,
                 int extent_norm, double stick_out_factor, int depth, uint64_t* keys, uint64_t* keys_lo,
                 void* records, void* stream)
{
    BT_PROF("bt_make_keys", (cudaStream_t)stream);
    // two-word keys: `keys` holds all the levels one word can, keys_lo the rest up to level 31
    const int Dh = keys_lo ? bt_max_key_level(dim) : key_depth(dim, depth);
    const int D = keys_lo ? bt_max_tree_level(dim) : Dh;
    if (keys_lo && D <= Dh) return BT_ERR_BAD_ARG;
    BT_DISPATCH(dtype, dim, make_keys_impl, p, bbox_min, bbox_max, extent_norm, stick_out_factor, D, Dh,
                (unsigned long long*)keys, (unsigned long long*)keys_lo, records, (cudaStream_t)stream);
}

int bt_sort_particles(int64_t n, int dim, int have_extent, int depth, uint64_t* keys, uint64_t* keys_alt,
                      uint32_t* ids, uint32_t* ids_alt, int* result_in_alt, void* stream)
{
    BT_PROF("bt_sort_particles", (cudaStream_t)stream);
    const int D = key_depth(dim, depth);
    const int begin_bit = have_extent ? 0 : bt::kStopBits;
    const int end_bit = bt::kStopBits + D * dim;
    return bt::radix_sort_pairs(n, (unsigned long long*)keys, (unsigned long long*)keys_alt, ids, ids_alt,
                                1, begin_bit, end_bit, result_in_alt, (cudaStream_t)stream);
}

namespace bt {
__global__ void gather_u64_kernel(int64_t n, const unsigned long long* __restrict__ src,
                                  const unsigned* __restrict__ idx, unsigned long long* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = src[idx[i]];
}
}  // namespace bt

// Stable sort by the two-word key (keys, keys_lo): LSD over the low word, then over the high
// word.  On return ids / keys / keys_lo hold the sorted order (no ping-pong flag); *_tmp are
// scratch of the same sizes.
int bt_sort_particles_deep(int64_t n, int dim, int have_extent, uint64_t* keys, uint64_t* keys_tmp,
                           uint64_t* keys_lo, uint64_t* keys_lo_tmp, uint32_t* ids, uint32_t* ids_tmp,
                           void* stream)
{
    BT_PROF("bt_sort_particles", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= 0) return BT_OK;
    typedef unsigned long long u64;
    const int Dh = bt_max_key_level(dim), Dl = bt_max_tree_level(dim) - Dh;
    if (Dl <= 0) return BT_ERR_BAD_ARG;
    const int grid = bt::grid_for(n, 256, 8);
    // 1. ids ordered by the low word (a copy of it is sorted, the original is gathered from later)
    BT_CHECK(cudaMemcpyAsync(keys_lo_tmp, keys_lo, sizeof(u64) * n, cudaMemcpyDeviceToDevice, s));
    u64* scratch = nullptr;
    BT_CHECK(bt::temp_alloc((void**)&scratch, sizeof(u64) * n, s));
    int in_alt = 0;
    BT_TRY(bt::radix_sort_pairs(n, (u64*)keys_lo_tmp, scratch, ids, ids_tmp, 1, have_extent ? 0 : bt::kStopBits,
                                bt::kStopBits + Dl * dim, &in_alt, s));
    uint32_t* ids1 = in_alt ? ids_tmp : ids;
    uint32_t* ids1_alt = in_alt ? ids : ids_tmp;
    // 2. the high words in that order, 3. stable sort by them (ids ride along)
    bt::gather_u64_kernel<<<grid, 256, 0, s>>>(n, (const u64*)keys, ids1, (u64*)keys_tmp);
    BT_LAUNCH_CHECK();
    int in_alt2 = 0;
    BT_TRY(bt::radix_sort_pairs(n, (u64*)keys_tmp, scratch, ids1, ids1_alt, 0, have_extent ? 0 : bt::kStopBits,
                                bt::kStopBits + Dh * dim, &in_alt2, s));
    const uint32_t* ids2 = in_alt2 ? ids1_alt : ids1;
    const u64* hi_sorted = in_alt2 ? scratch : (u64*)keys_tmp;
    // 4. results into the caller's primary buffers
    bt::gather_u64_kernel<<<grid, 256, 0, s>>>(n, (const u64*)keys_lo, ids2, (u64*)keys_lo_tmp);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaMemcpyAsync(keys_lo, keys_lo_tmp, sizeof(u64) * n, cudaMemcpyDeviceToDevice, s));
    BT_CHECK(cudaMemcpyAsync(keys, hi_sorted, sizeof(u64) * n, cudaMemcpyDeviceToDevice, s));
    if (ids2 != ids) BT_CHECK(cudaMemcpyAsync(ids, ids2, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
    BT_CHECK(cudaFreeAsync(scratch, s));
    return BT_OK;
}

int bt_weight_prefix(int64_t n, const uint32_t* sorted_ids, const int32_t* weights, int64_t* wprefix,
                     void* stream)
{
    BT_PROF("bt_weight_prefix", (cudaStream_t)stream);
    bt::WeightIn in{sorted_ids, weights};
    bt::WeightOut out{(long long*)wprefix, n};
    return bt::scan_exclusive(n, nullptr, in, out, (cudaStream_t)stream);
}

int bt_pool_init(int dtype, int dim, const bt_pool* pool, int64_t n, int have_extent,
                 const uint64_t* keys, const double* root_center, int32_t* ctl, void* stream)
{
    BT_PROF("bt_pool_init", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, pool_init_impl, pool, n, have_extent, (const unsigned long long*)keys,
                root_center, ctl, (cudaStream_t)stream);
}

int bt_level_step(int dtype, int dim, const bt_pool* pool, const uint64_t* keys, const int64_t* wprefix,
                  int32_t* ctl, int32_t* split_list, uint8_t* flag, int lo, int nboxes, int level,
                  int maxw, int adaptive, int level_restrict, int have_extent, int skip_if_no_regular,
                  double root_extent, int phases, int depth, const uint64_t* keys_lo, void* stream)
{
    BT_PROF("bt_level_step", (cudaStream_t)stream);
    const int Dh = keys_lo ? bt_max_key_level(dim) : key_depth(dim, depth);
    BT_DISPATCH(dtype, dim, level_step_impl, pool, (const unsigned long long*)keys,
                (const unsigned long long*)keys_lo, keys_lo ? bt_max_tree_level(dim) - Dh : 0,
                (const long long*)wprefix, ctl, split_list, flag, lo, nboxes, level,
                Dh, maxw, adaptive, level_restrict, have_extent, skip_if_no_regular,
                root_extent, phases, (cudaStream_t)stream);
}

int bt_level_restrict(int dtype, int dim, const bt_pool* pool, int32_t* ctl, int built_level,
                      int nboxes_upper, double root_extent, void* stream)
{
    BT_PROF("bt_level_restrict", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, level_restrict_impl, pool, ctl, built_level, nboxes_upper, root_extent,
                (cudaStream_t)stream);
}

int bt_finalize_numbering(int dtype, int dim, const bt_pool* pool, int nboxes, int level_restrict,
                          int skip_prune, int32_t* ctl, int32_t* map_old2new, int32_t* src_of_new,
                          int32_t* level_start, void* stream)
{
    BT_PROF("bt_finalize_numbering", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, finalize_boxes_impl, pool, nboxes, level_restrict, skip_prune, 0, ctl,
                map_old2new, src_of_new, level_start, 0, 0, 0, nullptr, (cudaStream_t)stream);
}

int bt_gather_boxes(int dtype, int dim, const bt_pool* pool, int have_extent, const int32_t* src_of_new,
                    const int32_t* map_old2new, int nfinal, int aligned, const bt_box_out* out,
                    void* stream)
{
    BT_PROF("bt_gather_boxes", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, finalize_boxes_impl, pool, 0, 0, 0, have_extent, nullptr,
                (int*)map_old2new, (int*)src_of_new, nullptr, 1, nfinal, aligned, out,
                (cudaStream_t)stream);
}

int bt_leaf_fixup(int nboxes, const int32_t* box_start, const int32_t* box_count,
                  const uint8_t* real_children, uint32_t* ids, int32_t* ctl, int32_t* big_list,
                  int big_cap, int32_t* huge_list, void* stream)
{
    BT_PROF("bt_leaf_fixup", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    bt::leaf_fixup_kernel<<<bt::grid_for((int64_t)nboxes, bt::kFixBlock, 8), bt::kFixBlock, 0, s>>>(
        box_start, box_count, real_children, nboxes, ids, big_list, ctl, big_cap);
    BT_LAUNCH_CHECK();
    bt::leaf_fixup_big_kernel<<<bt::kNumSMs * 2, bt::kFixBlock, 0, s>>>(
        box_start, box_count, big_list, ctl, big_cap, ids, huge_list);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_sort_u32_segment(int64_t n, uint32_t* ids, void* stream)
{
    BT_PROF("bt_sort_u32_segment", (cudaStream_t)stream);
    // ascending sort of one oversized leaf's ids (keys = ids widened to 64 bit)
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 2) return BT_OK;
    unsigned long long* k = nullptr; unsigned* v = nullptr;
    BT_CHECK(bt::temp_alloc((void**)&k, sizeof(unsigned long long) * n * 2, s));
    BT_CHECK(bt::temp_alloc((void**)&v, sizeof(unsigned) * n * 2, s));
    bt::widen_ids_kernel<<<bt::grid_for(n, 256), 256, 0, s>>>(ids, (int)n, k);
    BT_LAUNCH_CHECK();
    int in_alt = 0;
    BT_TRY(bt::radix_sort_pairs(n, k, k + n, v, v + n, 1, 0, 32, &in_alt, s));
    bt::narrow_ids_kernel<<<bt::grid_for(n, 256), 256, 0, s>>>(in_alt ? (k + n) : k, (int)n, ids);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(k, s));
    BT_CHECK(cudaFreeAsync(v, s));
    return BT_OK;
}

int bt_split_sources_targets(int64_t n, int64_t nsources, const uint32_t* sorted_ids,
                             int32_t* source_numbers, int32_t* user_source_ids,
                             int32_t* srcntgt_target_ids, int32_t* sorted_target_ids, void* stream)
{
    BT_PROF("bt_split_sources_targets", (cudaStream_t)stream);
    bt::SourceIn in{sorted_ids, (unsigned)nsources};
    bt::SourceOut out{sorted_ids, (unsigned)nsources, source_numbers, user_source_ids,
                      srcntgt_target_ids, sorted_target_ids, n};
    return bt::scan_exclusive(n, nullptr, in, out, (cudaStream_t)stream);
}

int bt_reverse_index(int64_t n, const uint32_t* ids, int32_t* out, void* stream)
{
    BT_PROF("bt_reverse_index", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::reverse_index_kernel<<<bt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(ids, (int)n, out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_permute(int dtype, int dim, const bt_particles* p, const void* records, const int32_t* from_ids,
               int64_t n, void* const* outs, void* out_radii, void* stream)
{
    BT_PROF("bt_permute", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, permute_impl, p, records, from_ids, (int)n, outs, out_radii, (cudaStream_t)stream);
}

int bt_box_info(int nboxes, int sources_are_targets, int have_extent, const int32_t* box_start,
                const int32_t* box_count, const int32_t* box_nonchild, const uint8_t* has_children,
                const int32_t* source_numbers, int32_t* src_starts, int32_t* src_nonchild,
                int32_t* src_cumul, int32_t* tgt_starts, int32_t* tgt_nonchild, int32_t* tgt_cumul,
                uint8_t* box_flags, void* stream)
{
    BT_PROF("bt_box_info", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::box_info_kernel<<<bt::grid_for(nboxes, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nboxes, sources_are_targets, have_extent, box_start, box_count, box_nonchild, has_children,
        source_numbers, src_starts, src_nonchild, src_cumul, tgt_starts, tgt_nonchild, tgt_cumul,
        box_flags);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_box_extents(int dtype, int dim, int nboxes, int aligned, int nlevels,
                   const int32_t* level_start_box_nrs_host, const int32_t* box_child_ids,
                   const void* box_centers, const int32_t* pstarts, const int32_t* pcounts,
                   void* const* particles, const void* radii, void* bb_min, void* bb_max, void* stream)
{
    BT_PROF("bt_box_extents", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, box_extents_impl, nboxes, aligned, nlevels, level_start_box_nrs_host,
                box_child_ids, box_centers, pstarts, pcounts, particles, radii, bb_min, bb_max, 3,
                (cudaStream_t)stream);
}

int bt_box_extents_phase(int dtype, int dim, int nboxes, int aligned, int nlevels,
                         const int32_t* level_start_box_nrs_host, const int32_t* box_child_ids,
                         const void* box_centers, const int32_t* pstarts, const int32_t* pcounts,
                         void* const* particles, const void* radii, void* bb_min, void* bb_max,
                         int phases, void* stream)
{
    BT_PROF("bt_box_extents", (cudaStream_t)stream);
    BT_DISPATCH(dtype, dim, box_extents_impl, nboxes, aligned, nlevels, level_start_box_nrs_host,
                box_child_ids, box_centers, pstarts, pcounts, particles, radii, bb_min, bb_max, phases,
                (cudaStream_t)stream);
}

int bt_box_info_local(int nboxes, int have_extent, const int32_t* local_start, const int32_t* local_count,
                      const int32_t* local_nonchild, const uint8_t* has_children,
                      const int32_t* source_numbers, int32_t* src3, int32_t* src_starts,
                      int32_t* src_nonchild, int32_t* src_cumul, int32_t* tgt_starts,
                      int32_t* tgt_nonchild, int32_t* tgt_cumul, void* stream)
{
    BT_PROF("bt_box_info", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::box_info_local_kernel<<<bt::grid_for(nboxes, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nboxes, have_extent, local_start, local_count, local_nonchild, has_children, source_numbers,
        src3, src_starts, src_nonchild, src_cumul, tgt_starts, tgt_nonchild, tgt_cumul);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_box_info_global(int nboxes, int have_extent, const int32_t* box_start, const int32_t* box_count,
                       const int32_t* box_nonchild, const uint8_t* has_children, const int32_t* src3,
                       int32_t* src_starts, int32_t* src_nonchild, int32_t* src_cumul,
                       int32_t* tgt_starts, int32_t* tgt_nonchild, int32_t* tgt_cumul,
                       uint8_t* box_flags, void* stream)
{
    BT_PROF("bt_box_info", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::box_info_global_kernel<<<bt::grid_for(nboxes, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nboxes, have_extent, box_start, box_count, box_nonchild, has_children, src3, src_starts,
        src_nonchild, src_cumul, tgt_starts, tgt_nonchild, tgt_cumul, box_flags);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // extern "C"
