// FMM traversal list builders for sm_100a.
//
// Every builder follows the reference's ListOfListsBuilder protocol (count
// pass, exclusive scan to `starts`, write pass; row content in the APPEND order
// of the row's depth-first, Morton-ordered walk) so the CSR arrays are
// identical to the reference's.  The walks restate boxtree/traversal.py:98-160
// (walk_init / walk_advance / walk_push) and the per-list generate() bodies
// cited at each kernel.  List 3 is produced for all source levels by ONE walk
// that bins appends by the level of the appended box (the reference repeats the
// walk nlevels(+1) times, traversal.py:2203-2228).
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"
#include "../../include/boxtree_b200.h"

namespace bt {

constexpr int kMaxWalkLevels = 40;
constexpr int kTravBlock = 128;

template <typename T, int DIM>
struct TreeView {
    const T* centers; const unsigned char* levels; const int* child_ids; const int* child_t;
    const unsigned char* flags; const int* parents;
    int aligned; int nboxes; int nlevels; T root_extent; int n_away;
    __device__ __forceinline__ void center(int b, T* c) const
    {
#pragma unroll
        for (int a = 0; a < DIM; ++a) c[a] = centers[aligned * a + b];
    }
    __device__ __forceinline__ int child(int parent, int mnr) const
    { return child_t ? child_t[((int64_t)parent << DIM) + mnr] : child_ids[mnr * aligned + parent]; }
};

template <typename T, int DIM>
static TreeView<T, DIM> make_view(const bt_tree_view* v)
{
    TreeView<T, DIM> t;
    t.centers = (const T*)v->box_centers; t.levels = v->box_levels; t.child_ids = v->box_child_ids;
    t.flags = v->box_flags; t.parents = v->box_parent_ids; t.aligned = v->aligned_nboxes;
    t.nboxes = v->nboxes; t.nlevels = v->nlevels; t.root_extent = (T)v->root_extent;
    t.n_away = v->well_sep_is_n_away;
    t.child_t = v->box_child_ids_t;
    return t;
}

// LEVEL_TO_RAD(level) = root_extent * 1 / (coord_t)(1 << (level + 1)), traversal.py:234-235
template <typename T>
__device__ __forceinline__ void fill_rad_table(T* rad, T root_extent)
{
    for (int l = threadIdx.x; l < kMaxWalkLevels; l += blockDim.x)
        rad[l] = (root_extent * 1 / (T)(1 << ((l + 1) & 31)));
    __syncthreads();
}

// is_adjacent_or_overlapping_with_neighborhood, traversal.py:279-305
template <typename T, int DIM>
__device__ __forceinline__ bool adj_nbhd(const T* rad, const T* tc, int tl, T nbhd, const T* sc, int sl)
{
    const T target_rad = rad[tl], source_rad = rad[sl];
    const T rad_sum = ((2 * (nbhd - 1) + 1) * target_rad + source_rad);
    const T slack = rad_sum + fmin(target_rad, source_rad);
    T l_inf_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) l_inf_dist = fmax(l_inf_dist, fabs(tc[a] - sc[a]));
    return l_inf_dist <= slack;
}

struct Walk {
    int box_stack[kMaxWalkLevels]; signed char mnr_stack[kMaxWalkLevels];
    int ssize, parent, mnr; bool cont;
    __device__ __forceinline__ void init(int start) { ssize = 0; parent = start; mnr = 0; cont = true; }
    __device__ __forceinline__ void push(int nb)
    { box_stack[ssize] = parent; mnr_stack[ssize] = (signed char)mnr; ++ssize; parent = nb; mnr = 0; }
    template <int NB> __device__ __forceinline__ void advance()
    {
        while (true) {
            ++mnr;
            if (mnr < NB) break;
            cont = (ssize > 0);
            if (cont) { --ssize; parent = box_stack[ssize]; mnr = mnr_stack[ssize]; }
            else break;
        }
    }
};

// emitters
struct CountEmit {
    int c0 = 0, c1 = 0;
    __device__ __forceinline__ void e0(int) { ++c0; }
    __device__ __forceinline__ void e1(int) { ++c1; }
};
struct FillEmit {
    int* p0; int* p1;
    __device__ __forceinline__ void e0(int v) { *p0++ = v; }
    __device__ __forceinline__ void e1(int v) { if (p1) *p1++ = v; }
};

// ---- b3 colleagues: traversal.py:398-464 -----------------------------------
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_colleagues(const TreeView<T, DIM>& t, const T* rad, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    if (box_id == 0) return;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    const T nbhd = (T)t.n_away;
    Walk w; w.init(0);
    while (w.cont) {
        const int wb = t.child(w.parent, w.mnr);
        if (wb) {
            T wc[DIM]; t.center(wb, wc);
            // The reference also descends into box_id itself (traversal.py:438-452) and
            // walks its whole subtree, which can never append (deeper levels only):
            // skipping that descent leaves the output unchanged and removes a serial
            // walk over up to nboxes/2^d boxes by a single thread.
            if (wb != box_id && adj_nbhd<T, DIM>(rad, center, level, nbhd, wc, t.levels[wb])) {
                if (w.ssize + 1 == level) e.e0(wb);
                else { w.push(wb); continue; }
            }
        }
        w.template advance<NB>();
    }
}

// ---- b4 list 1: traversal.py:470-550 ---------------------------------------
constexpr int kVisitPush = 1, kVisitEmit = 2, kVisitClose = 4, kVisitNear = 8;

// one child visit of the list-1 walk (traversal.py:508-539)
template <typename T, int DIM>
__device__ __forceinline__ int list1_visit(const TreeView<T, DIM>& t, const T* rad, const T* center,
                                           int level, int wb)
{
    if (!wb) return 0;
    T wc[DIM]; t.center(wb, wc);
    if (!adj_nbhd<T, DIM>(rad, center, level, (T)1, wc, t.levels[wb])) return 0;
    const unsigned char fl = t.flags[wb];
    return ((fl & BT_BOX_IS_SOURCE_BOX) ? kVisitEmit : 0) |
           ((fl & BT_BOX_HAS_SOURCE_CHILD_BOXES) ? kVisitPush : 0);
}

// returns false when the walk exceeded `budget` child visits (row becomes "heavy")
template <typename T, int DIM, class E>
__device__ __forceinline__ bool gen_list1(const TreeView<T, DIM>& t, const T* rad, int box_id, E& e,
                                          int budget)
{
    constexpr int NB = 1 << DIM;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    if (t.flags[0] & BT_BOX_IS_SOURCE_BOX) e.e0(0);
    Walk w; w.init(0);
    int visits = 0;
    while (w.cont) {
        if (++visits > budget) return false;
        const int wb = t.child(w.parent, w.mnr);
        const int act = list1_visit<T, DIM>(t, rad, center, level, wb);
        if (act & kVisitEmit) e.e0(wb);
        if (act & kVisitPush) { w.push(wb); continue; }
        w.template advance<NB>();
    }
    return true;
}

// ---- cooperative walks --------------------------------------------------------
// The reference walks a row with one work-item.  Here a row is walked by a GROUP of 2^d
// lanes (32 / 2^d rows per warp): when a walk node is expanded, lane m evaluates child m
// (the same child visit as the reference, `list*_visit`), ballots turn the results into
// bit masks, and the masks are consumed in Morton order -- a child's append precedes the
// descent into it, later children wait in a per-group shared-memory stack -- so the
// APPEND order is exactly the order of the reference's depth-first walk while the
// dependent-load chain per row is ~2^d times shorter and appends of sibling boxes are
// written by several lanes at once.
// Which mapping each builder uses (bit set = group/warp-cooperative, clear = one thread per
// row); the defaults are the faster choice measured on B200 (profiles/README.md).
constexpr int kModeColl = 1, kModeList1 = 2, kModeList3 = 4, kModeList3Auto = 8,
              kModeList2Count = 16, kModeList2Fill = 32, kModeCollTopDown = 64, kModeFused13 = 128,
              kModeNearCapZero = 256;   // testing: rows with near-field boxes above their level go heavy
// (bit 512, heavy rows by radix sort instead of the position map, is read by the host:
//  it leaves bt_heavy_ws.hrow_base NULL)
int g_walk_mode = kModeList1 | kModeList3Auto | kModeList2Fill | kModeCollTopDown | kModeFused13;

struct CoopFrame { int parent; unsigned bits; };

// Policy P (one object per lane, group-uniform state):
//   void init(int row, bool valid);               (called by ALL lanes of the warp)
//   bool next_root(int& parent); int visit(int wb);
//   void emit(unsigned eb, unsigned cb, unsigned qb, int c, int gl, int parent);   (ALL lanes;
//        eb/cb/qb = children whose visit returned kVisitEmit / kVisitClose / kVisitNear)
//   void finish(int row, bool valid, bool ok, int gl);                (ALL lanes)
template <typename T, int DIM, class P>
__device__ __forceinline__ void coop_walk_rows(const TreeView<T, DIM>& t, P& pol, int nrows, int budget,
                                               CoopFrame* frames /* [groups per block][kMaxWalkLevels] */)
{
    constexpr int NB = 1 << DIM;
    constexpr int GPW = 32 / NB;
    constexpr unsigned gmask = (1u << NB) - 1u;
    const int lane = threadIdx.x & 31, g = lane / NB, gl = lane % NB, gshift = g * NB;
    CoopFrame* stack = frames + (threadIdx.x / NB) * kMaxWalkLevels;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int max_exp = budget / NB + 1;
    for (int rbase = warp_global * GPW; rbase < nrows; rbase += nwarps * GPW) {
        const int row = rbase + g;
        bool active = row < nrows, ok = true, need_eval = false, reload = false;
        int parent = 0, sp = 0, c = 0, expansions = 0;
        unsigned emit = 0, close = 0, push = 0, near = 0;
        pol.init(row, active);
        if (active) {
            active = pol.next_root(parent);
            need_eval = active;
        }
        while (__any_sync(0xffffffffu, active)) {
            int act = 0;
            if (active && (need_eval || reload)) c = t.child(parent, gl);
            if (active && need_eval) act = pol.visit(c);
            reload = false;
            const unsigned be = (__ballot_sync(0xffffffffu, act & kVisitEmit) >> gshift) & gmask;
            const unsigned bc = (__ballot_sync(0xffffffffu, act & kVisitClose) >> gshift) & gmask;
            const unsigned bp = (__ballot_sync(0xffffffffu, act & kVisitPush) >> gshift) & gmask;
            const unsigned bq = (__ballot_sync(0xffffffffu, act & kVisitNear) >> gshift) & gmask;
            if (active && need_eval) {
                emit = be; close = bc; push = bp; near = bq; need_eval = false;
                if (++expansions > max_exp) { ok = false; active = false; }
            }
            const int first_push = push ? (__ffs(push) - 1) : NB;
            const unsigned lowmask = (first_push >= NB) ? gmask : ((2u << first_push) - 1u);
            const unsigned eb = active ? (emit & lowmask) : 0u, cb = active ? (close & lowmask) : 0u;
            const unsigned qb = active ? (near & lowmask) : 0u;
            pol.emit(eb, cb, qb, c, gl, parent);
            emit &= ~eb; close &= ~cb; near &= ~qb;
            const int cm = __shfl_sync(0xffffffffu, c, gshift + (first_push < NB ? first_push : 0));
            if (active) {
                if (first_push < NB) {
                    push &= ~(1u << first_push);
                    if (emit | close | push | near) {
                        stack[sp].parent = parent;
                        stack[sp].bits = emit | (close << 8) | (push << 16) | (near << 24);
                        ++sp;
                    }
                    emit = close = push = near = 0;
                    parent = cm; need_eval = true;
                } else if (sp > 0) {
                    --sp;
                    parent = stack[sp].parent;
                    const unsigned b = stack[sp].bits;
                    emit = b & 0xffu; close = (b >> 8) & 0xffu; push = (b >> 16) & 0xffu; near = b >> 24;
                    reload = true;
                } else {
                    active = pol.next_root(parent);
                    need_eval = active;
                }
            }
        }
        pol.finish(row, row < nrows, ok, gl);
    }
}

// colleagues (traversal.py:398-464)
template <typename T, int DIM, bool FILL>
struct CollPolicy {
    const TreeView<T, DIM>& t; const T* rad; int* starts; int* lists; const signed char* row_mask;
    T center[DIM]; int box, level, count; bool rooted; T nbhd; int* out;
    __device__ CollPolicy(const TreeView<T, DIM>& t_, const T* rad_, int* st, int* li, const signed char* rm)
        : t(t_), rad(rad_), starts(st), lists(li), row_mask(rm) {}
    __device__ __forceinline__ void init(int row, bool valid)
    {
        count = 0; rooted = false; nbhd = (T)t.n_away; box = 0; level = 0; out = nullptr;
        if (!valid) return;
        box = row; t.center(box, center); level = t.levels[box];
        out = FILL ? lists + starts[row] : nullptr;
        if (row_mask && !row_mask[box]) rooted = true;      // row not needed: empty list
    }
    __device__ __forceinline__ bool next_root(int& parent)
    { if (rooted || box == 0) return false; rooted = true; parent = 0; return true; }
    __device__ __forceinline__ int visit(int wb) const
    {
        if (!wb || wb == box) return 0;     // no descent into the box's own subtree (see gen_colleagues)
        T wc[DIM]; t.center(wb, wc);
        const int wl = t.levels[wb];
        if (!adj_nbhd<T, DIM>(rad, center, level, nbhd, wc, wl)) return 0;
        return (wl == level) ? kVisitEmit : kVisitPush;
    }
    __device__ __forceinline__ void emit(unsigned eb, unsigned, unsigned, int c, int gl, int)
    {
        if (FILL && ((eb >> gl) & 1u)) out[count + __popc(eb & ((1u << gl) - 1u))] = c;
        count += __popc(eb);
    }
    __device__ __forceinline__ void finish(int row, bool valid, bool, int gl)
    { if (valid && !FILL && gl == 0) starts[row] = count; }
};

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
coll_coop_kernel(TreeView<T, DIM> t, int nrows, int* __restrict__ starts, int* __restrict__ lists,
                 const signed char* __restrict__ row_mask)
{
    __shared__ T rad[kMaxWalkLevels];
    __shared__ CoopFrame frames[(kTravBlock >> DIM) * kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    CollPolicy<T, DIM, FILL> pol(t, rad, starts, lists, row_mask);
    coop_walk_rows<T, DIM>(t, pol, nrows, 0x7ffffff0, frames);
}

// ---- b5 list 2: traversal.py:556-601 ---------------------------------------
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list2(const TreeView<T, DIM>& t, const T* rad, const int* coll_starts,
                                          const int* coll_lists, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    const int parent = t.parents[box_id];
    if (parent == box_id) return;
    const T nbhd = (T)t.n_away;
    const int s = coll_starts[parent], en = coll_starts[parent + 1];
    for (int i = s; i < en; ++i) {
        const int parent_nf = coll_lists[i];
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const int sib = t.child(parent_nf, m);
            if (sib == 0) continue;
            T sc[DIM]; t.center(sib, sc);
            if (!adj_nbhd<T, DIM>(rad, center, level, nbhd, sc, t.levels[sib])) e.e0(sib);
        }
    }
}

// list 2, one warp per row: lane k tests candidate k = (colleague of the parent, Morton
// child) -- the order of the reference's two loops -- and the separated ones are written
// with one coalesced store per 32 candidates.
template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(256)
list2_warp_kernel(TreeView<T, DIM> t, const int* __restrict__ row_boxes, const int* __restrict__ coll_starts,
                  const int* __restrict__ coll_lists, int nrows, int* __restrict__ starts,
                  int* __restrict__ lists)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const T nbhd = (T)t.n_away;
    for (int r = w; r < nrows; r += nw) {
        const int box = row_boxes[r];
        const int parent = t.parents[box];
        int pos = FILL ? starts[r] : 0;
        if (parent != box) {
            T center[DIM]; t.center(box, center);
            const int level = t.levels[box];
            const int s = coll_starts[parent];
            const int ncand = (coll_starts[parent + 1] - s) * NB;
            for (int k0 = 0; k0 < ncand; k0 += 32) {
                const int k = k0 + lane;
                bool sep = false; int sib = 0;
                if (k < ncand) {
                    sib = t.child(coll_lists[s + k / NB], k % NB);
                    if (sib) {
                        T sc[DIM]; t.center(sib, sc);
                        sep = !adj_nbhd<T, DIM>(rad, center, level, nbhd, sc, t.levels[sib]);
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sep);
                if (FILL && sep) lists[pos + __popc(bal & ((1u << lane) - 1u))] = sib;
                pos += __popc(bal);
            }
        }
        if (!FILL && lane == 0) starts[r] = pos;
    }
}

static int counts_to_starts(int* a, int64_t n, long long* total_out, cudaStream_t s);

// ---- b3 colleagues, top-down (+ list-2 counts) -------------------------------
// The reference finds the colleagues of every box with a walk from the root
// (traversal.py:398-464).  The set it produces is {c on b's level, c != b, adjacent to b with
// the n-away neighbourhood}, in depth-first (Morton) order.  Every such c is a child of a
// colleague of parent(b) or of parent(b) itself (|i_b - i_c| <= n implies the same for the
// parents), so the lists are built level by level from the parent's list: one warp per box
// first applies the reference's descend test to the |coll(parent)| + 1 parent-level boxes
// (one per lane), then tests the children of those that pass -- densely packed over the lanes
// -- with the reference's predicate; parent(b) is merged into its colleague list at its
// depth-first rank, so candidates are visited -- and appended -- in the reference's order.
// The candidates that are NOT adjacent and stem from a colleague of the parent (all children
// of the parent-level boxes that fail the descend test, plus the failing children of the
// others) are exactly list 2 of b (traversal.py:556-601): count and membership mask come for free.  Rows are staged with a fixed stride of (2n+1)^d - 1 entries
// and compacted to CSR afterwards.
constexpr int kXfCollSource = 1;   // b or one of its colleagues is a source box
constexpr int kXfHasChild = 2;     // b has at least one child

constexpr int kCollMaskWordsMax = 40;     // (2n+1)^d * 2^d / 32 + 1 for n <= 2 in 3-D

// One warp per PARENT box p: the candidates of all 2^d children of p are the same boxes (the
// children of p and of p's colleagues), so each candidate's id and centre is loaded once and
// tested against every child of p.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
coll_topdown_kernel(TreeView<T, DIM> t, const int* __restrict__ level_start, int lev, int stride,
                    int* __restrict__ tmp, int* __restrict__ counts, const int* __restrict__ dfs_rank,
                    const signed char* __restrict__ row_mask, int* __restrict__ l2cnt,
                    unsigned char* __restrict__ xflags, unsigned* __restrict__ l2mask, int mask_words)
{
    constexpr int NB = 1 << DIM;
    constexpr unsigned nbmask = (1u << NB) - 1u;
    __shared__ T rad[kMaxWalkLevels];
    __shared__ T bcen_all[8][NB][DIM];            // centres of p's children, per warp
    __shared__ unsigned char padj_all[8][128];    // per parent-level box: which children of p it can touch
    fill_rad_table(rad, t.root_extent);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    T (*bcen)[DIM] = bcen_all[wib];
    unsigned char* padj = padj_all[wib];
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const T nbhd = (T)t.n_away;
    if (lev == 0) {                                // the root: no colleagues
        if (w == 0) {
            const int mychild = (lane < NB) ? t.child(0, lane) : 0;
            const bool has_child = __any_sync(0xffffffffu, mychild != 0);
            if (lane == 0) {
                counts[0] = 0; if (l2cnt) l2cnt[0] = 0;
                xflags[0] = (unsigned char)((has_child ? kXfHasChild : 0) |
                                            ((t.flags[0] & BT_BOX_IS_SOURCE_BOX) ? kXfCollSource : 0));
            }
        }
        return;
    }
    const int lo = level_start[lev - 1], hi = level_start[lev];     // parents
    for (int p = lo + w; p < hi; p += nw) {
        // children of p (lane m < 2^d holds child m), their own child bit and source bit
        const int bch = (lane < NB) ? t.child(p, lane) : 0;
        const unsigned chm = __ballot_sync(0xffffffffu, bch != 0) & nbmask;
        if (!chm) continue;
        // row wanted (not masked out)?  With a row mask (distributed setup: 7 of 8 parents on 8
        // ranks) a parent none of whose children's rows are wanted costs two loads: the empty
        // counts and the flags the walks read of ANY box are coll_init_kernel's
        const bool brow = bch && !(row_mask && !row_mask[bch]);
        const unsigned rowm = __ballot_sync(0xffffffffu, brow) & nbmask;
        if (!rowm) continue;
        unsigned char bxf = 0;
        if (bch) {
            bool hc = false;
#pragma unroll
            for (int m = 0; m < NB; ++m) hc = hc || (t.child(bch, m) != 0);
            bxf = (unsigned char)((hc ? kXfHasChild : 0) | ((t.flags[bch] & BT_BOX_IS_SOURCE_BOX) ? kXfCollSource : 0));
#pragma unroll
            for (int a = 0; a < DIM; ++a) bcen[lane][a] = t.centers[t.aligned * a + bch];
        }
        const int level = lev;
        const int np = counts[p];
        const int* prow = tmp + (int64_t)p * stride;
        const int prank = dfs_rank[p];
        int pos = 0;
        for (int j = lane; j < np; j += 32) pos += (dfs_rank[prow[j]] < prank) ? 1 : 0;
        pos = __reduce_add_sync(0xffffffffu, pos);
        const int nparents = np + 1;               // p merged into its colleagues at its rank
        __syncwarp();
        // (1) which children of p can a parent-level box P touch?  The reference's descend test
        //     (:429-452) of P against every child
        for (int j = lane; j < nparents; j += 32) {
            unsigned bits = 0;
            if (j == pos) bits = nbmask;
            else {
                const int P = prow[j - (j > pos ? 1 : 0)];
                T pc[DIM]; t.center(P, pc);
#pragma unroll
                for (int m = 0; m < NB; ++m)
                    if ((rowm >> m) & 1u)
                        bits |= adj_nbhd<T, DIM>(rad, bcen[m], level, nbhd, pc, level - 1) ? (1u << m) : 0u;
            }
            padj[j] = (unsigned char)bits;
        }
        __syncwarp();
        int out[NB], n2[NB];
        unsigned srcany = 0;
#pragma unroll
        for (int m = 0; m < NB; ++m) { out[m] = 0; n2[m] = 0; }
        const int ncand = nparents * NB;
        for (int k0 = 0; k0 < ncand; k0 += 32) {
            // (2) candidate k = (k / 2^d-th box of the merged parent list, Morton child k % 2^d)
            const int k = k0 + lane;
            const int j = k / NB;
            int c = 0; bool fromcoll = false; unsigned mypadj = 0;
            if (k < ncand) {
                const int P = (j == pos) ? p : prow[j - (j > pos ? 1 : 0)];
                fromcoll = (j != pos);
                c = t.child(P, k % NB);
                mypadj = padj[j];
            }
            T cc[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) cc[a] = c ? t.centers[t.aligned * a + c] : (T)0;
            const bool csrc = c && (t.flags[c] & BT_BOX_IS_SOURCE_BOX);
            const unsigned exm = __ballot_sync(0xffffffffu, c != 0 && fromcoll);   // list-2 bits of far children
            // children of p some candidate of this chunk can touch
            unsigned touch = mypadj;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) touch |= __shfl_xor_sync(0xffffffffu, touch, o);
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (!((rowm >> m) & 1u)) continue;
                const int b = __shfl_sync(0xffffffffu, bch, m);
                unsigned bs;
                if ((touch >> m) & 1u) {
                    bool adj = false;
                    if (c && c != b && ((mypadj >> m) & 1u))
                        adj = adj_nbhd<T, DIM>(rad, bcen[m], level, nbhd, cc, level);
                    const unsigned ba = __ballot_sync(0xffffffffu, adj);
                    if (adj) {
                        const int slot = out[m] + __popc(ba & ((1u << lane) - 1u));
                        if (slot < stride) tmp[(int64_t)b * stride + slot] = c;
                    }
                    out[m] += __popc(ba);
                    if (__ballot_sync(0xffffffffu, adj && csrc)) srcany |= 1u << m;
                    bs = exm & ~ba;                 // candidates of colleagues that are not adjacent
                } else bs = exm;
                n2[m] += __popc(bs);
                if (l2mask && lane == 0) l2mask[(int64_t)b * mask_words + (k0 >> 5)] = bs;
            }
        }
        // results of child m are written by lane m
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (lane == m && bch) {
                const bool row = (rowm >> m) & 1u;
                counts[bch] = row ? (out[m] < stride ? out[m] : stride) : 0;
                if (l2cnt) l2cnt[bch] = row ? n2[m] : 0;
                xflags[bch] = (unsigned char)(bxf | (((srcany >> m) & 1u) ? kXfCollSource : 0));
                if (l2mask && row) l2mask[(int64_t)bch * mask_words + mask_words - 1] = (unsigned)pos;
            }
        }
        __syncwarp();
    }
}

// With a row mask: empty rows and the structural flags (has a child; is a source box) of EVERY box
// in one streaming pass, so that the per-level kernel can leave a parent whose children are all
// masked out after testing the mask.  Rows that are wanted overwrite their entries.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
coll_init_kernel(TreeView<T, DIM> t, int* __restrict__ counts, int* __restrict__ l2cnt,
                 unsigned char* __restrict__ xflags)
{
    constexpr int NB = 1 << DIM;
    const int stride = gridDim.x * blockDim.x;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < t.nboxes; b += stride) {
        bool hc = false;
#pragma unroll
        for (int m = 0; m < NB; ++m) hc = hc || (t.child(b, m) != 0);
        counts[b] = 0;
        if (l2cnt) l2cnt[b] = 0;
        xflags[b] = (unsigned char)((hc ? kXfHasChild : 0) | ((t.flags[b] & BT_BOX_IS_SOURCE_BOX) ? kXfCollSource : 0));
    }
}

// staged rows -> CSR lists: one thread per staged slot (box, j): the reads of the staging area
// and -- rows being consecutive in the lists -- the writes are contiguous across a warp; a slot
// beyond its row's length (all of them for a row masked out) costs two cached loads of `starts`
__global__ void __launch_bounds__(256)
coll_compact_kernel(int nboxes, int stride, const int* __restrict__ tmp, const int* __restrict__ starts,
                    int* __restrict__ lists)
{
    const int64_t total = (int64_t)nboxes * stride;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
        const int b = (int)(i / stride), j = (int)(i - (int64_t)b * stride);
        const int s = starts[b];
        if (j < starts[b + 1] - s) lists[s + j] = tmp[i];
    }
}

// the same with one lane per box, for a pass with a row mask (distributed setup: most rows are
// empty, and a thread per staged slot would mostly find nothing to copy)
__global__ void __launch_bounds__(256)
coll_compact_rows_kernel(int nboxes, int stride, const int* __restrict__ tmp, const int* __restrict__ starts,
                         int* __restrict__ lists)
{
    const int gstride = gridDim.x * blockDim.x;
    for (int b0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); b0 < nboxes; b0 += gstride) {
        const int b = b0 + (threadIdx.x & 31);
        int s = 0, n = 0;
        if (b < nboxes) { s = starts[b]; n = starts[b + 1] - s; }
        int nmax = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, nmax, o); nmax = y > nmax ? y : nmax; }
        const int* row = tmp + (int64_t)b * stride;
        for (int j = 0; j < nmax; ++j)
            if (j < n) lists[s + j] = row[j];
    }
}

// list 2 from the separation masks the colleague pass left behind: no geometry, the set bits
// are the candidates (colleague of the parent, Morton child) in the reference's loop order
template <int DIM>
__global__ void __launch_bounds__(256)
list2_masked_fill_kernel(int nrows, const int* __restrict__ row_boxes, const int* __restrict__ parents,
                         const int* __restrict__ coll_starts, const int* __restrict__ coll_lists,
                         const int* __restrict__ child_t, const unsigned* __restrict__ l2mask, int mask_words,
                         const int* __restrict__ starts, int* __restrict__ lists)
{
    constexpr int NB = 1 << DIM;
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = w; r < nrows; r += nw) {
        const int box = row_boxes[r];
        const int p = parents[box];
        if (p == box) continue;
        const int cs = coll_starts[p];
        const int ncand = (coll_starts[p + 1] - cs + 1) * NB;
        const unsigned* mrow = l2mask + (int64_t)box * mask_words;
        const int pos = (int)mrow[mask_words - 1];
        int out = starts[r];
        for (int k0 = 0; k0 < ncand; k0 += 32) {
            const unsigned word = mrow[k0 >> 5];
            if (!word) continue;
            if ((word >> lane) & 1u) {
                const int k = k0 + lane, j = k / NB, m = k % NB;
                const int P = coll_lists[cs + j - (j > pos ? 1 : 0)];
                lists[out + __popc(word & ((1u << lane) - 1u))] = child_t[((int64_t)P << DIM) + m];
            }
            out += __popc(word);
        }
    }
}

struct GatherCountIn {
    const int* by_box; const int* rows;
    __device__ int operator()(int64_t i) const { return by_box[rows[i]]; }
};

template <typename T, int DIM>
static int colleagues_topdown_impl(int phase, const bt_tree_view* tv, const int* level_start,
                                   const int* dfs_rank, const signed char* row_mask, int stride, int* tmp,
                                   int* starts, int* lists, int* l2cnt, unsigned char* xflags,
                                   unsigned* l2mask, int mask_words, long long* totals, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nboxes <= 0) return BT_OK;
    if (phase == 0) {
        const int grid = grid_for((int64_t)t.nboxes * 32, 256, 8);
        if (row_mask) {
            coll_init_kernel<T, DIM><<<grid_for(t.nboxes, 256, 8), 256, 0, s>>>(t, starts, l2cnt, xflags);
            BT_LAUNCH_CHECK();
        }
        for (int lev = 0; lev < t.nlevels; ++lev) {
            coll_topdown_kernel<T, DIM><<<grid, 256, 0, s>>>(t, level_start, lev, stride, tmp, starts, dfs_rank,
                                                             row_mask, l2cnt, xflags, l2mask, mask_words);
            BT_LAUNCH_CHECK();
        }
        BT_TRY(counts_to_starts(starts, t.nboxes, totals, s));
    } else {
        if (row_mask)
            coll_compact_rows_kernel<<<grid_for((int64_t)t.nboxes, 256, 8), 256, 0, s>>>(t.nboxes, stride, tmp,
                                                                                        starts, lists);
        else
            coll_compact_kernel<<<grid_for((int64_t)t.nboxes * stride, 256, 8), 256, 0, s>>>(t.nboxes, stride, tmp,
                                                                                            starts, lists);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

// ---- b7 list 4 (+close): traversal.py:931-1146 -----------------------------
template <typename T, int DIM>
__device__ __forceinline__ bool meets_sep_bigger(const T* rad, const T* tc, int tl, const T* sc, int sl,
                                                 T stick_out_factor)
{   // traversal.py:933-972
    const T target_rad = rad[tl], source_rad = rad[sl];
    const T max_allowed = (3 * (1 + stick_out_factor) * target_rad + source_rad);
    T l_inf_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) l_inf_dist = fmax(l_inf_dist, fabs(tc[a] - sc[a]));
    return l_inf_dist >= max_allowed * (1 - 8 * CoordTraits<T>::eps());
}

template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list4(const TreeView<T, DIM>& t, const T* rad, const int* coll_starts,
                                          const int* coll_lists, int with_extent, T stick_out_factor,
                                          int tgt_ibox, E& e)
{
    T tc[DIM]; t.center(tgt_ibox, tc);
    const int tgt_level = t.levels[tgt_ibox];
    if (tgt_level == 0) return;
    const int tgt_parent = t.parents[tgt_ibox];
    const int tgt_parent_level = tgt_level - 1;
    T pc[DIM]; t.center(tgt_parent, pc);
    const unsigned char tgt_flags = t.flags[tgt_ibox];
    int walk_level, cur;
    if (t.n_away == 1) { walk_level = tgt_level - 1; cur = tgt_parent; }
    else { walk_level = tgt_level; cur = tgt_ibox; }
    for (; walk_level != 0; --walk_level, cur = t.parents[cur]) {
        const int s = coll_starts[cur], en = coll_starts[cur + 1];
        for (int i = s; i < en; ++i) {
            const int sb = coll_lists[i];
            if (!(t.flags[sb] & BT_BOX_IS_SOURCE_BOX)) continue;
            T sc[DIM]; t.center(sb, sc);
            if (adj_nbhd<T, DIM>(rad, tc, tgt_level, (T)1, sc, walk_level)) continue;
            if (with_extent) {
                if (!meets_sep_bigger<T, DIM>(rad, tc, tgt_level, sc, walk_level, stick_out_factor)) {
                    if (tgt_flags & BT_BOX_IS_TARGET_BOX) e.e1(sb);
                    continue;
                }
            }
            const bool in_parent_list_1 = adj_nbhd<T, DIM>(rad, pc, tgt_parent_level, (T)1, sc, walk_level);
            bool would_be_in_parent_list_4 = !in_parent_list_1;
            if (t.n_away > 1) would_be_in_parent_list_4 = would_be_in_parent_list_4 && (walk_level < tgt_level);
            if (would_be_in_parent_list_4) {
                if (with_extent &&
                    !meets_sep_bigger<T, DIM>(rad, pc, tgt_parent_level, sc, walk_level, stick_out_factor))
                    e.e0(sb);
            } else e.e0(sb);
        }
    }
}

// ---- N3 peer lists: area_query.py:393-475 (PEER_LIST_FINDER_TEMPLATE) ---------
// b_k is a peer of b_j if it is adjacent to b_j, at least as large, and none of its children
// satisfies both (the reference's level-restriction test reads these lists).
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_peers(const TreeView<T, DIM>& t, const T* rad, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    if (box_id == 0) { e.e0(0); return; }          // peer of root = self
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    Walk w; w.init(0);
    while (w.cont) {
        const int wb = t.child(w.parent, w.mnr);
        if (wb) {
            T wc[DIM]; t.center(wb, wc);
            // wb lives on level ssize + 1
            if (adj_nbhd<T, DIM>(rad, center, level, (T)1, wc, w.ssize + 1)) {
                if (w.ssize + 1 == level) e.e0(wb);
                else if (!(t.flags[wb] & (BT_BOX_HAS_SOURCE_CHILD_BOXES | BT_BOX_HAS_TARGET_CHILD_BOXES))) e.e0(wb);
                else {
                    bool must_be_peer = true;
                    for (int m = 0; must_be_peer && m < NB; ++m) {
                        const int c = t.child(wb, m);
                        if (c) {
                            T cc[DIM]; t.center(c, cc);
                            must_be_peer = must_be_peer && !adj_nbhd<T, DIM>(rad, center, level, (T)1, cc, w.ssize + 2);
                        }
                    }
                    if (must_be_peer) e.e0(wb);
                    else { w.push(wb); continue; }
                }
            }
        }
        w.template advance<NB>();
    }
}

// ---- N3 area query: area_query.py:168-392 ------------------------------------
// One row per l^inf ball: find the "guiding box" (the box around the clamped centre whose radius
// brackets the clamped ball radius, descending with the Morton-digit expression of the tree
// build), then emit the leaves among its peers' subtrees that overlap the ball
// (check_l_infty_ball_overlap, traversal.py:200-214).
template <typename T, int DIM>
struct Balls { const T* c[3]; const T* r; T bbox_min[3]; };

template <typename T, int DIM>
__device__ __forceinline__ bool ball_overlaps(const TreeView<T, DIM>& t, const T* rad, int b, const T* bc, T br)
{
    T c[DIM]; t.center(b, c);
    const T size_sum = rad[t.levels[b]] + br;
    T max_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) max_dist = fmax(max_dist, fabs(bc[a] - c[a]));
    return max_dist <= size_sum;
}

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
area_query_kernel(TreeView<T, DIM> t, Balls<T, DIM> balls, const int* __restrict__ peer_starts,
                  const int* __restrict__ peer_lists, int nballs, int* __restrict__ starts,
                  int* __restrict__ lists)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    constexpr unsigned char haschild = BT_BOX_HAS_SOURCE_CHILD_BOXES | BT_BOX_HAS_TARGET_CHILD_BOXES;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nballs; i += stride) {
        T bc[DIM], qc[DIM], bbox_max[DIM];
        const T br = balls.r[i];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            bc[a] = balls.c[a][i];
            bbox_max[a] = balls.bbox_min[a] + (T)((double)t.root_extent / (1 + 1e-4));
            qc[a] = fmin(bbox_max[a], fmax(balls.bbox_min[a], bc[a]));
        }
        T qr = 0;
#pragma unroll
        for (int mnr = 0; mnr < NB; ++mnr) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const T off = ((1 << (DIM - 1 - a)) & mnr) ? +br : -br;
                const T corner = fmin(bbox_max[a], fmax(balls.bbox_min[a], bc[a] + off));
                qr = fmax(qr, fabs(corner - qc[a]));
            }
        }
        int box = 0;
        if (rad[0] / 2 >= qr) {
            for (unsigned box_level = 0;; ++box_level) {
                if (!(t.flags[box] & haschild) || (rad[box_level] / 2 < qr && qr <= rad[box_level])) break;
                int morton = 0;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const T off_scaled = (qc[a] - balls.bbox_min[a]) / t.root_extent;
                    const unsigned bits = (unsigned)(off_scaled * (T)(1U << ((1 + box_level) & 31)));
                    morton |= (int)(bits & 1U) << (DIM - 1 - a);
                }
                const int next = t.child(box, morton);
                if (next) box = next; else break;
            }
        }
        int n = 0;
        int* out = FILL ? lists + starts[i] : nullptr;
        for (int pi = peer_starts[box]; pi < peer_starts[box + 1]; ++pi) {
            const int peer = peer_lists[pi];
            if (!(t.flags[peer] & haschild)) {
                if (ball_overlaps<T, DIM>(t, rad, peer, bc, br)) { if (FILL) out[n] = peer; ++n; }
            } else {
                Walk w; w.init(peer);
                while (w.cont) {
                    const int wb = t.child(w.parent, w.mnr);
                    if (wb) {
                        if (!(t.flags[wb] & haschild)) {
                            if (ball_overlaps<T, DIM>(t, rad, wb, bc, br)) { if (FILL) out[n] = wb; ++n; }
                        } else { w.push(wb); continue; }
                    }
                    w.template advance<NB>();
                }
            }
        }
        if (!FILL) starts[i] = n;
    }
}

template <typename T, int DIM>
static int area_query_impl(int phase, const bt_tree_view* tv, const int* peer_starts, const int* peer_lists,
                           int nballs, void* const* ball_centers, const void* ball_radii,
                           const double* bbox_min, int* starts, int* lists, long long* totals, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nlevels > kMaxWalkLevels) return BT_ERR_UNSUPPORTED;
    Balls<T, DIM> b;
    for (int a = 0; a < 3; ++a) {
        b.c[a] = a < DIM ? (const T*)ball_centers[a] : nullptr;
        b.bbox_min[a] = a < DIM ? (T)bbox_min[a] : (T)0;
    }
    b.r = (const T*)ball_radii;
    if (nballs > 0) {
        const int grid = grid_for(nballs, kTravBlock, 16);
        if (phase == 0) area_query_kernel<T, DIM, false><<<grid, kTravBlock, 0, s>>>(t, b, peer_starts, peer_lists, nballs, starts, lists);
        else area_query_kernel<T, DIM, true><<<grid, kTravBlock, 0, s>>>(t, b, peer_starts, peer_lists, nballs, starts, lists);
        BT_LAUNCH_CHECK();
    }
    if (phase == 0) BT_TRY(counts_to_starts(starts, nballs, totals, s));
    return BT_OK;
}

// generic row kernel; KIND: 0 colleagues, 2 list2, 4 list4, 5 peers (list 1 and 3 have their own)
template <typename T, int DIM, int KIND, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list_kernel(TreeView<T, DIM> t, const int* __restrict__ row_boxes, const int* __restrict__ coll_starts,
            const int* __restrict__ coll_lists, int with_extent, T stick_out_factor, int nrows,
            int* __restrict__ starts, int* __restrict__ lists, int* __restrict__ close_starts,
            int* __restrict__ close_lists, const signed char* __restrict__ row_mask)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
        const int box = row_boxes ? row_boxes[r] : r;
        if (row_mask && !row_mask[box]) {          // row not needed by the caller: empty list
            if (!FILL) { starts[r] = 0; if (close_starts) close_starts[r] = 0; }
            continue;
        }
        if (FILL) {
            FillEmit e{lists + starts[r], close_lists ? close_lists + close_starts[r] : nullptr};
            if (KIND == 0) gen_colleagues<T, DIM>(t, rad, box, e);
            if (KIND == 2) gen_list2<T, DIM>(t, rad, coll_starts, coll_lists, box, e);
            if (KIND == 4) gen_list4<T, DIM>(t, rad, coll_starts, coll_lists, with_extent, stick_out_factor, box, e);
            if (KIND == 5) gen_peers<T, DIM>(t, rad, box, e);
        } else {
            CountEmit e;
            if (KIND == 0) gen_colleagues<T, DIM>(t, rad, box, e);
            if (KIND == 2) gen_list2<T, DIM>(t, rad, coll_starts, coll_lists, box, e);
            if (KIND == 4) gen_list4<T, DIM>(t, rad, coll_starts, coll_lists, with_extent, stick_out_factor, box, e);
            if (KIND == 5) gen_peers<T, DIM>(t, rad, box, e);
            starts[r] = e.c0;
            if (close_starts) close_starts[r] = e.c1;
        }
    }
}

// distributed build: rows of lists 2, 4 and 4-close of the boxes whose (restricted) flags carry
// a target bit, marked straight into the box masks of partition.py:197-297 -- no lists are
// built: list 2 -> multipole mask, list 4 and list 4 close -> point-source mask
struct MarkEmit {
    signed char* m0; signed char* m1;
    __device__ __forceinline__ void e0(int v) { m0[v] = 1; }
    __device__ __forceinline__ void e1(int v) { if (m1) m1[v] = 1; }
};
template <typename T, int DIM>
__global__ void __launch_bounds__(kTravBlock)
mark_rows_kernel(TreeView<T, DIM> t, const int* __restrict__ coll_starts, const int* __restrict__ coll_lists,
                 int with_extent, T stick_out_factor, signed char* __restrict__ point_src_mask,
                 signed char* __restrict__ mpole_mask)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int stride = gridDim.x * blockDim.x;
    for (int box = blockIdx.x * blockDim.x + threadIdx.x; box < t.nboxes; box += stride) {
        if (!(t.flags[box] & (BT_BOX_IS_TARGET_BOX | BT_BOX_HAS_TARGET_CHILD_BOXES))) continue;
        MarkEmit e2{mpole_mask, nullptr};
        gen_list2<T, DIM>(t, rad, coll_starts, coll_lists, box, e2);
        MarkEmit e4{point_src_mask, point_src_mask};
        gen_list4<T, DIM>(t, rad, coll_starts, coll_lists, with_extent, stick_out_factor, box, e4);
    }
}
template <typename T, int DIM>
static int mark_rows_impl(const bt_tree_view* tv, const bt_list_args* a, signed char* point_src_mask,
                          signed char* mpole_mask, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nboxes <= 0) return BT_OK;
    mark_rows_kernel<T, DIM><<<grid_for(t.nboxes, kTravBlock, 16), kTravBlock, 0, s>>>(
        t, a->coll_starts, a->coll_lists, a->with_extent, (T)a->stick_out_factor, point_src_mask, mpole_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

// counts -> starts (in place), total to starts[n] and totals[slot]
struct InPlaceIn {
    const int* a;
    __device__ int operator()(int64_t i) const { return a[i]; }
};
struct InPlaceOut {
    int* a; long long* total_out; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { a[i] = (int)excl; }
    __device__ void total(long long t) const { a[n] = (int)t; if (total_out) *total_out = t; }
};

static int counts_to_starts(int* a, int64_t n, long long* total_out, cudaStream_t s)
{
    InPlaceIn in{a};
    InPlaceOut out{a, total_out, n};
    return scan_exclusive(n, nullptr, in, out, s);
}

template <typename T, int DIM>
static int build_list_impl(int kind, int phase, const bt_tree_view* tv, const bt_list_args* a, int nrows,
                           int* starts, int* lists, int* close_starts, int* close_lists,
                           long long* totals, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    const int grid = grid_for(nrows, kTravBlock, 16);
    const T sof = (T)a->stick_out_factor;
#define BT_LAUNCH_LIST(KIND, FILL)                                                              \
    list_kernel<T, DIM, KIND, FILL><<<grid, kTravBlock, 0, s>>>(                                \
        t, a->row_boxes, a->coll_starts, a->coll_lists, a->with_extent, sof, nrows, starts, lists, \
        close_starts, close_lists, (const signed char*)a->row_mask)
    if (nrows > 0) {
        if (phase == 0) {
            switch (kind) {
            case 0:
                if (g_walk_mode & kModeColl) coll_coop_kernel<T, DIM, false><<<grid_for((int64_t)nrows << DIM, kTravBlock, 16), kTravBlock, 0, s>>>(t, nrows, starts, lists, (const signed char*)a->row_mask);
                else BT_LAUNCH_LIST(0, false);
                break;
            case 2:
                if (g_walk_mode & kModeList2Count) list2_warp_kernel<T, DIM, false><<<grid_for((int64_t)nrows * 32, 256, 8), 256, 0, s>>>(t, a->row_boxes, a->coll_starts, a->coll_lists, nrows, starts, lists);
                else BT_LAUNCH_LIST(2, false);
                break;
            case 4: BT_LAUNCH_LIST(4, false); break;
            case 5: BT_LAUNCH_LIST(5, false); break;
            default: return BT_ERR_BAD_ARG;
            }
        } else {
            switch (kind) {
            case 0:
                if (g_walk_mode & kModeColl) coll_coop_kernel<T, DIM, true><<<grid_for((int64_t)nrows << DIM, kTravBlock, 16), kTravBlock, 0, s>>>(t, nrows, starts, lists, (const signed char*)a->row_mask);
                else BT_LAUNCH_LIST(0, true);
                break;
            case 2:
                if (g_walk_mode & kModeList2Fill) list2_warp_kernel<T, DIM, true><<<grid_for((int64_t)nrows * 32, 256, 8), 256, 0, s>>>(t, a->row_boxes, a->coll_starts, a->coll_lists, nrows, starts, lists);
                else BT_LAUNCH_LIST(2, true);
                break;
            case 4: BT_LAUNCH_LIST(4, true); break;
            case 5: BT_LAUNCH_LIST(5, true); break;
            default: return BT_ERR_BAD_ARG;
            }
        }
        BT_LAUNCH_CHECK();
    }
#undef BT_LAUNCH_LIST
    if (phase == 0) {
        BT_TRY(counts_to_starts(starts, nrows, totals, s));
        if (close_starts) BT_TRY(counts_to_starts(close_starts, nrows, totals + 1, s));
    }
    return BT_OK;
}

// ---- b6 list 3: traversal.py:607-875 ---------------------------------------
template <typename T, int DIM>
struct List3Args {
    const int* target_boxes; const int* coll_starts; const int* coll_lists;
    T stick_out_factor; int targets_have_extent, sources_have_extent, crit;
    const T* bb_min; const T* bb_max; const int* box_source_counts_cumul; int min_nsources_cumul;
};

// per-target-box constants of the list-3 walk (traversal.py:614-630)
template <typename T, int DIM>
struct L3Ctx { T tc[DIM]; int tgt_level; T stickout; T ext_center[DIM]; T radii_vec[DIM]; };

template <typename T, int DIM>
__device__ __forceinline__ void l3_make_ctx(const TreeView<T, DIM>& t, const T* rad,
                                            const List3Args<T, DIM>& x, int tgt_box_id, L3Ctx<T, DIM>& c)
{
    t.center(tgt_box_id, c.tc);
    c.tgt_level = t.levels[tgt_box_id];
    c.stickout = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) { c.ext_center[a] = 0; c.radii_vec[a] = 0; }
    if (x.targets_have_extent) {
        if (x.crit == 0 || x.crit == 2)
            c.stickout = (1 + x.stick_out_factor) * rad[c.tgt_level];
        else {   // load_true_box_extent, traversal.py:177-198
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const T mn = x.bb_min[a * t.aligned + tgt_box_id], mx = x.bb_max[a * t.aligned + tgt_box_id];
                c.ext_center[a] = ((T)0.5) * (mn + mx);
                c.radii_vec[a] = ((T)0.5) * (mx - mn);
            }
        }
    }
}

// one child visit of the list-3 walk (traversal.py:673-870), all source levels at once:
// kVisitEmit appends wb to the list of source level levels[wb], kVisitClose to list 3 close.
template <typename T, int DIM>
__device__ __forceinline__ int list3_visit(const TreeView<T, DIM>& t, const T* rad,
                                           const List3Args<T, DIM>& x, const L3Ctx<T, DIM>& c, int wb)
{
    const unsigned char cfl = t.flags[wb];
    if (!(wb && (cfl & (BT_BOX_IS_SOURCE_BOX | BT_BOX_HAS_SOURCE_CHILD_BOXES)))) return 0;
    T wc[DIM]; t.center(wb, wc);
    const int walk_level = t.levels[wb];
    if (adj_nbhd<T, DIM>(rad, c.tc, c.tgt_level, (T)1, wc, walk_level))   // what list 1 does here (:508-539)
        return ((cfl & BT_BOX_HAS_SOURCE_CHILD_BOXES) ? kVisitPush : 0) |
               ((cfl & BT_BOX_IS_SOURCE_BOX) ? kVisitNear : 0);
    const T two_minus = (2 - 8 * CoordTraits<T>::eps());
    bool meets;
    if (!x.targets_have_extent) meets = true;
    else if (x.crit == 0) {
        const T source_rad = rad[walk_level];
        T d = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) d = fmax(d, fabs(c.tc[a] - wc[a]) - c.stickout - source_rad);
        meets = d >= two_minus * source_rad;
    } else if (x.crit == 1) {
        const T source_rad = rad[walk_level];
        T d = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a)
            d = fmax(d, fabs(c.ext_center[a] - wc[a]) - c.radii_vec[a] - source_rad);
        meets = d >= two_minus * source_rad;
    } else {
        const T source_rad = rad[walk_level];
        T l2sq = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) l2sq = l2sq + (c.tc[a] - wc[a]) * (c.tc[a] - wc[a]);
        const T rhs = sqrt(l2sq) - sqrt((T)DIM) * c.stickout - source_rad;
        meets = (two_minus * source_rad <= rhs);
    }
    const bool close_lists_exist = x.sources_have_extent || x.targets_have_extent;
    const bool force_close = close_lists_exist && (x.box_source_counts_cumul[wb] < x.min_nsources_cumul);
    if (meets && !force_close) return kVisitEmit;
    if (!close_lists_exist) return 0;
    return ((cfl & BT_BOX_IS_SOURCE_BOX) ? kVisitClose : 0) |
           ((cfl & BT_BOX_HAS_SOURCE_CHILD_BOXES) ? kVisitPush : 0);
}

// E must provide append(level, box) and close(box); false = budget exceeded
template <typename T, int DIM, class E>
__device__ __forceinline__ bool gen_list3(const TreeView<T, DIM>& t, const T* rad, const List3Args<T, DIM>& x,
                                          int tgt_box_id, E& e, int budget)
{
    constexpr int NB = 1 << DIM;
    L3Ctx<T, DIM> c; l3_make_ctx<T, DIM>(t, rad, x, tgt_box_id, c);
    int visits = 0;
    const int s = x.coll_starts[tgt_box_id], en = x.coll_starts[tgt_box_id + 1];
    for (int i = s; i < en; ++i) {
        const int same_lev_nws_box = x.coll_lists[i];
        if (same_lev_nws_box == tgt_box_id) continue;
        Walk w; w.init(same_lev_nws_box);
        while (w.cont) {
            if (++visits > budget) return false;
            const int wb = t.child(w.parent, w.mnr);
            const int act = list3_visit<T, DIM>(t, rad, x, c, wb);
            if (act & kVisitEmit) e.append(t.levels[wb], wb);
            if (act & kVisitClose) e.close(wb);
            if (act & kVisitPush) { w.push(wb); continue; }
            w.template advance<NB>();
        }
    }
    return true;
}

struct L3Count {
    int cnt[kMaxWalkLevels + 1]; int nlevels;
    __device__ __forceinline__ void append(int level, int) { ++cnt[level]; }
    __device__ __forceinline__ void close(int) { ++cnt[nlevels]; }
};
struct L3Fill {
    int cur[kMaxWalkLevels + 1]; int nlevels; int* lists;
    __device__ __forceinline__ void append(int level, int v) { lists[cur[level]++] = v; }
    __device__ __forceinline__ void close(int v) { lists[cur[nlevels]++] = v; }
};

// ---- heavy rows -------------------------------------------------------------
// A row whose walk needs more than `walk_budget` child visits (an upper-level box with its
// own targets can have a list 1 of ~1e5..1e6 entries) is taken out of the one-thread-per-row
// kernels.  All heavy rows are expanded together by a level-synchronous BFS over the same
// child visits (same predicates, hence the same set of appended boxes); the append ORDER of
// the reference's depth-first, Morton-ordered walk is the global DFS pre-order of the tree,
// so the appended boxes are sorted by (list, row, pre-order rank) with the one-sweep radix
// sort and copied to their CSR positions.
struct HeavyWs {
    int budget; unsigned char* row_heavy; int* heavy_rows; int* hctl; long long* heavy_total;
    unsigned long long* frontier[2]; long long frontier_cap; const int* dfs_rank;
    unsigned long long* ekeys[2]; unsigned* evals[2]; long long ecap;
    const signed char* row_mask;     // optional [nboxes]: rows of boxes with mask 0 stay empty
    unsigned* stage; int stage_cap; int* stage_count;   // fused walk: entries staged by the count pass
    const int* dfs_order;            // fused walk: box of every depth-first rank (keys-only heavy sort)
    // fused walk, map mode (heavy rows without a sort), see "heavy rows by position map" below
    const int* subtree_size; long long* hrow_base; long long* hplan; int seg_stride;
    int* hseg_rank; int* hseg_prefix; unsigned char* hseg_kind; int* hseg_n;
    unsigned char* hmap; long long hmap_cap; int* chunk_cnt; void* hctx;
};
constexpr int kHctlNHeavy = 0, kHctlOverflow = 1, kHctlECount = 2, kHctlNWalk = 3, kHctlFrontier = 8;
// row states left by the count pass of the fused walk (row_heavy[]): 0 = the fill pass walks
// the row again, 1 = heavy row (grid-wide path), 2 = entries staged, the fill pass copies them
constexpr int kStageTagShift = 27;

static HeavyWs make_ws(const bt_heavy_ws* w)
{
    HeavyWs h;
    h.budget = w->walk_budget; h.row_heavy = w->row_heavy; h.heavy_rows = w->heavy_rows;
    h.hctl = w->hctl; h.heavy_total = (long long*)w->heavy_total;
    h.frontier[0] = (unsigned long long*)w->frontier[0]; h.frontier[1] = (unsigned long long*)w->frontier[1];
    h.frontier_cap = w->frontier_cap; h.dfs_rank = w->dfs_rank;
    h.ekeys[0] = (unsigned long long*)w->ekeys[0]; h.ekeys[1] = (unsigned long long*)w->ekeys[1];
    h.evals[0] = w->evals[0]; h.evals[1] = w->evals[1]; h.ecap = w->ecap;
    h.row_mask = (const signed char*)w->row_mask;
    h.stage = w->stage; h.stage_cap = w->stage_cap; h.stage_count = w->stage_count;
    h.dfs_order = w->dfs_order;
    h.subtree_size = w->subtree_size; h.hrow_base = (long long*)w->hrow_base; h.hplan = (long long*)w->hplan;
    h.seg_stride = w->seg_stride; h.hseg_rank = w->hseg_rank; h.hseg_prefix = w->hseg_prefix;
    h.hseg_kind = w->hseg_kind; h.hseg_n = w->hseg_n; h.hmap = w->hmap; h.hmap_cap = w->hmap_cap;
    h.chunk_cnt = w->chunk_cnt; h.hctx = w->hctx;
    return h;
}

__device__ __forceinline__ int bits_for(int n)   // bits needed for values in [0, n)
{ return n <= 1 ? 1 : 32 - __clz(n - 1); }

// warp-aggregated append: returns the slot of this lane's item (or -1)
__device__ __forceinline__ long long warp_append(bool want, int* counter)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return -1;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return want ? (long long)base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u)) : -1;
}

// light pass of list 1: one thread per row, budgeted
template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list1_kernel(TreeView<T, DIM> t, const int* __restrict__ target_boxes, int nrows, int* __restrict__ starts,
             int* __restrict__ lists, HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
        const int box = target_boxes[r];
        if (ws.row_mask && !ws.row_mask[box]) {
            if (!FILL) { starts[r] = 0; ws.row_heavy[r] = 0; }
            continue;
        }
        if (FILL) {
            if (ws.row_heavy[r]) continue;
            FillEmit e{lists + starts[r], nullptr};
            gen_list1<T, DIM>(t, rad, box, e, 0x7fffffff);
        } else {
            CountEmit e;
            const bool ok = gen_list1<T, DIM>(t, rad, box, e, ws.budget);
            starts[r] = ok ? e.c0 : 0;
            ws.row_heavy[r] = ok ? 0 : 1;
            if (!ok) ws.heavy_rows[atomicAdd(ws.hctl + kHctlNHeavy, 1)] = r;
        }
    }
}

// list 1, cooperative light pass
template <typename T, int DIM, bool FILL>
struct L1Policy {
    const TreeView<T, DIM>& t; const T* rad; const int* target_boxes; int* starts; int* lists; HeavyWs ws;
    T center[DIM]; int box, level, count; bool rooted, skip; int* out;
    __device__ L1Policy(const TreeView<T, DIM>& t_, const T* rad_, const int* tb, int* st, int* li, const HeavyWs& w)
        : t(t_), rad(rad_), target_boxes(tb), starts(st), lists(li), ws(w) {}
    __device__ __forceinline__ void init(int row, bool valid)
    {
        count = 0; rooted = false; skip = true; box = 0; level = 0; out = nullptr;
        if (!valid) return;
        box = target_boxes[row]; t.center(box, center); level = t.levels[box];
        skip = (FILL && ws.row_heavy[row]) || (ws.row_mask && !ws.row_mask[box]);
        out = FILL ? lists + starts[row] : nullptr;
    }
    __device__ __forceinline__ bool next_root(int& parent)
    {
        if (rooted || skip) return false;
        rooted = true; parent = 0;
        if (t.flags[0] & BT_BOX_IS_SOURCE_BOX) {        // the root is checked up front (:489-495)
            if (FILL && (threadIdx.x & ((1 << DIM) - 1)) == 0) out[0] = 0;
            count = 1;
        }
        return true;
    }
    __device__ __forceinline__ int visit(int wb) const { return list1_visit<T, DIM>(t, rad, center, level, wb); }
    __device__ __forceinline__ void emit(unsigned eb, unsigned, unsigned, int c, int gl, int)
    {
        if (FILL && ((eb >> gl) & 1u)) out[count + __popc(eb & ((1u << gl) - 1u))] = c;
        count += __popc(eb);
    }
    __device__ __forceinline__ void finish(int row, bool valid, bool ok, int gl)
    {
        if (FILL || gl != 0 || !valid) return;
        starts[row] = ok ? count : 0;
        ws.row_heavy[row] = ok ? 0 : 1;
        if (!ok) ws.heavy_rows[atomicAdd(ws.hctl + kHctlNHeavy, 1)] = row;
    }
};

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list1_coop_kernel(TreeView<T, DIM> t, const int* __restrict__ target_boxes, int nrows,
                  int* __restrict__ starts, int* __restrict__ lists, HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    __shared__ CoopFrame frames[(kTravBlock >> DIM) * kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    L1Policy<T, DIM, FILL> pol(t, rad, target_boxes, starts, lists, ws);
    coop_walk_rows<T, DIM>(t, pol, nrows, FILL ? 0x7ffffff0 : ws.budget, frames);
}

// BFS seed: one frontier item (row, walk parent = root) per heavy row + the root's own append
template <typename T, int DIM, bool FILL>
__global__ void list1_heavy_seed_kernel(TreeView<T, DIM> t, int nrows, int* __restrict__ starts, HeavyWs ws)
{
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int rank_bits = bits_for(t.nboxes);
    const int stride = gridDim.x * blockDim.x;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += stride) {
        const int r = ws.heavy_rows[h];
        ws.frontier[0][h] = ((unsigned long long)r << 32);      // walk parent = box 0
        if (t.flags[0] & BT_BOX_IS_SOURCE_BOX) {
            if (FILL) {
                const int k = atomicAdd(ws.hctl + kHctlECount, 1);
                ws.ekeys[0][k] = ((unsigned long long)r << rank_bits) | (unsigned)ws.dfs_rank[0];
                ws.evals[0][k] = 0;
            } else atomicAdd(starts + r, 1);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.hctl[kHctlFrontier] = nheavy;
}

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(256)
list1_heavy_step_kernel(TreeView<T, DIM> t, const int* __restrict__ target_boxes, int step,
                        int* __restrict__ starts, HeavyWs ws)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const unsigned long long* fin = ws.frontier[step & 1];
    unsigned long long* fout = ws.frontier[(step + 1) & 1];
    long long nitems = ws.hctl[kHctlFrontier + step];
    if (nitems > ws.frontier_cap) nitems = ws.frontier_cap;
    const long long total = nitems * NB;
    const int rank_bits = bits_for(t.nboxes);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < ((total + 31) & ~31ll);
         tid += stride) {
        int act = 0, wb = 0, r = 0;
        if (tid < total) {
            const unsigned long long item = fin[tid / NB];
            r = (int)(item >> 32);
            const int parent = (int)(unsigned)item, m = (int)(tid % NB);
            const int box = target_boxes[r];
            T center[DIM]; t.center(box, center);
            wb = t.child(parent, m);
            act = list1_visit<T, DIM>(t, rad, center, t.levels[box], wb);
        }
        if (FILL) {
            const long long k = warp_append(act & kVisitEmit, ws.hctl + kHctlECount);
            if (k >= 0 && k < ws.ecap) {
                ws.ekeys[0][k] = ((unsigned long long)r << rank_bits) | (unsigned)ws.dfs_rank[wb];
                ws.evals[0][k] = (unsigned)wb;
            }
        } else if (act & kVisitEmit) atomicAdd(starts + r, 1);
        const long long q = warp_append(act & kVisitPush, ws.hctl + kHctlFrontier + step + 1);
        if (q >= 0) {
            if (q < ws.frontier_cap) fout[q] = ((unsigned long long)r << 32) | (unsigned)wb;
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

// heavy_total = sum of the counts of heavy rows (sizes the sort buffers of the fill phase)
__global__ void heavy_total_kernel(const int* __restrict__ counts, int64_t rowlen, int nslots, HeavyWs ws)
{
    const int nheavy = ws.hctl[kHctlNHeavy];
    long long acc = 0;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += gridDim.x * blockDim.x) {
        const int r = ws.heavy_rows[h];
        for (int sl = 0; sl < nslots; ++sl) acc += counts[sl * rowlen + r];
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long*)ws.heavy_total, (unsigned long long)acc);
}

// sorted (list slot, row, pre-order rank) records -> CSR positions
__global__ void __launch_bounds__(256)
heavy_scatter_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals,
                     const int* __restrict__ ecount_dev, int rank_bits, int row_bits, int64_t rowlen,
                     const int* __restrict__ dest_base, int* __restrict__ lists)
{
    const int n = *ecount_dev;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        const unsigned long long group = (key >> rank_bits) << rank_bits;
        int lo = 0, hi = i;                       // first record of this (slot, row) group
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < group) lo = mid + 1; else hi = mid; }
        const long long row = (long long)((key >> rank_bits) & ((1ull << row_bits) - 1ull));
        const long long slot = (long long)(key >> (rank_bits + row_bits));
        lists[dest_base[slot * rowlen + row] + (i - lo)] = (int)vals[i];
    }
}

static int heavy_sort_and_scatter(const HeavyWs& ws, long long ecount_host, int nboxes, int nrows,
                                  int nslots, int64_t rowlen, const int* dest_base, int* lists,
                                  cudaStream_t s)
{
    if (ecount_host <= 0) return BT_OK;
    int rank_bits = 1; while ((1ll << rank_bits) < nboxes) ++rank_bits;
    int row_bits = 1; while ((1ll << row_bits) < nrows) ++row_bits;
    int slot_bits = 0; while ((1ll << slot_bits) < nslots) ++slot_bits;
    if (rank_bits + row_bits + slot_bits > 64) return BT_ERR_UNSUPPORTED;
    int in_alt = 0;
    BT_TRY(radix_sort_pairs(ecount_host, ws.ekeys[0], ws.ekeys[1], ws.evals[0], ws.evals[1], 0, 0,
                            rank_bits + row_bits + slot_bits, &in_alt, s));
    heavy_scatter_kernel<<<grid_for(ecount_host, 256, 8), 256, 0, s>>>(
        ws.ekeys[in_alt], ws.evals[in_alt], ws.hctl + kHctlECount, rank_bits, row_bits, rowlen,
        dest_base, lists);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

template <typename T, int DIM>
static int list1_impl(int phase, const bt_tree_view* tv, const int* target_boxes, int nrows, int* starts,
                      int* lists, long long* totals, const bt_heavy_ws* w, long long heavy_total_host,
                      cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    HeavyWs ws = make_ws(w);
    const int grid = grid_for(nrows, kTravBlock, 16);
    const int cgrid = grid_for((int64_t)nrows << DIM, kTravBlock, 16);
    const int nsteps = t.nlevels;
    if (phase == 0) {
        BT_CHECK(cudaMemsetAsync(ws.hctl, 0, sizeof(int) * BT_HCTL_SIZE, s));
        BT_CHECK(cudaMemsetAsync(ws.heavy_total, 0, sizeof(long long), s));
        if (nrows > 0) {
            if (g_walk_mode & kModeList1)
                list1_coop_kernel<T, DIM, false><<<cgrid, kTravBlock, 0, s>>>(t, target_boxes, nrows, starts, nullptr, ws);
            else
                list1_kernel<T, DIM, false><<<grid, kTravBlock, 0, s>>>(t, target_boxes, nrows, starts, nullptr, ws);
            BT_LAUNCH_CHECK();
            list1_heavy_seed_kernel<T, DIM, false><<<kNumSMs, 256, 0, s>>>(t, nrows, starts, ws);
            BT_LAUNCH_CHECK();
            for (int st = 0; st < nsteps; ++st) {
                list1_heavy_step_kernel<T, DIM, false><<<kNumSMs * 8, 256, 0, s>>>(t, target_boxes, st, starts, ws);
                BT_LAUNCH_CHECK();
            }
            heavy_total_kernel<<<kNumSMs, 256, 0, s>>>(starts, 0, 1, ws);
            BT_LAUNCH_CHECK();
        }
        BT_TRY(counts_to_starts(starts, nrows, totals, s));
        return BT_OK;
    }
    if (nrows <= 0) return BT_OK;
    if (g_walk_mode & kModeList1)
        list1_coop_kernel<T, DIM, true><<<cgrid, kTravBlock, 0, s>>>(t, target_boxes, nrows, starts, lists, ws);
    else
        list1_kernel<T, DIM, true><<<grid, kTravBlock, 0, s>>>(t, target_boxes, nrows, starts, lists, ws);
    BT_LAUNCH_CHECK();
    if (heavy_total_host > 0) {
        BT_CHECK(cudaMemsetAsync(ws.hctl + kHctlECount, 0, sizeof(int) * (BT_HCTL_SIZE - kHctlECount), s));
        list1_heavy_seed_kernel<T, DIM, true><<<kNumSMs, 256, 0, s>>>(t, nrows, starts, ws);
        BT_LAUNCH_CHECK();
        for (int st = 0; st < nsteps; ++st) {
            list1_heavy_step_kernel<T, DIM, true><<<kNumSMs * 8, 256, 0, s>>>(t, target_boxes, st, starts, ws);
            BT_LAUNCH_CHECK();
        }
        BT_TRY(heavy_sort_and_scatter(ws, heavy_total_host, t.nboxes, nrows, 1, 0, starts, lists, s));
    }
    return BT_OK;
}

// light pass of list 3
template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list3_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int ntgt, int* __restrict__ G /*[nlevels+1][ntgt+1]*/,
             int* __restrict__ lists, HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int nl = t.nlevels;
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < ntgt; r += stride) {
        const int box = x.target_boxes[r];
        if (ws.row_mask && !ws.row_mask[box]) {
            if (!FILL) {
                for (int l = 0; l <= nl; ++l) G[l * rowlen + r] = 0;
                ws.row_heavy[r] = 0;
            }
            continue;
        }
        if (FILL) {
            if (ws.row_heavy[r]) continue;
            L3Fill e; e.nlevels = nl; e.lists = lists;
            for (int l = 0; l <= nl; ++l) e.cur[l] = G[l * rowlen + r];
            gen_list3<T, DIM>(t, rad, x, box, e, 0x7fffffff);
        } else {
            L3Count e; e.nlevels = nl;
            for (int l = 0; l <= nl; ++l) e.cnt[l] = 0;
            const bool ok = gen_list3<T, DIM>(t, rad, x, box, e, ws.budget);
            for (int l = 0; l <= nl; ++l) G[l * rowlen + r] = ok ? e.cnt[l] : 0;
            ws.row_heavy[r] = ok ? 0 : 1;
            if (!ok) ws.heavy_rows[atomicAdd(ws.hctl + kHctlNHeavy, 1)] = r;
        }
    }
}

// list 3, cooperative light pass: per-group slot counters / cursors live in shared memory
template <typename T, int DIM, bool FILL>
struct L3Policy {
    const TreeView<T, DIM>& t; const T* rad; const List3Args<T, DIM>& x; int ntgt; int* G; int* lists;
    HeavyWs ws; int* slots;      // [nlevels + 1] of this group (shared memory)
    L3Ctx<T, DIM> c; int box, icoll, ecoll; bool skip;
    __device__ L3Policy(const TreeView<T, DIM>& t_, const T* rad_, const List3Args<T, DIM>& x_, int n, int* g,
                        int* li, const HeavyWs& w, int* sl)
        : t(t_), rad(rad_), x(x_), ntgt(n), G(g), lists(li), ws(w), slots(sl) {}
    __device__ __forceinline__ void init(int row, bool valid)
    {
        box = 0; icoll = ecoll = 0; skip = true;
        if (valid) {
            box = x.target_boxes[row];
            l3_make_ctx<T, DIM>(t, rad, x, box, c);
            icoll = x.coll_starts[box]; ecoll = x.coll_starts[box + 1];
            skip = (FILL && ws.row_heavy[row]) || (ws.row_mask && !ws.row_mask[box]);
        }
        const int gl = threadIdx.x & ((1 << DIM) - 1);
        const int64_t rowlen = (int64_t)ntgt + 1;
        __syncwarp();
        if (valid)
            for (int l = gl; l <= t.nlevels; l += (1 << DIM)) slots[l] = FILL ? G[l * rowlen + row] : 0;
        __syncwarp();
    }
    __device__ __forceinline__ bool next_root(int& parent)
    {
        if (skip) return false;
        while (icoll < ecoll) {
            const int cb = x.coll_lists[icoll++];
            if (cb == box) continue;
            parent = cb;
            return true;
        }
        return false;
    }
    __device__ __forceinline__ int visit(int wb) const { return list3_visit<T, DIM>(t, rad, x, c, wb); }
    __device__ __forceinline__ void emit(unsigned eb, unsigned cb, unsigned, int ch, int gl, int parent)
    {
        // called by every lane of the warp (eb = cb = 0 for idle groups)
        int lev = 0, base_e = 0, base_c = 0;
        if (eb | cb) {
            lev = t.levels[parent] + 1;          // siblings share their level
            base_e = slots[lev]; base_c = slots[t.nlevels];
        }
        __syncwarp();
        if (FILL) {
            if ((eb >> gl) & 1u) lists[base_e + __popc(eb & ((1u << gl) - 1u))] = ch;
            if ((cb >> gl) & 1u) lists[base_c + __popc(cb & ((1u << gl) - 1u))] = ch;
        }
        if (gl == 0 && (eb | cb)) { slots[lev] = base_e + __popc(eb); slots[t.nlevels] = base_c + __popc(cb); }
        __syncwarp();
    }
    __device__ __forceinline__ void finish(int row, bool valid, bool ok, int gl)
    {
        __syncwarp();
        if (FILL || !valid) return;
        const int64_t rowlen = (int64_t)ntgt + 1;
        for (int l = gl; l <= t.nlevels; l += (1 << DIM)) G[l * rowlen + row] = ok ? slots[l] : 0;
        if (gl == 0) {
            ws.row_heavy[row] = ok ? 0 : 1;
            if (!ok) ws.heavy_rows[atomicAdd(ws.hctl + kHctlNHeavy, 1)] = row;
        }
    }
};

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list3_coop_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int ntgt, int* __restrict__ G,
                  int* __restrict__ lists, HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    __shared__ CoopFrame frames[(kTravBlock >> DIM) * kMaxWalkLevels];
    __shared__ int slots[(kTravBlock >> DIM) * (kMaxWalkLevels + 1)];
    fill_rad_table(rad, t.root_extent);
    L3Policy<T, DIM, FILL> pol(t, rad, x, ntgt, G, lists, ws,
                               slots + (threadIdx.x >> DIM) * (kMaxWalkLevels + 1));
    coop_walk_rows<T, DIM>(t, pol, ntgt, FILL ? 0x7ffffff0 : ws.budget, frames);
}

// BFS seed of list 3: one frontier item per (heavy row, colleague != row box)
template <typename T, int DIM>
__global__ void list3_heavy_seed_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, HeavyWs ws)
{
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int stride = gridDim.x * blockDim.x;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += stride) {
        const int r = ws.heavy_rows[h];
        const int box = x.target_boxes[r];
        for (int i = x.coll_starts[box]; i < x.coll_starts[box + 1]; ++i) {
            const int c = x.coll_lists[i];
            if (c == box) continue;
            const int q = atomicAdd(ws.hctl + kHctlFrontier, 1);
            if (q < ws.frontier_cap) ws.frontier[0][q] = ((unsigned long long)r << 32) | (unsigned)c;
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(256)
list3_heavy_step_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int ntgt, int step, int* __restrict__ G,
                        HeavyWs ws)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const unsigned long long* fin = ws.frontier[step & 1];
    unsigned long long* fout = ws.frontier[(step + 1) & 1];
    long long nitems = ws.hctl[kHctlFrontier + step];
    if (nitems > ws.frontier_cap) nitems = ws.frontier_cap;
    const long long total = nitems * NB;
    const int rank_bits = bits_for(t.nboxes), row_bits = bits_for(ntgt);
    const int64_t rowlen = (int64_t)ntgt + 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < ((total + 31) & ~31ll);
         tid += stride) {
        int act = 0, wb = 0, r = 0;
        if (tid < total) {
            const unsigned long long item = fin[tid / NB];
            r = (int)(item >> 32);
            const int parent = (int)(unsigned)item, m = (int)(tid % NB);
            L3Ctx<T, DIM> c; l3_make_ctx<T, DIM>(t, rad, x, x.target_boxes[r], c);
            wb = t.child(parent, m);
            act = list3_visit<T, DIM>(t, rad, x, c, wb);
        }
        const bool emits = (act & (kVisitEmit | kVisitClose)) != 0;
        const int slot = (act & kVisitEmit) ? (int)t.levels[wb] : t.nlevels;
        if (FILL) {
            const long long k = warp_append(emits, ws.hctl + kHctlECount);
            if (k >= 0 && k < ws.ecap) {
                ws.ekeys[0][k] = ((unsigned long long)slot << (rank_bits + row_bits))
                                 | ((unsigned long long)r << rank_bits) | (unsigned)ws.dfs_rank[wb];
                ws.evals[0][k] = (unsigned)wb;
            }
        } else if (emits) atomicAdd(G + slot * rowlen + r, 1);
        const long long q = warp_append(act & kVisitPush, ws.hctl + kHctlFrontier + step + 1);
        if (q >= 0) {
            if (q < ws.frontier_cap) fout[q] = ((unsigned long long)r << 32) | (unsigned)wb;
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

// after the flattened scan of G: C = flattened exclusive scan of row-nonempty flags
struct NonemptyIn {
    const int* G; int64_t total_len; int64_t rowlen;
    __device__ int operator()(int64_t i) const
    {
        if ((i % rowlen) == rowlen - 1) return 0;          // padding slot of each row
        const int nxt = G[i + 1];
        return (nxt - G[i]) > 0 ? 1 : 0;
    }
};
struct PlainOut {
    int* a; long long* total_out; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { a[i] = (int)excl; }
    __device__ void total(long long t) const { a[n] = (int)t; if (total_out) *total_out = t; }
};

__global__ void list3_summary_kernel(const int* __restrict__ G, const int* __restrict__ C, int nrows,
                                     int64_t rowlen, long long* __restrict__ summary)
{   // summary[l] = G[l][0] (l = 0..nrows, last = total), summary[nrows+1+l] = C[l][0]
    const int l = threadIdx.x;
    if (l <= nrows) {
        summary[l] = G[(int64_t)l * rowlen];
        summary[nrows + 1 + l] = C[(int64_t)l * rowlen];
    }
}

template <typename T, int DIM>
static int list3_impl(int phase, const bt_tree_view* tv, const bt_list3_args* a, int ntgt, int* G, int* C,
                      int* lists, long long* summary, const bt_heavy_ws* w, long long heavy_total_host,
                      cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nlevels > kMaxWalkLevels) return BT_ERR_UNSUPPORTED;
    HeavyWs ws = make_ws(w);
    List3Args<T, DIM> x{a->target_boxes, a->coll_starts, a->coll_lists, (T)a->stick_out_factor,
                        a->targets_have_extent, a->sources_have_extent, a->crit,
                        (const T*)a->box_target_bounding_box_min, (const T*)a->box_target_bounding_box_max,
                        a->box_source_counts_cumul, a->min_nsources_cumul};
    const int nrows = t.nlevels + 1;
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int64_t total_len = rowlen * nrows;
    const int grid = grid_for(ntgt, kTravBlock, 16);
    const int cgrid = grid_for((int64_t)ntgt << DIM, kTravBlock, 16);
    // with target extents the list-3 walks are deep (close lists): cooperative wins there
    const bool coop3 = (g_walk_mode & kModeList3) ||
                       ((g_walk_mode & kModeList3Auto) && a->targets_have_extent);
    const int nsteps = t.nlevels;
    if (phase == 0) {
        BT_CHECK(cudaMemsetAsync(G, 0, sizeof(int) * (total_len + 1), s));
        BT_CHECK(cudaMemsetAsync(ws.hctl, 0, sizeof(int) * BT_HCTL_SIZE, s));
        BT_CHECK(cudaMemsetAsync(ws.heavy_total, 0, sizeof(long long), s));
        if (ntgt > 0) {
            if (coop3)
                list3_coop_kernel<T, DIM, false><<<cgrid, kTravBlock, 0, s>>>(t, x, ntgt, G, nullptr, ws);
            else
                list3_kernel<T, DIM, false><<<grid, kTravBlock, 0, s>>>(t, x, ntgt, G, nullptr, ws);
            BT_LAUNCH_CHECK();
            list3_heavy_seed_kernel<T, DIM><<<kNumSMs, 256, 0, s>>>(t, x, ws);
            BT_LAUNCH_CHECK();
            for (int st = 0; st < nsteps; ++st) {
                list3_heavy_step_kernel<T, DIM, false><<<kNumSMs * 8, 256, 0, s>>>(t, x, ntgt, st, G, ws);
                BT_LAUNCH_CHECK();
            }
            heavy_total_kernel<<<kNumSMs, 256, 0, s>>>(G, rowlen, nrows, ws);
            BT_LAUNCH_CHECK();
        }
        InPlaceIn in{G};
        PlainOut out{G, nullptr, total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, in, out, s));
        NonemptyIn nin{G, total_len, rowlen};
        PlainOut nout{C, nullptr, total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, nin, nout, s));
        list3_summary_kernel<<<1, 64, 0, s>>>(G, C, nrows, rowlen, summary);
        BT_LAUNCH_CHECK();
    } else if (ntgt > 0) {
        if (coop3)
            list3_coop_kernel<T, DIM, true><<<cgrid, kTravBlock, 0, s>>>(t, x, ntgt, G, lists, ws);
        else
            list3_kernel<T, DIM, true><<<grid, kTravBlock, 0, s>>>(t, x, ntgt, G, lists, ws);
        BT_LAUNCH_CHECK();
        if (heavy_total_host > 0) {
            BT_CHECK(cudaMemsetAsync(ws.hctl + kHctlECount, 0, sizeof(int) * (BT_HCTL_SIZE - kHctlECount), s));
            list3_heavy_seed_kernel<T, DIM><<<kNumSMs, 256, 0, s>>>(t, x, ws);
            BT_LAUNCH_CHECK();
            for (int st = 0; st < nsteps; ++st) {
                list3_heavy_step_kernel<T, DIM, true><<<kNumSMs * 8, 256, 0, s>>>(t, x, ntgt, st, G, ws);
                BT_LAUNCH_CHECK();
            }
            BT_TRY(heavy_sort_and_scatter(ws, heavy_total_host, t.nboxes, ntgt, nrows, rowlen, G, lists, s));
        }
    }
    return BT_OK;
}

// ---- lists 1 and 3 from ONE walk ---------------------------------------------
// The list-3 walk of a target box b starts at b's colleagues and descends through the boxes
// adjacent to b (traversal.py:673-870); the list-1 walk (traversal.py:470-550) starts at the
// root and descends through the same adjacent boxes, appending the source boxes among them.
// Below b's level the two walks visit the same boxes with the same adjacency test, so one
// walk over the roots coll(b) U {b} (in depth-first order) yields both lists: an adjacent
// source box goes to list 1 (kVisitNear), a non-adjacent one is handled by the list-3 rules.
// What the list-1 walk appends ABOVE b's level -- source boxes among the colleagues of b's
// ancestors (and the ancestors themselves) that are adjacent to b -- is collected by climbing
// the parent chain (skipping ancestors whose xflags say there is no source box around) and
// merged in by depth-first rank between the roots' subtrees, which restores the append order
// of the reference's walk.  Slot nlevels+1 of the G array holds list 1.
constexpr int kNearMax = 24;      // near-field boxes above the target's level kept per row

template <typename T, int DIM, bool FILL>
struct L13Policy {
    static constexpr int NB = 1 << DIM;
    const TreeView<T, DIM>& t; const T* rad; const List3Args<T, DIM>& x; int ntgt; int* G; int* lists;
    HeavyWs ws; const unsigned char* xflags; int* slots; int* abox; int* arank;
    L3Ctx<T, DIM> c; int box, icoll, nroots, selfpos, jb, jb_next, my_cb, my_rank, nA, ia, l1cur, near_cap;
    unsigned nearm, expm, okm; bool skip, l1ok, aovf;
    // count pass: the row's entries are also staged, tagged with their slot, in append order
    unsigned* stage; int stg, stage_cap;
    __device__ L13Policy(const TreeView<T, DIM>& t_, const T* rad_, const List3Args<T, DIM>& x_, int n, int* g,
                         int* li, const HeavyWs& w, const unsigned char* xf, int* sl, int* ab, int* ar, int cap)
        : t(t_), rad(rad_), x(x_), ntgt(n), G(g), lists(li), ws(w), xflags(xf), slots(sl), abox(ab), arank(ar),
          near_cap(cap), stage(nullptr), stg(0), stage_cap(FILL ? 0 : w.stage_cap) {}
    // entries must come out in append order (fill pass, or count pass with staging)
    __device__ __forceinline__ bool ordered() const { return FILL || stage_cap > 0; }
    __device__ __forceinline__ void put_near(int idx_off, int v)
    {   // one near-field entry at offset idx_off from the current cursors (caller advances them)
        if (FILL) lists[l1cur + idx_off] = v;
        else if (stg + idx_off < stage_cap)
            stage[stg + idx_off] = ((unsigned)(t.nlevels + 1) << kStageTagShift) | (unsigned)v;
    }
    __device__ __forceinline__ void init(int row, bool valid)
    {
        const int lane = threadIdx.x & 31, gl = lane % NB;
        const unsigned gm = ((1u << NB) - 1u) << (lane - gl);
        box = 0; icoll = 0; nroots = 0; selfpos = 0; jb = 0; jb_next = 0; my_cb = 0; my_rank = 0;
        nA = 0; ia = 0; l1cur = 0; nearm = expm = okm = 0; stg = 0;
        skip = true; l1ok = false; aovf = false;
        if (valid) {
            box = x.target_boxes[row];
            l3_make_ctx<T, DIM>(t, rad, x, box, c);
            skip = (FILL && ws.row_heavy[row]) || (ws.row_mask && !ws.row_mask[box]);
            if (stage_cap > 0) stage = ws.stage + (int64_t)row * stage_cap;
        }
        const int64_t rowlen = (int64_t)ntgt + 1;
        __syncwarp();
        if (valid) {
            for (int l = gl; l <= t.nlevels; l += NB) slots[l] = FILL ? G[l * rowlen + row] : 0;
            if (FILL) l1cur = G[(int64_t)(t.nlevels + 1) * rowlen + row];
        }
        __syncwarp();
        if (!valid || skip) return;
        // from here on the groups of a warp diverge; every lane of a group does the same
        icoll = x.coll_starts[box];
        const int ncoll = x.coll_starts[box + 1] - icoll;
        nroots = ncoll + 1;
        const int brank = ws.dfs_rank[box];
        int pos = 0;
        for (int j = gl; j < ncoll; j += NB) pos += (ws.dfs_rank[x.coll_lists[icoll + j]] < brank) ? 1 : 0;
#pragma unroll
        for (int o = NB >> 1; o > 0; o >>= 1) pos += __shfl_xor_sync(gm, pos, o);
        selfpos = pos;
        // near-field source boxes above b's level: colleagues of the ancestors (+ the ancestors)
        int a = box;
        for (int lv = c.tgt_level - 1; lv >= 0; --lv) {
            a = t.parents[a];
            if (!(xflags[a] & kXfCollSource)) continue;
            const int cs = x.coll_starts[a], n = x.coll_starts[a + 1] - cs;
            for (int j0 = 0; j0 <= n; j0 += NB) {
                const int j = j0 + gl;
                const int sbox = (j < n) ? x.coll_lists[cs + j] : (j == n ? a : -1);
                bool take = false;
                if (sbox >= 0 && (t.flags[sbox] & BT_BOX_IS_SOURCE_BOX)) {
                    if (sbox == 0) take = true;             // the root is appended up front (:489-495)
                    else {
                        T sc[DIM]; t.center(sbox, sc);
                        take = adj_nbhd<T, DIM>(rad, c.tc, c.tgt_level, (T)1, sc, lv);
                    }
                }
                const unsigned m = (__ballot_sync(gm, take) >> (lane - gl)) & ((1u << NB) - 1u);
                if (take) {
                    const int idx = nA + __popc(m & ((1u << gl) - 1u));
                    if (idx < near_cap) { abox[idx] = sbox; arank[idx] = ws.dfs_rank[sbox]; }
                }
                nA += __popc(m);
            }
        }
        if (nA > near_cap) { aovf = true; skip = true; nA = 0; return; }   // -> heavy-row path
        __syncwarp(gm);
        if (ordered() && nA > 1) {
            if (gl == 0) {
                for (int i = 1; i < nA; ++i) {
                    const int rk = arank[i], bx = abox[i];
                    int j = i - 1;
                    while (j >= 0 && arank[j] > rk) { arank[j + 1] = arank[j]; abox[j + 1] = abox[j]; --j; }
                    arank[j + 1] = rk; abox[j + 1] = bx;
                }
            }
            __syncwarp(gm);
        }
    }
    __device__ __forceinline__ void flush_near(int limit, bool writer)
    {
        while (ia < nA && arank[ia] < limit) {
            if (writer) put_near(0, abox[ia]);
            ++l1cur; ++stg; ++ia;
        }
    }
    // the roots coll(b) U {b} are classified 2^d at a time, one root per lane of the group:
    // near = the root itself belongs to list 1, exp = its children have to be visited
    __device__ __forceinline__ void load_roots(int gl, unsigned gm, int gshift)
    {
        const int j = jb + gl;
        my_cb = 0; my_rank = 0;
        bool near = false, exp = false, ok = false;
        if (j < nroots) {
            const int cb = (j == selfpos) ? box : x.coll_lists[icoll + j - (j > selfpos ? 1 : 0)];
            my_cb = cb;
            const unsigned char fl = t.flags[cb];
            if (ordered() && nA > 0) my_rank = ws.dfs_rank[cb];
            bool adj = true;
            if (cb != box && t.n_away != 1) {
                T sc[DIM]; t.center(cb, sc);
                adj = adj_nbhd<T, DIM>(rad, c.tc, c.tgt_level, (T)1, sc, c.tgt_level);
            }
            near = adj && (fl & BT_BOX_IS_SOURCE_BOX);
            if (cb == box) {
                // the list-3 walk never enters b itself; list 1 does if there are sources below
                exp = (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES) != 0; ok = true;
            } else {
                exp = (xflags[cb] & kXfHasChild) != 0;       // a leaf: its walk visits nothing
                ok = adj && (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES);
            }
        }
        constexpr unsigned gmask = (1u << NB) - 1u;
        nearm = (__ballot_sync(gm, near) >> gshift) & gmask;
        expm = (__ballot_sync(gm, exp) >> gshift) & gmask;
        okm = (__ballot_sync(gm, ok) >> gshift) & gmask;
    }
    __device__ __forceinline__ bool next_root(int& parent)
    {
        if (skip) return false;
        constexpr unsigned gmask = (1u << NB) - 1u;
        const int lane = threadIdx.x & 31, gl = lane % NB, gshift = lane - gl;
        const unsigned gm = gmask << gshift;
        const bool writer = gl == 0;
        while (true) {
            if (!(nearm | expm)) {
                if (jb_next >= nroots) break;
                jb = jb_next; jb_next += NB;
                load_roots(gl, gm, gshift);
                continue;
            }
            const int first_exp = expm ? (__ffs(expm) - 1) : NB;
            const unsigned low = (first_exp >= NB) ? gmask : ((2u << first_exp) - 1u);
            const unsigned nb = nearm & low;     // near roots up to (and including) the next expanded one
            if (nb) {
                if (ordered() && ia < nA) {
                    // near-field boxes from above the level are pending: merge by depth-first rank
                    for (unsigned m = nb; m; m &= m - 1) {
                        const int i = __ffs(m) - 1;
                        const int rk = __shfl_sync(gm, my_rank, gshift + i);
                        const int cbi = __shfl_sync(gm, my_cb, gshift + i);
                        flush_near(rk, writer);
                        if (writer) put_near(0, cbi);
                        ++l1cur; ++stg;
                    }
                } else {
                    if ((nb >> gl) & 1u) put_near(__popc(nb & ((1u << gl) - 1u)), my_cb);
                    l1cur += __popc(nb); stg += __popc(nb);
                }
                nearm &= ~nb;
            }
            if (first_exp < NB) {
                expm &= ~(1u << first_exp);
                const int cbi = __shfl_sync(gm, my_cb, gshift + first_exp);
                if (ordered() && ia < nA) flush_near(__shfl_sync(gm, my_rank, gshift + first_exp), writer);
                l1ok = ((okm >> first_exp) & 1u) != 0;
                parent = cbi;
                return true;
            }
        }
        if (ordered()) flush_near(0x7fffffff, writer);
        else { l1cur += nA - ia; ia = nA; }
        return false;
    }
    __device__ __forceinline__ int visit(int wb) const
    {
        const int act = list3_visit<T, DIM>(t, rad, x, c, wb);
        return l1ok ? act : (act & ~kVisitNear);
    }
    __device__ __forceinline__ void emit(unsigned eb, unsigned cb, unsigned qb, int ch, int gl, int parent)
    {
        // called by every lane of the warp (all masks 0 for idle groups)
        int lev = 0, base_e = 0, base_c = 0;
        if (eb | cb) {
            lev = t.levels[parent] + 1;          // siblings share their level
            base_e = slots[lev]; base_c = slots[t.nlevels];
        }
        __syncwarp();
        if (FILL) {
            if ((eb >> gl) & 1u) lists[base_e + __popc(eb & ((1u << gl) - 1u))] = ch;
            if ((cb >> gl) & 1u) lists[base_c + __popc(cb & ((1u << gl) - 1u))] = ch;
            if ((qb >> gl) & 1u) lists[l1cur + __popc(qb & ((1u << gl) - 1u))] = ch;
        } else if (stage_cap > 0) {
            const unsigned all = eb | cb | qb;   // a child is in at most one of the three
            if ((all >> gl) & 1u) {
                const int idx = stg + __popc(all & ((1u << gl) - 1u));
                const int tag = ((eb >> gl) & 1u) ? lev : ((cb >> gl) & 1u) ? t.nlevels : t.nlevels + 1;
                if (idx < stage_cap) stage[idx] = ((unsigned)tag << kStageTagShift) | (unsigned)ch;
            }
            stg += __popc(all);
        }
        l1cur += __popc(qb);
        if (gl == 0 && (eb | cb)) { slots[lev] = base_e + __popc(eb); slots[t.nlevels] = base_c + __popc(cb); }
        __syncwarp();
    }
    __device__ __forceinline__ void finish(int row, bool valid, bool ok, int gl)
    {
        __syncwarp();
        if (FILL || !valid) return;
        const bool good = ok && !aovf;
        const int64_t rowlen = (int64_t)ntgt + 1;
        for (int l = gl; l <= t.nlevels; l += NB) G[l * rowlen + row] = good ? slots[l] : 0;
        if (gl == 0) {
            G[(int64_t)(t.nlevels + 1) * rowlen + row] = good ? l1cur : 0;
            const bool staged = good && stage_cap > 0 && stg <= stage_cap;
            ws.row_heavy[row] = good ? (staged ? 2 : 0) : 1;
            if (staged) ws.stage_count[row] = stg;
            if (!good) ws.heavy_rows[atomicAdd(ws.hctl + kHctlNHeavy, 1)] = row;
            else if (!staged) atomicAdd(ws.hctl + kHctlNWalk, 1);
        }
    }
};

// fill pass for staged rows: the entries of a row, tagged with their slot, are copied to the
// slot's position in append order (one warp per row, ranks by match_any)
__global__ void __launch_bounds__(256)
list13_unstage_kernel(int ntgt, int nslots, const int* __restrict__ G, const unsigned char* __restrict__ row_state,
                      const unsigned* __restrict__ stage, int stage_cap, const int* __restrict__ stage_count,
                      int* __restrict__ lists)
{
    __shared__ int cnt[8][kMaxWalkLevels + 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int64_t rowlen = (int64_t)ntgt + 1;
    for (int row = w; row < ntgt; row += nw) {
        if (row_state[row] != 2) continue;
        const int n = stage_count[row];
        if (n <= 0) continue;
        const unsigned* src = stage + (int64_t)row * stage_cap;
        for (int l = lane; l < nslots; l += 32) cnt[wib][l] = 0;
        __syncwarp();
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < n;
            const unsigned e = valid ? src[i] : 0u;
            const int sl = valid ? (int)(e >> kStageTagShift) : 63;
            const unsigned peers = __match_any_sync(0xffffffffu, sl);
            const int prior = valid ? cnt[wib][sl] : 0;
            __syncwarp();
            if (valid) {
                lists[G[sl * rowlen + row] + prior + __popc(peers & ((1u << lane) - 1u))] =
                    (int)(e & ((1u << kStageTagShift) - 1u));
                if (lane == __ffs(peers) - 1) cnt[wib][sl] = prior + __popc(peers);
            }
            __syncwarp();
        }
    }
}

// 8 blocks of 128 threads per SM (64 registers): the walk is latency/issue bound and gains from
// occupancy (count pass 3.25 -> 2.18 ms on 1e7 uniform points against the 94-register build)
#ifndef BT_L13_MIN_BLOCKS
#define BT_L13_MIN_BLOCKS 8
#endif
template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock, BT_L13_MIN_BLOCKS)
list13_coop_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, const unsigned char* __restrict__ xflags, int ntgt,
                   int* __restrict__ G, int* __restrict__ lists, HeavyWs ws, int near_cap)
{
    __shared__ T rad[kMaxWalkLevels];
    __shared__ CoopFrame frames[(kTravBlock >> DIM) * kMaxWalkLevels];
    __shared__ int slots[(kTravBlock >> DIM) * (kMaxWalkLevels + 1)];
    __shared__ int near_box[(kTravBlock >> DIM) * kNearMax], near_rank[(kTravBlock >> DIM) * kNearMax];
    fill_rad_table(rad, t.root_extent);
    const int grp = threadIdx.x >> DIM;
    L13Policy<T, DIM, FILL> pol(t, rad, x, ntgt, G, lists, ws, xflags, slots + grp * (kMaxWalkLevels + 1),
                                near_box + grp * kNearMax, near_rank + grp * kNearMax, near_cap);
    coop_walk_rows<T, DIM>(t, pol, ntgt, FILL ? 0x7ffffff0 : ws.budget, frames);
}

// heavy rows of the fused walk: frontier item = heavy-row index << 32 | near-ok << 31 | walk
// parent; sort key = slot | heavy-row index | depth-first rank (the index keeps the key short)
__device__ __forceinline__ void heavy_count_add(int* G, int64_t rowlen, bool emits, int slot, int r)
{   // one atomic per distinct (slot, row) of the warp; called by full warps
    const unsigned long long ckey = emits ? (((unsigned long long)slot << 32) | (unsigned)r) : ~0ull;
    const unsigned peers = __match_any_sync(0xffffffffu, ckey);
    if (emits && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(G + slot * rowlen + r, __popc(peers));
}

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(256)
list13_heavy_seed_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, const unsigned char* __restrict__ xflags,
                         int ntgt, int* __restrict__ G, HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int rank_bits = bits_for(t.nboxes), row_bits = bits_for(nheavy);
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int l1slot = t.nlevels + 1;
    const int stride = gridDim.x * blockDim.x;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += stride) {
        const int r = ws.heavy_rows[h];
        const int box = x.target_boxes[r];
        T tc[DIM]; t.center(box, tc);
        const int level = t.levels[box];
        auto near = [&](int sbox) {
            if (FILL) {
                const int k = atomicAdd(ws.hctl + kHctlECount, 1);
                if (k < ws.ecap)       // the box is recovered from its rank after the sort
                    ws.ekeys[0][k] = ((unsigned long long)l1slot << (rank_bits + row_bits))
                                     | ((unsigned long long)h << rank_bits) | (unsigned)ws.dfs_rank[sbox];
            } else atomicAdd(G + l1slot * rowlen + r, 1);
        };
        int a = box;
        for (int lv = level - 1; lv >= 0; --lv) {
            a = t.parents[a];
            if (!(xflags[a] & kXfCollSource)) continue;
            const int cs = x.coll_starts[a], n = x.coll_starts[a + 1] - cs;
            for (int j = 0; j <= n; ++j) {
                const int sbox = (j < n) ? x.coll_lists[cs + j] : a;
                if (!(t.flags[sbox] & BT_BOX_IS_SOURCE_BOX)) continue;
                bool take = (sbox == 0);
                if (!take) { T sc[DIM]; t.center(sbox, sc); take = adj_nbhd<T, DIM>(rad, tc, level, (T)1, sc, lv); }
                if (take) near(sbox);
            }
        }
        const int cs = x.coll_starts[box], n = x.coll_starts[box + 1] - cs;
        for (int j = 0; j <= n; ++j) {
            const int cb = (j < n) ? x.coll_lists[cs + j] : box;
            const unsigned char fl = t.flags[cb];
            bool adj = true;
            if (cb != box && t.n_away != 1) {
                T sc[DIM]; t.center(cb, sc);
                adj = adj_nbhd<T, DIM>(rad, tc, level, (T)1, sc, level);
            }
            if (adj && (fl & BT_BOX_IS_SOURCE_BOX)) near(cb);
            bool nearok;
            if (cb == box) { if (!(fl & BT_BOX_HAS_SOURCE_CHILD_BOXES)) continue; nearok = true; }
            else { if (!(xflags[cb] & kXfHasChild)) continue; nearok = adj && (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES); }
            const int q = atomicAdd(ws.hctl + kHctlFrontier, 1);
            if (q < ws.frontier_cap)
                ws.frontier[0][q] = ((unsigned long long)h << 32) | (nearok ? 0x80000000ull : 0ull) | (unsigned)cb;
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(256)
list13_heavy_step_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int ntgt, int step, int* __restrict__ G,
                         HeavyWs ws)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const unsigned long long* fin = ws.frontier[step & 1];
    unsigned long long* fout = ws.frontier[(step + 1) & 1];
    long long nitems = ws.hctl[kHctlFrontier + step];
    if (nitems > ws.frontier_cap) nitems = ws.frontier_cap;
    const long long total = nitems * NB;
    const int rank_bits = bits_for(t.nboxes), row_bits = bits_for(ws.hctl[kHctlNHeavy]);
    const int64_t rowlen = (int64_t)ntgt + 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < ((total + 31) & ~31ll);
         tid += stride) {
        int act = 0, wb = 0, r = 0, h = 0;
        unsigned long long nearok = 0;
        if (tid < total) {
            const unsigned long long item = fin[tid / NB];
            h = (int)(item >> 32);
            r = ws.heavy_rows[h];
            nearok = item & 0x80000000ull;
            const int parent = (int)(item & 0x7fffffffull), m = (int)(tid % NB);
            L3Ctx<T, DIM> c; l3_make_ctx<T, DIM>(t, rad, x, x.target_boxes[r], c);
            wb = t.child(parent, m);
            act = list3_visit<T, DIM>(t, rad, x, c, wb);
            if (!nearok) act &= ~kVisitNear;
        }
        const bool emits = (act & (kVisitEmit | kVisitClose | kVisitNear)) != 0;
        const int slot = (act & kVisitNear) ? t.nlevels + 1 : (act & kVisitEmit) ? (int)t.levels[wb] : t.nlevels;
        if (FILL) {
            const long long k = warp_append(emits, ws.hctl + kHctlECount);
            if (k >= 0 && k < ws.ecap)
                ws.ekeys[0][k] = ((unsigned long long)slot << (rank_bits + row_bits))
                                 | ((unsigned long long)h << rank_bits) | (unsigned)ws.dfs_rank[wb];
        } else heavy_count_add(G, rowlen, emits, slot, r);
        const long long q = warp_append(act & kVisitPush, ws.hctl + kHctlFrontier + step + 1);
        if (q >= 0) {
            if (q < ws.frontier_cap) fout[q] = ((unsigned long long)h << 32) | nearok | (unsigned)wb;
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

// start of every (slot, heavy row) group in the sorted entry array = exclusive scan of the
// groups' sizes, which the count phase left in G
struct HeavyGroupIn {
    const int* G; const int* heavy_rows; int nheavy; int64_t rowlen;
    __device__ int operator()(int64_t i) const
    {
        const int64_t slot = i / nheavy; const int r = heavy_rows[i % nheavy];
        return G[slot * rowlen + r + 1] - G[slot * rowlen + r];
    }
};

__global__ void __launch_bounds__(256)
heavy_scatter_groups_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ dfs_order,
                            const int* __restrict__ ecount_dev, int rank_bits, int row_bits, int nheavy,
                            const int* __restrict__ heavy_rows, const int* __restrict__ group_start,
                            int64_t rowlen, const int* __restrict__ G, int* __restrict__ lists)
{
    const int n = *ecount_dev;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = ld_stream_u64(keys + i);
        const int h = (int)((key >> rank_bits) & ((1ull << row_bits) - 1ull));
        const int64_t slot = (int64_t)(key >> (rank_bits + row_bits));
        lists[G[slot * rowlen + heavy_rows[h]] + (i - group_start[slot * nheavy + h])] =
            dfs_order[(int)(key & ((1ull << rank_bits) - 1ull))];
    }
}

static int heavy_sort_and_scatter_groups(const HeavyWs& ws, long long ecount_host, int nboxes, int nheavy,
                                         int nslots, int64_t rowlen, const int* G, int* lists, cudaStream_t s)
{
    if (ecount_host <= 0 || nheavy <= 0) return BT_OK;
    int rank_bits = 1; while ((1ll << rank_bits) < nboxes) ++rank_bits;
    int row_bits = 1; while ((1ll << row_bits) < nheavy) ++row_bits;
    int slot_bits = 0; while ((1ll << slot_bits) < nslots) ++slot_bits;
    if (rank_bits + row_bits + slot_bits > 64) return BT_ERR_UNSUPPORTED;
    int in_alt = 0;
    BT_TRY(radix_sort_pairs(ecount_host, ws.ekeys[0], ws.ekeys[1], nullptr, nullptr, 0, 0,
                            rank_bits + row_bits + slot_bits, &in_alt, s, "l13_heavy_sort_pass"));
    BT_PROF("l13_heavy_scatter", s);
    int* group_start = nullptr;
    const int64_t ngroups = (int64_t)nslots * nheavy;
    BT_CHECK(temp_alloc((void**)&group_start, sizeof(int) * (ngroups + 1), s));
    HeavyGroupIn in{G, ws.heavy_rows, nheavy, rowlen};
    PlainOut out{group_start, nullptr, ngroups};
    BT_TRY(scan_exclusive(ngroups, nullptr, in, out, s));
    heavy_scatter_groups_kernel<<<grid_for(ecount_host, 256, 8), 256, 0, s>>>(
        ws.ekeys[in_alt], ws.dfs_order, ws.hctl + kHctlECount, rank_bits, row_bits, nheavy, ws.heavy_rows,
        group_start, rowlen, G, lists);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(group_start, s));
    return BT_OK;
}

// ---- heavy rows by position map (no sort) ------------------------------------
// The entries of a heavy row are boxes below its roots coll(b) U {b} (plus the few near-field
// boxes above b's level), and their order in every list is the depth-first rank.  The rank
// ranges of the roots' subtrees are disjoint, so "depth-first order inside the row" is a
// position in the concatenation of those ranges: a row owns a byte map over that
// concatenation (row h: segments sorted by rank, each a root's subtree or one near-field box
// from above), the breadth-first expansion stores slot + 1 at the position of every box it
// appends -- one byte store, no atomics, no keys -- and one ordered pass over the map writes
// the lists.  Rows are padded to whole chunks of kMapChunk positions; chunk histograms give
// the per-(row, slot) counts (-> G) and the offsets of every chunk inside its row's lists.
constexpr int kMapChunk = 1024;
constexpr int kMapNearExtra = 136;       // >= 7 near-field boxes per level above b (3-D), 19 levels

// phase 0: size of every heavy row's map
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
heavy_map_size_kernel(List3Args<T, DIM> x, int* __restrict__ row_len, HeavyWs ws)
{
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int stride = gridDim.x * blockDim.x;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += stride) {
        const int box = x.target_boxes[ws.heavy_rows[h]];
        long long tot = ws.subtree_size[box] + kMapNearExtra;
        for (int i = x.coll_starts[box]; i < x.coll_starts[box + 1]; ++i) tot += ws.subtree_size[x.coll_lists[i]];
        tot = (tot + kMapChunk - 1) / kMapChunk * kMapChunk;
        row_len[h] = (int)(tot / kMapChunk);                 // in chunks
    }
}
struct RowLenIn {
    const int* len;
    __device__ int operator()(int64_t i) const { return len[i]; }
};
struct RowBaseOut {
    long long* base; long long* total_out; const int* n_dev;
    __device__ void operator()(int64_t i, long long excl) const { base[i] = excl * kMapChunk; }
    __device__ void total(long long t) const { base[*n_dev] = t * kMapChunk; *total_out = t * kMapChunk; }
};

// segment table of every heavy row: roots in depth-first order, near-field boxes from above
// (kind 1) inserted by rank; then the row's roots are seeded into the frontier
template <typename T, int DIM>
__global__ void __launch_bounds__(128)
heavy_map_plan_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, const unsigned char* __restrict__ xflags,
                      HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int S = ws.seg_stride;
    const int stride = gridDim.x * blockDim.x;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nheavy; h += stride) {
        const int r = ws.heavy_rows[h];
        const int box = x.target_boxes[r];
        T tc[DIM]; t.center(box, tc);
        const int level = t.levels[box];
        int* srank = ws.hseg_rank + (int64_t)h * S;
        int* spre = ws.hseg_prefix + (int64_t)h * S;
        unsigned char* skind = ws.hseg_kind + (int64_t)h * S;
        // roots, self merged in by rank
        const int cs = x.coll_starts[box], n = x.coll_starts[box + 1] - cs;
        const int brank = ws.dfs_rank[box];
        int ns = 0;
        bool self_done = false;
        for (int j = 0; j < n; ++j) {
            const int rk = ws.dfs_rank[x.coll_lists[cs + j]];
            if (!self_done && brank < rk) { srank[ns] = brank; skind[ns] = 0; ++ns; self_done = true; }
            srank[ns] = rk; skind[ns] = 0; ++ns;
        }
        if (!self_done) { srank[ns] = brank; skind[ns] = 0; ++ns; }
        // near-field boxes above the level (what list 1's walk appends before it reaches b's level)
        int a = box;
        for (int lv = level - 1; lv >= 0; --lv) {
            a = t.parents[a];
            if (!(xflags[a] & kXfCollSource)) continue;
            const int as = x.coll_starts[a], an = x.coll_starts[a + 1] - as;
            for (int j = 0; j <= an; ++j) {
                const int sbox = (j < an) ? x.coll_lists[as + j] : a;
                if (!(t.flags[sbox] & BT_BOX_IS_SOURCE_BOX)) continue;
                bool take = (sbox == 0);
                if (!take) { T sc[DIM]; t.center(sbox, sc); take = adj_nbhd<T, DIM>(rad, tc, level, (T)1, sc, lv); }
                if (!take || ns >= S) continue;
                const int rk = ws.dfs_rank[sbox];
                int k = ns;
                while (k > 0 && srank[k - 1] > rk) { srank[k] = srank[k - 1]; skind[k] = skind[k - 1]; --k; }
                srank[k] = rk; skind[k] = 1; ++ns;
            }
        }
        int pre = 0;
        for (int k = 0; k < ns; ++k) {
            spre[k] = pre;
            pre += skind[k] ? 1 : ws.subtree_size[ws.dfs_order[srank[k]]];
        }
        ws.hseg_n[h] = ns;
        // the row's walk constants, read by every child visit of the expansion
        static_assert(sizeof(L3Ctx<T, DIM>) <= 128, "bt_heavy_ws.hctx holds 128 bytes per heavy row");
        l3_make_ctx<T, DIM>(t, rad, x, box, reinterpret_cast<L3Ctx<T, DIM>*>(ws.hctx)[h]);
    }
}

// frontier item of the map expansion: heavy-row index (23 bits) | segment of the row's map the
// walk parent lies in (9 bits; a root's whole subtree is one segment, so the walk never leaves
// it and the position of a box needs no search) | near-ok | walk parent (31 bits)
constexpr int kMapSegBits = 9, kMapRowBits = 23;
__device__ __forceinline__ unsigned long long heavy_map_item(int h, int seg, bool nearok, int parent)
{
    return ((unsigned long long)h << (32 + kMapSegBits)) | ((unsigned long long)seg << 32) |
           (nearok ? 0x80000000ull : 0ull) | (unsigned)parent;
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
heavy_map_seed_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, const unsigned char* __restrict__ xflags,
                      HeavyWs ws)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int S = ws.seg_stride;
    const unsigned char l1code = (unsigned char)(t.nlevels + 2);        // slot nlevels + 1, plus 1
    // one thread per (heavy row, segment)
    const long long total = (long long)nheavy * S;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int h = (int)(i / S), k = (int)(i % S);
        if (k >= ws.hseg_n[h]) continue;
        const long long pos = ws.hrow_base[h] + ws.hseg_prefix[i];
        if (ws.hseg_kind[i]) { ws.hmap[pos] = l1code; continue; }        // near-field box from above
        const int cb = ws.dfs_order[ws.hseg_rank[i]];
        const int box = x.target_boxes[ws.heavy_rows[h]];
        const unsigned char fl = t.flags[cb];
        bool adj = true;
        if (cb != box && t.n_away != 1) {
            T tc[DIM], sc[DIM]; t.center(box, tc); t.center(cb, sc);
            const int level = t.levels[box];
            adj = adj_nbhd<T, DIM>(rad, tc, level, (T)1, sc, level);
        }
        if (adj && (fl & BT_BOX_IS_SOURCE_BOX)) ws.hmap[pos] = l1code;
        bool nearok;
        if (cb == box) { if (!(fl & BT_BOX_HAS_SOURCE_CHILD_BOXES)) continue; nearok = true; }
        else { if (!(xflags[cb] & kXfHasChild)) continue; nearok = adj && (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES); }
        const int q = atomicAdd(ws.hctl + kHctlFrontier, 1);
        if (q < ws.frontier_cap)
            ws.frontier[0][q] = heavy_map_item(h, k, nearok, cb);
        else ws.hctl[kHctlOverflow] = 1;
    }
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
heavy_map_step_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int step, HeavyWs ws)
{
    constexpr int NB = 1 << DIM;
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const unsigned long long* fin = ws.frontier[step & 1];
    unsigned long long* fout = ws.frontier[(step + 1) & 1];
    long long nitems = ws.hctl[kHctlFrontier + step];
    if (nitems > ws.frontier_cap) nitems = ws.frontier_cap;
    const long long total = nitems * NB;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < ((total + 31) & ~31ll);
         tid += stride) {
        int act = 0, wb = 0, h = 0, seg = 0;
        unsigned long long nearok = 0;
        if (tid < total) {
            const unsigned long long item = fin[tid / NB];
            h = (int)(item >> (32 + kMapSegBits));
            seg = (int)(item >> 32) & ((1 << kMapSegBits) - 1);
            nearok = item & 0x80000000ull;
            const int parent = (int)(item & 0x7fffffffull), m = (int)(tid % NB);
            const L3Ctx<T, DIM> c = reinterpret_cast<const L3Ctx<T, DIM>*>(ws.hctx)[h];
            wb = t.child(parent, m);
            act = list3_visit<T, DIM>(t, rad, x, c, wb);
            if (!nearok) act &= ~kVisitNear;
        }
        if (act & (kVisitEmit | kVisitClose | kVisitNear)) {
            const int slot = (act & kVisitNear) ? t.nlevels + 1 : (act & kVisitEmit) ? (int)t.levels[wb] : t.nlevels;
            const int64_t sg = (int64_t)h * ws.seg_stride + seg;
            ws.hmap[ws.hrow_base[h] + ws.hseg_prefix[sg] + (ws.dfs_rank[wb] - ws.hseg_rank[sg])] =
                (unsigned char)(slot + 1);
        }
        const long long q = warp_append(act & kVisitPush, ws.hctl + kHctlFrontier + step + 1);
        if (q >= 0) {
            if (q < ws.frontier_cap) fout[q] = heavy_map_item(h, seg, nearok != 0, wb);
            else ws.hctl[kHctlOverflow] = 1;
        }
    }
}

__device__ __forceinline__ int heavy_row_of_chunk(const HeavyWs& ws, int nheavy, long long chunk)
{
    const long long p = chunk * kMapChunk;
    int lo = 0, hi = nheavy;                              // last row with base <= p
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ws.hrow_base[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}

// per-chunk histogram of the slots (one warp per chunk)
__global__ void __launch_bounds__(256)
heavy_map_hist_kernel(int nslots, HeavyWs ws)
{
    __shared__ int cnt[8][kMaxWalkLevels + 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long nchunks = ws.hplan[0] / kMapChunk;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = w; c < nchunks; c += nw) {
        for (int l = lane; l < nslots; l += 32) cnt[wib][l] = 0;
        __syncwarp();
        const unsigned* src = reinterpret_cast<const unsigned*>(ws.hmap + c * kMapChunk);
        for (int i = lane; i < kMapChunk / 4; i += 32) {
            unsigned v = src[i];
#pragma unroll
            for (int b = 0; b < 4; ++b) { const unsigned code = (v >> (8 * b)) & 0xffu; if (code) atomicAdd(&cnt[wib][code - 1], 1); }
        }
        __syncwarp();
        for (int l = lane; l < nslots; l += 32) ws.chunk_cnt[c * nslots + l] = cnt[wib][l];
        __syncwarp();
    }
}

// per heavy row and slot: chunk counts -> exclusive offsets inside the row; row totals -> G.  One
// warp per (row, slot), 32 chunks per step (lane = chunk).  (A warp per ROW looping over the
// slots left one SM scanning the few rows with thousands of chunks -- an upper-level box's
// map covers a large part of the tree -- long after the others were done.)
__global__ void __launch_bounds__(256)
heavy_map_rowscan_kernel(int nslots, int64_t rowlen, int* __restrict__ G, HeavyWs ws)
{
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long idx = w; idx < (long long)nheavy * nslots; idx += nw) {
        const int h = (int)(idx / nslots), sl = (int)(idx % nslots);
        const long long c0 = ws.hrow_base[h] / kMapChunk, c1 = ws.hrow_base[h + 1] / kMapChunk;
        int carry = 0;
        for (long long cb = c0; cb < c1; cb += 32) {
            const long long c = cb + lane;
            const int v = (c < c1) ? ws.chunk_cnt[c * nslots + sl] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (c < c1) ws.chunk_cnt[c * nslots + sl] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) G[sl * rowlen + ws.heavy_rows[h]] = carry;
    }
}

// fill: ordered pass over the map (one warp per chunk).  Most positions are empty (the walk only
// visits the shell around the target box), so the chunk's entries are first compacted, in position
// order, into shared memory (16 map bytes per lane and load); ranking (match_any), the segment
// search and the stores then run on full warps.
__global__ void __launch_bounds__(256)
heavy_map_extract_kernel(int nslots, int64_t rowlen, const int* __restrict__ G, int* __restrict__ lists,
                         HeavyWs ws)
{
    static_assert(kMapChunk == 1024 && kMaxWalkLevels + 2 < 64, "entry = position (10 bits) | code << 10");
    __shared__ int run[8][kMaxWalkLevels + 2];
    __shared__ unsigned short buf[8][kMapChunk];
    const int nheavy = ws.hctl[kHctlNHeavy];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long nchunks = ws.hplan[0] / kMapChunk;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = w; c < nchunks; c += nw) {
        const long long base = c * kMapChunk;
        // compaction: lane l of load i holds positions (32 i + l) * 16 .. + 15
        const uint4* src = reinterpret_cast<const uint4*>(ws.hmap + base);
        int n = 0;
#pragma unroll
        for (int i = 0; i < kMapChunk / 16 / 32; ++i) {
            const uint4 v = src[i * 32 + lane];
            const unsigned wd[4] = {v.x, v.y, v.z, v.w};
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) cnt += ((wd[k] >> (8 * b)) & 0xffu) ? 1 : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            int off = n + inc - cnt;
            if (cnt) {
                const unsigned p0 = (unsigned)(i * 32 + lane) * 16u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const unsigned code = (wd[k] >> (8 * b)) & 0xffu;
                        if (code) buf[wib][off++] = (unsigned short)((p0 + 4 * k + b) | (code << 10));
                    }
            }
            n += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (n == 0) continue;                 // (warp-uniform)
        const int h = heavy_row_of_chunk(ws, nheavy, c);
        const int r = ws.heavy_rows[h];
        const int S = ws.seg_stride, ns = ws.hseg_n[h];
        const int* spre = ws.hseg_prefix + (int64_t)h * S;
        const int* srank = ws.hseg_rank + (int64_t)h * S;
        const int local0 = (int)(base - ws.hrow_base[h]);
        // segment of the chunk's first position (warp-uniform search); the entries of one chunk
        // span a segment or two, so each entry only steps forward from there
        int seg0 = 0;
        { int hi = ns; while (hi - seg0 > 1) { const int mid = (seg0 + hi) >> 1; if (spre[mid] <= local0) seg0 = mid; else hi = mid; } }
        // next output position of every slot: the row's start in the slot + the chunk's offset
        for (int l = lane; l < nslots; l += 32) run[wib][l] = G[l * rowlen + r] + ws.chunk_cnt[c * nslots + l];
        __syncwarp();
        for (int j0 = 0; j0 < n; j0 += 32) {
            const bool valid = j0 + lane < n;
            const unsigned e = valid ? buf[wib][j0 + lane] : 0u;
            const int sl = valid ? (int)(e >> 10) - 1 : 63;
            const unsigned peers = __match_any_sync(0xffffffffu, sl);
            const int prior = valid ? run[wib][sl] : 0;
            __syncwarp();
            if (valid) {
                const int local = local0 + (int)(e & 1023u);
                int lo = seg0;                        // last segment with prefix <= local
                while (lo + 1 < ns && spre[lo + 1] <= local) ++lo;
                lists[prior + __popc(peers & ((1u << lane) - 1u))] = ws.dfs_order[srank[lo] + (local - spre[lo])];
                if (lane == __ffs(peers) - 1) run[wib][sl] = prior + __popc(peers);
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

template <typename T, int DIM>
static int list13_impl(int phase, const bt_tree_view* tv, const bt_list3_args* a, const unsigned char* xflags,
                       int ntgt, int* G, int* C, int* lists, long long* summary, const bt_heavy_ws* w,
                       long long heavy_total_host, int nheavy_host, int nwalk_host, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nlevels + 1 > kMaxWalkLevels) return BT_ERR_UNSUPPORTED;
    if (w->stage_cap > 0 && (t.nlevels + 2 > 31 || t.nboxes >= (1 << bt::kStageTagShift))) return BT_ERR_BAD_ARG;
    HeavyWs ws = make_ws(w);
    List3Args<T, DIM> x{a->target_boxes, a->coll_starts, a->coll_lists, (T)a->stick_out_factor,
                        a->targets_have_extent, a->sources_have_extent, a->crit,
                        (const T*)a->box_target_bounding_box_min, (const T*)a->box_target_bounding_box_max,
                        a->box_source_counts_cumul, a->min_nsources_cumul};
    const int nrows = t.nlevels + 2;             // source levels, close list, list 1
    const int near_cap = (g_walk_mode & kModeNearCapZero) ? 0 : kNearMax;
    const bool use_map = ws.hrow_base != nullptr;   // heavy rows by position map (else: radix sort)
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int64_t total_len = rowlen * nrows;
    // two waves of resident blocks (measured: 2.9 ms against 3.0 ms with one wave on config 3 --
    // the rows' walks differ widely in length and the second wave evens the SMs out)
    static const int walk_bps = [] { const char* e = getenv("BT_L13_BLOCKS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
    const int cgrid = grid_for((int64_t)ntgt << DIM, kTravBlock, walk_bps);
    const int fgrid = cgrid;
    const int nsteps = t.nlevels;
    // counts -> global offsets (one flattened scan), non-empty ranks, summary for the host
    auto finish_counts = [&]() -> int {
        InPlaceIn in{G};
        PlainOut out{G, summary + 2 * (nrows + 1), total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, in, out, s));
        NonemptyIn nin{G, total_len, rowlen};
        PlainOut nout{C, nullptr, total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, nin, nout, s));
        list3_summary_kernel<<<1, 64, 0, s>>>(G, C, nrows, rowlen, summary);
        BT_LAUNCH_CHECK();
        return BT_OK;
    };
    if (phase == 0) {
        BT_CHECK(cudaMemsetAsync(G, 0, sizeof(int) * (total_len + 1), s));
        BT_CHECK(cudaMemsetAsync(ws.hctl, 0, sizeof(int) * BT_HCTL_SIZE, s));
        BT_CHECK(cudaMemsetAsync(ws.heavy_total, 0, sizeof(long long), s));
        if (use_map) BT_CHECK(cudaMemsetAsync(ws.hplan, 0, 2 * sizeof(long long), s));
        if (ntgt > 0) {
            {
                BT_PROF("l13_walk_count", s);
                list13_coop_kernel<T, DIM, false><<<cgrid, kTravBlock, 0, s>>>(t, x, xflags, ntgt, G, nullptr, ws, near_cap);
                BT_LAUNCH_CHECK();
            }
            if (use_map) {
                // sizes of the heavy rows' maps; the host reads (nheavy, total) and calls phase 2
                BT_PROF("l13_heavy_plan", s);
                int* row_len = nullptr;
                BT_CHECK(temp_alloc((void**)&row_len, sizeof(int) * (size_t)ntgt, s));
                heavy_map_size_kernel<T, DIM><<<kNumSMs, 256, 0, s>>>(x, row_len, ws);
                BT_LAUNCH_CHECK();
                RowLenIn in{row_len};
                RowBaseOut out{ws.hrow_base, ws.hplan, ws.hctl + kHctlNHeavy};
                BT_TRY(scan_exclusive(ntgt, ws.hctl + kHctlNHeavy, in, out, s));
                BT_CHECK(cudaFreeAsync(row_len, s));
                return BT_OK;
            }
            BT_PROF("l13_heavy_steps_count", s);
            list13_heavy_seed_kernel<T, DIM, false><<<kNumSMs, 256, 0, s>>>(t, x, xflags, ntgt, G, ws);
            BT_LAUNCH_CHECK();
            for (int st = 0; st < nsteps; ++st) {
                list13_heavy_step_kernel<T, DIM, false><<<kNumSMs * 8, 256, 0, s>>>(t, x, ntgt, st, G, ws);
                BT_LAUNCH_CHECK();
            }
            heavy_total_kernel<<<kNumSMs, 256, 0, s>>>(G, rowlen, nrows, ws);
            BT_LAUNCH_CHECK();
        }
        if (!use_map) BT_TRY(finish_counts());
    } else if (phase == 2) {
        // map mode: expand the heavy rows ONCE into their position maps, count from the maps
        if (!use_map) return BT_ERR_BAD_ARG;
        if (nheavy_host >= (1 << kMapRowBits) || ws.seg_stride > (1 << kMapSegBits)) return BT_ERR_UNSUPPORTED;
        if (ntgt > 0 && nheavy_host > 0) {
            BT_PROF("l13_heavy_expand", s);
            {
                BT_PROF("l13h_plan_seed", s);
                BT_CHECK(cudaMemsetAsync(ws.hmap, 0, (size_t)ws.hmap_cap, s));
                heavy_map_plan_kernel<T, DIM><<<grid_for(nheavy_host, 128, 8), 128, 0, s>>>(t, x, xflags, ws);
                BT_LAUNCH_CHECK();
                heavy_map_seed_kernel<T, DIM><<<grid_for((int64_t)nheavy_host * ws.seg_stride, 256, 8), 256, 0, s>>>(t, x, xflags, ws);
                BT_LAUNCH_CHECK();
            }
            {
                BT_PROF("l13h_steps", s);
                const int step_grid = grid_resident(heavy_map_step_kernel<T, DIM>, (int64_t)1 << 40, 256);
                for (int st = 0; st < nsteps; ++st) {
                    heavy_map_step_kernel<T, DIM><<<step_grid, 256, 0, s>>>(t, x, st, ws);
                    BT_LAUNCH_CHECK();
                }
            }
            {
                BT_PROF("l13h_counts", s);
                const long long nchunks = ws.hmap_cap / kMapChunk;
                heavy_map_hist_kernel<<<grid_resident(heavy_map_hist_kernel, nchunks * 32, 256), 256, 0, s>>>(nrows, ws);
                BT_LAUNCH_CHECK();
                heavy_map_rowscan_kernel<<<grid_for((int64_t)nheavy_host * nrows * 32, 256, 8), 256, 0, s>>>(nrows, rowlen, G, ws);
                BT_LAUNCH_CHECK();
                heavy_total_kernel<<<kNumSMs, 256, 0, s>>>(G, rowlen, nrows, ws);
                BT_LAUNCH_CHECK();
            }
        }
        BT_TRY(finish_counts());
    } else if (ntgt > 0) {
        if (ws.stage_cap > 0) {
            BT_PROF("l13_unstage", s);
            list13_unstage_kernel<<<grid_for((int64_t)ntgt * 32, 256, 8), 256, 0, s>>>(
                ntgt, nrows, G, ws.row_heavy, ws.stage, ws.stage_cap, ws.stage_count, lists);
            BT_LAUNCH_CHECK();
        }
        if (nwalk_host > 0) {        // rows the count pass could not stage are walked again
            BT_PROF("l13_walk_fill", s);
            list13_coop_kernel<T, DIM, true><<<fgrid, kTravBlock, 0, s>>>(t, x, xflags, ntgt, G, lists, ws, near_cap);
            BT_LAUNCH_CHECK();
        }
        if (use_map) {
            if (nheavy_host > 0) {
                BT_PROF("l13_heavy_extract", s);
                const long long nchunks = ws.hmap_cap / kMapChunk;
                heavy_map_extract_kernel<<<grid_resident(heavy_map_extract_kernel, nchunks * 32, 256), 256, 0, s>>>(
                    nrows, rowlen, G, lists, ws);
                BT_LAUNCH_CHECK();
            }
        } else if (heavy_total_host > 0) {
            {
                BT_PROF("l13_heavy_steps_fill", s);
                BT_CHECK(cudaMemsetAsync(ws.hctl + kHctlECount, 0, sizeof(int) * (BT_HCTL_SIZE - kHctlECount), s));
                list13_heavy_seed_kernel<T, DIM, true><<<kNumSMs, 256, 0, s>>>(t, x, xflags, ntgt, G, ws);
                BT_LAUNCH_CHECK();
                for (int st = 0; st < nsteps; ++st) {
                    list13_heavy_step_kernel<T, DIM, true><<<kNumSMs * 8, 256, 0, s>>>(t, x, ntgt, st, G, ws);
                    BT_LAUNCH_CHECK();
                }
            }
            BT_TRY(heavy_sort_and_scatter_groups(ws, heavy_total_host, t.nboxes, nheavy_host, nrows, rowlen, G,
                                                 lists, s));
        }
    }
    return BT_OK;
}

// global DFS pre-order rank of every box (children in Morton order): the append order of
// every reference walk is this order restricted to the appended boxes
template <int DIM>
__global__ void subtree_size_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                    const int* __restrict__ child_ids, int* __restrict__ size)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        int sz = 1;
#pragma unroll
        for (int m = 0; m < NB; ++m) { const int c = child_ids[m * aligned + b]; if (c) sz += size[c]; }
        size[b] = sz;
    }
}
template <int DIM>
__global__ void dfs_rank_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                const int* __restrict__ child_ids, const int* __restrict__ size,
                                int* __restrict__ rank)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        int r = (b == 0 ? 0 : rank[b]) + 1;
        if (b == 0) rank[0] = 0;
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const int c = child_ids[m * aligned + b];
            if (c) { rank[c] = r; r += size[c]; }
        }
    }
}

// per-level compressed CSR (eliminate_empty_output_lists) -- one kernel for all levels
__global__ void __launch_bounds__(256)
list3_compress_kernel(int nlevels, int ntgt, const int* __restrict__ G, const int* __restrict__ C,
                      const int* __restrict__ target_boxes, int* __restrict__ cstarts,
                      int* __restrict__ nonempty_indices, int* __restrict__ tb_nonempty,
                      int* __restrict__ compressed_indices /*[nlevels][ntgt+1]*/,
                      int* __restrict__ close_starts /*[ntgt+1] or null*/,
                      int* __restrict__ list1_starts /*[ntgt+1] or null (G has a list-1 row)*/)
{
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int64_t total = rowlen * (nlevels + (list1_starts ? 2 : 1));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int l = (int)(i / rowlen), tt = (int)(i % rowlen);
        const int g0 = G[(int64_t)l * rowlen];
        const int local_start = G[i] - g0;
        if (l == nlevels) { if (close_starts) close_starts[tt] = local_start; continue; }
        if (l == nlevels + 1) { list1_starts[tt] = local_start; continue; }
        const int c0 = C[(int64_t)l * rowlen];
        const int ci = C[i] - c0;
        compressed_indices[i] = ci;
        const int soff = c0 + l;             // sum over previous levels of (nonempty + 1)
        if (tt < ntgt) {
            if (G[i + 1] - G[i] > 0) {
                nonempty_indices[c0 + ci] = tt;
                tb_nonempty[c0 + ci] = target_boxes[tt];
                cstarts[soff + ci] = local_start;
            }
        } else cstarts[soff + ci] = local_start;   // total of the level
    }
}

// ---- box lists, level starts -------------------------------------------------
struct BoxListIn {
    const unsigned char* flags; const signed char* mask; int which;
    __device__ int operator()(int64_t b) const
    {
        const unsigned char fl = flags[b];
        bool k;
        if (which == 0) k = (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES) && (!mask || mask[b]);
        else if (which == 1) k = (fl & BT_BOX_IS_SOURCE_BOX) && (!mask || mask[b]);
        else if (which == 2) k = (fl & (BT_BOX_HAS_TARGET_CHILD_BOXES | BT_BOX_IS_TARGET_BOX));
        else k = (fl & BT_BOX_IS_TARGET_BOX);
        return k ? 1 : 0;
    }
};
struct BoxListOut {
    BoxListIn in; int* out; int* count;
    __device__ void operator()(int64_t b, long long excl) const { if (in(b)) out[excl] = (int)b; }
    __device__ void total(long long t) const { *count = (int)t; }
};

__global__ void level_starts_kernel(int nlevels, const int* __restrict__ level_start_box_nrs,
                                    const int* __restrict__ list, int nlist, int* __restrict__ out)
{
    const int lev = blockIdx.x * blockDim.x + threadIdx.x;
    if (lev > nlevels) return;
    if (lev == nlevels) { out[lev] = nlist; return; }
    const int key = level_start_box_nrs[lev];
    int lo = 0, hi = nlist;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] < key) lo = mid + 1; else hi = mid; }
    out[lev] = lo;
}

// ---- list merger: traversal.py:1153-1214 ------------------------------------
struct MergeArgs { const int* starts[3]; const int* lists[3]; int nlists; };

struct MergeIn {
    MergeArgs m; const int* o2i;
    __device__ int operator()(int64_t i) const
    {
        const int ibox = o2i ? o2i[i] : (int)i;
        int tot = 0;
        for (int l = 0; l < m.nlists; ++l) tot += m.starts[l][ibox + 1] - m.starts[l][ibox];
        return tot;
    }
};

__global__ void __launch_bounds__(256)
merge_write_kernel(MergeArgs m, const int* __restrict__ o2i, int nout, const int* __restrict__ new_starts,
                   int* __restrict__ new_lists)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += stride) {
        const int ibox = o2i ? o2i[i] : i;
        int cur = new_starts[i];
        for (int l = 0; l < m.nlists; ++l) {
            const int s = m.starts[l][ibox], c = m.starts[l][ibox + 1] - s;
            for (int j = 0; j < c; ++j) new_lists[cur++] = m.lists[l][s + j];
        }
    }
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

void bt_set_walk_mode(int mode) { bt::g_walk_mode = mode; }
int bt_get_walk_mode(void) { return bt::g_walk_mode; }

int bt_trav_colleagues(int dtype, int phase, const bt_tree_view* tree, const int32_t* level_start_box_nrs,
                       const int32_t* dfs_rank, const int8_t* row_mask, int stride, int32_t* staging,
                       int32_t* starts, int32_t* lists, int32_t* list2_count_by_box, uint8_t* xflags,
                       uint32_t* list2_masks, int mask_words, int64_t* totals_dev, void* stream)
{
    BT_PROF(phase ? "trav_colleagues_fill" : "trav_colleagues_count", (cudaStream_t)stream);
    if (list2_masks && mask_words < ((stride + 1) * (1 << tree->dim) + 31) / 32 + 1) return BT_ERR_BAD_ARG;
    if (((stride + 1) * (1 << tree->dim) + 31) / 32 + 1 > bt::kCollMaskWordsMax || stride + 1 > 128)
        return BT_ERR_UNSUPPORTED;     // well_sep_is_n_away > 2 in 3-D: use the walk-based builder
    BT_DISPATCH(dtype, tree->dim, colleagues_topdown_impl, phase, tree, level_start_box_nrs, dfs_rank,
                (const signed char*)row_mask, stride, staging, starts, lists, list2_count_by_box, xflags,
                list2_masks, mask_words, (long long*)totals_dev, (cudaStream_t)stream);
}

int bt_trav_list2_fill_masked(int dim, int nrows, const int32_t* row_boxes, const int32_t* box_parent_ids,
                              const int32_t* coll_starts, const int32_t* coll_lists,
                              const int32_t* box_child_ids_t, const uint32_t* list2_masks, int mask_words,
                              const int32_t* starts, int32_t* lists, void* stream)
{
    BT_PROF("trav_list2_fill", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = bt::grid_for((int64_t)nrows * 32, 256, 8);
    if (dim == 1) bt::list2_masked_fill_kernel<1><<<grid, 256, 0, s>>>(nrows, row_boxes, box_parent_ids, coll_starts, coll_lists, box_child_ids_t, list2_masks, mask_words, starts, lists);
    else if (dim == 2) bt::list2_masked_fill_kernel<2><<<grid, 256, 0, s>>>(nrows, row_boxes, box_parent_ids, coll_starts, coll_lists, box_child_ids_t, list2_masks, mask_words, starts, lists);
    else if (dim == 3) bt::list2_masked_fill_kernel<3><<<grid, 256, 0, s>>>(nrows, row_boxes, box_parent_ids, coll_starts, coll_lists, box_child_ids_t, list2_masks, mask_words, starts, lists);
    else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_trav_list2_starts(int nrows, const int32_t* row_boxes, const int32_t* list2_count_by_box,
                         int32_t* starts, int64_t* totals_dev, void* stream)
{
    BT_PROF("trav_list2_count", (cudaStream_t)stream);
    bt::GatherCountIn in{list2_count_by_box, row_boxes};
    bt::PlainOut out{starts, (long long*)totals_dev, nrows};
    return bt::scan_exclusive(nrows, nullptr, in, out, (cudaStream_t)stream);
}

int bt_trav_box_list(int which, int nboxes, const uint8_t* box_flags, const int8_t* mask,
                     int32_t* out_list, int32_t* count_dev, void* stream)
{
    BT_PROF("bt_trav_box_list", (cudaStream_t)stream);
    bt::BoxListIn in{box_flags, (const signed char*)mask, which};
    bt::BoxListOut out{in, out_list, count_dev};
    return bt::scan_exclusive(nboxes, nullptr, in, out, (cudaStream_t)stream);
}

int bt_trav_level_starts(int nlevels, const int32_t* level_start_box_nrs, const int32_t* box_list,
                         int nlist, int32_t* out, void* stream)
{
    BT_PROF("bt_trav_level_starts", (cudaStream_t)stream);
    bt::level_starts_kernel<<<(nlevels + 1 + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        nlevels, level_start_box_nrs, box_list, nlist, out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_trav_build_list(int dtype, int kind, int phase, const bt_tree_view* tree, const bt_list_args* args,
                       int nrows, int32_t* starts, int32_t* lists, int32_t* close_starts,
                       int32_t* close_lists, int64_t* totals_dev, void* stream)
{
    static const char* const kNames[2][6] = {
        {"trav_colleagues_count", "?", "trav_list2_count", "?", "trav_list4_count", "peer_lists_count"},
        {"trav_colleagues_fill", "?", "trav_list2_fill", "?", "trav_list4_fill", "peer_lists_fill"}};
    BT_PROF(kNames[phase ? 1 : 0][(kind >= 0 && kind <= 5) ? kind : 3], (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, build_list_impl, kind, phase, tree, args, nrows, starts, lists,
                close_starts, close_lists, (long long*)totals_dev, (cudaStream_t)stream);
}

int bt_trav_mark_rows(int dtype, const bt_tree_view* tree, const bt_list_args* args, int8_t* point_src_mask,
                      int8_t* multipole_mask, void* stream)
{
    BT_PROF("bt_trav_mark_rows", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, mark_rows_impl, tree, args, (signed char*)point_src_mask,
                (signed char*)multipole_mask, (cudaStream_t)stream);
}

int bt_trav_list1(int dtype, int phase, const bt_tree_view* tree, const int32_t* target_boxes,
                  int ntarget_boxes, int32_t* starts, int32_t* lists, int64_t* totals_dev,
                  const bt_heavy_ws* ws, int64_t heavy_total, void* stream)
{
    BT_PROF(phase ? "trav_list1_fill" : "trav_list1_count", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, list1_impl, phase, tree, target_boxes, ntarget_boxes, starts, lists,
                (long long*)totals_dev, ws, (long long)heavy_total, (cudaStream_t)stream);
}

int bt_trav_list3(int dtype, int phase, const bt_tree_view* tree, const bt_list3_args* args,
                  int ntarget_boxes, int32_t* G, int32_t* C, int32_t* lists, int64_t* summary_dev,
                  const bt_heavy_ws* ws, int64_t heavy_total, void* stream)
{
    BT_PROF(phase ? "trav_list3_fill" : "trav_list3_count", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, list3_impl, phase, tree, args, ntarget_boxes, G, C, lists,
                (long long*)summary_dev, ws, (long long)heavy_total, (cudaStream_t)stream);
}

__global__ void transpose_children_kernel(int nb, int aligned, const int* __restrict__ in, int* __restrict__ out)
{
    const int64_t total = (int64_t)aligned * nb;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        // i enumerates the OUTPUT (coalesced stores); the nb input rows are read at stride `aligned`
        const int64_t b = i / nb; const int m = (int)(i % nb);
        out[i] = in[(int64_t)m * aligned + b];
    }
}

int bt_trav_transpose_children(int dim, int aligned_nboxes, const int32_t* box_child_ids,
                               int32_t* box_child_ids_t, void* stream)
{
    BT_PROF("bt_trav_transpose_children", (cudaStream_t)stream);
    if (aligned_nboxes <= 0) return BT_OK;
    const int nb = 1 << dim;
    transpose_children_kernel<<<bt::grid_for((int64_t)aligned_nboxes * nb, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nb, aligned_nboxes, box_child_ids, box_child_ids_t);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_area_query(int dtype, int phase, const bt_tree_view* tree, const int32_t* peer_list_starts,
                  const int32_t* peer_lists, int nballs, void* const* ball_centers, const void* ball_radii,
                  const double* bbox_min, int32_t* starts, int32_t* lists, int64_t* totals_dev, void* stream)
{
    BT_PROF(phase ? "area_query_fill" : "area_query_count", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, area_query_impl, phase, tree, peer_list_starts, peer_lists, nballs, ball_centers,
                ball_radii, bbox_min, starts, lists, (long long*)totals_dev, (cudaStream_t)stream);
}

int bt_trav_list13(int dtype, int phase, const bt_tree_view* tree, const bt_list3_args* args,
                   const uint8_t* xflags, int ntarget_boxes, int32_t* G, int32_t* C, int32_t* lists,
                   int64_t* summary_dev, const bt_heavy_ws* ws, int64_t heavy_total, int nheavy,
                   int nwalk, void* stream)
{
    BT_PROF(phase == 1 ? "trav_list13_fill" : "trav_list13_count", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, list13_impl, phase, tree, args, xflags, ntarget_boxes, G, C, lists,
                (long long*)summary_dev, ws, (long long)heavy_total, nheavy, nwalk, (cudaStream_t)stream);
}

int bt_trav_dfs_rank(int dim, int nboxes, int aligned_nboxes, int nlevels,
                     const int32_t* level_start_box_nrs, const int32_t* box_child_ids,
                     int32_t* subtree_size, int32_t* dfs_rank, void* stream)
{
    BT_PROF("bt_trav_dfs_rank", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    const int grid = bt::grid_for(nboxes, 256, 4);
    for (int lev = nlevels - 1; lev >= 0; --lev) {
        if (dim == 1) bt::subtree_size_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else if (dim == 2) bt::subtree_size_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else bt::subtree_size_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        BT_LAUNCH_CHECK();
    }
    for (int lev = 0; lev < nlevels; ++lev) {
        if (dim == 1) bt::dfs_rank_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, dfs_rank);
        else if (dim == 2) bt::dfs_rank_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, dfs_rank);
        else bt::dfs_rank_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, dfs_rank);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

int bt_trav_list3_compress(int nlevels, int ntarget_boxes, const int32_t* G, const int32_t* C,
                           const int32_t* target_boxes, int32_t* compressed_starts,
                           int32_t* nonempty_indices, int32_t* target_boxes_nonempty,
                           int32_t* compressed_indices, int32_t* close_starts, int32_t* list1_starts,
                           void* stream)
{
    BT_PROF("bt_trav_list3_compress", (cudaStream_t)stream);
    const int64_t total = ((int64_t)ntarget_boxes + 1) * (nlevels + (list1_starts ? 2 : 1));
    bt::list3_compress_kernel<<<bt::grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nlevels, ntarget_boxes, G, C, target_boxes, compressed_starts, nonempty_indices,
        target_boxes_nonempty, compressed_indices, close_starts, list1_starts);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

__global__ void bt_gather_i32_kernel(int64_t n, const int* __restrict__ src, const int* __restrict__ idx,
                                     int* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

int bt_gather_i32(int64_t n, const int32_t* src, const int32_t* idx, int32_t* out, void* stream)
{
    BT_PROF("bt_gather_i32", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt_gather_i32_kernel<<<bt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, src, idx, out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_trav_merge_lists(int phase, int noutput, const int32_t* output_to_input_box, int nlists,
                        const int32_t* const* starts, const int32_t* const* lists, int32_t* new_starts,
                        int32_t* new_lists, int64_t* totals_dev, void* stream)
{
    BT_PROF("bt_trav_merge_lists", (cudaStream_t)stream);
    if (nlists < 1 || nlists > 3) return BT_ERR_BAD_ARG;
    bt::MergeArgs m;
    m.nlists = nlists;
    for (int l = 0; l < 3; ++l) { m.starts[l] = l < nlists ? starts[l] : nullptr; m.lists[l] = l < nlists ? lists[l] : nullptr; }
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::MergeIn in{m, output_to_input_box};
        bt::PlainOut out{new_starts, (long long*)totals_dev, noutput};
        return bt::scan_exclusive(noutput, nullptr, in, out, s);
    }
    if (noutput > 0) {
        bt::merge_write_kernel<<<bt::grid_for(noutput, 256, 8), 256, 0, s>>>(m, output_to_input_box, noutput,
                                                                             new_starts, new_lists);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

}  // extern "C"
