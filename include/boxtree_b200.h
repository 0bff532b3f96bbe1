/*
 * boxtree_b200.h -- C ABI of libboxtree_b200.so (sm_100a kernels for the
 * TreeBuilder -> Tree -> FMMTraversalBuilder -> FMMTraversalInfo path).
 *
 * The natural FFI seam of the reference is its two records of compiled device
 * kernels: boxtree/tree_build_kernels.py:130-150 (`_KernelInfo`, 13 tree-build
 * kernels + bounding box + gappy copy) and boxtree/traversal.py:1710-1718
 * (`_KernelInfo`, 7 list builders).  Each entry point below names the reference
 * kernel(s) it replaces (paths relative to the reference checkout).
 *
 * Conventions: every pointer is a DEVICE pointer unless stated otherwise; sizes
 * are element counts; `stream` is a cudaStream_t passed as void*; every call
 * only ENQUEUES work on `stream` and returns 0 on success, a cudaError_t value
 * (< 10000) or a BT_ERR_* code otherwise.  No torch types cross this boundary.
 * dtype: BT_F32 / BT_F64 coordinate type; dim: 1, 2 or 3.
 */
#ifndef BOXTREE_B200_H
#define BOXTREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_F32 0
#define BT_F64 1

/* box flag bits -- boxtree/tree.py:133-142 */
#define BT_BOX_IS_SOURCE_BOX 1
#define BT_BOX_IS_TARGET_BOX 2
#define BT_BOX_HAS_SOURCE_CHILD_BOXES 4
#define BT_BOX_HAS_TARGET_CHILD_BOXES 8

/* slots of the int32 device control block `ctl` (BT_CTL_SIZE entries) */
#define BT_CTL_NBOXES 0          /* boxes in the pool */
#define BT_CTL_NSPLIT 1          /* boxes split in this level iteration */
#define BT_CTL_OVERSIZE 2        /* have_oversize_split_box (tree_build.py:641) */
#define BT_CTL_OVERFLOW 3        /* pool capacity too small; grow and call again */
#define BT_CTL_NSPLIT_REGULAR 4  /* splits of boxes on the level above the new one */
#define BT_CTL_COMMITTED 5
#define BT_CTL_NBOXES_FINAL 6
#define BT_CTL_NBIG 7
#define BT_CTL_NHUGE 8
#define BT_CTL_LR_FOUND 16       /* [16 + level]: have_upper_level_split_box per level */
#define BT_CTL_LR_SLOTS 48
#define BT_CTL_SIZE 64

typedef struct {
    const void *sources[3];     /* per-axis coordinate arrays, [nsources] */
    const void *targets[3];     /* per-axis, [ntargets]; NULL when sources are targets */
    const void *source_radii;   /* [nsources] or NULL */
    const void *target_radii;   /* [ntargets] or NULL */
    int64_t nsources;
    int64_t ntargets;           /* 0 when sources are targets */
} bt_particles;

/* creation-order box pool used while the level loop runs */
typedef struct {
    int32_t *start;             /* first particle (sorted order); 0 for empty boxes */
    int32_t *count;             /* cumulative particle count */
    uint8_t *level;
    int32_t *parent;
    int32_t *child0;            /* id of Morton child 0 (children are contiguous); 0 = none */
    uint8_t *has_children;
    uint8_t *force_split;
    int32_t *nonchild;          /* particles that stop in this box (extents) */
    void *center[3];
    int32_t capacity;
    /* Distributed build (boxtree_b200/distributed/tree_build.py): start/count/nonchild above
     * are the box's range in THIS rank's sorted particles; gstart/gcount/gnonchild hold the
     * sums over all ranks (the global tree's values) and xch [2 * 2^dim * nsplit] receives
     * (count, nonchild) of the children created by BT_STEP_CREATE, which the host all-reduces
     * before BT_STEP_COMMIT.  All four NULL on one GPU (global == local). */
    int32_t *gstart;
    int32_t *gcount;
    int32_t *gnonchild;
    int32_t *xch;
} bt_pool;

typedef struct {
    int32_t *box_start;         /* box_srcntgt_starts */
    int32_t *box_count;         /* box_srcntgt_counts_cumul */
    int32_t *box_nonchild;      /* box_srcntgt_counts_nonchild (extents) */
    uint8_t *box_levels;
    int32_t *box_parent_ids;
    int32_t *box_child_ids;     /* [2^dim, aligned_nboxes] */
    void *box_centers;          /* [dim, aligned_nboxes] */
    uint8_t *has_children;
    uint8_t *real_children;
    int32_t *local_start;       /* distributed build only (else NULL): the box's range in */
    int32_t *local_count;       /* this rank's sorted particles                            */
    int32_t *local_nonchild;
} bt_box_out;

/* instrumentation used by bench.py: kernels launched so far; optional CUDA-event
 * timing of every entry point / sort pass on its launching stream */
long long bt_launch_count(void);
void bt_prof_enable(int on);
void bt_prof_reset(void);
int bt_prof_report(char *buf, int len);   /* "name\tcalls\ttotal_ms" lines */
/* Bit mask choosing, per builder, the group/warp-cooperative mapping (bit set) or one thread
 * per row (the reference's mapping): 1 colleagues, 2 list 1, 4 list 3, 8 list 3 only with
 * target extents, 16 list-2 count, 32 list-2 fill; 64 = colleagues built top-down
 * (bt_trav_colleagues) instead of by walks, 128 = lists 1 and 3 from one fused walk
 * (bt_trav_list13), 512 = its heavy rows by radix sort instead of the position map (the host
 * chooses by leaving bt_heavy_ws.hrow_base NULL); 256 is a test switch.  Same output. */
void bt_set_walk_mode(int mode);
int bt_get_walk_mode(void);

/* deepest level ONE 64-bit sort key resolves for `dim`; deeper trees use two-word keys
 * (keys_lo below) up to bt_max_tree_level = 31, the deepest level the reference's
 * `1U << (1 + level)` digit expression (tree_build_kernels.py:374-376) can resolve:
 * beyond it both raise MaxLevelsExceeded */
int bt_max_key_level(int dim);
int bt_max_tree_level(int dim);

/* bounding_box.py:54-122 (BBOX_REDUCTION_TPL).  out_minmax: [2*dim] coords,
 * laid out min_x, max_x, min_y, max_y, ... like make_bounding_box_dtype (:35-52) */
int bt_bounding_box(int dtype, int dim, const bt_particles *p, void *out_minmax, void *stream);

/* Digit computation of morton_scan (tree_build_kernels.py:308-470) for all levels
 * at once.  bbox_min/bbox_max are HOST arrays [dim].  extent_norm: 0 none, 1 linf, 2 l2.
 * depth: levels the key resolves (0 or >= bt_max_key_level(dim): all it can hold); a tree that
 * turns out deeper must be rebuilt with a larger depth.  keys_lo (optional): two-word keys --
 * `keys` then resolves bt_max_key_level(dim) levels and keys_lo the levels above, up to
 * bt_max_tree_level(dim) (depth is ignored).  records (optional, [n][4] coords): x, y, z,
 * radius of every particle side by side for bt_permute. */
int bt_make_keys(int dtype, int dim, const bt_particles *p, const double *bbox_min,
                 const double *bbox_max, int extent_norm, double stick_out_factor, int depth,
                 uint64_t *keys, uint64_t *keys_lo, void *records, void *stream);

/* Stable radix sort of (key, particle id); replaces the per-level partition
 * morton_scan + renumber_particles (tree_build_kernels.py:247-508, 717-819).
 * ids are generated (identity) by the first pass.  *result_in_alt (HOST) tells
 * which buffer pair holds the result. */
int bt_sort_particles(int64_t n, int dim, int have_extent, int depth, uint64_t *keys,
                      uint64_t *keys_alt, uint32_t *ids, uint32_t *ids_alt, int *result_in_alt,
                      void *stream);

/* the same for two-word keys: on return keys, keys_lo and ids are in sorted order (the *_tmp
 * arrays are scratch of the same sizes) */
int bt_sort_particles_deep(int64_t n, int dim, int have_extent, uint64_t *keys, uint64_t *keys_tmp,
                           uint64_t *keys_lo, uint64_t *keys_lo_tmp, uint32_t *ids,
                           uint32_t *ids_tmp, void *stream);

/* refine weights in sorted order -> exclusive int64 prefix, wprefix[n] = total
 * (the pwt fields of the morton_scan struct, tree_build_kernels.py:462-466) */
int bt_weight_prefix(int64_t n, const uint32_t *sorted_ids, const int32_t *weights,
                     int64_t *wprefix, void *stream);

/* root box and control block (tree_build.py:586-618, 656-668).  root_center: HOST [dim] */
int bt_pool_init(int dtype, int dim, const bt_pool *pool, int64_t n, int have_extent,
                 const uint64_t *keys, const double *root_center, int32_t *ctl, void *stream);

/* One iteration of the level loop (tree_build.py:698-1121) on per-box data:
 * split_box_id_scan (tree_build_kernels.py:514-640) + box_splitter (:646-711).
 * Boxes [lo, nboxes) of the pool are examined.  phases: BT_STEP_DECIDE (split decisions),
 * BT_STEP_CREATE (children; repeat alone after the caller enlarged the pool on
 * BT_CTL_OVERFLOW), BT_STEP_COMMIT (children become part of the pool; in a distributed build
 * after the host all-reduced pool->xch).  One GPU: BT_STEP_ALL. */
#define BT_STEP_DECIDE 1
#define BT_STEP_CREATE 2
#define BT_STEP_COMMIT 4
#define BT_STEP_ALL 7
int bt_level_step(int dtype, int dim, const bt_pool *pool, const uint64_t *keys,
                  const int64_t *wprefix, int32_t *ctl, int32_t *split_list, uint8_t *flag, int lo,
                  int nboxes, int level, int maxw, int adaptive, int level_restrict,
                  int have_extent, int skip_if_no_regular, double root_extent, int phases,
                  int depth, const uint64_t *keys_lo, void *stream);

/* level_restrict kernel + upper-level sweep (tree_build_kernels.py:825-915,
 * tree_build.py:1145-1200); the sweep's early exit is a device-side flag chain. */
int bt_level_restrict(int dtype, int dim, const bt_pool *pool, int32_t *ctl, int built_level,
                      int nboxes_upper, double root_extent, void *stream);

/* prune / renumber: find_prune_indices_scan, find_level_box_counts_scan
 * (tree_build_kernels.py:1697-1742) and the gappy copies of tree_build.py:1389-1431 */
int bt_finalize_numbering(int dtype, int dim, const bt_pool *pool, int nboxes, int level_restrict,
                          int skip_prune, int32_t *ctl, int32_t *map_old2new, int32_t *src_of_new,
                          int32_t *level_start, void *stream);
int bt_gather_boxes(int dtype, int dim, const bt_pool *pool, int have_extent,
                    const int32_t *src_of_new, const int32_t *map_old2new, int nfinal, int aligned,
                    const bt_box_out *out, void *stream);

/* restores ascending-user-id order inside never-partitioned boxes (the order the
 * reference's stable partition leaves, tree_build_kernels.py:766-798) */
int bt_leaf_fixup(int nboxes, const int32_t *box_start, const int32_t *box_count,
                  const uint8_t *real_children, uint32_t *ids, int32_t *ctl, int32_t *big_list,
                  int big_cap, int32_t *huge_list, void *stream);
int bt_sort_u32_segment(int64_t n, uint32_t *ids, void *stream);

/* source_counter + find_source_and_target_indices, particle part
 * (tree_build_kernels.py:1770-1782, 1151-1161); source_numbers is [n+1] */
int bt_split_sources_targets(int64_t n, int64_t nsources, const uint32_t *sorted_ids,
                             int32_t *source_numbers, int32_t *user_source_ids,
                             int32_t *srcntgt_target_ids, int32_t *sorted_target_ids, void *stream);
/* tools.py:81-109 (reverse_index_array) */
int bt_reverse_index(int64_t n, const uint32_t *ids, int32_t *out, void *stream);

/* srcntgt_permuter + cl_array.take (tree_build_kernels.py:1170-1186, tree_build.py:1609-1616).
 * outs: HOST array of dim device pointers; records: the array bt_make_keys wrote, or NULL
 * (then the coordinates are gathered from *p) */
int bt_permute(int dtype, int dim, const bt_particles *p, const void *records,
               const int32_t *from_ids, int64_t n, void *const *outs, void *out_radii, void *stream);

/* find_source_and_target_indices box part + box_info (tree_build_kernels.py:1062-1147, 1192-1305) */
int bt_box_info(int nboxes, int sources_are_targets, int have_extent, const int32_t *box_start,
                const int32_t *box_count, const int32_t *box_nonchild, const uint8_t *has_children,
                const int32_t *source_numbers, int32_t *src_starts, int32_t *src_nonchild,
                int32_t *src_cumul, int32_t *tgt_starts, int32_t *tgt_nonchild, int32_t *tgt_cumul,
                uint8_t *box_flags, void *stream);

/* find_box_extents for all levels (tree_build_kernels.py:1311-1399, launched per level
 * bottom-up at tree_build.py:1751-1802): own-particle min/max for every box in one launch,
 * then one child-merge launch per level.  particles: HOST array of dim device pointers
 * (tree-ordered coordinates); level_start_box_nrs_host: HOST [nlevels+1] */
int bt_box_extents(int dtype, int dim, int nboxes, int aligned, int nlevels,
                   const int32_t *level_start_box_nrs_host, const int32_t *box_child_ids,
                   const void *box_centers, const int32_t *pstarts, const int32_t *pcounts,
                   void *const *particles, const void *radii, void *bb_min, void *bb_max,
                   void *stream);

/* Distributed build (no kernel counterpart in the reference, whose tree is built on one rank and
 * broadcast, distributed/__init__.py:185-203): bt_box_extents split into its two phases
 * (phases & 1: own-particle min/max per box -- with phases & 4 one lane per box, for boxes
 * that hold a particle or two --, phases & 2: child merge per level) so that the
 * host can all-reduce min/max between them; bt_box_info split into the per-rank source counts
 * src3 = [3][nboxes] (sources before / in / stopping in the box's range of THIS rank's
 * particles, plus the rank-local range arrays) and the global ranges + flags from the global
 * srcntgt ranges and the all-reduced src3. */
int bt_box_extents_phase(int dtype, int dim, int nboxes, int aligned, int nlevels,
                         const int32_t *level_start_box_nrs_host, const int32_t *box_child_ids,
                         const void *box_centers, const int32_t *pstarts, const int32_t *pcounts,
                         void *const *particles, const void *radii, void *bb_min, void *bb_max,
                         int phases, void *stream);
int bt_box_info_local(int nboxes, int have_extent, const int32_t *local_start,
                      const int32_t *local_count, const int32_t *local_nonchild,
                      const uint8_t *has_children, const int32_t *source_numbers, int32_t *src3,
                      int32_t *src_starts, int32_t *src_nonchild, int32_t *src_cumul,
                      int32_t *tgt_starts, int32_t *tgt_nonchild, int32_t *tgt_cumul, void *stream);
int bt_box_info_global(int nboxes, int have_extent, const int32_t *box_start,
                       const int32_t *box_count, const int32_t *box_nonchild,
                       const uint8_t *has_children, const int32_t *src3, int32_t *src_starts,
                       int32_t *src_nonchild, int32_t *src_cumul, int32_t *tgt_starts,
                       int32_t *tgt_nonchild, int32_t *tgt_cumul, uint8_t *box_flags, void *stream);

/* ---------------------------------------------------------------- traversal */

typedef struct {
    int32_t dim;
    int32_t nboxes;
    int32_t aligned_nboxes;
    int32_t nlevels;
    double root_extent;
    const void *box_centers;            /* [dim, aligned] */
    const uint8_t *box_levels;
    const int32_t *box_child_ids;       /* [2^dim, aligned] */
    const uint8_t *box_flags;
    const int32_t *box_parent_ids;
    int32_t well_sep_is_n_away;
    const int32_t *box_child_ids_t;     /* optional scratch copy [aligned, 2^dim] (bt_trav_transpose_children):
                                           the 2^dim children of a box share one 32-byte sector */
} bt_tree_view;

/* box_child_ids [2^dim, aligned] -> [aligned, 2^dim] */
int bt_trav_transpose_children(int dim, int aligned_nboxes, const int32_t *box_child_ids,
                               int32_t *box_child_ids_t, void *stream);

/* sources_parents_and_targets (traversal.py:326-355): which = 0 source_parent_boxes,
 * 1 source_boxes, 2 target_or_target_parent_boxes, 3 target_boxes.  Writes the
 * compacted ascending box list and its length (*count_dev, device int32). */
int bt_trav_box_list(int which, int nboxes, const uint8_t *box_flags, const int8_t *mask,
                     int32_t *out_list, int32_t *count_dev, void *stream);

/* extract_level_start_box_nrs + host fix-up (traversal.py:361-392, 2073-2098) */
int bt_trav_level_starts(int nlevels, const int32_t *level_start_box_nrs, const int32_t *box_list,
                         int nlist, int32_t *out /*[nlevels+1]*/, void *stream);

/* List builders (pyopencl ListOfListsBuilder: count, scan, write).
 * kind: 0 same_level_non_well_sep_boxes (traversal.py:398-464)
 *       2 from_sep_siblings / list 2     (:556-601)
 *       4 from_sep_bigger / list 4 (+close) (:931-1146)
 *       5 peer lists of all boxes (area_query.py:393-475, PeerListFinder)
 * phase 0 writes per-row counts then turns them into starts[nrows+1] in place and
 * stores the total at totals_dev[0] (and [1] for the close list), int64; phase 1 fills. */
typedef struct {
    const int32_t *row_boxes;           /* target_boxes / target_or_target_parent_boxes; NULL = all boxes */
    const int32_t *coll_starts;         /* same_level_non_well_sep_boxes */
    const int32_t *coll_lists;
    double stick_out_factor;
    int32_t with_extent;
    const int8_t *row_mask;             /* optional [nboxes]: rows of boxes with mask 0 stay empty */
} bt_list_args;

int bt_trav_build_list(int dtype, int kind, int phase, const bt_tree_view *tree,
                       const bt_list_args *args, int nrows, int32_t *starts, int32_t *lists,
                       int32_t *close_starts, int32_t *close_lists, int64_t *totals_dev,
                       void *stream);

/* distributed build (no reference counterpart): rows of lists 2, 4 and 4-close
 * (traversal.py:556-601, 931-1146) of every box whose tree->box_flags carry a target bit,
 * marked straight into the masks of partition.py:197-297 -- list 2 into multipole_mask,
 * lists 4 / 4-close into point_src_mask -- without building the lists.  args: colleague CSR,
 * stick_out_factor, with_extent. */
int bt_trav_mark_rows(int dtype, const bt_tree_view *tree, const bt_list_args *args,
                      int8_t *point_src_mask, int8_t *multipole_mask, void *stream);

/* same_level_non_well_sep_boxes built top-down instead of by the reference's walk from the
 * root (traversal.py:398-464; same set, same depth-first order): the colleagues of b are the
 * adjacent children of parent(b) and of parent(b)'s colleagues.  One launch per level.
 * phase 0: rows staged in `staging` [nboxes * stride], stride = (2n+1)^d - 1; starts[nboxes+1]
 *          (total at totals_dev[0]); list2_count_by_box[nboxes] = number of from_sep_siblings
 *          entries of every box (traversal.py:556-601, the non-adjacent candidates);
 *          xflags[nboxes]: bit 0 = the box or one of its colleagues is a source box,
 *          bit 1 = the box has a child; list2_masks (optional) [nboxes, mask_words],
 *          mask_words >= ceil((stride + 1) * 2^d / 32) + 1: bit k of a row = candidate k
 *          (k / 2^d-th box of the parent's colleagues with the parent merged in at its
 *          depth-first position, Morton child k % 2^d) belongs to list 2; the row's last word is
 *          the parent's position.
 * phase 1: staging -> lists.   dfs_rank from bt_trav_dfs_rank. */
int bt_trav_colleagues(int dtype, int phase, const bt_tree_view *tree,
                       const int32_t *level_start_box_nrs, const int32_t *dfs_rank,
                       const int8_t *row_mask, int stride, int32_t *staging, int32_t *starts,
                       int32_t *lists, int32_t *list2_count_by_box, uint8_t *xflags,
                       uint32_t *list2_masks, int mask_words, int64_t *totals_dev, void *stream);
/* from_sep_siblings lists from the masks of bt_trav_colleagues (no geometry is re-evaluated) */
int bt_trav_list2_fill_masked(int dim, int nrows, const int32_t *row_boxes,
                              const int32_t *box_parent_ids, const int32_t *coll_starts,
                              const int32_t *coll_lists, const int32_t *box_child_ids_t,
                              const uint32_t *list2_masks, int mask_words, const int32_t *starts,
                              int32_t *lists, void *stream);
/* starts[nrows+1] of from_sep_siblings from the per-box counts of bt_trav_colleagues */
int bt_trav_list2_starts(int nrows, const int32_t *row_boxes, const int32_t *list2_count_by_box,
                         int32_t *starts, int64_t *totals_dev, void *stream);

/* Workspace of the "heavy row" path of lists 1 and 3.  A row whose walk needs more than
 * walk_budget child visits (an upper-level box holding its own targets can have ~1e6 list
 * entries) is expanded by a grid-wide breadth-first pass over the same child visits; its
 * entries are then ordered by the tree's global DFS pre-order rank (= the append order of
 * the reference's walk) with the radix sort and copied into the CSR arrays. */
#define BT_HCTL_SIZE 64
#define BT_HCTL_NHEAVY 0
#define BT_HCTL_OVERFLOW 1
#define BT_HCTL_NWALK 3
typedef struct {
    int32_t walk_budget;
    uint8_t *row_heavy;          /* [nrows] */
    int32_t *heavy_rows;         /* [nrows] */
    int32_t *hctl;               /* [BT_HCTL_SIZE] */
    int64_t *heavy_total;        /* [1]: entries of all heavy rows (count phase output) */
    uint64_t *frontier[2];       /* [frontier_cap] each, frontier_cap >= nrows */
    int64_t frontier_cap;
    const int32_t *dfs_rank;     /* [nboxes], from bt_trav_dfs_rank */
    uint64_t *ekeys[2];          /* fill phase: [heavy_total] each */
    uint32_t *evals[2];
    int64_t ecap;
    const int8_t *row_mask;      /* optional [nboxes]: rows of boxes with mask 0 stay empty */
    /* bt_trav_list13 only: the count pass also stages every light row's entries (tagged with
     * their slot, in append order) so that the fill pass copies instead of walking again.
     * stage [nrows * stage_cap] (0 = off; needs nboxes < 2^27, nlevels <= 29), stage_count
     * [nrows].  Rows with more than stage_cap entries are walked again (hctl[3] counts them). */
    uint32_t *stage;
    int32_t stage_cap;
    int32_t *stage_count;
    const int32_t *dfs_order;    /* bt_trav_list13: box id of every depth-first rank (inverse of
                                    dfs_rank); its heavy rows sort keys only, evals stay unused */
    /* bt_trav_list13, heavy rows by position map (hrow_base != NULL; no sort): a heavy row owns
     * a byte map over the concatenated depth-first rank ranges of its roots' subtrees; the
     * breadth-first expansion (phase 2, run once) stores slot + 1 at the position of every
     * appended box, an ordered pass over the map (phase 1) writes the lists.
     * phase 0 fills hrow_base [ntarget_boxes + 1] (map offset of every heavy row, multiples of
     * 1024) and hplan[0] = total map bytes; the host reads hctl[0] (heavy rows) and hplan[0],
     * allocates hseg_* [nheavy * seg_stride] (seg_stride >= (2n+1)^d + 136), hseg_n [nheavy],
     * hmap [hmap_cap = hplan[0]], chunk_cnt [hmap_cap / 1024 * (nlevels + 2)], calls phase 2. */
    const int32_t *subtree_size; /* from bt_trav_dfs_rank */
    int64_t *hrow_base;
    int64_t *hplan;              /* [2] */
    int32_t seg_stride;
    int32_t *hseg_rank;
    int32_t *hseg_prefix;
    uint8_t *hseg_kind;
    int32_t *hseg_n;
    uint8_t *hmap;
    int64_t hmap_cap;
    int32_t *chunk_cnt;
    void *hctx;                  /* [nheavy * 128 bytes]: per-row walk constants */
} bt_heavy_ws;

/* pre-order (depth first, children in Morton order) rank of every box */
int bt_trav_dfs_rank(int dim, int nboxes, int aligned_nboxes, int nlevels,
                     const int32_t *level_start_box_nrs, const int32_t *box_child_ids,
                     int32_t *subtree_size, int32_t *dfs_rank, void *stream);

/* neighbor_source_boxes / list 1 (traversal.py:470-550).  phase 0: starts[ntarget_boxes+1],
 * total at totals_dev[0], ws->heavy_total; phase 1 (heavy_total = HOST copy): lists. */
int bt_trav_list1(int dtype, int phase, const bt_tree_view *tree, const int32_t *target_boxes,
                  int ntarget_boxes, int32_t *starts, int32_t *lists, int64_t *totals_dev,
                  const bt_heavy_ws *ws, int64_t heavy_total, void *stream);

/* from_sep_smaller for ALL source levels in one walk (+ list 3 close), traversal.py:607-875.
 * G, C: int32 [nlevels + 1, ntarget_boxes + 1] (+1 trailing entry); row l < nlevels is
 * source level l, row nlevels is the close list.
 * phase 0: per-(level,row) counts, then ONE flattened exclusive scan in place (G holds
 *          global offsets into the concatenated lists), C = flattened exclusive scan of
 *          the non-empty flags; summary_dev (int64 [2*(nlevels+2)]) = G[l][0], C[l][0].
 * phase 1: fill `lists` (all levels concatenated, level l at offset G[l][0]). */
typedef struct {
    const int32_t *target_boxes;
    const int32_t *coll_starts;
    const int32_t *coll_lists;
    double stick_out_factor;
    int32_t targets_have_extent;
    int32_t sources_have_extent;
    int32_t crit;                       /* 0 static_linf, 1 precise_linf, 2 static_l2 */
    const void *box_target_bounding_box_min;
    const void *box_target_bounding_box_max;
    const int32_t *box_source_counts_cumul;
    int32_t min_nsources_cumul;
} bt_list3_args;

int bt_trav_list3(int dtype, int phase, const bt_tree_view *tree, const bt_list3_args *args,
                  int ntarget_boxes, int32_t *G, int32_t *C, int32_t *lists, int64_t *summary_dev,
                  const bt_heavy_ws *ws, int64_t heavy_total, void *stream);

/* eliminate_empty_output_lists bookkeeping for all levels in one launch: compressed
 * starts (level l at offset C[l][0] + l), nonempty_indices and
 * target_boxes[nonempty_indices] (level l at offset C[l][0]; traversal.py:2211-2215),
 * compressed_indices [nlevels, ntarget_boxes + 1] and the close list's starts (and list 1's
 * when G carries the extra row of bt_trav_list13; NULL otherwise). */
int bt_trav_list3_compress(int nlevels, int ntarget_boxes, const int32_t *G, const int32_t *C,
                           const int32_t *target_boxes, int32_t *compressed_starts,
                           int32_t *nonempty_indices, int32_t *target_boxes_nonempty,
                           int32_t *compressed_indices, int32_t *close_starts,
                           int32_t *list1_starts, void *stream);

/* Lists 1 and 3 (+ list 3 close) from ONE walk per target box: below the target's level the
 * reference's list-1 walk (traversal.py:470-550) and list-3 walk (:607-875) visit the same
 * boxes; the near-field boxes above that level come from the colleagues of the ancestors
 * and are merged in by depth-first rank.  Needs bt_trav_colleagues' lists and xflags.
 * G, C: int32 [nlevels + 2, ntarget_boxes + 1] (+1): rows as in bt_trav_list3, row nlevels+1
 * is list 1.  summary_dev: int64 [2*(nlevels+3) + 1] = G[l][0], C[l][0], then the grand
 * total in 64 bits.  Phases: 0 = walk (+ counts and scans, or with the position map of
 * bt_heavy_ws only the map sizes), 2 = position-map mode: expansion of the heavy rows +
 * counts and scans, 1 = fill.  bt_trav_list3_compress (list1_starts != NULL) yields the list-1 starts;
 * its lists are lists[G[nlevels+1][0] .. G[nlevels+2][0]). */
int bt_trav_list13(int dtype, int phase, const bt_tree_view *tree, const bt_list3_args *args,
                   const uint8_t *xflags, int ntarget_boxes, int32_t *G, int32_t *C,
                   int32_t *lists, int64_t *summary_dev, const bt_heavy_ws *ws,
                   int64_t heavy_total, int nheavy, int nwalk /* HOST copies, phase 1 */,
                   void *stream);

/* AreaQueryBuilder (area_query.py:168-392, 657-807): per l^inf ball (centre, radius) the leaf
 * boxes overlapping it, found from the guiding box's peer lists (bt_trav_build_list kind 5).
 * ball_centers: HOST array of dim device pointers; bbox_min: HOST [dim] (tree.bounding_box[0]).
 * phase 0: starts[nballs+1], total at totals_dev[0]; phase 1: lists. */
int bt_area_query(int dtype, int phase, const bt_tree_view *tree, const int32_t *peer_list_starts,
                  const int32_t *peer_lists, int nballs, void *const *ball_centers,
                  const void *ball_radii, const double *bbox_min, int32_t *starts, int32_t *lists,
                  int64_t *totals_dev, void *stream);

/* _ListMerger (traversal.py:1153-1344): phase 0 -> new_starts[noutput+1], total at
 * totals_dev[0]; phase 1 -> new_lists.  starts/lists: HOST arrays of nlists device pointers. */
int bt_trav_merge_lists(int phase, int noutput, const int32_t *output_to_input_box, int nlists,
                        const int32_t *const *starts, const int32_t *const *lists,
                        int32_t *new_starts, int32_t *new_lists, int64_t *totals_dev, void *stream);

/* out[i] = src[idx[i]]  (the take / fancy-index glue of traversal.py:1298-1302) */
int bt_gather_i32(int64_t n, const int32_t *src, const int32_t *idx, int32_t *out, void *stream);

/* ------------------------------------------------- distributed setup (rows c1-c4) */

/* get_box_ids_dfs_order (distributed/partition.py:38-57): pre-order with the HIGHEST
 * Morton child first (the reference pops an explicit stack).  dfs_order[i] = box id */
int bt_dist_dfs_order(int dim, int nboxes, int aligned_nboxes, int nlevels,
                      const int32_t *level_start_box_nrs, const int32_t *box_child_ids,
                      int32_t *subtree_size, int32_t *rank_tmp, int32_t *dfs_order, void *stream);

/* The cut positions of boxtree/distributed/partition.py:81-116 for the default cost
 * 1 + own sources + own targets of a box: cuts[k-1], k = 1..nranks-1, is the first depth-first
 * position whose running cost exceeds k * total / nranks (float64 expression of the reference;
 * the integer running sums are exact); cuts[nranks-1] receives the total cost.
 * cuts: int64[nranks], device. */
int bt_dist_partition_cuts(int nboxes, int nranks, const int32_t *dfs_order,
                           const int32_t *box_source_counts_nonchild,
                           const int32_t *box_target_counts_nonchild, int64_t *cuts, void *stream);

/* get_box_masks (distributed/partition.py:124-357) building blocks: int8 masks [nboxes] */
int bt_dist_mask_from_list(int n, const int32_t *list, int8_t *mask, void *stream);
int bt_dist_ancestor_mask(int nboxes, const int8_t *responsible, const int32_t *box_parent_ids,
                          int8_t *ancestor_mask, void *stream);
/* add_interaction_list_boxes (:135-162): rows whose box is in mask_a | mask_b mark their entries */
int bt_dist_add_list_boxes(int nrows, const int32_t *box_list, const int8_t *mask_a,
                           const int8_t *mask_b, const int32_t *starts, const int32_t *lists,
                           int8_t *out_mask, void *stream);

/* construct_local_particles_and_lists (distributed/local_tree.py:70-151, 198-284) */
int bt_dist_particle_mask(int nboxes, const int8_t *box_mask, const int32_t *starts,
                          const int32_t *counts_nonchild, int32_t *particle_mask, void *stream);
int bt_dist_mask_scan(int64_t n, const int32_t *particle_mask, int32_t *global_to_local,
                      void *stream);
int bt_dist_fetch_local_particles(int dtype, int dim, int64_t n, const int32_t *particle_mask,
                                  const int32_t *global_to_local, void *const *particles,
                                  const void *radii, void *const *local_particles,
                                  void *local_radii, int64_t *particle_idx, void *stream);
int bt_dist_local_lists(int nboxes, const int8_t *box_mask, const int32_t *global_to_local,
                        const int32_t *starts, const int32_t *counts_nonchild,
                        const int32_t *counts_cumul, int32_t *local_starts,
                        int32_t *local_nonchild, int32_t *local_cumul, void *stream);
/* modify_target_flags (local_tree.py:163-185) */
int bt_dist_modify_target_flags(int nboxes, const int32_t *tgt_nonchild, const int32_t *tgt_cumul,
                                uint8_t *box_flags, void *stream);
/* sharded setup (no reference counterpart): box flags whose target bits survive only on
 * boxes of mask_a | mask_b, so that a traversal of the result has exactly the rows
 * get_box_masks reads; need_mask = mask_a | mask_b */
int bt_dist_restrict_target_flags(int nboxes, const uint8_t *box_flags, const int8_t *mask_a,
                                  const int8_t *mask_b, uint8_t *out_flags, int8_t *need_mask,
                                  void *stream);
/* add_interaction_list_boxes (partition.py:135-162) for a list ALL of whose rows qualify:
 * out_mask[lists[k]] = 1 for every entry */
int bt_dist_mark_list_boxes(int64_t nentries, const int32_t *lists, int8_t *out_mask, void *stream);
/* distributed build (no reference counterpart): out_flags keeps the source bits of the global
 * flags and, on boxes of mask_a | mask_b, the target bits that the rank's local flags lack --
 * the rows of the global traversal that get_box_masks reads (partition.py:197-297) but the
 * local traversal does not contain; *any_out != 0 if there is such a box */
int bt_dist_corner_flags(int nboxes, const uint8_t *global_flags, const uint8_t *local_flags,
                         const int8_t *mask_a, const int8_t *mask_b, uint8_t *out_flags,
                         int32_t *any_out, void *stream);
/* MaskCompressorKernel 2-D (tools.py:647-740) on the gathered multipole masks
 * [nranks, nboxes]: phase 0 starts[nboxes+1] + total, phase 1 lists (ascending ranks) */
int bt_dist_box_to_user_rank(int phase, int nboxes, int nranks, const int8_t *masks_all_ranks,
                             int32_t *starts, int32_t *lists, int64_t *total_dev, void *stream);


/* ---- particle exchange of the distributed build: the all-to-all that replaces the root's
 * fetch_local_particles + scatter (local_tree.py:124-151, 408-495).  Every rank holds the
 * global box arrays and its own input particles in tree order.
 * bt_dist_mask_bits: masks [nranks, nboxes] (entries tested with `& bitsel`) -> one bit per
 *   rank and box.
 * bt_dist_pack_count: box id of every local particle (particle_box [n], from the boxes' own
 *   ranges local_start / local_own) and a multisplit count: tile_counts / tile_offsets
 *   [nranks * bt_dist_pack_ntiles(n)] scratch, dest_offsets [nranks+1] (device) = first record
 *   of every destination's chunk.
 * bt_dist_pack_records: one record
 *   [coords (dim) | radius (if radii) | box id i32 | index in the box's own range i32]
 *   per local particle and destination d whose bit is set in dest_bits[box], in tree order
 *   inside the chunk of d.
 * bt_dist_compact_index: compact[b] = number of boxes of the mask below b; *nmasked_dev = total.
 * bt_dist_unpack_records: records of sender s occupy [chunk_offsets_host[s],
 *   chunk_offsets_host[s+1]) of recvbuf (HOST [nranks+1]); a record of box b lands at
 *   dst_start[b] + (records of b from lower senders) + index; particle_idx gets its position
 *   in the global tree order (box_global_start[b] + the same offset).  count_tmp: device
 *   scratch [nranks * nmasked].
 * bt_dist_local_ranges: per-box ranges of the local particle arrays (local_tree.py:249-284)
 *   from the masked own counts in box pre-order (own particles precede the children's). */
int bt_dist_mask_bits(int nboxes, int nranks, int bitsel, const int8_t *masks_all_ranks,
                      uint32_t *dest_bits, void *stream);
int bt_dist_pack_ntiles(int64_t n);
int bt_dist_pack_count(int nranks, int nboxes, int64_t n, const uint32_t *dest_bits,
                       const int32_t *local_start, const int32_t *local_own, int32_t *particle_box,
                       int32_t *tile_counts, int64_t *tile_offsets, int64_t *dest_offsets,
                       void *stream);
int bt_dist_pack_records(int dtype, int nranks, int dim, int64_t n, const int32_t *particle_box,
                         const uint32_t *dest_bits, const int64_t *tile_offsets,
                         void *const *particles, const void *radii, const int32_t *local_start,
                         void *sendbuf, void *stream);
int bt_dist_compact_index(int nboxes, const int8_t *box_mask, int32_t *compact,
                          int32_t *nmasked_dev, void *stream);
int bt_dist_unpack_records(int dtype, int nranks, int dim, int64_t nrec, int has_radii,
                           const void *recvbuf, const int64_t *chunk_offsets_host,
                           const int32_t *compact, int nmasked, int32_t *count_tmp,
                           const int32_t *dst_start, const int32_t *box_global_start,
                           void *const *local_particles, void *local_radii, int64_t *particle_idx,
                           void *stream);
int bt_dist_local_ranges(int nboxes, const int8_t *box_mask, const int32_t *own_counts,
                         const int32_t *preorder_rank, const int32_t *preorder_boxes,
                         const int32_t *subtree_size, int32_t *prefix_tmp, int32_t *local_starts,
                         int32_t *local_nonchild, int32_t *local_cumul, void *stream);
/* bt_dist_box_to_user_rank on masks whose entries are bit fields (tested with `& bitsel`) */
int bt_dist_box_to_user_rank_bits(int phase, int nboxes, int nranks, int bitsel,
                                  const int8_t *masks_all_ranks, int32_t *starts, int32_t *lists,
                                  int64_t *total_dev, void *stream);

/* ------------------------------------------- consumers of Tree / FMMTraversalInfo (rows N1, N2, N4)
 * bt_csr_row_sums: out[out_index ? out_index[i] : i] (+)= scale * sum of values[lists[k]] over row i
 *   -- every translation of the constant-one FMM (boxtree/constant_one.py:104-237: p2p, m2l, m2p,
 *   p2l are additions over an interaction list) and the per-box segmented sums of the cost model
 *   (boxtree/cost.py:445-525, 715-1262).  value_kind: 0 int64, 1 float64.
 * bt_range_sums_i64: out[b] = sum of values[starts[b] .. + counts[b]) (form_multipoles,
 *   constant_one.py:104-116); bt_add_to_ranges_i64: pot[j] += vals[i] over the own range of
 *   boxes[i] (the eval_* steps); bt_fmm_upward_i64 / bt_fmm_downward_i64: one level of
 *   coarsen_multipoles / refine_locals (constant_one.py:118-152, 208-224).
 * bt_gather_i64 / bt_gather_coords: out[i] = src[idx[i]] (reorder_sources / reorder_potentials,
 *   fmm.py:370-374, 525-528; cl_array.take); bt_widen_i32: int32 -> int64 / float64. */
int bt_csr_row_sums(int value_kind, int nrows, const int32_t *starts, const int32_t *lists,
                    const void *values, const int32_t *out_index, void *out, int accumulate,
                    double scale, void *stream);
int bt_range_sums_i64(int n, const int32_t *starts, const int32_t *counts, const int64_t *values,
                      int64_t *out, void *stream);
int bt_add_to_ranges_i64(int nrows, const int32_t *boxes, const int64_t *vals, const int32_t *starts,
                         const int32_t *counts, int64_t *pot, void *stream);
int bt_fmm_upward_i64(int dim, int nrows, const int32_t *boxes, const int32_t *box_child_ids,
                      int aligned_nboxes, int64_t *mpoles, void *stream);
int bt_fmm_downward_i64(int nrows, const int32_t *boxes, const int32_t *box_parent_ids,
                        int64_t *local, void *stream);
int bt_gather_i64(int64_t n, const int64_t *src, const int32_t *idx, int64_t *out, void *stream);
int bt_gather_coords(int dtype, int64_t n, const void *src, const int32_t *idx, void *out,
                     void *stream);
int bt_widen_i32(int value_kind, int64_t n, const int32_t *src, void *out, void *stream);
/* ParticleListFilter (boxtree/tree.py:1057-1239): user order = ListOfListsBuilder over boxes
 * (phase 0: counts -> starts + total, phase 1: lists); tree order = TREE_ORDER_TARGET_FILTER_*
 * (tree_build_kernels.py:1954-2021).  user_target_ids[tree position] = user id;
 * flags_user: int8 [ntargets] in user order. */
int bt_filter_targets_user_order(int phase, int nboxes, const int32_t *box_target_starts,
                                 const int32_t *box_target_counts_nonchild,
                                 const int32_t *user_target_ids, const int8_t *flags_user,
                                 int32_t *starts, int32_t *lists, int64_t *total_dev, void *stream);
int bt_filter_targets_tree_order(int nboxes, int64_t ntargets, const int32_t *box_target_starts,
                                 const int32_t *box_target_counts_nonchild,
                                 const int32_t *user_target_ids, const int8_t *flags_user,
                                 int32_t *filtered_from_unfiltered, int32_t *unfiltered_from_filtered,
                                 int32_t *nfiltered_dev, int32_t *filtered_starts,
                                 int32_t *filtered_counts, void *stream);
/* link_point_sources (boxtree/tree.py:773-955, POINT_SOURCE_LINKING_* of
 * tree_build_kernels.py:1872-1950).  phase 0: tree-order starts/counts + total (device);
 * phase 1: user ids of the point sources in tree order and the per-box ranges. */
int bt_link_point_sources(int phase, int nboxes, int64_t nsources,
                          const int32_t *point_source_starts_user, const int32_t *user_source_ids,
                          int32_t *tree_order_starts, int32_t *point_source_counts,
                          int32_t *npoint_sources_dev, int32_t *user_point_source_ids,
                          const int32_t *box_source_starts, const int32_t *box_source_counts_nonchild,
                          const int32_t *box_source_counts_cumul, int32_t *box_ps_starts,
                          int32_t *box_ps_nonchild, int32_t *box_ps_cumul, void *stream);
/* TRANSLATION_CLASS_FINDER_TEMPLATE (boxtree/translation_classes.py:60-196): the class of every
 * from_sep_siblings entry, used[class] = 1, *error_dev = 1 where the reference raises;
 * bt_remap_classes: classes[i] = used_map[classes[i]] (:424-426). */
int bt_translation_classes(int dtype, int dim, int nrows, const int32_t *row_boxes,
                           const int32_t *starts, const int32_t *lists, const void *box_centers,
                           int aligned_nboxes, const uint8_t *box_levels, double root_extent,
                           int well_sep_is_n_away, int per_level, int nclasses_per_level,
                           int64_t npairs, int32_t *classes, int32_t *used, int32_t *error_dev,
                           void *stream);
int bt_remap_classes(int64_t n, const int32_t *used_map, int32_t *classes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BOXTREE_B200_H */
