/* libbinary_cuda -- C ABI of the B200-native interval-overlap join.
 *
 * This is the one process/device boundary of the port. The reference (ylab-hi/BINARY) has NO FFI on
 * this path: its boundary is the C++ class template binary::algorithm::tree::IntervalTree. Each
 * entry point below names the reference interface it stands in for (paths relative to the
 * reference checkout, library/include/binary/algorithm/...):
 *
 *   bcu_index_build*      <- RbTree::insert_node(R&&) / insert_node(Args&&...)   rb_tree.hpp:111-117,145-149
 *                            + IntervalTree::insert_node_impl                   interval_tree.hpp:230-260
 *                            (one tree per chromosome: sv2nl mapper.hpp:147-162 -> `group`)
 *   bcu_query_count*      <- IntervalTree::find_overlaps(...).size()             interval_tree.hpp:161-168
 *   bcu_query_scatter*    <- IntervalTree::find_overlaps / find_overlaps_impl    interval_tree.hpp:306-334
 *   bcu_join*             <- the sv2nl hot loop: one find_overlaps per record    sv2nl mapper.hpp:207-218
 *   bcu_join_multi        <- the same loop spread over the pool's tasks          sv2nl mapper.hpp:238-246
 *   bcu_query_any*        <- IntervalTree::find_overlap(...).has_value()         interval_tree.hpp:290-304
 *   bcu_index_size        <- RbTree::size()                                      rb_tree.hpp:173-180
 *
 * Semantics (bit-exact with the reference): for query q and the multiset T of
 * inserted intervals, the hits are { i : group(q)==group(T[i]) && q.low <= T[i].high &&
 * T[i].low <= q.high } -- CLOSED intervals, evaluated exactly as written (interval_tree.hpp:119-121),
 * so inverted intervals (low > high; TraMapper inserts such, sv2nl mapper.cpp:127-142) are neither
 * rejected nor "fixed". Duplicates are distinct hits. target_id = 0-based insertion ordinal,
 * query_id = 0-based position in the batch. Output is CSR: offsets[n_q+1] (u64) + pairs sorted by
 * query_id; the order of targets inside one query is unspecified (the reference's is tree-shape
 * preorder) -- the parity contract is the sorted set of (query_id, target_id) pairs.
 *
 * Conventions: every function returns 0 (BCU_OK) or a negative bcu_status; the message of the last
 * failure on the calling thread is bcu_last_error(). No C++ types or exceptions cross the boundary.
 * Host-pointer functions: the caller owns all host buffers (any memory; pinned memory from
 * bcu_host_alloc is fastest), the library owns all device memory. `_dev` functions take DEVICE
 * pointers on the index's device plus a cudaStream_t (passed as void*; NULL = default stream), are
 * asynchronous with respect to the host and allocate scratch from the stream-ordered pool. An index
 * is immutable after build and may be queried concurrently from several host threads/streams.
 * There is no CPU fallback: without a usable CUDA device every call fails with BCU_E_CUDA.
 */
#ifndef BINARY_CUDA_H_
#define BINARY_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bcu_status {
  BCU_OK = 0,
  BCU_E_INVALID = -1,  /* bad argument                                              */
  BCU_E_CUDA = -2,     /* CUDA runtime/driver error (message has the CUDA string)   */
  BCU_E_NOMEM = -3,    /* host or device allocation failed                          */
  BCU_E_CAPACITY = -4, /* caller-provided pair buffer too small; *total = required  */
  BCU_E_LIMIT = -5     /* input exceeds a documented limit (n_t <= 2^31-1, n_q <= 2^32-2) */
} bcu_status;

typedef struct bcu_index bcu_index; /* opaque; lives on one device */

typedef struct bcu_index_info {
  uint64_t n_targets;
  uint32_t n_groups;     /* distinct group values among the targets                         */
  uint32_t n_components; /* sub-lists of the AIList-style decomposition (1 = plain list)     */
  uint32_t bin_shift;    /* directory bin width = 1 << bin_shift coordinate units            */
  uint32_t sort_passes;  /* radix passes the build ran                                       */
  uint64_t n_bins;       /* directory entries over all groups/components                     */
  uint64_t device_bytes; /* device memory held by the index                                  */
  int32_t device;
  int32_t binned_tiles;  /* shared-memory sized coordinate tiles of the index (0 = no bin layout: large batches
                            take the general path too); see bcu_join_dev */
} bcu_index_info;

/* Optional pair filter, evaluated on the device on top of the overlap predicate so that rejected pairs are
 * never counted or written. The two kinds are sv2nl's post-filters (query = validated non-linear record,
 * target = validated SV record; standalone/sv2nl/source/mapper.cpp in the reference):
 *   BCU_FILTER_SV2NL_DUP  DupMapper::check_condition  mapper.cpp:50-55:  target contains query and both ends
 *                         are within `diff` of each other
 *   BCU_FILTER_SV2NL_INV  InvMapper::check_condition  mapper.cpp:57-79:  neither contains the other, both ends
 *                         within `diff`, and (if use_strand) strand1 && !strand2 when q.low <= t.low,
 *                         !strand1 && strand2 otherwise; qstrand[i] bit0 = strand1 is '+', bit1 = strand2 is '+' */
typedef enum bcu_filter_kind { BCU_FILTER_NONE = 0, BCU_FILTER_SV2NL_DUP = 1, BCU_FILTER_SV2NL_INV = 2 } bcu_filter_kind;
typedef struct bcu_filter {
  uint32_t kind;       /* bcu_filter_kind */
  uint32_t diff;       /* sv2nl --dis */
  uint32_t use_strand; /* INV only */
  uint32_t reserved;
} bcu_filter;

/* ---- library / device ------------------------------------------------------------------------- */
const char* bcu_version(void);
const char* bcu_last_error(void);
int bcu_device_count(int* n);

/* Pinned host memory (cudaHostAlloc); optional, for fast host<->device copies. */
int bcu_host_alloc(void** ptr, size_t bytes);
int bcu_host_free(void* ptr);

/* ---- build: K1 radix sort by (group, low) + K2 flat augmented index ----------------------------
 * group may be NULL (= one group, id 0). Arrays are SoA u32, length n_t; target id = position. */
int bcu_index_build(int device, uint64_t n_t, const uint32_t* group, const uint32_t* low,
                    const uint32_t* high, bcu_index** out);
int bcu_index_build_dev(int device, uint64_t n_t, const uint32_t* d_group, const uint32_t* d_low,
                        const uint32_t* d_high, void* stream, bcu_index** out);
/* Lifetime: bcu_index_free synchronises the index's device before releasing its memory, so *_dev work
 * still queued on ANY stream of that device finishes first (the call therefore blocks; it is not for
 * hot loops). Do not call it concurrently with a query on the same index from another host thread. */
int bcu_index_free(bcu_index* index);
int bcu_index_size(const bcu_index* index, uint64_t* n_t);
int bcu_index_get_info(const bcu_index* index, bcu_index_info* info);

/* ---- query, host buffers (H2D/D2H inside the call) ----------------------------------------------
 * qgroup may be NULL (= group 0). offsets has n_q+1 entries. */
int bcu_query_count(const bcu_index* index, uint64_t n_q, const uint32_t* qgroup,
                    const uint32_t* qlow, const uint32_t* qhigh, uint64_t* offsets,
                    uint64_t* total);
int bcu_query_scatter(const bcu_index* index, uint64_t n_q, const uint32_t* qgroup,
                      const uint32_t* qlow, const uint32_t* qhigh, const uint64_t* offsets,
                      uint32_t* hit_query, uint32_t* hit_target);
/* One call for the whole join (what the sv2nl loop switches to): the batch is cut into chunks that are
 * pipelined over copy-in / compute / copy-out streams. pair_capacity = entries available in
 * hit_query/hit_target; if the join produces more, returns BCU_E_CAPACITY with *total = required entries
 * (offsets are complete and valid in that case, pairs are not). hit_query may be NULL: the column is
 * redundant with the offsets and skipping it saves a fifth of the device-to-host traffic on sparse joins.
 * Staging buffers are cached per host thread; bcu_trim() releases them. */
int bcu_join(const bcu_index* index, uint64_t n_q, const uint32_t* qgroup, const uint32_t* qlow,
             const uint32_t* qhigh, uint64_t* offsets, uint64_t pair_capacity, uint32_t* hit_query,
             uint32_t* hit_target, uint64_t* total);
int bcu_trim(void);
/* The same join on SEVERAL GPUs in one call (the reference's counterpart: one thread-pool task per chromosome over
 * shared trees, sv2nl mapper.hpp:238-246, mapper.cpp:136-140). indexes[d] are replicas of ONE target set, one per
 * device (build them with bcu_index_build*, or ship one with bcu_index_export_dev / import_dev). The batch is cut
 * into n_dev contiguous query ranges, range d is joined on the device of indexes[d] by a worker thread that owns
 * that device's streams and pinned-buffer staging (created on first use, kept for the life of the process); the
 * devices never exchange data -- the only exchange is each range's hit total, on the host. Result: the CSR of
 * bcu_join over the whole batch, written in place (query ids and offsets are global). Either `offsets` (u64,
 * n_q + 1) or `counts` (u32 hits per query, n_q; half the bytes over PCIe) or both; hit_query may be NULL. */
int bcu_join_multi(const bcu_index* const* indexes, int n_dev, uint64_t n_q, const uint32_t* qgroup,
                   const uint32_t* qlow, const uint32_t* qhigh, uint64_t* offsets, uint32_t* counts,
                   uint64_t pair_capacity, uint32_t* hit_query, uint32_t* hit_target, uint64_t* total);
/* bcu_join with a pair filter (see bcu_filter). qstrand: n_q bytes, may be NULL unless kind is INV with
 * use_strand. */
int bcu_join_filtered(const bcu_index* index, const bcu_filter* filter, uint64_t n_q, const uint32_t* qgroup,
                      const uint32_t* qlow, const uint32_t* qhigh, const uint8_t* qstrand, uint64_t* offsets,
                      uint64_t pair_capacity, uint32_t* hit_query, uint32_t* hit_target, uint64_t* total);
/* any[i] = 1 iff query i overlaps at least one target (shape-independent part of find_overlap). */
int bcu_query_any(const bcu_index* index, uint64_t n_q, const uint32_t* qgroup,
                  const uint32_t* qlow, const uint32_t* qhigh, uint8_t* any);

/* ---- query, device buffers (no PCIe; what the roofline is measured on) --------------------------
 * d_offsets: u64[n_q+1] exclusive prefix sums, d_offsets[n_q] = total. */
int bcu_query_count_dev(const bcu_index* index, uint64_t n_q, const uint32_t* d_qgroup,
                        const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                        void* stream);
int bcu_query_scatter_dev(const bcu_index* index, uint64_t n_q, const uint32_t* d_qgroup,
                          const uint32_t* d_qlow, const uint32_t* d_qhigh,
                          const uint64_t* d_offsets, uint32_t* d_hit_query, uint32_t* d_hit_target,
                          void* stream);
/* Count + prefix sum + scatter (probe and emit kernels back to back). d_total: u64[1] on device,
 * receives the number of pairs the join has (also when it exceeds pair_capacity, in which case nothing
 * is written at or beyond the capacity, offsets and total are complete, and the pairs below the capacity
 * are complete only on the general path -- treat them as unspecified and retry with a larger buffer).
 * query_id_base is added to every emitted query id (sharding). d_hit_query may be NULL (column not produced).
 * Large batches (>= 2 Mi queries) against an index with a bin layout that exceeds L2 are answered by the
 * binned path (queries routed to shared-memory sized tiles of the index; binned_join.cu); the result is the
 * same CSR. BCU_BINNED=0 / =1 in the environment disables / forces that path. */
int bcu_join_dev(const bcu_index* index, uint64_t n_q, const uint32_t* d_qgroup,
                 const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                 uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                 uint64_t* d_total, uint32_t query_id_base, void* stream);
int bcu_join_filtered_dev(const bcu_index* index, const bcu_filter* filter, uint64_t n_q,
                          const uint32_t* d_qgroup, const uint32_t* d_qlow, const uint32_t* d_qhigh,
                          const uint8_t* d_qstrand, uint64_t* d_offsets, uint64_t pair_capacity,
                          uint32_t* d_hit_query, uint32_t* d_hit_target, uint64_t* d_total,
                          uint32_t query_id_base, void* stream);
int bcu_query_any_dev(const bcu_index* index, uint64_t n_q, const uint32_t* d_qgroup,
                      const uint32_t* d_qlow, const uint32_t* d_qhigh, uint8_t* d_any,
                      void* stream);

/* ---- index image (build once, broadcast over NVLink) -----------------------------
 * The reference rebuilds its trees in every task (mapper.hpp:147-162, mapper.cpp:128-140 builds ONE tree and
 * shares it between threads). Across GPUs the equivalent of "share" is: build on one device, export the
 * index into one contiguous device buffer, move that buffer with any device-to-device transport (NCCL
 * broadcast, cudaMemcpyPeer), and import it on the receiving device. Export and import synchronise `stream`.
 * The image is position-independent, specific to this library version, and only valid between devices of
 * the same process group (no endianness / version negotiation). */
int bcu_index_image_size(const bcu_index* index, uint64_t* bytes);
/* d_image: `bytes` >= bcu_index_image_size bytes on the index's device (BCU_E_CAPACITY if smaller). */
int bcu_index_export_dev(const bcu_index* index, void* d_image, uint64_t bytes, void* stream);
/* d_image: an exported image resident on `device`; the new index owns copies, d_image may be freed after. */
int bcu_index_import_dev(int device, const void* d_image, uint64_t bytes, void* stream, bcu_index** out);

/* ---- sv2nl: one mapper from its queries to the lines that are written ------
 * The reference, per NL record of a mapper (standalone/sv2nl): find_overlaps -> filter(check_condition) ->
 * SV2NL_USE_CACHE duplicate-key rule -> writer (include/mapper.hpp:194-236, source/mapper.cpp:86-126). This
 * entry runs the (optionally filtered) join AND those rules on the device; only the final CSR returns:
 *   - `filter` (may be NULL): DupMapper / InvMapper check_condition, fused into the join kernels (bcu_filter);
 *   - rules->tra: TraMapper::check_condition (source/mapper.cpp:144-156) for the RE-KEYED translocation join --
 *     group = ordered chromosome pair (x bucket of the second breakpoint), target = the point p1, query =
 *     [p1 - diff, p1 + diff]; record r owns the `probes_per_record` consecutive queries r*p .. r*p + p - 1 (one
 *     per bucket a partner's p2 can fall into). Kept are the pairs with |p1 - p1'| <= diff, |p2 - p2'| <= diff
 *     whose raw intervals overlap as in the reference's tree of unvalidated BND records (mapper.cpp:103,158-170);
 *   - rules->dedup: of the records with the same key (format_map_key, include/helper.hpp:84-91, as four words)
 *     only the FIRST one with at least one kept pair keeps its pairs (include/mapper.hpp:204-229).
 * offsets: n_rec + 1 entries, per RECORD. BCU_E_CAPACITY (offsets and *total complete, no pairs) if the kept
 * pairs exceed pair_capacity. All pointers are host pointers; tgt_* are indexed by target id (insertion order). */
typedef struct bcu_sv2nl_rules {
  uint32_t probes_per_record; /* >= 1 */
  uint32_t tra;               /* != 0: apply the TRA rule */
  uint32_t diff;              /* sv2nl --dis */
  uint32_t dedup;             /* != 0: apply the duplicate-key rule */
  const uint32_t* rec_p1;     /* tra: [n_rec] ordered breakpoints (get_2chroms_with_pos, helper.hpp:76-82) */
  const uint32_t* rec_p2;     /*      of the VALIDATED NL record */
  const uint32_t* tgt_p1;     /* tra: [n_t] ordered breakpoints of the SV record (not validated) */
  const uint32_t* tgt_p2;
  const uint32_t* tgt_pos;    /* tra: [n_t] its POS / END as read: the interval the reference inserts */
  const uint32_t* tgt_end;
  const uint32_t* rec_key;    /* dedup: [4 * n_rec] */
} bcu_sv2nl_rules;
int bcu_sv2nl_join(const bcu_index* index, const bcu_filter* filter, const bcu_sv2nl_rules* rules, uint64_t n_rec,
                   const uint32_t* qgroup, const uint32_t* qlow, const uint32_t* qhigh, const uint8_t* qstrand,
                   uint64_t* offsets, uint64_t pair_capacity, uint32_t* hit_target, uint64_t* total);

/* Number of kernel launches this library has issued on the calling process (all threads). */
uint64_t bcu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BINARY_CUDA_H_ */
