// Shared host/device helpers for libbinary_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/binary_cuda.h"

namespace bcu {

// ---- error plumbing: thread-local message behind bcu_last_error() ---------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define BCU_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::bcu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (_e == cudaErrorMemoryAllocation) ? BCU_E_NOMEM : BCU_E_CUDA;                  \
    }                                                                                       \
  } while (0)

#define BCU_TRY(expr)       \
  do {                      \
    int _s = (expr);        \
    if (_s != BCU_OK) return _s; \
  } while (0)

#define BCU_LAUNCHED()                                  \
  do {                                                  \
    ::bcu::g_launches.fetch_add(1, std::memory_order_relaxed); \
    BCU_CUDA(cudaGetLastError());                       \
  } while (0)

// Device-side bounds assertions. compute-sanitizer is closed on this GPU pool, so memory safety is checked by a
// second build of the library (make check -> libbinary_cuda_check.so, -DBCU_BOUNDS_CHECK) whose kernels trap on any
// index outside the extent of the array it addresses; tests/test_gpu_bounds.py runs the parity shapes through it.
#ifdef BCU_BOUNDS_CHECK
#define BCU_DEV_ASSERT(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      printf("BCU_BOUNDS_CHECK failed: %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
      __trap();                                                                     \
    }                                                                               \
  } while (0)
#else
#define BCU_DEV_ASSERT(cond) \
  do {                       \
  } while (0)
#endif

// RAII guard: make `device` current for the scope, restore on exit.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- index layout in HBM --------------------------------------------------------------------------
// One descriptor per distinct group value (sorted by gval). Rows [row_begin,row_end) of the sorted
// arrays belong to the group; its directory slice is dir[bin_base .. bin_base+nb).
struct GroupDesc {
  uint32_t gval;
  uint32_t row_begin;
  uint32_t row_end;
  uint32_t nb;  // number of bins = directory entries = (cmax >> shift) + 1, cmax = max coordinate in the segment
  uint64_t bin_base;
  uint32_t proper;  // 1 = every row of the segment has low <= high (enables the O(1) long-range count)
  uint32_t shift;   // log2 of the bin width W of this segment's length class
};

constexpr int kMaxSmemGroups = 256;

// One directory entry = 16 bytes, read with a single 128-bit load: the two bounds of bin b plus the
// first candidate row, so a sparse query whose candidate range has one row needs ONE dependent access.
struct __align__(16) DirEntry {
  uint32_t lb;    // first row of the group with runmax >= b*W
  uint32_t ub;    // first row of the group with low >= (b+1)*W
  uint32_t low0;  // {low, high} of row lb (unspecified past the group's end; guarded by ub - lb)
  uint32_t high0;
};
constexpr uint32_t kInlineRows = 1;

// ---- bin layout: the index cut into shared-memory sized coordinate tiles (binned_join.cu) ----------------
// A BIN is a run of consecutive coordinate cells (width 1 << cell_shift) of one group. Its OWN rows are the rows
// whose low lies inside it -- a contiguous piece of the (group, low)-sorted arrays bn_low / bn_high / bn_id.
// The bin's coordinate range is cut into sub-cells of width 1 << ls; for every sub-cell the layout stores
//   sub_start[g]  the first own row with low >= the sub-cell's start, and
//   cov[g]        the {high, id} of every proper row (any bin, any length) that starts BEFORE the sub-cell's
//                 start and reaches it: low < start(g) <= high  ("coverage list", the rows a stabbing query at
//                 start(g) returns),
// so that the hits of a query q with q.low in sub-cell g are exactly
//   { r in cov[g] : high_r >= q.low }  +  { own rows r >= sub_start[g] with low_r <= q.high : high_r >= q.low }.
// No length classes and no running max are involved: a target of any length costs one entry per sub-cell start
// it covers. Queries are routed to the bin of their `low`.
constexpr uint32_t kBinTileBytes = 184320;  // own rows (12 B each) + the bin's blob must fit this much shared memory
constexpr uint32_t kBinMaxCells = 16384;    // cells over all groups: the routing table is staged in shared memory
constexpr uint32_t kBinMaxBins = 4095;      // bin ids are u16; kBinNull = no bin (unknown group / beyond every row)
constexpr uint32_t kBinNull = 0xFFFFu;

struct BinDesc {     // 64 bytes
  uint32_t group;    // index of the group in the sorted group table
  uint32_t x_begin;  // first coordinate of the bin
  uint32_t x_end;    // first coordinate past the bin; 0 = the group's last bin (no upper limit)
  uint32_t ls;       // log2 of the sub-cell width
  uint32_t nsub;     // sub-cells; sub_start and cov_rel have nsub + 1 entries (padded to a multiple of 8)
  uint32_t row0;     // first row of bn_low/high/id copied into the tile (multiple of 4: bulk copies move 16 B units)
  uint32_t n_copy;   // rows copied (multiple of 4)
  uint32_t lo, hi;   // the bin's own rows, tile-relative: [lo, hi)
  uint32_t n_cov;    // coverage entries of the bin (padded to a multiple of 4)
  uint32_t blob_bytes;  // bytes of the bin's blob (multiple of 16)
  uint32_t pad0;
  uint64_t blob;     // byte offset of the blob in bn_blob: sub_start u16[pad8(nsub+1)] | cov_rel u16[pad8(nsub+1)] |
                     // cov_high u32[n_cov] | cov_id u32[n_cov]
  uint64_t pad1;
};
static_assert(sizeof(BinDesc) == 64, "BinDesc is copied word by word");
struct BinGroup {    // query routing, per group (same order as the GroupDesc table)
  uint32_t gval;
  uint32_t cell_base;  // first entry of the group in cell2bin
  uint32_t n_cells;    // cells of the group; a query whose low lies beyond them cannot hit anything
  uint32_t pad;
};

}  // namespace bcu

struct bcu_index {
  int device = 0;
  uint64_t n = 0;
  uint32_t n_groups = 0;
  uint32_t n_comp = 1;          // length-class slots a query probes: 1, 2 or 4 (join.cu: virtual queries)
  uint32_t class_base_len = 0;  // class c holds lengths < class_base_len * 4^c
  uint32_t shift = 0;           // bin shift of length class 0 (reported by bcu_index_get_info)
  uint32_t shifts = 0;          // bin shift of class c in byte c: sparse classes get wider bins, i.e. small directories
  uint32_t max_gval = 0;  // largest group value (selects the direct group map in join.cu)
  uint32_t sort_passes = 0;
  uint64_t n_bins = 0;
  uint64_t bytes = 0;
  uint2* d_lowhigh = nullptr;        // [n+2] {low, high} of the targets sorted by (group, low, id)
  uint32_t* d_high = nullptr;        // [n+4] `high` of the same rows as a plain column (long-range scans)
  uint32_t* d_id = nullptr;          // [n+4] insertion ordinal of each sorted row
  uint32_t* d_runmax = nullptr;      // [n]   running max of high inside the group (max-end array)
  bcu::GroupDesc* d_groups = nullptr;  // [n_comp][n_groups], empty slots have nb == 0
  bcu::DirEntry* d_dir = nullptr;    // [n_bins] see DirEntry
  uint32_t* d_hs = nullptr;          // [n]   `high` of each segment's rows sorted ASCENDING (same row ranges)
  uint32_t* d_dirh = nullptr;        // [n_bins] first index of the segment's slice of d_hs with value >= b*W
  uint32_t max_len = 0;              // longest proper target (high - low): window bound of the top length class
  // bin layout (binned_join.cu); bn_bins == 0: the index is not eligible for the binned path
  uint32_t bn_bins = 0, bn_cells = 0, bn_cell_shift = 0;
  uint32_t* d_bn_low = nullptr;        // [n+4] the targets sorted by (group, low, id) -- BEFORE the length-class
  uint32_t* d_bn_high = nullptr;       //       permutation -- as plain columns; high/id alias d_high/d_id when
  uint32_t* d_bn_id = nullptr;         //       the index has one class (bn_owns_rows == false)
  bool bn_owns_rows = false;
  unsigned char* d_bn_blob = nullptr;  // per bin: sub-cell tables + coverage lists (BinDesc::blob)
  bcu::BinDesc* d_bn_desc = nullptr;   // [bn_bins]
  bcu::BinGroup* d_bn_groups = nullptr;  // [n_groups]
  uint16_t* d_bn_cell2bin = nullptr;   // [bn_cells] (used while building the coverage lists)
  uint32_t* d_bn_cellbits = nullptr;   // [2 * ceil(bn_cells / 32)] routing table of the sort kernel: bit i = cell i is the
                                       // first cell of a bin, then the number of such bits before every word
  // stab lists of the long-range emit (join.cu emit_long_kernel), only built for densely covered indexes:
  // per segment and coordinate bin of width 2^lc_shift, the rows with low < bin start <= high ({high, id} columns)
  // and the first row with low >= bin start. lc_bins == 0: none.
  uint32_t lc_shift = 0;
  uint64_t lc_bins = 0, lc_entries = 0;
  uint2* d_lc_seg = nullptr;       // [n_comp][n_groups] {first bin slot, number of bins (0 = no lists)}
  uint32_t* d_lc_off = nullptr;    // [lc_bins + 1] list of bin slot s = entries [off[s], off[s+1])
  uint32_t* d_lc_row0 = nullptr;   // [lc_bins]
  uint32_t* d_lc_high = nullptr;   // [lc_entries]
  uint32_t* d_lc_id = nullptr;     // [lc_entries]
};

namespace bcu {

// K1: stable LSD radix sort of (key64, val32) pairs on `stream`; on return *out_keys/*out_vals point at
// whichever of the two buffers holds the result. Only the digits in which the keys differ are run.
int radix_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint64_t n,
                     uint64_t varying_bits, cudaStream_t stream, uint64_t** out_keys,
                     uint32_t** out_vals, uint32_t* passes);

// Generic decoupled look-back scans (scan.cu)
int exclusive_sum_u32(const uint32_t* d_in, uint32_t* d_out, uint64_t n, cudaStream_t stream);
int segmented_running_max(const uint64_t* d_segkey /* equal within a segment */, const uint32_t* d_val,
                          uint32_t* d_out, uint64_t n, cudaStream_t stream);

// K3/K4 (join.cu)
enum JoinMode { kModeCount = 0, kModeScatter = 1, kModeFused = 2, kModeAny = 3 };
int launch_join(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup,
                const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                uint64_t* d_total, uint8_t* d_any, uint32_t query_id_base, cudaStream_t stream,
                const uint64_t* d_offset_base = nullptr,  // device u64 added to every offset (chunked joins)
                const bcu_filter* filter = nullptr, const uint8_t* d_qstrand = nullptr,
                uint64_t* total_mapped = nullptr);  // device-visible pinned host u64 that also receives the total

// bin layout of a finished index (index_build.cu). group_begin[g] .. group_begin[g+1] = rows of group g in the
// (group, low)-sorted arrays ix->d_bn_*; group_cmax[g] = largest coordinate of the group
int build_bin_layout(bcu_index* ix, const uint32_t* group_gval, const uint32_t* group_begin,
                     const uint32_t* group_cmax, cudaStream_t stream);
// stab lists of a finished index (index_build.cu); a no-op unless the index is densely covered
int build_long_lists(bcu_index* ix, cudaStream_t stream);
// the binned join (binned_join.cu); returns BCU_NOT_TAKEN when the call is not eligible: the caller then runs
// the general path
constexpr int BCU_NOT_TAKEN = 1;
int launch_join_binned(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup, const uint32_t* d_qlow,
                       const uint32_t* d_qhigh, uint64_t* d_offsets, uint64_t pair_capacity, uint32_t* d_hit_query,
                       uint32_t* d_hit_target, uint64_t* d_total, uint32_t query_id_base, cudaStream_t stream,
                       uint64_t* total_mapped);

}  // namespace bcu
