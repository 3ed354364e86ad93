// pb_kernels.cuh — kernel-side parameter blocks and launcher prototypes (internal).
#pragma once
#include <atomic>

#include "pb_common.cuh"

namespace pb {

// per-batch slot table, passed by value in kernel-parameter memory
struct SlotsDev {
  uint64_t prefix[PB_MAX_SLOTS];
  uint32_t occ_off[PB_MAX_SLOTS + 1];  // slot boundaries in the flat id array
  uint8_t sqrt_scaling[PB_MAX_SLOTS];
  uint32_t n_slots;
  uint64_t spacing;  // 2^(64-prefix_bit) - 1
  uint32_t spacing_bits;  // 64 - prefix_bit
  uint32_t null_sign;     // owner-mode contexts: PB_NULL_SIGN entries are padding, not lookups
  uint32_t uniform;       // occurrences per slot when every slot holds the same number (one id per sample), else 0
};

// per-batch gradient table (GradientBatch, persia-core/src/backward.rs:74-106)
struct GradsDev {
  const void* ptr[PB_MAX_SLOTS];  // nullptr = skipped slot
  float inv_scale[PB_MAX_SLOTS];  // 1/scale_factor
  uint8_t do_scale[PB_MAX_SLOTS]; // |scale-1| > f32::EPSILON (mod.rs:751)
  const float* adam_pow;          // Adam: accumulated (beta1^t, beta2^t) pairs on the device, one per feature group
  uint8_t pow_idx[PB_MAX_SLOTS];  // the slot's pair
};

// arguments of the backward segment kernels (see pb_update.cu)
struct SegArgs {
  const uint32_t* skey;        // first occurrence of the occurrence's sign (n = no storage), sorted
  const uint32_t* occ_row;     // row number of every occurrence (ROW_NONE = no storage)
  const uint32_t* sval;        // occurrence position | slot << 24
  const uint32_t* occ_outrow;  // nullptr: one id per sample per slot (output row == occurrence)
  const uint32_t* row_off;
  const uint32_t* tick_ptr;
  const uint32_t* nan_tick;
  float* partials;  // 2 rows of dim floats per PIECE-block
  float* vw_stage;
  uint32_t n, batch, piece, shared_groups, quiet_miss;
};
constexpr uint32_t PB_PIECE = 32;

// raw slots (pb_raw.cu): per-batch scratch set of distinct signs and its workspace
struct RawCell {
  uint64_t key;
  uint32_t first;  // first occurrence of the sign in the flat id array
  uint32_t rank;   // number of the sign among the batch's distinct signs (first-occurrence order)
};
struct RawWork {
  RawCell* set;      // set_mask + 2 cells
  uint32_t set_mask;
  uint32_t* occ_set;        // set cell of every occurrence
  uint32_t* flag;           // scan input / second scan output
  uint32_t* rank;           // scan output
  uint32_t* tiles;          // scan spine
  uint32_t* distinct_cell;  // index cell of every distinct sign (forward -> backward)
  uint32_t* counts;         // [0] distinct signs, [1] ids placed in `index`
};

void launch_fill_cells(Cell* cells, uint64_t n, cudaStream_t st);
void launch_begin_batch(const TableDev& t, uint32_t* ctx_tick, cudaStream_t st);
void launch_probe(int mode, bool prefix, const TableDev& t, const HyperDev& hy, const OptimDev& op, const SlotsDev& sl,
                  const uint64_t* ids, uint32_t n, uint32_t* occ_cell, cudaStream_t st);
void launch_gather(const TableDev& t, const SlotsDev& sl, const uint32_t* occ_cell, const uint32_t* row_off,
                   uint32_t n_out, uint32_t batch, void* out, bool out_f32, cudaStream_t st);
void launch_elect(const TableDev& t, const uint32_t* occ_cell, uint32_t n, uint32_t* occ_row, uint32_t* zero,
                  uint32_t zero_words, cudaStream_t st);
void launch_find_heads(const SegArgs& a, uint4* heads, uint2* owners, uint32_t* counts, cudaStream_t st);
void launch_copy_entries(bool write, const TableDev& t, const uint32_t* occ_cell, uint32_t n, float* entries,
                         uint8_t* found, cudaStream_t st);
void launch_nan_scan(const GradsDev& gr, uint32_t n_slots, uint32_t elems_per_slot, bool f16, const uint32_t* tick,
                     uint32_t* nan_tick, int32_t* status, cudaStream_t st);
void launch_reduce_update(const TableDev& t, const OptimDev& op, const HyperDev& hy, const SlotsDev& sl,
                          const GradsDev& gr, bool f16, const SegArgs& a, uint4* heads, uint2* owners,
                          uint32_t* counts, cudaStream_t st);
// n_ptr (optional): the live count on the device (<= n); tick/nan_tick (optional): skip everything when equal
void launch_update_direct(const TableDev& t, const OptimDev& op, const HyperDev& hy, const uint32_t* occ_cell,
                          const float* grads, uint32_t n, const float* adam_pair, cudaStream_t st,
                          const uint32_t* n_ptr = nullptr, const uint32_t* tick = nullptr,
                          const uint32_t* nan_tick = nullptr);
constexpr uint32_t PB_ADAM_KEYS = 256;  // beta-power pairs per table; the last one serves pb_update
struct AdamKeys {
  uint8_t idx[PB_MAX_SLOTS];
  uint32_t n;
};
void launch_adam_fill(float* pow, float b1, float b2, cudaStream_t st);
void launch_adam_advance(float* pow, const AdamKeys& keys, float b1, float b2, cudaStream_t st);
uint32_t raw_scan_tiles(uint32_t n);
void launch_raw_forward(const TableDev& t, const SlotsDev& sl, const uint64_t* ids, uint32_t n,
                        const uint32_t* row_off, const uint32_t* occ_sample, uint32_t batch, uint32_t fixed,
                        const uint32_t* occ_cell, const RawWork& w, void* table_f16, long long* index,
                        long long* non_empty, uint32_t* sample_id_num, cudaStream_t st);
void launch_raw_nan(const void* grad, bool f16, const uint32_t* n_distinct, uint32_t dim, const uint32_t* tick,
                    uint32_t* nan_tick, cudaStream_t st);
void launch_raw_stage(const void* grad, bool f16, const uint32_t* n_distinct, uint32_t dim, float inv_scale,
                      bool do_scale, float* out, cudaStream_t st);
void launch_slot_status(const GradsDev& gr, uint32_t n_slots, const uint32_t* tick, const uint32_t* nan_tick,
                        int32_t* status, cudaStream_t st);
uint32_t radix_tile(uint32_t n);
uint32_t radix_hist_words();
uint32_t radix_hist_zero_words(uint32_t n);
int launch_radix_sort_leader(const TableDev& t, const uint32_t* occ_row, uint32_t n, const SlotsDev& sl, uint32_t* keys_a,
                             uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* zero4,
                             cudaStream_t st);
void launch_zero_words(uint32_t* p, uint32_t n_words, cudaStream_t st);
uint64_t partition_workspace_bytes(uint32_t n);
void launch_partition_by_shard(const uint64_t* signs, uint32_t n, uint32_t R, uint32_t* perm, uint32_t* counts,
                               uint32_t* work, cudaStream_t st);
void launch_expand_rows(const uint32_t* row_off, uint32_t n_out, uint32_t* occ_outrow, cudaStream_t st);
void launch_add_prefix(const SlotsDev& sl, const uint64_t* ids, uint32_t n, uint64_t* out, cudaStream_t st);
void launch_shard_of(const uint64_t* signs, uint32_t n, uint32_t R, uint32_t* shard, uint64_t* hash, cudaStream_t st);
void launch_permute_rows(const void* src, const uint32_t* perm, uint32_t n, uint32_t row_bytes, int scatter, void* out,
                         cudaStream_t st);
void launch_permute_u64(const uint64_t* src, const uint32_t* perm, uint32_t n, uint64_t* out, cudaStream_t st);
void launch_pack_signs(const uint64_t* signs, const uint32_t* perm, const uint32_t* counts, uint32_t R, uint32_t cap,
                       uint64_t* out, uint32_t* overflow, cudaStream_t st);
void launch_frame_rows(const void* src, const uint32_t* perm, const uint32_t* counts, uint32_t R, uint32_t cap,
                       uint32_t row_bytes, int pack, void* out, cudaStream_t st);
void launch_export_signs(const TableDev& t, uint64_t* signs, uint32_t* recency, uint32_t max_n, uint32_t* count,
                         cudaStream_t st);
void launch_evict(const TableDev& t, uint32_t low_water, uint32_t target_free, uint32_t keep, uint32_t* ev, cudaStream_t st);
void launch_p2p_exchange(const void* src, const uint64_t* peer_ptrs, uint32_t R, uint32_t my_rank, uint32_t cap,
                         uint32_t row_bytes, cudaStream_t st);
void launch_p2p_barrier(const uint64_t* flag_ptrs, uint32_t* epoch, uint32_t R, uint32_t my_rank, uint32_t* err, cudaStream_t st);
uint64_t launch_count();
enum { FAM_PROBE = 0, FAM_COMBINE, FAM_GATHER, FAM_NAN, FAM_SORT, FAM_UPDATE, FAM_OTHER, FAM_COUNT };
void profile_enable(uint32_t family_mask);
void profile_read(double* ms, uint64_t* count, int n_families);

enum { MODE_FIND = 0, MODE_TRAIN = 1, MODE_SET = 2 };

}  // namespace pb
