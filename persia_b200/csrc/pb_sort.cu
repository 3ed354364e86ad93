// pb_sort.cu — stable LSD radix partition of a batch's id occurrences (SURVEY.md §8a rows A3, A8 grouping).
//
// "Warp-radix partition": ranks inside a warp come from __match_any_sync, across warps from per-warp digit
// counters in shared memory, across blocks from a block-major digit histogram that every scatter block folds
// itself (no separate scan kernel).  A scatter pass also builds the histogram of the next pass with global
// REDs on each element's destination tile, so a k-pass sort is 1 + k launches.  Digits are 9 bits.
//
// Two uses:
//  * backward grouping: key = position of the first occurrence of the occurrence's sign in the batch (elected
//    per row by the forward pass, materialised by the histogram pass), payload = position | slot << 24.
//    After the sort the occurrences of one sign are adjacent, ordered by slot and then by ascending position —
//    the order FeatureBatch::new pushed them (persia-common/src/lib.rs:45-82) — and signs follow each other in
//    first-seen order; the key does not depend on thread timing, so neither does anything derived from it.
//  * pb_partition_by_shard: key = farmhash64(sign) % R, one pass (indices_to_sharded_indices, mod.rs:454-479).
#include "pb_device.cuh"

namespace pb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 2;                    // keys per thread held in registers
constexpr int RS_SUB = RS_THREADS * RS_ITEMS;  // 512 keys per sub-tile: warp w owns keys [64 w, 64 w + 64)
constexpr int RS_COARSE = 16;                  // tiles per coarse histogram row
constexpr int RS_MAX_TILES = 256;
constexpr int RS_COARSE_ROWS = RS_MAX_TILES / RS_COARSE;
// One pass's histogram region: [RS_COARSE_ROWS coarse rows][RS_MAX_TILES fine rows] x RS_BINS words.  A scatter
// block folds <= 16 coarse rows + <= 15 fine rows instead of every tile's row.
constexpr int RS_BITS = 9;
constexpr int RS_BINS = 1 << RS_BITS;          // 512: two bins per thread

// key sources of the histogram pass -----------------------------------------------------------------
struct SrcLeader {  // occurrence -> row -> first occurrence of the sign in this batch; no storage sorts last
  const uint32_t* occ_row;
  const unsigned long long* row_lead;
  uint32_t n;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
    uint32_t row = occ_row[i];
    return row == ROW_NONE ? n : ~(uint32_t)row_lead[row];
  }
};
struct SrcShard {  // sign_to_shard_modulo (mod.rs:341-345)
  const uint64_t* signs;
  uint32_t R;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return (uint32_t)(farmhash64_u64(signs[i]) % R); }
};

struct ValIdentity {
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i; }
};
struct ValOccSlot {
  SlotsDev sl;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i | (slot_of_occ(sl, i) << 24); }
};

// Pass-0 histogram, block-major hist[tile][bin].  Blocks cover 512 keys each (more blocks than tiles, so
// the dependent key loads are spread over the whole chip); counts go to the tile's row with global REDs —
// the row must be zero on entry.  Also writes the materialised keys, clears the later passes' rows and
// (block 0) the segment-list counters of the backward pass.
constexpr int RH_KEYS = 512;
__device__ __forceinline__ uint32_t* fine_row(uint32_t* h, uint32_t tile_idx) { return h + (RS_COARSE_ROWS + tile_idx) * RS_BINS; }
__device__ __forceinline__ const uint32_t* fine_row(const uint32_t* h, uint32_t tile_idx) { return h + (RS_COARSE_ROWS + tile_idx) * RS_BINS; }
__device__ __forceinline__ uint32_t* coarse_row(uint32_t* h, uint32_t tile_idx) { return h + (tile_idx / RS_COARSE) * RS_BINS; }
template <typename SRC>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(SRC src, uint32_t n, uint32_t tile,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ hist,
                                                           uint32_t* __restrict__ z1, uint32_t* __restrict__ z2,
                                                           uint32_t* __restrict__ z3, uint32_t* __restrict__ zero4) {
  __shared__ uint32_t cnt[RS_BINS];
  cnt[threadIdx.x] = 0;
  cnt[threadIdx.x + RS_THREADS] = 0;
  if (zero4 && blockIdx.x == 0 && threadIdx.x < 4) zero4[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t beg = blockIdx.x * RH_KEYS;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t k[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t i = beg + r * RS_THREADS + threadIdx.x;
    k[r] = (i < n) ? src(i) : 0u;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t i = beg + r * RS_THREADS + threadIdx.x;
    bool valid = i < n;
    if (valid) keys_out[i] = k[r];
    uint32_t digit = valid ? (k[r] & (RS_BINS - 1)) : RS_BINS + lane;
    uint32_t peers = __match_any_sync(0xffffffffu, digit);  // one shared-memory atomic per distinct digit of the warp
    if (valid && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[digit], __popc(peers));
  }
  __syncthreads();
  const uint32_t row = beg / tile;  // RH_KEYS divides the tile size
  const bool first_of_tile = beg % tile == 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t bin = threadIdx.x + h * RS_THREADS;
    if (cnt[bin]) {
      atomicAdd(fine_row(hist, row) + bin, cnt[bin]);
      atomicAdd(coarse_row(hist, row) + bin, cnt[bin]);
    }
    if (first_of_tile) {
      uint32_t* z[3] = {z1, z2, z3};
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (z[k]) {
          fine_row(z[k], row)[bin] = 0;
          if (row % RS_COARSE == 0) coarse_row(z[k], row)[bin] = 0;
        }
    }
  }
}

// One scatter pass.  A block owns one tile; inside a 1024-key sub-tile warp w owns 128 consecutive keys
// (4 rounds of 32, kept in registers).  Phase 1: every warp counts its own digits (warp-private shared
// counters, __match_any_sync per round).  Phase 2: one sweep turns the counters into each warp's first
// output slot per bin.  Phase 3: every warp walks its keys again in order and writes them out, bumping
// its private cursors.  Two block barriers per sub-tile.
template <typename VALOP>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                              const uint32_t* __restrict__ vals_in,
                                                              uint32_t* __restrict__ keys_out,
                                                              uint32_t* __restrict__ vals_out, uint32_t n, uint32_t shift,
                                                              uint32_t tile, VALOP vop, const uint32_t* __restrict__ hist,
                                                              uint32_t* __restrict__ hist_next) {
  __shared__ uint32_t base[RS_BINS];
  __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
  __shared__ uint32_t wsum[2][RS_WARPS];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, d = threadIdx.x;
  // this block's first output slot per bin = (keys of smaller bins anywhere) + (same bin in earlier blocks);
  // thread d owns bins d and d + 256
  uint32_t below[2] = {0, 0}, total[2] = {0, 0};
  const uint32_t nb = gridDim.x;
  const uint32_t ncoarse = (nb + RS_COARSE - 1) / RS_COARSE, cb = blockIdx.x / RS_COARSE;
  {
    uint32_t v[RS_COARSE_ROWS][2];
#pragma unroll
    for (int c = 0; c < RS_COARSE_ROWS; ++c) {
      bool in = (uint32_t)c < ncoarse;
      v[c][0] = in ? hist[c * RS_BINS + d] : 0u;
      v[c][1] = in ? hist[c * RS_BINS + d + RS_THREADS] : 0u;
    }
    uint32_t f[RS_COARSE][2];
#pragma unroll
    for (int u = 0; u < RS_COARSE; ++u) {
      uint32_t b = cb * RS_COARSE + u;
      bool in = b < blockIdx.x;
      f[u][0] = in ? fine_row(hist, b)[d] : 0u;
      f[u][1] = in ? fine_row(hist, b)[d + RS_THREADS] : 0u;
    }
#pragma unroll
    for (int c = 0; c < RS_COARSE_ROWS; ++c) {
      total[0] += v[c][0];
      total[1] += v[c][1];
      if ((uint32_t)c < cb) {
        below[0] += v[c][0];
        below[1] += v[c][1];
      }
    }
#pragma unroll
    for (int u = 0; u < RS_COARSE; ++u) {
      below[0] += f[u][0];
      below[1] += f[u][1];
    }
  }
  uint32_t x[2] = {total[0], total[1]};  // inclusive scans of the two halves over the 256 threads
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y0 = __shfl_up_sync(0xffffffffu, x[0], o), y1 = __shfl_up_sync(0xffffffffu, x[1], o);
    if (lane >= o) {
      x[0] += y0;
      x[1] += y1;
    }
  }
  if (lane == 31) {
    wsum[0][warp] = x[0];
    wsum[1][warp] = x[1];
  }
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    wcnt[w][d] = 0;
    wcnt[w][d + RS_THREADS] = 0;
  }
  __syncthreads();
  uint32_t woff[2] = {0, 0}, lower_total = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    lower_total += wsum[0][w];
    if (w < (int)warp) {
      woff[0] += wsum[0][w];
      woff[1] += wsum[1][w];
    }
  }
  base[d] = woff[0] + x[0] - total[0] + below[0];
  base[d + RS_THREADS] = lower_total + woff[1] + x[1] - total[1] + below[1];
  __syncthreads();

  const uint32_t beg = blockIdx.x * tile, end = min(n, beg + tile);
  const uint32_t next_shift = shift + RS_BITS;
  for (uint32_t sub = beg; sub < end; sub += RS_SUB) {
    uint32_t key[RS_ITEMS], val[RS_ITEMS];
    const uint32_t w0 = sub + warp * (32 * RS_ITEMS) + lane;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {  // all loads of the sub-tile are in flight together
      uint32_t i = w0 + r * 32;
      if (i < end) {
        key[r] = keys_in[i];
        val[r] = vals_in ? vals_in[i] : vop(i);
      } else {
        key[r] = 0;
        val[r] = 0;
      }
    }
    // phase 1: warp-private digit counts
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      bool valid = w0 + r * 32 < end;
      uint32_t digit = valid ? ((key[r] >> shift) & (RS_BINS - 1)) : RS_BINS + lane;  // invalid lanes match nobody
      uint32_t peers = __match_any_sync(0xffffffffu, digit);
      if (valid && (peers & ((1u << lane) - 1u)) == 0) wcnt[warp][digit] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    // phase 2: counts -> first output slot of each warp per bin; the block cursor moves past the sub-tile
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t bin = d + h * RS_THREADS, o = base[bin];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        uint32_t c = wcnt[w][bin];
        wcnt[w][bin] = o;
        o += c;
      }
      base[bin] = o;
    }
    __syncthreads();
    // phase 3: ordered placement
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      bool valid = w0 + r * 32 < end;
      uint32_t digit = valid ? ((key[r] >> shift) & (RS_BINS - 1)) : RS_BINS + lane;
      uint32_t peers = __match_any_sync(0xffffffffu, digit);
      uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t pos = 0;
      if (valid) pos = wcnt[warp][digit] + rank;
      __syncwarp();
      if (valid && rank == 0) wcnt[warp][digit] += __popc(peers);
      __syncwarp();
      if (valid) {
        if (keys_out) keys_out[pos] = key[r];
        vals_out[pos] = val[r];
        if (hist_next) {
          const uint32_t nd = (key[r] >> next_shift) & (RS_BINS - 1), nt = pos / tile;
          atomicAdd(fine_row(hist_next, nt) + nd, 1u);
          atomicAdd(coarse_row(hist_next, nt) + nd, 1u);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {  // cursors -> zeroed counters for the next sub-tile
      wcnt[w][d] = 0;
      wcnt[w][d + RS_THREADS] = 0;
    }
    __syncthreads();
  }
}

// per-shard group sizes of a single-pass partition: column sums of the block-major histogram
__global__ void k_counts_from_hist(const uint32_t* __restrict__ hist, uint32_t n_blocks, uint32_t R,
                                   uint32_t* __restrict__ counts) {
  uint32_t d = threadIdx.x;
  if (d >= R) return;
  uint32_t t = 0;
  for (uint32_t c = 0; c * RS_COARSE < n_blocks; ++c) t += hist[c * RS_BINS + d];  // coarse rows hold the column sums
  counts[d] = t;
}

// ------------------------------------------------------------------------------------------------
// launchers (host)
// ------------------------------------------------------------------------------------------------
uint32_t radix_tile(uint32_t n) {
  // tiles are multiples of the 1024-key sub-tile; at most 256 of them so that folding the block-major
  // histogram inside every scatter block stays cheap
  uint32_t tile = RS_SUB;
  while (cdiv(n, tile) > RS_MAX_TILES) tile += RS_SUB;
  return tile;
}
uint32_t radix_hist_words() { return (uint32_t)(RS_COARSE_ROWS + RS_MAX_TILES) * RS_BINS; }  // one pass's region
uint32_t radix_hist_zero_words(uint32_t n) {  // what must be zero before the histogram pass of an n-key sort
  return (uint32_t)(RS_COARSE_ROWS + cdiv(n ? n : 1, radix_tile(n ? n : 1))) * RS_BINS;
}

void launch_zero_words(uint32_t* p, uint32_t n_words, cudaStream_t st);

// Sorts the occurrences of a batch by the first occurrence of their sign (stable), i.e. groups them per
// sign in first-seen order.  keys_a receives the materialised keys; the result alternates between the
// (keys_b, vals_b) and (keys_a, vals_a) pairs; returns 0 if it ends in the a pair, 1 if in the b pair.
// hist: 4 x radix_hist_words() u32, the first pass's rows zero on entry.
int launch_radix_sort_leader(const TableDev& t, const uint32_t* occ_row, uint32_t n, const SlotsDev& sl, uint32_t* keys_a,
                             uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* zero4,
                             cudaStream_t st) {
  if (!n) return 0;
  uint32_t bits = 1;
  while ((1ull << bits) <= (uint64_t)n) ++bits;  // keys are in [0, n]
  uint32_t passes = (bits + RS_BITS - 1) / RS_BITS;
  uint32_t tile = radix_tile(n), nb = cdiv(n, tile);
  const uint32_t W = radix_hist_words();
  uint32_t* h[4] = {hist, hist + W, hist + 2 * W, hist + 3 * W};
  SrcLeader src{occ_row, t.row_lead, n};
  PB_LAUNCH_F(FAM_SORT, (k_radix_hist<SrcLeader>), cdiv(n, RH_KEYS), RS_THREADS, 0, st, src, n, tile, keys_a, h[0],
              passes > 1 ? h[1] : nullptr, passes > 2 ? h[2] : nullptr, passes > 3 ? h[3] : nullptr, zero4);
  const uint32_t* kin = keys_a;
  const uint32_t* vin = nullptr;
  ValOccSlot vop{sl};
  int cur = 0;  // pair holding the current input keys (a after the histogram pass)
  for (uint32_t p = 0; p < passes; ++p) {
    cur ^= 1;
    uint32_t* kout = cur == 0 ? keys_a : keys_b;
    uint32_t* vout = cur == 0 ? vals_a : vals_b;
    uint32_t* hn = (p + 1 < passes) ? h[p + 1] : nullptr;
    PB_LAUNCH_F(FAM_SORT, (k_radix_scatter<ValOccSlot>), nb, RS_THREADS, 0, st, kin, vin, kout, vout, n, p * RS_BITS, tile,
                vop, h[p], hn);
    kin = kout;
    vin = vout;
  }
  return cur;
}

__global__ void k_zero_words(uint32_t* p, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}
void launch_zero_words(uint32_t* p, uint32_t n_words, cudaStream_t st) {
  if (n_words) PB_LAUNCH(k_zero_words, cdiv(n_words, 1024) < 148 ? cdiv(n_words, 1024) : 148, 256, 0, st, p, n_words);
}

// d_work layout: [one pass's histogram region][n materialised shard ids]
uint64_t partition_workspace_bytes(uint32_t n) { return ((uint64_t)radix_hist_words() + n) * sizeof(uint32_t); }

void launch_partition_by_shard(const uint64_t* signs, uint32_t n, uint32_t R, uint32_t* perm, uint32_t* counts,
                               uint32_t* work, cudaStream_t st) {
  uint32_t tile = radix_tile(n ? n : 1), nb = cdiv(n ? n : 1, tile);
  uint32_t* hist = work;
  uint32_t* keys = work + radix_hist_words();
  SrcShard src{signs, R};
  launch_zero_words(hist, radix_hist_zero_words(n), st);
  if (n)
    PB_LAUNCH((k_radix_hist<SrcShard>), cdiv(n, RH_KEYS), RS_THREADS, 0, st, src, n, tile, keys, hist, (uint32_t*)nullptr,
              (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr);
  PB_LAUNCH(k_counts_from_hist, 1, 256, 0, st, hist, nb, R, counts);
  if (n)
    PB_LAUNCH((k_radix_scatter<ValIdentity>), nb, RS_THREADS, 0, st, keys, (const uint32_t*)nullptr, (uint32_t*)nullptr,
              perm, n, 0u, tile, ValIdentity(), hist, (uint32_t*)nullptr);
}

}  // namespace pb
