// pb_group.cuh — device bodies of the stable radix partition (histogram pass, scatter pass), written against a
// virtual block number (internal).  Used by pb_partition_by_shard (pb_sort.cu).
#pragma once
#include "pb_device.cuh"

namespace pb {

// ---- radix partition (see pb_sort.cu) --------------------------------------------------------------------------
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
struct SrcShard {  // sign_to_shard_modulo (mod.rs:341-345)
  const uint64_t* signs;
  uint32_t R;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return (uint32_t)(farmhash64_u64(signs[i]) % R); }
};

struct ValIdentity {
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i; }
};

// Pass-0 histogram, block-major hist[tile][bin].  Blocks cover 512 keys each (more blocks than tiles, so
// the dependent key loads are spread over the whole chip); counts go to the tile's row with global REDs —
// the row must be zero on entry.  Also writes the materialised keys, clears the later passes' rows and
// (block 0) the segment-list counters of the backward pass.
constexpr int RH_KEYS = 512;
__device__ __forceinline__ uint32_t* fine_row(uint32_t* h, uint32_t tile_idx) { return h + (RS_COARSE_ROWS + tile_idx) * RS_BINS; }
__device__ __forceinline__ const uint32_t* fine_row(const uint32_t* h, uint32_t tile_idx) { return h + (RS_COARSE_ROWS + tile_idx) * RS_BINS; }
__device__ __forceinline__ uint32_t* coarse_row(uint32_t* h, uint32_t tile_idx) { return h + (tile_idx / RS_COARSE) * RS_BINS; }
// vb = the (virtual) block: the block number of the stand-alone kernel, a loop index of the fused grouping kernel
template <typename SRC>
__device__ __forceinline__ void radix_hist_body(uint32_t vb, SRC src, uint32_t n, uint32_t tile,
                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ hist,
                                                uint32_t* __restrict__ z1, uint32_t* __restrict__ z2,
                                                uint32_t* __restrict__ z3, uint32_t* __restrict__ zero4) {
  __shared__ uint32_t cnt[RS_BINS];
  cnt[threadIdx.x] = 0;
  cnt[threadIdx.x + RS_THREADS] = 0;
  if (zero4 && vb == 0 && threadIdx.x < 4) zero4[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t beg = vb * RH_KEYS;
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

// One scatter pass.  A block owns one tile; inside a 512-key sub-tile warp w owns 64 consecutive keys
// (2 rounds of 32, kept in registers).  Phase 1: every warp counts its own digits (warp-private shared
// counters, __match_any_sync per round).  Phase 2: one sweep turns the counters into each warp's first
// output slot per bin.  Phase 3: every warp walks its keys again in order and writes them out, bumping
// its private cursors.  Two block barriers per sub-tile.
template <typename VALOP>
__device__ __forceinline__ void radix_scatter_body(uint32_t vb, uint32_t nb, const uint32_t* __restrict__ keys_in,
                                                   const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
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
  const uint32_t ncoarse = (nb + RS_COARSE - 1) / RS_COARSE, cb = vb / RS_COARSE;
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
      bool in = b < vb;
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

  const uint32_t beg = vb * tile, end = min(n, beg + tile);
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


}  // namespace pb
