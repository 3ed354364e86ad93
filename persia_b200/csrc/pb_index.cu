// pb_index.cu — the shard's hash index and row store: find-or-admit, row initialisation, gather + pool
//  and the small id-preprocessing kernels (A2, A3 hash).
#include "pb_group.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------------
// A4 (admission part): initialise a newly admitted row, eight lanes cooperating.
// emb_entry.rs:28-68 + optim.rs:299-302.  The value stream restates rand 0.8.4 SmallRng (Xoshiro256++
// seeded through rand_core's PCG32 expansion) + UniformFloat<f32> — PARITY UNPINNED (no reference test
// asserts an initial value); the oracle carries the same restatement.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void init_row(const TableDev& t, const HyperDev& hy, const OptimDev& op, uint64_t seed,
                                      uint32_t row_idx, uint32_t sub) {
  // rand_core::SeedableRng::seed_from_u64 (PCG32 stream) -> 4 x u64 state
  uint64_t st = seed;
  uint32_t wds[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    st = st * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
    uint32_t rot = (uint32_t)(st >> 59);
    wds[k] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
  uint64_t s0 = wds[0] | ((uint64_t)wds[1] << 32), s1 = wds[2] | ((uint64_t)wds[3] << 32);
  uint64_t s2 = wds[4] | ((uint64_t)wds[5] << 32), s3 = wds[6] | ((uint64_t)wds[7] << 32);
  float* row = t.rows + (size_t)row_idx * t.stride;
  // every lane walks the whole stream (it is sequential) and keeps the elements it owns
  for (uint32_t e = 0; e < t.dim; ++e) {
    uint64_t sum = s0 + s3;
    uint64_t r = ((sum << 23) | (sum >> 41)) + s0;
    uint64_t tt = s1 << 17;
    s2 ^= s0;
    s3 ^= s1;
    s1 ^= s2;
    s0 ^= s3;
    s2 ^= tt;
    s3 = (s3 << 45) | (s3 >> 19);
    if ((e & (BUCKET - 1)) == sub) {
      uint32_t bits = ((uint32_t)(r >> 32) >> 9) | 0x3f800000u;
      float v01 = __fsub_rn(__uint_as_float(bits), 1.0f);
      row[e] = __fadd_rn(__fmul_rn(v01, hy.scale), hy.lo);
    }
  }
  float sv = (op.kind == PB_OPT_ADAGRAD || op.kind == PB_OPT_ADAGRAD_VW) ? op.init_acc : 0.0f;
  for (uint32_t e = t.dim + sub; e < t.stride; e += BUCKET) row[e] = (e < t.dim + t.state_floats) ? sv : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// A2 + A4 (index part): eight lanes per id occurrence; a group reads one 128 B bucket per step.
//   MODE_FIND   read-only probe (inference lookup, update, get_rows)
//   MODE_TRAIN  find, refresh recency, admit on miss (training lookup)
//   MODE_SET    find or force-admit without initialisation (set_embedding)
// Output: the index cell of every occurrence (n_cells + N_SPECIAL when the sign has no storage).  The row number
// is read from the cell by the kernels that follow, so nothing here ever waits on another thread.
// The group that admits a sign also initialises its row (emb_entry.rs:28-68 + optim.rs:299-302).
// Recency (get_refresh, eviction_map.rs:48-60) is not written here: thousands of occurrences of one hot
// sign would all store to the same cell; k_elect_leaders records it per row after a block-level dedup.
// Invariant that makes "an empty cell in the bucket => the sign is absent" true: a sign is stored no later
// in its probe sequence than the first bucket that had an EMPTY cell when it was admitted, and a cell never
// returns to EMPTY: eviction leaves a tombstone, which lookups walk past and admissions reuse (after having
// seen an EMPTY cell further on, i.e. knowing the sign is absent).
// ------------------------------------------------------------------------------------------------
#ifndef PB_PROBE_BLOCKS
#define PB_PROBE_BLOCKS 5  // resident blocks per SM the probe is compiled for (48 registers; 6 and 8 measured no faster)
#endif
template <int MODE, bool PREFIX>
__global__ void __launch_bounds__(256, PB_PROBE_BLOCKS) k_probe(TableDev t, HyperDev hy, OptimDev op, SlotsDev sl,
                                               const uint64_t* __restrict__ ids, uint32_t n,
                                               uint32_t* __restrict__ occ_cell) {
  const uint32_t tick = t.counters[CTR_TICK];
  const uint32_t i = (blockIdx.x * 256 + threadIdx.x) / BUCKET;
  const uint32_t sub = threadIdx.x % BUCKET;
  const uint32_t gshift = (threadIdx.x & 31) & ~(BUCKET - 1);  // this group's bit offset in a warp ballot
  const uint32_t h_none = t.n_cells + N_SPECIAL;
  bool valid = i < n;
  // lane 0 of the group derives the sign (prefix arithmetic, hash); the other seven take it by shuffle
  uint64_t sign = 0ULL;
  uint32_t bucket = 0;
  if (valid && sub == 0) {
    sign = ids[i];
    if (PREFIX) {
      uint64_t p = sl.prefix[slot_of_occ(sl, i)];
      if (p) sign = mod_mersenne(sign, sl.spacing_bits) + p;  // indices_add_prefix, mod.rs:402-429
    }
    bucket = (uint32_t)(mix64(sign)) & t.bucket_mask;
  }
  sign = __shfl_sync(0xffffffffu, sign, gshift);
  bucket = __shfl_sync(0xffffffffu, bucket, gshift);
  const bool null_sign = sl.null_sign && sign == PB_NULL_SIGN;  // padding of a framed exchange: no lookup, reads as zeros
  const bool special = (sign >= KEY_TOMB);  // the three signs that collide with a marker have their own cells
  const uint32_t special_cell = t.n_cells + (uint32_t)(KEY_EMPTY - sign);
  const unsigned long long stored = special ? 0ULL : sign;
  bool admit = true;
  if (MODE == MODE_TRAIN && hy.admit_p < 1.0f) {  // reference: unseeded thread_rng draw (unpinned)
    float u = (float)(mix64(sign ^ (0x9E3779B97F4A7C15ULL * (tick + 1))) >> 40) * (1.0f / 16777216.0f);
    admit = u < hy.admit_p;
  }
  uint32_t result = h_none;
  bool done = !valid || null_sign;
  const uint32_t home = bucket;
  uint32_t tomb_cell = 0xFFFFFFFFu;  // first tombstone met on the probe path (admissions reuse it)
  for (uint32_t step = 0; step <= 2u * (t.bucket_mask + 1u); ++step) {
    if (!__any_sync(0xffffffffu, !done)) break;
    const bool look = !done && (!special || sub == 0);
    const uint32_t cell = special ? special_cell : bucket * BUCKET + sub;
    uint4 c = make_uint4(0xFFFFFFFEu, 0xFFFFFFFFu, ROW_PENDING, 0u);  // neither empty, tombstone nor any sign's low word pair
    if (look) c = __ldcg(reinterpret_cast<const uint4*>(t.cells + cell));
    const unsigned long long kk = (unsigned long long)c.x | ((unsigned long long)c.y << 32);
    const uint32_t mm = (__ballot_sync(0xffffffffu, look && kk == stored) >> gshift) & 0xffu;
    const uint32_t em = (__ballot_sync(0xffffffffu, look && kk == KEY_EMPTY) >> gshift) & 0xffu;
    const uint32_t tm = (__ballot_sync(0xffffffffu, look && kk == KEY_TOMB && !special) >> gshift) & 0xffu;
    const uint32_t lm = mm ? __ffs(mm) - 1 : 0;  // lane of the match
    if (!done && !mm && tm && tomb_cell == 0xFFFFFFFFu) tomb_cell = bucket * BUCKET + (__ffs(tm) - 1);
    // an EMPTY cell in this bucket (and no match so far) proves the sign absent: admit it into the first tombstone
    // seen on the way, else into the first empty cell here
    const bool try_ins = !done && !mm && em && MODE != MODE_FIND && admit;
    const bool use_tomb = try_ins && tomb_cell != 0xFFFFFFFFu;
    const uint32_t le = em ? __ffs(em) - 1 : 0;  // lane of the first free cell
    const uint32_t free_cell = special ? special_cell : (use_tomb ? tomb_cell : bucket * BUCKET + le);
    unsigned long long old = 0ULL;
    if (try_ins && sub == le) old = atomicCAS(&t.cells[free_cell].key, use_tomb ? KEY_TOMB : KEY_EMPTY, stored);
    old = __shfl_sync(0xffffffffu, old, gshift + le);
    const bool won_cas = try_ins && old == (use_tomb ? KEY_TOMB : KEY_EMPTY);  // lane `le` of this group admitted the sign
    uint32_t row = 0;
    if (won_cas && sub == le) {
      // storage: a row released by eviction if there is one, else the next never-used row
      uint32_t f = atomicSub(&t.counters[CTR_FREE], 1u);
      if (f > 0 && f <= t.capacity) {
        row = t.free_rows[f - 1];
      } else {
        atomicAdd(&t.counters[CTR_FREE], 1u);
        row = atomicAdd(&t.counters[CTR_ROWS], 1u);
      }
      if (row >= t.capacity) {
        row = ROW_NONE;
        atomicAdd(&t.counters[CTR_FULL], 1u);
      } else {
        atomicAdd(&t.counters[CTR_ADMIT], 1u);
        t.row_lead[row] = (unsigned long long)tick << 32;  // recency of a fresh row (training lookups raise it)
      }
      *reinterpret_cast<volatile uint32_t*>(&t.cells[free_cell].row) = row;
    }
    row = __shfl_sync(0xffffffffu, row, gshift + le);
    if (MODE == MODE_TRAIN && won_cas && row != ROW_NONE) init_row(t, hy, op, sign, row, sub);  // all 8 lanes
    if (!done) {
      if (mm) {
        result = special ? special_cell : bucket * BUCKET + lm;
        done = true;
      } else if (em) {
        if (!try_ins) {
          done = true;  // absent (and not admitted)
        } else if (won_cas) {
          result = (row == ROW_NONE) ? h_none : free_cell;
          done = true;
        } else if (old == stored) {  // a duplicate occurrence won the race for the same cell
          result = free_cell;
          done = true;
        } else {
          // another sign took the cell (possibly this very sign took a different one): search again from home
          bucket = home;
          tomb_cell = 0xFFFFFFFFu;
        }
      } else {
        if (special) done = true;  // cannot happen: a reserved cell only ever holds its sign
        bucket = (bucket + 1) & t.bucket_mask;  // no match and no empty cell: next line
      }
    }
  }
  if (valid && sub == 0) {
    if (MODE != MODE_SET && result == h_none && !null_sign) atomicAdd(&t.counters[CTR_MISS], 1u);
    occ_cell[i] = result;
  }
}

// ------------------------------------------------------------------------------------------------
// A4 + A5: gather + pool.  A group of G lanes owns one output row (slot s, sample b); lanes stride over
// VEC-float chunks of the embedding.  f32 accumulate in sample order, optional 1/sqrt(max(n,1)), RNE to
// f16 (mod.rs:547-579, persia-common lib.rs:157-161).  OUT_F32 writes plain f32 rows (pb_lookup).
// ------------------------------------------------------------------------------------------------
template <int VEC, bool OUT_F32>
__device__ __forceinline__ void store_out(void* out, size_t o, const float (&acc)[VEC], float scale) {
  if (OUT_F32) {
    float* dst = reinterpret_cast<float*>(out) + o;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[VEC > 1 ? 2 : 0], acc[VEC > 1 ? 3 : 0]);
    else dst[0] = acc[0];
  } else {
    __half* dst = reinterpret_cast<__half*>(out) + o;
    if (VEC == 4) {
      __half2 a = __floats2half2_rn(__fmul_rn(acc[0], scale), __fmul_rn(acc[VEC > 1 ? 1 : 0], scale));
      __half2 b = __floats2half2_rn(__fmul_rn(acc[VEC > 1 ? 2 : 0], scale), __fmul_rn(acc[VEC > 1 ? 3 : 0], scale));
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(dst) = pk;
    } else {
      dst[0] = __float2half_rn(__fmul_rn(acc[0], scale));
    }
  }
}

constexpr int GATHER_ROWS = 4;  // output rows per group in the one-id-per-sample layout (independent loads in flight)

template <int VEC, int G, bool OUT_F32>
__global__ void __launch_bounds__(256) k_gather_pool(TableDev t, SlotsDev sl, const uint32_t* __restrict__ occ_cell,
                                                     const uint32_t* __restrict__ row_off, uint32_t n_out,
                                                     uint32_t batch, void* __restrict__ out) {
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const uint32_t lane = threadIdx.x % G;
  const uint32_t nvec = t.dim / VEC;
  if (!row_off) {
    // one occurrence per output row: GATHER_ROWS rows per group, every stage issued for all rows before use
    const uint32_t r0 = group * GATHER_ROWS;
    if (r0 >= n_out) return;
    uint32_t cell[GATHER_ROWS], row[GATHER_ROWS];
#pragma unroll
    for (int k = 0; k < GATHER_ROWS; ++k) cell[k] = (r0 + k < n_out) ? occ_cell[r0 + k] : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < GATHER_ROWS; ++k) row[k] = (cell[k] < t.n_cells + N_SPECIAL) ? t.cells[cell[k]].row : ROW_NONE;
    for (uint32_t c = lane; c < nvec; c += G) {
      float v[GATHER_ROWS][VEC];
#pragma unroll
      for (int k = 0; k < GATHER_ROWS; ++k) {
        if (row[k] < t.capacity) {
          load_vec<VEC>(t.rows + (size_t)row[k] * t.stride + c * VEC, v[k]);
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[k][e] = 0.0f;
        }
      }
#pragma unroll
      for (int k = 0; k < GATHER_ROWS; ++k)
        if (r0 + k < n_out) {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (!OUT_F32) v[k][e] = __fadd_rn(0.0f, v[k][e]);  // the EW adds into a zeroed row (mod.rs:555-561)
          store_out<VEC, OUT_F32>(out, (size_t)(r0 + k) * t.dim + c * VEC, v[k], 1.0f);  // 1/sqrt(max(1,1)) = 1
        }
    }
    return;
  }
  const uint32_t gid = group;
  if (gid >= n_out) return;
  const uint32_t beg = row_off[gid], end = row_off[gid + 1];
  float scale = 1.0f;
  if (!OUT_F32 && batch && sl.sqrt_scaling[gid / batch]) {
    uint32_t cnt = end - beg;
    scale = __fdiv_rn(1.0f, __fsqrt_rn((float)(cnt > 1 ? cnt : 1)));
  }
  for (uint32_t c = lane; c < nvec; c += G) {
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
    for (uint32_t j = beg; j < end; ++j) {
      uint32_t h = occ_cell[j];
      uint32_t row = (h < t.n_cells + N_SPECIAL) ? t.cells[h].row : ROW_NONE;
      if (row >= t.capacity) continue;
      float v[VEC];
      load_vec<VEC>(t.rows + (size_t)row * t.stride + c * VEC, v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], v[k]);
    }
    store_out<VEC, OUT_F32>(out, (size_t)gid * t.dim + c * VEC, acc, scale);
  }
}

// ------------------------------------------------------------------------------------------------
// Leader election (training).  The backward pass groups the occurrences of a batch by the position of the
// first occurrence of their sign — a key that does not depend on thread timing.  Every row keeps
// (batch number << 32 | ~position) in TableDev::row_lead and occurrences race with atomicMax; a hot sign
// (tiny-cardinality slots repeat an id thousands of times) would serialise thousands of atomics on one
// address, so a block first reduces its 256 occurrences in shared memory and only distinct rows go to
// global memory.  The high half doubles as the row's recency (get_refresh, eviction_map.rs:48-60).
// Also records the row of every occurrence and clears the first radix histogram (side job).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_elect_leaders(TableDev t, const uint32_t* __restrict__ occ_cell, uint32_t n,
                                                       uint32_t* __restrict__ occ_row, uint32_t* __restrict__ zero,
                                                       uint32_t zero_words) {
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < zero_words; w += gridDim.x * blockDim.x) zero[w] = 0;
  elect_body(blockIdx.x, t, occ_cell, n, occ_row);
}

// set_embedding / get_rows: whole entries (emb ++ state), one group per sign.
template <bool WRITE>
__global__ void __launch_bounds__(256) k_copy_entries(TableDev t, const uint32_t* __restrict__ occ_cell, uint32_t n,
                                                      float* __restrict__ entries, uint8_t* __restrict__ found) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  uint32_t h = occ_cell[warp];
  uint32_t elen = t.dim + t.state_floats;
  uint32_t row = (h < t.n_cells + N_SPECIAL) ? t.cells[h].row : ROW_NONE;
  bool ok = row < t.capacity;
  if (!WRITE && found && lane == 0) found[warp] = ok ? 1 : 0;
  float* e = entries + (size_t)warp * elen;
  if (WRITE) {
    if (!ok) return;
    float* dst = t.rows + (size_t)row * t.stride;
    for (uint32_t i = lane; i < t.stride; i += 32) dst[i] = (i < elen) ? e[i] : 0.0f;
  } else {
    const float* src = t.rows + (size_t)row * t.stride;
    for (uint32_t i = lane; i < elen; i += 32) e[i] = ok ? src[i] : 0.0f;
  }
}

// CSR row offsets -> output row of every occurrence (multi-id slots)
__global__ void __launch_bounds__(256) k_expand_rows(const uint32_t* __restrict__ row_off, uint32_t n_out,
                                                     uint32_t* __restrict__ occ_outrow) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_out) return;
  for (uint32_t j = row_off[r]; j < row_off[r + 1]; ++j) occ_outrow[j] = r;
}

__global__ void __launch_bounds__(256) k_add_prefix(SlotsDev sl, const uint64_t* __restrict__ ids, uint32_t n,
                                                    uint64_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t p = sl.prefix[slot_of_occ(sl, i)];
  uint64_t v = ids[i];
  out[i] = p ? mod_mersenne(v, sl.spacing_bits) + p : v;
}

__global__ void __launch_bounds__(256) k_shard_of(const uint64_t* __restrict__ signs, uint32_t n, uint32_t R,
                                                  uint32_t* __restrict__ shard, uint64_t* __restrict__ hash) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t h = farmhash64_u64(signs[i]);
  if (shard) shard[i] = (uint32_t)(h % R);
  if (hash) hash[i] = h;
}

// Opens a training request on the device: bumps the table's batch number (recency, leader election) and the
// context's own request number, by which the backward of this batch recognises its NaN marks.  The latter never
// repeats within a context, whatever tables it serves and whenever they are cleared.
__global__ void k_begin_batch(uint32_t* counters, uint32_t* ctx_tick) {
  counters[CTR_TICK] = counters[CTR_TICK] + 1;
  if (ctx_tick) *ctx_tick = *ctx_tick + 1;
}

__global__ void k_fill_cells(Cell* cells, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  Cell e;
  e.key = KEY_EMPTY;
  e.row = ROW_PENDING;
  e.aux = 0;
  for (; i < n; i += stride) cells[i] = e;
}

// Row permutations around the shard exchange (the EW regroups signs per parameter server and puts the
// returned rows back in batch order, mod.rs:886-919): out[i] = src[perm[i]] (gather) or out[perm[i]] = src[i]
// (scatter), rows of `row_words` 16-byte words, one group of lanes per row.
__global__ void __launch_bounds__(256) k_permute_rows(const uint4* __restrict__ src, const uint32_t* __restrict__ perm,
                                                      uint32_t n, uint32_t row_words, uint32_t lanes, int scatter,
                                                      uint4* __restrict__ out) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const uint32_t l = threadIdx.x % lanes;
  if (g >= n) return;
  const uint32_t p = perm[g];
  const size_t from = (size_t)(scatter ? g : p) * row_words, to = (size_t)(scatter ? p : g) * row_words;
  for (uint32_t w = l; w < row_words; w += lanes) out[to + w] = src[from + w];
}

__global__ void __launch_bounds__(256) k_permute_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm,
                                                     uint32_t n, uint64_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[perm[i]];
}

// ------------------------------------------------------------------------------------------------
// Capacity: the reference's EvictionMap drops the least recently used entry whenever an insert pushes it over
// capacity (eviction_map.rs:76-97).  Here recency is the batch number kept in row_lead's high half and
// eviction is a sweep run between batches when free storage is low: (1) histogram of row ages over the occupied
// cells, (2) the age threshold that releases `want` rows (never rows touched in the last `keep` batches),
// (3) tombstone those cells and push their rows on the free stack.  ev[0] = threshold age, ev[1] = want,
// ev[2] = go flag, ev[3..] = age histogram (EV_BINS bins, last bin = older).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t EV_BINS = 1024;
__global__ void k_evict_plan(TableDev t, uint32_t low_water, uint32_t target_free, uint32_t* __restrict__ ev) {
  // one block: decide whether a sweep is needed and clear the histogram
  uint32_t used = min(t.counters[CTR_ROWS], t.capacity);
  uint32_t free_now = (t.capacity - used) + t.counters[CTR_FREE];
  bool go = free_now < low_water;
  for (uint32_t i = threadIdx.x; i < EV_BINS; i += blockDim.x) ev[3 + i] = 0;
  if (threadIdx.x == 0) {
    ev[2] = go ? 1u : 0u;
    ev[1] = go ? (target_free > free_now ? target_free - free_now : 0u) : 0u;
    ev[0] = 0xFFFFFFFFu;
  }
}

__global__ void __launch_bounds__(256) k_evict_hist(TableDev t, uint32_t* __restrict__ ev) {
  if (!ev[2]) return;
  const uint32_t tick = t.counters[CTR_TICK];
  const uint64_t n = (uint64_t)t.n_cells + N_SPECIAL;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Cell c = t.cells[i];
    bool occupied = (i < t.n_cells) ? (c.key < KEY_TOMB) : (c.key == 0ULL);
    if (!occupied || c.row >= t.capacity) continue;
    uint32_t age = tick - (uint32_t)(t.row_lead[c.row] >> 32);
    atomicAdd(&ev[3 + min(age, EV_BINS - 1)], 1u);
  }
}

__global__ void k_evict_threshold(uint32_t keep, uint32_t* __restrict__ ev) {
  if (threadIdx.x || !ev[2]) return;
  uint32_t want = ev[1], got = 0, thr = 0xFFFFFFFFu;
  for (uint32_t a = EV_BINS; a-- > 0 && a > keep;) {  // oldest first; ages <= keep are protected
    got += ev[3 + a];
    thr = a;
    if (got >= want) break;
  }
  ev[0] = got ? thr : 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) k_evict_sweep(TableDev t, uint32_t* __restrict__ ev) {
  if (!ev[2] || ev[0] == 0xFFFFFFFFu) return;
  const uint32_t tick = t.counters[CTR_TICK], thr = ev[0];
  const uint64_t n = (uint64_t)t.n_cells + N_SPECIAL;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Cell c = t.cells[i];
    bool occupied = (i < t.n_cells) ? (c.key < KEY_TOMB) : (c.key == 0ULL);
    if (!occupied || c.row >= t.capacity) continue;
    uint32_t age = tick - (uint32_t)(t.row_lead[c.row] >> 32);
    if (age < thr) continue;
    t.cells[i].key = (i < t.n_cells) ? KEY_TOMB : KEY_EMPTY;  // reserved cells have no probe chain behind them
    t.cells[i].row = ROW_PENDING;
    t.row_lead[c.row] = 0ULL;
    uint32_t f = atomicAdd(&t.counters[CTR_FREE], 1u);
    t.free_rows[f] = c.row;
    atomicAdd(&t.counters[CTR_EVICT], 1u);
  }
}

// Fixed-capacity framing of the shard exchange: every (source, destination) pair owns `cap` slots, so the
// all-to-all has static shapes (no split sizes on the host, CUDA-graph capturable).  Unused slots carry
// PB_NULL_SIGN (which the owner's probe ignores) / zero rows.  counts[r] > cap raises *overflow.
__device__ __forceinline__ bool frame_slot(const uint32_t* __restrict__ counts, uint32_t R, uint32_t cap, uint32_t idx,
                                           uint32_t& src_pos, uint32_t* overflow) {
  const uint32_t r = idx / cap, k = idx % cap;
  uint32_t off = 0;
  for (uint32_t q = 0; q < r; ++q) off += counts[q];
  const uint32_t c = counts[r];
  if (k == 0 && c > cap && overflow) *overflow = 1;
  src_pos = off + k;
  return k < c;
}

__global__ void __launch_bounds__(256) k_pack_signs(const uint64_t* __restrict__ signs, const uint32_t* __restrict__ perm,
                                                    const uint32_t* __restrict__ counts, uint32_t R, uint32_t cap,
                                                    uint64_t* __restrict__ out, uint32_t* __restrict__ overflow) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * cap) return;
  uint32_t pos;
  out[idx] = frame_slot(counts, R, cap, idx, pos, overflow) ? signs[perm[pos]] : PB_NULL_SIGN;
}

// pack != 0: framed[r*cap + k] = rows[perm[off_r + k]] (zero rows in the padding); pack == 0: the inverse scatter
__global__ void __launch_bounds__(256) k_frame_rows(const uint4* __restrict__ src, const uint32_t* __restrict__ perm,
                                                    const uint32_t* __restrict__ counts, uint32_t R, uint32_t cap,
                                                    uint32_t row_words, uint32_t lanes, int pack, uint4* __restrict__ out) {
  const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const uint32_t l = threadIdx.x % lanes;
  if (idx >= R * cap) return;
  uint32_t pos;
  const bool valid = frame_slot(counts, R, cap, idx, pos, nullptr);
  if (pack) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    const size_t from = valid ? (size_t)perm[pos] * row_words : 0, to = (size_t)idx * row_words;
    for (uint32_t w = l; w < row_words; w += lanes) out[to + w] = valid ? src[from + w] : z;
  } else if (valid) {
    const size_t from = (size_t)idx * row_words, to = (size_t)perm[pos] * row_words;
    for (uint32_t w = l; w < row_words; w += lanes) out[to + w] = src[from + w];
  }
}

// ------------------------------------------------------------------------------------------------
// Shard exchange over NVLink peer memory (no NCCL, no host): the R segments of a framed buffer are stored
// straight into the peers' receive buffers (peer q gets segment q at slot `my_rank`), then a barrier kernel
// over peer-mapped flag words orders the step.  Buffers and flags live in symmetric memory mapped by the host
// side (torch symmetric memory); everything here is plain stores / loads on peer pointers.
// ------------------------------------------------------------------------------------------------
struct PeerPtrs {
  uint64_t p[16];
};

__global__ void __launch_bounds__(256) k_p2p_exchange(const uint4* __restrict__ src, PeerPtrs peers, uint32_t R,
                                                      uint32_t my_rank, uint32_t cap, uint32_t row_words, uint32_t lanes) {
  const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const uint32_t l = threadIdx.x % lanes;
  if (idx >= R * cap) return;
  const uint32_t q = idx / cap, k = idx % cap;
  uint4* dst = reinterpret_cast<uint4*>(peers.p[q]) + ((size_t)my_rank * cap + k) * row_words;
  const uint4* from = src + (size_t)idx * row_words;
  for (uint32_t w = l; w < row_words; w += lanes) dst[w] = from[w];
}

// flags: every rank owns an array of 16 u32; rank r's word [q] is written by rank q.  *epoch counts barriers.
__global__ void k_p2p_barrier(PeerPtrs flags, uint32_t* __restrict__ epoch, uint32_t R, uint32_t my_rank,
                              uint32_t* __restrict__ err) {
  __shared__ uint32_t e_s;
  if (threadIdx.x == 0) {
    e_s = *epoch + 1;
    *epoch = e_s;
  }
  __syncthreads();
  const uint32_t e = e_s;
  __threadfence_system();  // everything this GPU stored to its peers before the barrier is visible first
  if (threadIdx.x < R) {
    volatile uint32_t* theirs = reinterpret_cast<volatile uint32_t*>(flags.p[threadIdx.x]) + my_rank;
    *theirs = e;
    volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(flags.p[my_rank]) + threadIdx.x;
    uint32_t spins = 0;
    while ((int32_t)(*mine - e) < 0) {
      if (++spins > (1u << 27)) {  // a peer is not coming: give up instead of hanging the GPU, and say so
        if (err) *err = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// launchers (host)
// ------------------------------------------------------------------------------------------------
void launch_fill_cells(Cell* cells, uint64_t n, cudaStream_t st) { PB_LAUNCH(k_fill_cells, 148 * 8, 256, 0, st, cells, n); }

void launch_begin_batch(const TableDev& t, uint32_t* ctx_tick, cudaStream_t st) {
  PB_LAUNCH(k_begin_batch, 1, 1, 0, st, t.counters, ctx_tick);
}

void launch_probe(int mode, bool prefix, const TableDev& t, const HyperDev& hy, const OptimDev& op, const SlotsDev& sl,
                  const uint64_t* ids, uint32_t n, uint32_t* occ_cell, cudaStream_t st) {
  if (!n) return;
  uint32_t g = cdiv((uint64_t)n * BUCKET, 256);
  if (mode == MODE_FIND) {
    if (prefix) PB_LAUNCH_F(FAM_PROBE, (k_probe<MODE_FIND, true>), g, 256, 0, st, t, hy, op, sl, ids, n, occ_cell);
    else PB_LAUNCH_F(FAM_PROBE, (k_probe<MODE_FIND, false>), g, 256, 0, st, t, hy, op, sl, ids, n, occ_cell);
  } else if (mode == MODE_TRAIN) {
    if (prefix) PB_LAUNCH_F(FAM_PROBE, (k_probe<MODE_TRAIN, true>), g, 256, 0, st, t, hy, op, sl, ids, n, occ_cell);
    else PB_LAUNCH_F(FAM_PROBE, (k_probe<MODE_TRAIN, false>), g, 256, 0, st, t, hy, op, sl, ids, n, occ_cell);
  } else {
    PB_LAUNCH_F(FAM_PROBE, (k_probe<MODE_SET, false>), g, 256, 0, st, t, hy, op, sl, ids, n, occ_cell);
  }
}

template <int VEC, bool F32>
static void gather_dispatch(int G, const TableDev& t, const SlotsDev& sl, const uint32_t* occ_cell,
                            const uint32_t* row_off, uint32_t n_out, uint32_t batch, void* out, cudaStream_t st) {
  uint32_t grid;
#define PB_G(GG)                                                                                              \
  case GG:                                                                                                    \
    grid = cdiv((uint64_t)(row_off ? n_out : cdiv(n_out, GATHER_ROWS)) * GG, 256);                            \
    PB_LAUNCH_F(FAM_GATHER, (k_gather_pool<VEC, GG, F32>), grid, 256, 0, st, t, sl, occ_cell, row_off, n_out, batch, out);  \
    break;
  switch (G) {
    PB_G(1) PB_G(2) PB_G(4) PB_G(8) PB_G(16) PB_G(32)
  }
#undef PB_G
}

void launch_gather(const TableDev& t, const SlotsDev& sl, const uint32_t* occ_cell, const uint32_t* row_off,
                   uint32_t n_out, uint32_t batch, void* out, bool out_f32, cudaStream_t st) {
  if (!n_out) return;
  int vec, G;
  vec_group(t.dim, vec, G);
  if (vec == 4) {
    if (out_f32) gather_dispatch<4, true>(G, t, sl, occ_cell, row_off, n_out, batch, out, st);
    else gather_dispatch<4, false>(G, t, sl, occ_cell, row_off, n_out, batch, out, st);
  } else {
    if (out_f32) gather_dispatch<1, true>(G, t, sl, occ_cell, row_off, n_out, batch, out, st);
    else gather_dispatch<1, false>(G, t, sl, occ_cell, row_off, n_out, batch, out, st);
  }
}

void launch_elect(const TableDev& t, const uint32_t* occ_cell, uint32_t n, uint32_t* occ_row, uint32_t* zero,
                  uint32_t zero_words, cudaStream_t st) {
  if (n) PB_LAUNCH_F(FAM_SORT, k_elect_leaders, cdiv(n, 256), 256, 0, st, t, occ_cell, n, occ_row, zero, zero_words);
}

void launch_copy_entries(bool write, const TableDev& t, const uint32_t* occ_cell, uint32_t n, float* entries,
                         uint8_t* found, cudaStream_t st) {
  if (!n) return;
  uint32_t grid = cdiv((uint64_t)n * 32, 256);
  if (write) PB_LAUNCH(k_copy_entries<true>, grid, 256, 0, st, t, occ_cell, n, entries, found);
  else PB_LAUNCH(k_copy_entries<false>, grid, 256, 0, st, t, occ_cell, n, entries, found);
}

void launch_expand_rows(const uint32_t* row_off, uint32_t n_out, uint32_t* occ_outrow, cudaStream_t st) {
  if (n_out) PB_LAUNCH(k_expand_rows, cdiv(n_out, 256), 256, 0, st, row_off, n_out, occ_outrow);
}

void launch_add_prefix(const SlotsDev& sl, const uint64_t* ids, uint32_t n, uint64_t* out, cudaStream_t st) {
  if (n) PB_LAUNCH(k_add_prefix, cdiv(n, 256), 256, 0, st, sl, ids, n, out);
}

void launch_permute_rows(const void* src, const uint32_t* perm, uint32_t n, uint32_t row_bytes, int scatter, void* out,
                         cudaStream_t st) {
  if (!n) return;
  uint32_t words = row_bytes / 16, lanes = 1;
  while (lanes < words && lanes < 32) lanes <<= 1;
  PB_LAUNCH(k_permute_rows, cdiv((uint64_t)n * lanes, 256), 256, 0, st, (const uint4*)src, perm, n, words, lanes, scatter,
            (uint4*)out);
}

void launch_permute_u64(const uint64_t* src, const uint32_t* perm, uint32_t n, uint64_t* out, cudaStream_t st) {
  if (n) PB_LAUNCH(k_permute_u64, cdiv(n, 256), 256, 0, st, src, perm, n, out);
}

void launch_pack_signs(const uint64_t* signs, const uint32_t* perm, const uint32_t* counts, uint32_t R, uint32_t cap,
                       uint64_t* out, uint32_t* overflow, cudaStream_t st) {
  PB_LAUNCH(k_pack_signs, cdiv((uint64_t)R * cap, 256), 256, 0, st, signs, perm, counts, R, cap, out, overflow);
}

void launch_frame_rows(const void* src, const uint32_t* perm, const uint32_t* counts, uint32_t R, uint32_t cap,
                       uint32_t row_bytes, int pack, void* out, cudaStream_t st) {
  uint32_t words = row_bytes / 16, lanes = 1;
  while (lanes < words && lanes < 32) lanes <<= 1;
  PB_LAUNCH(k_frame_rows, cdiv((uint64_t)R * cap * lanes, 256), 256, 0, st, (const uint4*)src, perm, counts, R, cap, words,
            lanes, pack, (uint4*)out);
}

// Checkpointing (persia-model-manager dump_internal_shard_embeddings, lib.rs:242-257): the resident signs and the
// batch number each was last used in (the reference walks its LRU list; this is the same order up to ties inside
// a batch).  Writes at most `max_n` pairs in index order; *count receives the number of resident signs.
__global__ void __launch_bounds__(256) k_export_signs(TableDev t, uint64_t* __restrict__ signs,
                                                      uint32_t* __restrict__ recency, uint32_t max_n,
                                                      uint32_t* __restrict__ count) {
  const uint64_t n = (uint64_t)t.n_cells + N_SPECIAL;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Cell c = t.cells[i];
    const bool occupied = (i < t.n_cells) ? (c.key < KEY_TOMB) : (c.key == 0ULL);
    if (!occupied || c.row >= t.capacity) continue;
    const uint32_t k = atomicAdd(count, 1u);
    if (k < max_n) {
      signs[k] = (i < t.n_cells) ? c.key : KEY_EMPTY - (i - t.n_cells);  // the three marker-valued signs
      recency[k] = (uint32_t)(t.row_lead[c.row] >> 32);
    }
  }
}

void launch_export_signs(const TableDev& t, uint64_t* signs, uint32_t* recency, uint32_t max_n, uint32_t* count,
                         cudaStream_t st) {
  cudaMemsetAsync(count, 0, sizeof(uint32_t), st);
  PB_LAUNCH(k_export_signs, 148 * 8, 256, 0, st, t, signs, recency, max_n, count);
}

void launch_evict(const TableDev& t, uint32_t low_water, uint32_t target_free, uint32_t keep, uint32_t* ev, cudaStream_t st) {
  PB_LAUNCH(k_evict_plan, 1, 256, 0, st, t, low_water, target_free, ev);
  PB_LAUNCH(k_evict_hist, 148 * 8, 256, 0, st, t, ev);
  PB_LAUNCH(k_evict_threshold, 1, 32, 0, st, keep, ev);
  PB_LAUNCH(k_evict_sweep, 148 * 8, 256, 0, st, t, ev);
}

void launch_p2p_exchange(const void* src, const uint64_t* peer_ptrs, uint32_t R, uint32_t my_rank, uint32_t cap,
                         uint32_t row_bytes, cudaStream_t st) {
  PeerPtrs pp;
  for (uint32_t i = 0; i < 16; ++i) pp.p[i] = i < R ? peer_ptrs[i] : 0;
  uint32_t words = row_bytes / 16, lanes = 1;
  while (lanes < words && lanes < 32) lanes <<= 1;
  PB_LAUNCH(k_p2p_exchange, cdiv((uint64_t)R * cap * lanes, 256), 256, 0, st, (const uint4*)src, pp, R, my_rank, cap, words, lanes);
}

void launch_p2p_barrier(const uint64_t* flag_ptrs, uint32_t* epoch, uint32_t R, uint32_t my_rank, uint32_t* err, cudaStream_t st) {
  PeerPtrs pp;
  for (uint32_t i = 0; i < 16; ++i) pp.p[i] = i < R ? flag_ptrs[i] : 0;
  PB_LAUNCH(k_p2p_barrier, 1, 32, 0, st, pp, epoch, R, my_rank, err);
}

void launch_shard_of(const uint64_t* signs, uint32_t n, uint32_t R, uint32_t* shard, uint64_t* hash, cudaStream_t st) {
  if (n) PB_LAUNCH(k_shard_of, cdiv(n, 256), 256, 0, st, signs, n, R, shard, hash);
}

}  // namespace pb
