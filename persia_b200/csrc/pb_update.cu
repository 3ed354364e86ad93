// pb_update.cu — backward of the sparse path: NaN rule, in-order gradient segment reduce (A8) fused with the
// optimizer step and weight bound on the resident rows (A9).  SURVEY.md §8a.
#include <cstdlib>

#include "pb_group.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------------
// The optimizer step on one VEC-chunk (persia-simd/src/lib.rs, persia-common/src/optim.rs:227-307).
// The reference runs 8-wide AVX2 FMAs on elements [0, 8*floor(len/8)) and an UNFUSED scalar tail after
// that; both forms are reproduced per element so that SGD is bit-exact and Adagrad differs from the
// reference only by its _mm256_rsqrt_ps approximation (exact 1/sqrt here, as in the reference's tail).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bound(float w, const HyperDev& hy) {
  return hy.enable_wb ? fminf(fmaxf(w, -hy.wb), hy.wb) : w;
}

__device__ __forceinline__ void sgd_elem(float& w, float g, bool fused, const OptimDev& op) {
  if (fused) {
    float dg = __fmaf_rn(op.wd, w, g);
    w = __fmaf_rn(-op.lr, dg, w);
  } else {
    float dg = __fadd_rn(g, __fmul_rn(w, op.wd));
    w = __fsub_rn(w, __fmul_rn(op.lr, dg));
  }
}

__device__ __forceinline__ void adagrad_elem(float& w, float& s, float g, bool fused, const OptimDev& op) {
  float sq = __fmul_rn(g, g);
  float r = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(s, op.eps)));
  float scaled = __fmul_rn(g, r);
  if (fused) {
    w = __fmaf_rn(-op.lr, scaled, w);
    s = __fmaf_rn(s, op.mom, sq);
  } else {
    w = __fadd_rn(__fmul_rn(-op.lr, scaled), w);
    s = __fadd_rn(__fmul_rn(s, op.mom), sq);
  }
}

// adam_avx2 (persia-simd/src/lib.rs:147-228); b1p/b2p = accumulated beta powers of the feature group.
__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, bool fused, const OptimDev& op,
                                          float r1, float r2) {
  float omb1 = __fsub_rn(1.0f, op.b1), omb2 = __fsub_rn(1.0f, op.b2);
  float um, uv;
  if (fused) {
    um = __fmaf_rn(op.b1, m, __fmul_rn(omb1, g));
    uv = __fmaf_rn(op.b2, v, __fmul_rn(omb2, __fmul_rn(g, g)));
  } else {
    um = __fadd_rn(__fmul_rn(op.b1, m), __fmul_rn(omb1, g));
    uv = __fadd_rn(__fmul_rn(op.b2, v), __fmul_rn(__fmul_rn(omb2, g), g));
  }
  float mc = __fmul_rn(um, r1), vc = __fmul_rn(uv, r2);
  float descent = __fdiv_rn(mc, __fadd_rn(op.eps, __fsqrt_rn(vc)));
  w = fused ? __fmaf_rn(-op.lr, descent, w) : __fsub_rn(w, __fmul_rn(op.lr, descent));
  m = um;
  v = uv;
}

// One VEC-chunk of a resident row: load (issued before the gradient reduce so that both fetches overlap),
// optimizer step + weight bound given the reduced gradient, store.
template <int VEC>
struct RowChunk {
  float w[VEC], s1[VEC], s2[VEC];
  __device__ __forceinline__ void load(const float* row, uint32_t c, const TableDev& t, const OptimDev& op) {
    load_vec<VEC>(row + c * VEC, w);
    if (op.kind == PB_OPT_ADAGRAD || op.kind == PB_OPT_ADAM) load_vec<VEC>(row + t.dim + c * VEC, s1);
    if (op.kind == PB_OPT_ADAM) load_vec<VEC>(row + 2 * t.dim + c * VEC, s2);
  }
  __device__ __forceinline__ void step(uint32_t c, const float (&g)[VEC], const TableDev& t, const OptimDev& op,
                                       const HyperDev& hy, float vw_state, float r1, float r2) {
    const uint32_t fused_end = (t.dim / 8) * 8;
    if (op.kind == PB_OPT_SGD) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        sgd_elem(w[k], g[k], c * VEC + k < fused_end, op);
        w[k] = bound(w[k], hy);
      }
    } else if (op.kind == PB_OPT_ADAGRAD) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        adagrad_elem(w[k], s1[k], g[k], c * VEC + k < fused_end, op);
        w[k] = bound(w[k], hy);
      }
    } else if (op.kind == PB_OPT_ADAGRAD_VW) {  // emb step with the OLD scalar state (lib.rs:81-121)
      float r = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(vw_state, op.eps)));
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float scaled = __fmul_rn(g[k], r);
        w[k] = (c * VEC + k < fused_end) ? __fmaf_rn(-op.lr, scaled, w[k]) : __fadd_rn(__fmul_rn(-op.lr, scaled), w[k]);
        w[k] = bound(w[k], hy);
      }
    } else {  // Adam
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        adam_elem(w[k], s1[k], s2[k], g[k], c * VEC + k < fused_end, op, r1, r2);
        w[k] = bound(w[k], hy);
      }
    }
  }
  __device__ __forceinline__ void store(float* row, uint32_t c, const TableDev& t, const OptimDev& op) const {
    store_vec<VEC>(row + c * VEC, w);
    if (op.kind == PB_OPT_ADAGRAD || op.kind == PB_OPT_ADAM) store_vec<VEC>(row + t.dim + c * VEC, s1);
    if (op.kind == PB_OPT_ADAM) store_vec<VEC>(row + 2 * t.dim + c * VEC, s2);
  }
};

template <int VEC>
__device__ __forceinline__ void apply_chunk(float* row, uint32_t c, const float (&g)[VEC], const TableDev& t,
                                            const OptimDev& op, const HyperDev& hy, float vw_state, float r1,
                                            float r2) {
  RowChunk<VEC> rc;
  rc.load(row, c, t, op);
  rc.step(c, g, t, op, hy, vw_state, r1, r2);
  rc.store(row, c, t, op);
}

// ndarray 0.15 unrolled_dot order (8 partial sums, pairwise fold, scalar tail), serial per row: only
// lane 0 of the group calls it, reading the reduced gradient the group staged in shared memory.
__device__ float vw_dot(const float* g, uint32_t n) {
  float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t i = 0;
  for (; i + 8 <= n; i += 8)
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = __fadd_rn(p[k], __fmul_rn(g[i + k], g[i + k]));
  float sum = 0.0f;
  sum = __fadd_rn(sum, __fadd_rn(p[0], p[4]));
  sum = __fadd_rn(sum, __fadd_rn(p[1], p[5]));
  sum = __fadd_rn(sum, __fadd_rn(p[2], p[6]));
  sum = __fadd_rn(sum, __fadd_rn(p[3], p[7]));
  for (; i < n; ++i) sum = __fadd_rn(sum, __fmul_rn(g[i], g[i]));
  return sum;
}

// ------------------------------------------------------------------------------------------------
// A8 (NaN rule): a slot whose gradient holds any NaN is skipped whole (mod.rs:731-746).
// grid.y = slot.  status[s] = tick when a NaN was seen (no reset needed between batches).
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256) k_nan_scan(GradsDev gr, uint32_t elems_per_slot,
                                                  const uint32_t* __restrict__ tick_ptr,
                                                  uint32_t* __restrict__ nan_tick) {
  uint32_t s = blockIdx.y;
  const void* base = gr.ptr[s];
  if (!base) return;
  bool bad = false;
  uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (F16) {
    // 8 halves per 16 B load; every slot tensor is at least 16 B aligned (torch allocations are 512 B)
    const uint4* p = reinterpret_cast<const uint4*>(base);
    uint32_t nv = elems_per_slot / 8;
    for (uint32_t j = i; j < nv; j += stride) {
      uint4 v = p[j];
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t lo = w[k] & 0x7fffu, hi = (w[k] >> 16) & 0x7fffu;
        bad |= (lo > 0x7c00u) | (hi > 0x7c00u);
      }
    }
    const uint16_t* q = reinterpret_cast<const uint16_t*>(base);
    for (uint32_t j = nv * 8 + i; j < elems_per_slot; j += stride) bad |= (q[j] & 0x7fffu) > 0x7c00u;
  } else {
    const float* p = reinterpret_cast<const float*>(base);
    for (uint32_t j = i; j < elems_per_slot; j += stride) bad |= isnan(p[j]);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) nan_tick[s] = *tick_ptr;
}

__global__ void k_slot_status(GradsDev gr, uint32_t n_slots, const uint32_t* __restrict__ tick_ptr,
                              const uint32_t* __restrict__ nan_tick, int32_t* __restrict__ status) {
  uint32_t s = threadIdx.x;
  const uint32_t tick = *tick_ptr;
  if (s < n_slots) status[s] = !gr.ptr[s] ? 1 : (nan_tick[s] == tick ? 2 : 0);
}

// ------------------------------------------------------------------------------------------------
// A8 + A9.  Input: the occurrence list sorted by index cell (stable), so the occurrences of one sign are
// adjacent, ordered by slot and then by ascending position — the order FeatureBatch::new pushed them.
// A "segment" is one (sign, slot) run: the reference reduces it to one gradient (mod.rs:786-812) and the PS
// performs one optimizer step with it (PS mod.rs:380-398).
//
// k_find_heads — marks piece heads / cut-segment owners and compacts them into lists.
// k_reduce_update — one group of G lanes per piece head.  Pieces cut segments at multiples of PIECE positions so that a sign repeated thousands
//   of times in a batch (tiny-cardinality slots) is reduced by many groups at once.  A segment that fits
//   in one piece (the vast majority) is summed in reference order and the optimizer step + weight bound
//   are applied with the reduced gradient still in registers: bit-exact w.r.t. the reference order.
//   Other pieces store their partial sum (<= 2 per PIECE-block).
// k_combine_update — one block per cut segment (owner record): its lane groups add contiguous ranges of the
//   segment's partial sums in position order, group 0 adds the range sums and performs the step.  Deterministic, but
//   the f32 association differs from the reference's strictly sequential sum (documented tolerance); piece == 0
//   ("strict") disables cutting and restores the sequential order for any length.
// k_update_shared — only when two slots of one feature group can hold the same sign: such a sign gets
//   one step per slot, sequentially in slot order (mod.rs:720-822); one group walks the whole run.
// ------------------------------------------------------------------------------------------------
// One gradient chunk of one occurrence, already clamped / unscaled / sqrt-scaled as the EW does before summing
// (persia-common lib.rs:163-180, mod.rs:751-778).
struct PieceCtx {
  const void* gbase;
  float inv_scale;
  bool do_scale, sqrt_sc;
  uint32_t slot_row0;
};

template <int VEC, bool F16>
__device__ __forceinline__ void load_grad(float (&g)[VEC], const PieceCtx& pc, const SegArgs& a, const TableDev& t,
                                          uint32_t orow, uint32_t c) {
  size_t off = (size_t)(orow - pc.slot_row0) * t.dim + c * VEC;
  if (F16) {
    const __half* gp = reinterpret_cast<const __half*>(pc.gbase) + off;
    if (VEC == 4) {
      uint2 raw = *reinterpret_cast<const uint2*>(gp);
      // +-inf -> +-65504 (persia-common lib.rs:163-180), two halves per instruction; finite halves are inside already
      const __half2 lim = __floats2half2_rn(65504.0f, 65504.0f);
      __half2 h0 = __hmin2(__hmax2(*reinterpret_cast<__half2*>(&raw.x), __hneg2(lim)), lim);
      __half2 h1 = __hmin2(__hmax2(*reinterpret_cast<__half2*>(&raw.y), __hneg2(lim)), lim);
      float2 x = __half22float2(h0);
      float2 y = __half22float2(h1);
      g[0] = x.x; g[VEC > 1 ? 1 : 0] = x.y; g[VEC > 1 ? 2 : 0] = y.x; g[VEC > 1 ? 3 : 0] = y.y;
    } else {
      g[0] = fminf(fmaxf(__half2float(gp[0]), -65504.0f), 65504.0f);
    }
  } else {
    load_vec<VEC>(reinterpret_cast<const float*>(pc.gbase) + off, g);
  }
}

template <int VEC, bool F16, bool PLAIN>
__device__ __forceinline__ void add_grad(float (&acc)[VEC], const float (&g)[VEC], const PieceCtx& pc, const SegArgs& a,
                                         uint32_t orow) {
  if (PLAIN) {  // no loss scale, no sqrt scaling: the gradient is summed as it is
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], g[k]);
    return;
  }
  float f = 1.0f;
  if (pc.sqrt_sc) {  // mirror of the forward scaling, without its max(.,1) (mod.rs:757-768)
    uint32_t cnt = a.row_off ? a.row_off[orow + 1] - a.row_off[orow] : 1u;
    f = __fdiv_rn(1.0f, __fsqrt_rn((float)cnt));
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float v = g[k];
    if (pc.do_scale) v = __fmul_rn(v, pc.inv_scale);    // x 1/scale_factor
    if (pc.sqrt_sc) v = __fmul_rn(v, f);
    acc[k] = __fadd_rn(acc[k], v);
  }
}

// sum of the (scaled) gradients of occurrences [j0, j1) of `slot`, chunk c, in position order.
// Full batches of 8 and 4 occurrences have all their loads issued together; the adds stay sequential.
template <int VEC, bool F16, bool PLAIN>
__device__ __forceinline__ void reduce_piece_t(float (&acc)[VEC], const SegArgs& a, const TableDev& t, const SlotsDev& sl,
                                               const GradsDev& gr, uint32_t slot, uint32_t j0, uint32_t j1, uint32_t c) {
  PieceCtx pc;
  pc.gbase = gr.ptr[slot];
  pc.inv_scale = gr.inv_scale[slot];
  pc.do_scale = gr.do_scale[slot];
  pc.sqrt_sc = sl.sqrt_scaling[slot];
  pc.slot_row0 = slot * a.batch;
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
  uint32_t j = j0;
  for (; j + 8 <= j1; j += 8) {
    uint32_t orow[8];
    float g[8][VEC];
#pragma unroll
    for (int u = 0; u < 8; ++u) orow[u] = val_occ(a.sval[j + u]);
    if (!PLAIN && a.occ_outrow) {
#pragma unroll
      for (int u = 0; u < 8; ++u) orow[u] = a.occ_outrow[orow[u]];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) load_grad<VEC, F16>(g[u], pc, a, t, orow[u], c);
#pragma unroll
    for (int u = 0; u < 8; ++u) add_grad<VEC, F16, PLAIN>(acc, g[u], pc, a, orow[u]);
  }
  if (j + 4 <= j1) {
    uint32_t orow[4];
    float g[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) orow[u] = val_occ(a.sval[j + u]);
    if (!PLAIN && a.occ_outrow) {
#pragma unroll
      for (int u = 0; u < 4; ++u) orow[u] = a.occ_outrow[orow[u]];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) load_grad<VEC, F16>(g[u], pc, a, t, orow[u], c);
#pragma unroll
    for (int u = 0; u < 4; ++u) add_grad<VEC, F16, PLAIN>(acc, g[u], pc, a, orow[u]);
    j += 4;
  }
  for (; j < j1; ++j) {
    uint32_t orow = val_occ(a.sval[j]);
    if (!PLAIN && a.occ_outrow) orow = a.occ_outrow[orow];
    float g[VEC];
    load_grad<VEC, F16>(g, pc, a, t, orow, c);
    add_grad<VEC, F16, PLAIN>(acc, g, pc, a, orow);
  }
}

// sv0 = a.sval[j0], carried by the head record: a piece of one occurrence (the majority) loads no index at all
template <int VEC, bool F16>
__device__ __forceinline__ void reduce_piece(float (&acc)[VEC], const SegArgs& a, const TableDev& t, const SlotsDev& sl,
                                             const GradsDev& gr, uint32_t slot, uint32_t j0, uint32_t j1, uint32_t c,
                                             uint32_t sv0) {
  // the common case (one id per sample, loss scale 1, no sqrt scaling) gets a branch-free body
  const bool plain = !a.occ_outrow && !gr.do_scale[slot] && !sl.sqrt_scaling[slot];
  if (plain && j1 - j0 == 1) {
    PieceCtx pc;
    pc.gbase = gr.ptr[slot];
    pc.slot_row0 = slot * a.batch;
    float g[VEC];
    load_grad<VEC, F16>(g, pc, a, t, val_occ(sv0), c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(0.0f, g[k]);  // the reference adds into a zeroed row (-0 -> +0)
    return;
  }
  if (plain) reduce_piece_t<VEC, F16, true>(acc, a, t, sl, gr, slot, j0, j1, c);
  else reduce_piece_t<VEC, F16, false>(acc, a, t, sl, gr, slot, j0, j1, c);
}

// the optimizer step of one segment given per-lane reduced chunks produced by `reduce(c, acc)`
template <int VEC, int G, typename REDUCE>
__device__ __forceinline__ void step_segment(const TableDev& t, const OptimDev& op, const HyperDev& hy, const GradsDev& gr,
                                             uint32_t slot, uint32_t row, uint32_t lane, float* stage, REDUCE reduce) {
  float* prow = t.rows + (size_t)row * t.stride;
  const uint32_t nvec = t.dim / VEC;
  float vw_state = 0.0f, r1 = 0.0f, r2 = 0.0f;
  if (op.kind == PB_OPT_ADAGRAD_VW) vw_state = prow[t.dim];
  if (op.kind == PB_OPT_ADAM) {
    const float* pw = gr.adam_pow + 2u * gr.pow_idx[slot];
    r1 = __fdiv_rn(1.0f, __fsub_rn(1.0f, pw[0]));
    r2 = __fdiv_rn(1.0f, __fsub_rn(1.0f, pw[1]));
  }
  for (uint32_t c = lane; c < nvec; c += G) {
    RowChunk<VEC> rc;
    rc.load(prow, c, t, op);  // in flight while the gradients are fetched and summed
    float acc[VEC];
    reduce(c, acc);
    if (op.kind == PB_OPT_ADAGRAD_VW) store_vec<VEC>(stage + c * VEC, acc);
    rc.step(c, acc, t, op, hy, vw_state, r1, r2);
    rc.store(prow, c, t, op);
  }
  if (op.kind == PB_OPT_ADAGRAD_VW) {  // state = state*mom + dot(g,g)/dim (optim.rs:280-283)
    const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    __syncwarp(gmask);  // the staged gradient of every lane of the group is visible to lane 0
    if (lane == 0) {
      float gs = __fdiv_rn(vw_dot(stage, t.dim), (float)t.dim);
      prow[t.dim] = __fadd_rn(__fmul_rn(vw_state, op.mom), gs);
    }
    __syncwarp(gmask);
  }
}

__device__ __forceinline__ float* partial_slot(const SegArgs& a, const TableDev& t, uint32_t head) {
  uint32_t blk = head / a.piece;
  return a.partials + ((size_t)2 * blk + (head % a.piece ? 1 : 0)) * t.dim;
}

__global__ void __launch_bounds__(256) k_find_heads(SegArgs a, uint4* __restrict__ heads, uint2* __restrict__ owners,
                                                    uint32_t* __restrict__ counts) {
  find_heads_body(blockIdx.x, a, heads, owners, counts);
}

#ifndef PB_REDUCE_BLOCKS
#define PB_REDUCE_BLOCKS 3  // resident blocks per SM k_reduce_update is compiled for
#endif
// One group of G lanes per piece.  Every group starts on the piece of its own number and then takes pieces
// off a device-side counter (requested before the current piece is worked on, consumed after it), so long
// pieces, which come first in the list, never pile up on one group.
template <int VEC, int G, bool F16>
__global__ void __launch_bounds__(256, PB_REDUCE_BLOCKS) k_reduce_update(TableDev t, OptimDev op, HyperDev hy, SlotsDev sl, GradsDev gr,
                                                          SegArgs a, const uint4* __restrict__ heads,
                                                          uint32_t* __restrict__ counts) {
  const uint32_t lane = threadIdx.x % G;
  const uint32_t n_groups = gridDim.x * (blockDim.x / G);
  const uint32_t n_long = counts[0], n_work = n_long + counts[2];
  const uint32_t tick = *a.tick_ptr;
  const uint32_t wl = threadIdx.x & 31;
  const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
  // slots whose gradient is skipped or holds a NaN, as a bit mask per block (instead of a dependent load per piece)
  __shared__ uint32_t dead[PB_MAX_SLOTS / 32];
  if (threadIdx.x < PB_MAX_SLOTS / 32) dead[threadIdx.x] = 0u;
  __syncthreads();
  if (threadIdx.x < PB_MAX_SLOTS && (!gr.ptr[threadIdx.x] || a.nan_tick[threadIdx.x] == tick))
    atomicOr(&dead[threadIdx.x >> 5], 1u << (threadIdx.x & 31));
  __syncthreads();
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  while (w < n_work) {
    uint32_t next = 0;
    if (lane == 0) next = n_groups + atomicAdd(&counts[3], 1u);
    const uint4 hd = heads[w < n_long ? w : a.n - 1u - (w - n_long)];
    const uint32_t j = hd.x, e = hd.y & 0x7FFFFFFFu;
    const bool whole = hd.y >> 31;
    const uint32_t row = hd.z;
    const uint32_t slot = val_slot(hd.w);
    const bool live = !((dead[slot >> 5] >> (slot & 31)) & 1u);  // skipped / NaN slot: nothing is applied
    if (live && whole) {
      if (row >= t.capacity) {
        if (lane == 0 && !a.quiet_miss) atomicAdd(&t.counters[CTR_GRAD_MISS], 1u);  // gradient_id_miss_count (PS mod.rs:401-403)
      } else {
        float* stage = a.vw_stage ? a.vw_stage + (size_t)j * t.dim : nullptr;
        step_segment<VEC, G>(t, op, hy, gr, slot, row, lane, stage, [&](uint32_t c, float (&acc)[VEC]) {
          reduce_piece<VEC, F16>(acc, a, t, sl, gr, slot, j, e, c, hd.w);
        });
      }
    } else if (live && row < t.capacity) {  // a miss is counted once by k_combine_update
      float* dst = partial_slot(a, t, j);
      const uint32_t nvec = t.dim / VEC;
      for (uint32_t c = lane; c < nvec; c += G) {
        float acc[VEC];
        reduce_piece<VEC, F16>(acc, a, t, sl, gr, slot, j, e, c, hd.w);
        store_vec<VEC>(dst + c * VEC, acc);
      }
    }
    w = __shfl_sync(gmask, next, wl / G * G);
  }
}

// One block per cut segment (owner list).  The segment's partial sums (first piece, then one per boundary)
// are split into contiguous ranges, one per lane group; every group adds its range in position order, then
// group 0 adds the group sums in order and performs the step.  A fixed association, so deterministic.
template <int VEC, int G>
__global__ void __launch_bounds__(256) k_combine_update(TableDev t, OptimDev op, HyperDev hy, GradsDev gr, SegArgs a,
                                                        const uint2* __restrict__ owners,
                                                        const uint32_t* __restrict__ counts, uint32_t n_groups_used) {
  extern __shared__ float gsum[];  // n_groups_used x dim
  const uint32_t lane = threadIdx.x % G;
  const uint32_t grp = threadIdx.x / G;
  const uint32_t n_own = counts[1];
  const uint32_t tick = *a.tick_ptr;
  const uint32_t nvec = t.dim / VEC;
  for (uint32_t idx = blockIdx.x; idx < n_own; idx += gridDim.x) {
    const uint32_t b = owners[idx].x;
    const uint32_t j0 = owners[idx].y;  // segment start: fewer than PIECE positions before its first boundary
    const uint32_t key = a.skey[b];
    const uint32_t slot = val_slot(a.sval[b]);
    if (!gr.ptr[slot] || a.nan_tick[slot] == tick) continue;
    // last boundary of the segment: the list is sorted by (leader, slot), so binary-search the boundaries
    uint32_t lo = b / a.piece + 1, hi = (a.n - 1) / a.piece;  // boundary numbers; lo is known to be inside
    while (lo < hi) {
      uint32_t mid = (lo + hi + 1) >> 1;
      uint32_t p = mid * a.piece;
      uint32_t k2 = a.skey[p], s2 = val_slot(a.sval[p]);
      bool inside = (k2 < key) || (k2 == key && s2 <= slot);  // positions > b never sort before the segment
      if (inside) lo = mid; else hi = mid - 1;
    }
    const uint32_t q_last = lo * a.piece;
    if (a.shared_groups) {  // a run holding several slots belongs to k_update_shared
      uint32_t e = q_last + 1;
      while (e < a.n && same_seg(a, e, key, slot)) ++e;
      if ((j0 > 0 && a.skey[j0 - 1] == key) || (e < a.n && a.skey[e] == key)) continue;
    }
    const uint32_t row = key < a.n ? a.occ_row[key] : ROW_NONE;
    if (row >= t.capacity) {
      if (threadIdx.x == 0 && !a.quiet_miss) atomicAdd(&t.counters[CTR_GRAD_MISS], 1u);
      continue;
    }
    const uint32_t q0 = (j0 < b) ? b : b + a.piece;   // first boundary after the first piece
    const uint32_t P = 1 + (q_last - q0) / a.piece + 1;  // first piece + one partial per boundary q0..q_last
    float* first = partial_slot(a, t, j0);
    const uint32_t per = (P + n_groups_used - 1) / n_groups_used;
    if (grp < n_groups_used) {
      const uint32_t k0 = grp * per, k1 = min(P, k0 + per);
      for (uint32_t c = lane; c < nvec; c += G) {
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
        for (uint32_t k = k0; k < k1; k += 4) {
          float p[4][VEC];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < k1) {
              const float* src = (k + u == 0) ? first : a.partials + (size_t)2 * ((q0 + (k + u - 1) * a.piece) / a.piece) * t.dim;
              load_vec<VEC>(src + c * VEC, p[u]);
            }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < k1) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc[e] = (k + u == k0) ? p[u][e] : __fadd_rn(acc[e], p[u][e]);
            }
        }
        store_vec<VEC>(gsum + (size_t)grp * t.dim + c * VEC, acc);
      }
    }
    __syncthreads();
    if (grp == 0) {
      const uint32_t used = (P + per - 1) / per;  // groups that had a non-empty range
      step_segment<VEC, G>(t, op, hy, gr, slot, row, lane, first, [&](uint32_t c, float (&acc)[VEC]) {
        load_vec<VEC>(gsum + c * VEC, acc);
        for (uint32_t g2 = 1; g2 < used; ++g2) {
          float p[VEC];
          load_vec<VEC>(gsum + (size_t)g2 * t.dim + c * VEC, p);
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[e] = __fadd_rn(acc[e], p[e]);
        }
      });
    }
    __syncthreads();
  }
}

template <int VEC, int G, bool F16>
__global__ void __launch_bounds__(256) k_update_shared(TableDev t, OptimDev op, HyperDev hy, SlotsDev sl, GradsDev gr,
                                                       SegArgs a) {
  uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  uint32_t lane = threadIdx.x % G;
  if (j >= a.n) return;
  uint32_t key = a.skey[j];
  if (j > 0 && a.skey[j - 1] == key) return;  // not the head of its sign's run
  uint32_t slot0 = val_slot(a.sval[j]);
  uint32_t end = j + 1;
  bool multi = false;
  while (end < a.n && a.skey[end] == key) {
    multi |= val_slot(a.sval[end]) != slot0;
    ++end;
  }
  if (!multi) return;  // single-slot runs were handled by the two kernels above
  uint32_t row = key < a.n ? a.occ_row[key] : ROW_NONE;  // the sort key is the sign's first occurrence
  uint32_t j0 = j;
  while (j0 < end) {
    uint32_t slot = val_slot(a.sval[j0]);
    uint32_t j1 = j0 + 1;
    while (j1 < end && val_slot(a.sval[j1]) == slot) ++j1;
    bool active = gr.ptr[slot] && a.nan_tick[slot] != *a.tick_ptr;
    if (active && row >= t.capacity) {
      if (lane == 0) atomicAdd(&t.counters[CTR_GRAD_MISS], 1u);
      active = false;
    }
    if (active) {
      float* stage = a.vw_stage ? a.vw_stage + (size_t)j0 * t.dim : nullptr;
      step_segment<VEC, G>(t, op, hy, gr, slot, row, lane, stage,
                           [&](uint32_t c, float (&acc)[VEC]) { reduce_piece<VEC, F16>(acc, a, t, sl, gr, slot, j0, j1, c, a.sval[j0]); });
    }
    j0 = j1;
  }
}

// pb_update: distinct signs with explicit f32 gradients (update_gradient_mixed, PS mod.rs:359-427).
template <int VEC, int G>
__global__ void __launch_bounds__(256) k_update_direct(TableDev t, OptimDev op, HyperDev hy,
                                                       const uint32_t* __restrict__ occ_cell,
                                                       const float* __restrict__ grads, uint32_t n,
                                                       const float* __restrict__ adam_pair,
                                                       const uint32_t* __restrict__ n_ptr,
                                                       const uint32_t* __restrict__ tick_ptr,
                                                       const uint32_t* __restrict__ nan_tick) {
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  uint32_t lane = threadIdx.x % G;
  if (n_ptr) n = *n_ptr;
  if (gid >= n) return;
  if (nan_tick && nan_tick[0] == *tick_ptr) return;  // NaN rule: the whole gradient is dropped
  uint32_t h = occ_cell[gid];
  uint32_t row = (h < t.n_cells + N_SPECIAL) ? t.cells[h].row : ROW_NONE;
  if (row >= t.capacity) {
    if (lane == 0) atomicAdd(&t.counters[CTR_GRAD_MISS], 1u);
    return;
  }
  float* prow = t.rows + (size_t)row * t.stride;
  const uint32_t nvec = t.dim / VEC;
  float vw_state = (op.kind == PB_OPT_ADAGRAD_VW) ? prow[t.dim] : 0.0f;
  float r1 = 0.0f, r2 = 0.0f;
  if (op.kind == PB_OPT_ADAM) {
    r1 = __fdiv_rn(1.0f, __fsub_rn(1.0f, adam_pair[0]));
    r2 = __fdiv_rn(1.0f, __fsub_rn(1.0f, adam_pair[1]));
  }
  const float* g0 = grads + (size_t)gid * t.dim;
  for (uint32_t c = lane; c < nvec; c += G) {
    float g[VEC];
    load_vec<VEC>(g0 + c * VEC, g);
    apply_chunk<VEC>(prow, c, g, t, op, hy, vw_state, r1, r2);
  }
  if (op.kind == PB_OPT_ADAGRAD_VW && lane == 0) {
    float gs = __fdiv_rn(vw_dot(g0, t.dim), (float)t.dim);
    prow[t.dim] = __fadd_rn(__fmul_rn(vw_state, op.mom), gs);
  }
}

// ------------------------------------------------------------------------------------------------
// launchers (host)
// ------------------------------------------------------------------------------------------------
void launch_nan_scan(const GradsDev& gr, uint32_t n_slots, uint32_t elems_per_slot, bool f16, const uint32_t* tick,
                     uint32_t* nan_tick, int32_t* status, cudaStream_t st) {
  uint32_t per = f16 ? elems_per_slot / 8 : elems_per_slot;
  uint32_t gx = cdiv(per ? per : 1, 256 * 4);
  if (gx > 148 * 4) gx = 148 * 4;
  dim3 grid(gx, n_slots);
  if (f16) PB_LAUNCH_F(FAM_NAN, k_nan_scan<true>, grid, 256, 0, st, gr, elems_per_slot, tick, nan_tick);
  else PB_LAUNCH_F(FAM_NAN, k_nan_scan<false>, grid, 256, 0, st, gr, elems_per_slot, tick, nan_tick);
  if (status) PB_LAUNCH(k_slot_status, 1, PB_MAX_SLOTS, 0, st, gr, n_slots, tick, nan_tick, status);
}

// Adam's batch-level state (optim.rs:99-131, 155-197): one (beta1^t, beta2^t) pair per feature group, kept on the
// device so that a captured backward advances it on every replay.
__global__ void k_adam_fill(float* pow, float b1, float b2) {
  uint32_t i = threadIdx.x;
  if (i < PB_ADAM_KEYS) {
    pow[2 * i] = b1;
    pow[2 * i + 1] = b2;
  }
}
__global__ void k_adam_advance(float* pow, AdamKeys keys, float b1, float b2) {
  uint32_t i = threadIdx.x;
  if (i < keys.n) {
    float* p = pow + 2u * keys.idx[i];
    p[0] = __fmul_rn(p[0], b1);
    p[1] = __fmul_rn(p[1], b2);
  }
}
void launch_adam_fill(float* pow, float b1, float b2, cudaStream_t st) { PB_LAUNCH(k_adam_fill, 1, PB_ADAM_KEYS, 0, st, pow, b1, b2); }
void launch_adam_advance(float* pow, const AdamKeys& keys, float b1, float b2, cudaStream_t st) {
  if (keys.n) PB_LAUNCH(k_adam_advance, 1, PB_MAX_SLOTS, 0, st, pow, keys, b1, b2);
}

void launch_slot_status(const GradsDev& gr, uint32_t n_slots, const uint32_t* tick, const uint32_t* nan_tick,
                        int32_t* status, cudaStream_t st) {
  PB_LAUNCH(k_slot_status, 1, PB_MAX_SLOTS, 0, st, gr, n_slots, tick, nan_tick, status);
}

template <int VEC, bool F16>
static void reduce_dispatch(int G, const TableDev& t, const OptimDev& op, const HyperDev& hy, const SlotsDev& sl,
                            const GradsDev& gr, const SegArgs& a, const uint4* heads, const uint2* owners,
                            uint32_t* counts, cudaStream_t st) {
  // persistent-style grids: the list lengths live on the device
  const uint32_t full = cdiv((uint64_t)a.n * G, 256);
  static const uint32_t tune_grid = getenv("PB_REDUCE_GRID") ? (uint32_t)atoi(getenv("PB_REDUCE_GRID")) : 148u * PB_REDUCE_BLOCKS;
  const uint32_t grid = full < tune_grid ? full : tune_grid;
  const uint32_t n_bound = a.piece ? cdiv(a.n, a.piece) : 0;
  const uint32_t gridc = n_bound < 148u * 2u ? (n_bound ? n_bound : 1) : 148u * 2u;  // one block per cut segment
  uint32_t ng = 256u / (uint32_t)G;
  while (ng > 1 && (size_t)ng * t.dim * sizeof(float) > 40 * 1024) ng >>= 1;
  const size_t smemc = (size_t)ng * t.dim * sizeof(float);
#define PB_G(GG)                                                                                                    \
  case GG:                                                                                                          \
    PB_LAUNCH_F(FAM_UPDATE, (k_reduce_update<VEC, GG, F16>), grid, 256, 0, st, t, op, hy, sl, gr, a, heads, counts); \
    if (n_bound > 1) PB_LAUNCH_F(FAM_COMBINE, (k_combine_update<VEC, GG>), gridc, 256, smemc, st, t, op, hy, gr, a, owners, counts, ng); \
    if (a.shared_groups) PB_LAUNCH_F(FAM_UPDATE, (k_update_shared<VEC, GG, F16>), full, 256, 0, st, t, op, hy, sl, gr, a);  \
    break;
  switch (G) {
    PB_G(1) PB_G(2) PB_G(4) PB_G(8) PB_G(16) PB_G(32)
  }
#undef PB_G
}

void launch_reduce_update(const TableDev& t, const OptimDev& op, const HyperDev& hy, const SlotsDev& sl,
                          const GradsDev& gr, bool f16, const SegArgs& a, uint4* heads, uint2* owners,
                          uint32_t* counts, cudaStream_t st) {
  if (!a.n) return;
  int vec, G;
  vec_group(t.dim, vec, G);
  if (vec == 4) {
    if (f16) reduce_dispatch<4, true>(G, t, op, hy, sl, gr, a, heads, owners, counts, st);
    else reduce_dispatch<4, false>(G, t, op, hy, sl, gr, a, heads, owners, counts, st);
  } else {
    if (f16) reduce_dispatch<1, true>(G, t, op, hy, sl, gr, a, heads, owners, counts, st);
    else reduce_dispatch<1, false>(G, t, op, hy, sl, gr, a, heads, owners, counts, st);
  }
}

void launch_find_heads(const SegArgs& a, uint4* heads, uint2* owners, uint32_t* counts, cudaStream_t st) {
  if (a.n) PB_LAUNCH_F(FAM_SORT, k_find_heads, cdiv((uint64_t)cdiv(a.n, 32) * 32, 256), 256, 0, st, a, heads, owners, counts);
}

void launch_update_direct(const TableDev& t, const OptimDev& op, const HyperDev& hy, const uint32_t* occ_cell,
                          const float* grads, uint32_t n, const float* adam_pair, cudaStream_t st,
                          const uint32_t* n_ptr, const uint32_t* tick, const uint32_t* nan_tick) {
  if (!n) return;
  int vec, G;
  vec_group(t.dim, vec, G);
  uint32_t grid = cdiv((uint64_t)n * G, 256);
#define PB_U(V, GG)                                                                                          \
  if (vec == V && G == GG)                                                                                   \
    PB_LAUNCH_F(FAM_UPDATE, (k_update_direct<V, GG>), grid, 256, 0, st, t, op, hy, occ_cell, grads, n, adam_pair, n_ptr, \
                tick, nan_tick);
  PB_U(4, 1) PB_U(4, 2) PB_U(4, 4) PB_U(4, 8) PB_U(4, 16) PB_U(4, 32)
  PB_U(1, 1) PB_U(1, 2) PB_U(1, 4) PB_U(1, 8) PB_U(1, 16) PB_U(1, 32)
#undef PB_U
}

}  // namespace pb
