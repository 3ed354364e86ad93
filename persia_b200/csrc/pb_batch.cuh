// pb_batch.cuh — bookkeeping of a batch's work lists, shared by the single-GPU forward (pb_dedup.cu) and the sharded
// requester (pb_shard.cu) (internal).
#pragma once
#include "pb_device.cuh"

namespace pb {

// block-level list bookkeeping shared by k_probe_items and k_route_items (pb_shard.cu): every thread of the block
// calls it once, converged.  `head`: this thread speaks for an item; cnt its multiplicity (0 = keep nothing);
// returns through the references the item's list position and the base of its occurrence list.
// One global atomic per block and list instead of one per warp (they all hit the same few words).
struct ItemSlots {
  uint32_t cls;   // 0 none, 1 cold, 2 warm, 3 hot
  uint32_t pos;   // position in the class's list
  uint32_t base;  // first entry of the occurrence list (cnt > 1)
};
__device__ __forceinline__ ItemSlots block_item_slots(const BatchDev& b, bool head, uint32_t cnt) {
  __shared__ uint32_t s_n[4], s_g[4];  // cold, warm, hot, occurrence-list entries
  if (threadIdx.x < 4) s_n[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  ItemSlots r;
  r.cls = !head || cnt == 0 ? 0u : (cnt == 1 ? 1u : (cnt <= PB_WARM_MAX ? 2u : 3u));
  r.pos = r.base = 0;
  uint32_t off = 0, seg = 0;
#pragma unroll
  for (uint32_t c = 1; c <= 3; ++c) {
    const uint32_t m = __ballot_sync(0xffffffffu, r.cls == c);
    uint32_t w = 0;
    if (m && lane == 0) w = atomicAdd(&s_n[c - 1], (uint32_t)__popc(m));
    w = __shfl_sync(0xffffffffu, w, 0);
    if (r.cls == c) off = w + __popc(m & ((1u << lane) - 1u));
  }
  if (r.cls >= 2) seg = atomicAdd(&s_n[3], cnt);
  __syncthreads();
  if (threadIdx.x < 4 && s_n[threadIdx.x]) {
    const uint32_t which = threadIdx.x == 0 ? BC_COLD : threadIdx.x == 1 ? BC_WARM : threadIdx.x == 2 ? BC_HOT : BC_SEG;
    s_g[threadIdx.x] = atomicAdd(&b.cnt[which], s_n[threadIdx.x]);
  }
  __syncthreads();
  if (r.cls) r.pos = s_g[r.cls - 1] + off;
  if (r.cls >= 2) r.base = s_g[3] + seg;
  return r;
}

}  // namespace pb
