// Peer-memory exchange of the per-shard best records (multi-GPU step of the evaluation cycle).
//
// Every rank owns one mailbox (PeerBox) in its own HBM; the other ranks of the box map it over NVLink / NVSwitch (CUDA IPC
// between processes, plain peer access inside one process). The argmax kernel's last block publishes the shard winner with
// three tagged 8-byte stores INTO EVERY PEER'S mailbox, then polls its own mailbox (local HBM reads) until
// every rank's record of this step has arrived and merges them — no collective library on the data path, no host round trip,
// no extra launch. Four slots (step % 4) make the pipelined form safe, in which a step's records are collected one call later
// (see peer_exchange_block).
#ifndef SIPP_PEER_CUH_
#define SIPP_PEER_CUH_

#include <math_constants.h>

#include "sipp_internal.h"

__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void peer_st_u64(void* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_ld_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by ALL threads of one block (blockDim.x >= SIPP_PEER_MAX_WORLD). `mine` must hold the same record in every thread.
// mode PEER_SYNC:      publish `mine` as step s, wait for every rank's record of step s, merge -> *out.
// mode PEER_PIPELINED: publish `mine` as step s, then collect step s-1 (published one call ago, so the records are normally
//                      already here: the exchange latency hides under the caller's next kernel) -> *out; the very first call
//                      has nothing to collect and writes {0, -1}. peer_flush (PEER_COLLECT_ONLY) collects the last step.
// Thread r < world writes into rank r's mailbox and waits for rank r's record in the own mailbox; thread 0 merges with the
// semantics of k_merge_best (largest value, lowest global index on ties, NaN / empty shards never win; if nothing is
// comparable rank 0's record is returned). On a timeout *out = {NaN, -3} and the mailbox status is set.
// Slot reuse: a rank publishes step s only after it collected s-2, i.e. after EVERY rank published s-2, and a rank publishes
// s-2 only after it collected s-4 — so nobody still needs the slot of step s-4 that step s overwrites.
enum { PEER_SYNC = 0, PEER_PIPELINED = 1, PEER_COLLECT_ONLY = 2 };

// The mailbox table travels BY VALUE in the kernel parameters and the two counters are read by thread 0 when the kernel
// starts (peer_preload; nothing else writes them while the kernel runs): the tail of the kernel then has no chain of
// dependent global loads in front of its stores (three of them cost ~2 us per step).
struct PeerCounters { unsigned long long published, collected; };
__device__ __forceinline__ PeerCounters peer_preload(const PeerDev& pd) {
  PeerCounters c;
  const PeerBox* own = pd.box[pd.rank];
  c.published = own->published;
  c.collected = own->collected;
  return c;
}

__device__ __forceinline__ void peer_exchange_block(const PeerDev& pd, BestRec mine, BestRec* out, int mode, PeerCounters cnt /* thread 0's */) {
  __shared__ BestRec s_rec[SIPP_PEER_MAX_WORLD];
  __shared__ int s_fail;
  __shared__ PeerCounters s_cnt;
  const int t = threadIdx.x;
  const int rank = pd.rank, world = pd.world;
  PeerBox* own = pd.box[rank];
  if (t == 0) { s_fail = 0; s_cnt = cnt; }
  __syncthreads();
  const unsigned long long pub = s_cnt.published + (mode == PEER_COLLECT_ONLY ? 0 : 1);   // after this call
  const unsigned long long col = s_cnt.collected + 1;                                      // the step to collect
  const bool do_collect = col <= pub && !(mode == PEER_PIPELINED && col == pub);
  if (t < world) {
    if (mode != PEER_COLLECT_ONLY) {
      PeerSlot* dst = &pd.box[t]->slots[pub % kPeerSlots][rank];
      const unsigned long long vb = (unsigned long long)__double_as_longlong(mine.value), tag = (pub & 0xFFFFFFFFull) << 32;
      peer_st_u64(&dst->w0, tag | (vb & 0xFFFFFFFFull));
      peer_st_u64(&dst->w1, tag | (vb >> 32));
      peer_st_u64(&dst->w2, ((pub & 0xFFFFull) << 48) | ((unsigned long long)mine.index & 0xFFFFFFFFFFFFull));
    }
    if (do_collect) {
      const PeerSlot* src = &own->slots[col % kPeerSlots][t];
      const unsigned long long t0 = peer_now_ns();
      unsigned spins = 0;
      bool ok = true;
      unsigned long long a, b, c;
      for (;;) {
        a = peer_ld_u64(&src->w0); b = peer_ld_u64(&src->w1); c = peer_ld_u64(&src->w2);
        if ((a >> 32) == (col & 0xFFFFFFFFull) && (b >> 32) == (col & 0xFFFFFFFFull) && (c >> 48) == (col & 0xFFFFull)) break;
        if (++spins > 64) {
          __nanosleep(100);
          if (peer_now_ns() - t0 > pd.timeout_ns) { ok = false; break; }
        }
      }
      if (ok) {
        s_rec[t].value = __longlong_as_double((long long)((b << 32) | (a & 0xFFFFFFFFull)));
        s_rec[t].index = ((long long)(c << 16)) >> 16;      // sign-extend the 48-bit index
      } else {
        atomicExch(&s_fail, 1);
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    if (!do_collect) {
      out->value = 0.0;
      out->index = -1;
    } else if (s_fail) {
      out->value = CUDART_NAN;
      out->index = -3;
      own->status = 1;
    } else {
      BestRec best = s_rec[0];
      bool have = best.index >= 0 && !(best.value != best.value);
      for (int r = 1; r < world; ++r) {
        const BestRec c = s_rec[r];
        if (c.index < 0 || c.value != c.value) continue;
        if (!have || c.value > best.value || (c.value == best.value && c.index < best.index)) { best = c; have = true; }
      }
      *out = best;
    }
    own->published = pub;
    if (do_collect) own->collected = col;
  }
}

#endif  // SIPP_PEER_CUH_
