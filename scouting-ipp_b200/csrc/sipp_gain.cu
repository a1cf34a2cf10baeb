// Batched footprint-gain kernels (hot-path rows A1-A8).
//
// One warp per candidate view. The candidate's (2h+1)^2 window of the packed one-byte cell records is brought
// into shared memory by a single TMA tile load (cp.async.bulk.tensor.2d, zero fill outside the map — exactly
// the reference's "skip off-map voxels", map_planexp.cpp:97), three tiles in flight per warp; lanes then read
// the tile with 128-bit shared loads, accumulate per-byte-lane SIMD counters and finish with one warp
// redux. No tensor cores: nothing here is a contraction.
//
// Reference semantics (all five evaluators erase the observed voxels, then reduce over the unknown ones):
//   footprint    MapPlanExp::getVoxelArea                advanced_planner_modules/src/map/map_planexp.cpp:81-101
//   RRTStar      computeGainFromVisibleVoxels            src/trajectory_evaluator/rrt_star_evaluator.cpp:36-96
//   GoalDistance                                         src/trajectory_evaluator/goal_distance_evaluator.cpp:33-66
//   ExpandReach.                                         src/trajectory_evaluator/expand_reachable_area_evaluator.cpp:33-67
//   AvgViewCost                                          src/trajectory_evaluator/average_view_cost_evaluator.cpp:31-61
//   ExpandLowCost                                        src/trajectory_evaluator/expand_low_cost_area_evaluator.cpp:33-72
// The centre cell is emitted twice by the reference (once with centre coordinates, once with corner
// coordinates); lane 0 adds the duplicate with the reference's own double-precision index math.
#include <cstdio>
#include <cstring>
#include <mutex>

#include "sipp_argmax.cuh"
#include "sipp_internal.h"
#include "sipp_peer.cuh"

namespace {

constexpr int kMaxWarpsPerCta = 16;
constexpr int kMaxStages = 8;       // tile slots per warp (runtime choice, GainArgs::stages)

// evaluator "reduction modes"
enum { MODE_COUNT = 0, MODE_RRT_DISCOUNT = 1, MODE_GOAL_DISTANCE = 2, MODE_AVG_VIEW_COST = 3, MODE_LOW_COST = 4 };

struct GainArgs {
  GainGeom g;
  int evaluator;
  int half_fov;
  int box_w;          // bytes per tile row of the rec tile (multiple of 16)
  int box_w16;        // elements per tile row of the uint16 aux tile (multiple of 8)
  int slot_bytes;     // bytes per pipeline slot (rec tile [+ aux tile]), multiple of 128
  int stages;         // tile slots per warp: stages - 1 TMA loads in flight while one tile is scanned
  int aux_off;        // byte offset of the aux tile inside a slot
  double npg;         // non_path_voxel_gain
  int use_discount, do_binary;
  double discount_offset, max_gain, inv_max_gain, discount_factor;
  double inv_max_gain2_lo, inv_max_gain2_hi;   // (1/max_gain)^2 -+ a safety band: outside it the squared test decides
  int use_bv;
  double bv_x_min, bv_x_max, bv_y_min, bv_y_max;
  int n;
  const double* xy;
  const double* start_xy;
  const double* cost;     // optional: value = gain / cost fused in the epilogue (flat tree)
  double* gain;
  uint32_t* counts;
  uint8_t* terminal;
  double* value;
  const uint8_t* rec_plane;   // packed scan: the byte record plane, read once per candidate for the duplicated centre voxel
  size_t rec_pitch;
  const uint8_t* rec2_plane;  // packed scan: the 2-bit plane of the evaluator (path or reachability flavour)
  size_t rec2_pitch;
  const uint16_t* label_plane;  // AverageViewCost: map_cost_ (raw labels) for the duplicated centre voxel — the tile holds labels masked by the
  size_t label_pitch16;         // CORNER voxels' bounding-volume test, the centre voxel has its own test on centre coordinates
  double* value2;         // optional second copy of value (the caller's pinned host array, written over PCIe by the kernel)
  double* fsum_out;       // optional: the candidate's floating-point accumulator before the epilogue (several views per segment: k_gain_combine)
  const double* recip;    // ExpandLowCost lookup
  const int* sel;         // optional: candidate k of this launch is node sel[k] of the caller's arrays (outputs are scattered)
  const int* n_dev;       // optional: number of candidates, read on the device (a.n is then only the upper bound for the grid)
  // fused argmax (flat tree, value computed here): every lane keeps the best (value, index) of its candidates, warps / CTAs
  // reduce, the last CTA to finish merges the per-CTA records — and, with peers attached, exchanges the winner over NVLink
  BestRec* best_out;      // non-null: fused argmax on
  SelectWs* sel_ws;
  long long index_offset;
  PeerDev peer;           // by value: the mailbox pointers sit in the kernel's parameter space
  int peer_mode;          // < 0: single GPU
};

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// carry-save adder on 32 bit columns at once: h:l = a + b + c per bit (two LOP3)
#define SIPP_CSA(h, l, a, b, c) { const uint32_t u_ = (a) ^ (b); h = ((a) & (b)) | (u_ & (c)); l = u_ ^ (c); }
__device__ __forceinline__ uint32_t bytesum(uint32_t v) { return __dp4a(v, 0x01010101u, 0u); }
// mask with bytes [lo, hi) of a 4-byte word set, 0 <= lo, hi <= 4 after clamping
__device__ __forceinline__ uint32_t byte_range_mask(int lo, int hi) {
  lo = max(lo, 0);
  hi = min(hi, 4);
  if (hi <= lo) return 0u;
  return (0xFFFFFFFFu >> (8 * (4 - hi))) & (0xFFFFFFFFu << (8 * lo));
}
// mask with bits [lo, hi) of a 32-bit word set, after clamping to [0, 32]
__device__ __forceinline__ uint32_t bit_range_mask(int lo, int hi) {
  lo = max(lo, 0);
  hi = min(hi, 32);
  if (hi <= lo) return 0u;
  const uint32_t upto_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
  return upto_hi & (0xFFFFFFFFu << lo);
}
// same for the two halfwords of a word, 0 <= lo, hi <= 2
__device__ __forceinline__ uint32_t half_range_mask(int lo, int hi) {
  lo = max(lo, 0);
  hi = min(hi, 2);
  if (hi <= lo) return 0u;
  return (0xFFFFFFFFu >> (16 * (2 - hi))) & (0xFFFFFFFFu << (16 * lo));
}

// ---- per-candidate window geometry (the reference's double-precision index math, restated) -------------
struct Window {
  int cid;           // index of the candidate in the caller's arrays
  int on_map;        // getVoxelCenter returned true (map_planexp.cpp:78,87)
  int cx, cy;        // centre cell of the window (center_voxel_loc, :89)
  double vcx, vcy;   // centre voxel coordinates ((loc + 0.5) * vs, :74-76)
  uint32_t crec;     // packed scan: byte record of the window's centre cell (the tile holds 2-bit records only)
};

__device__ __forceinline__ int trunc_to_int(double v) {
  // (int) cast of the reference; out-of-range / NaN inputs are UB there, we map them to a far off-map index
  if (!(v > -1.0e9 && v < 1.0e9)) return -1000000;
  return (int)v;
}

__device__ __forceinline__ Window make_window(const GainGeom& g, double px, double py) {
  Window w;
  const int lx = trunc_to_int(px * g.inv_vs), ly = trunc_to_int(py * g.inv_vs);   // position32pixelLocation :136-138
  const double clx = lx + 0.5, cly = ly + 0.5;                                     // :74
  w.vcx = clx * g.vs;                                                              // pixelLocation2position :142-144
  w.vcy = cly * g.vs;
  w.on_map = (0 <= (int)clx) && (0 <= (int)cly) && (g.W > (int)clx) && (g.H > (int)cly) && lx > -1000000 && ly > -1000000;  // isOnMap2 :222-224
  w.cx = trunc_to_int(w.vcx * g.inv_vs);                                           // :89
  w.cy = trunc_to_int(w.vcy * g.inv_vs);
  return w;
}

// number of window cells (corner voxels) that are on the map and inside the bounding-volume index range
__device__ __forceinline__ int clipped_extent(int c, int h, int lo, int hi) {
  const int a = max(c - h, lo), b = min(c + h, hi);
  return max(b - a + 1, 0);
}

// Work split: warp gw owns the CONTIGUOUS candidate range [gw*n/nw, (gw+1)*n/nw) and walks it in batches of up to 32.
// Per batch, lane j does the per-candidate scalar work of candidate j — the reference's double-precision index math before
// the scan, the duplicated centre voxel and the evaluator epilogue after it — so that work runs 32-wide instead of serially
// on lane 0, and pose loads / result stores are coalesced. Between the two, the warp scans the batch's tiles one after the
// other; lane j issues the TMA load of its own candidate `stages - 1` scans ahead (the next batch's windows are computed one
// batch early so the pipeline never drains).
//
// Lane layout over a tile whose rows are `ncr` 16-byte chunks wide: lane = lrow * ncr + lcc for the first
// rpi * ncr lanes (rpi = 32 / ncr rows are covered per iteration), the remaining lanes idle. A lane keeps its chunk
// column for the whole scan, so the column masks of the (unaligned) window are computed once per candidate.
// PACKED (count-type evaluators): the window is read from the 2-bit-per-cell plane (bit 0 UNKNOWN, bit 1 UNKNOWN & path /
// reachable; a 65-cell window row is 17 bytes instead of 65) with plain read-only loads straight into registers — no TMA, no
// shared-memory tile, no mbarrier. Measured: a cp.async.bulk.tensor costs the SM ~130 cycles of TMA-unit time however small the
// box (tools/gain_rows_probe.py: 30 us for 65536 candidates at 17 rows, 45 us at 65), which bounds the tiled form at one
// candidate per ~200 cycles per SM; the direct form has no per-candidate fixed cost beyond its own instructions and runs two
// CTAs per SM (no dynamic shared memory).
template <int MODE, bool PACKED>
__device__ __forceinline__ void gain_body(const CUtensorMap* tm_rec_p, const CUtensorMap* tm_aux_p, const GainArgs& a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[kMaxWarpsPerCta * kMaxStages];
  __shared__ uint4 mask_tbl[16 * 16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kStages = a.stages;
  const uint32_t slot_base = smem_u32(smem_raw) + (uint32_t)(warp * kStages) * (uint32_t)a.slot_bytes;
  const uint32_t bar_base = smem_u32(&bars[warp * kStages]);

  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar_base + 8u * s, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int wpc = blockDim.x >> 5;
  const int gw = blockIdx.x * wpc + warp;
  const int nw = gridDim.x * wpc;
  const int n_cand = a.n_dev ? min(a.n, *a.n_dev) : a.n;
  const int c0 = (int)(((long long)gw * n_cand) / nw), c1 = (int)(((long long)(gw + 1) * n_cand) / nw);
  const int total = c1 - c0;
  const int h = a.half_fov;
  const int side = 2 * h + 1;
  constexpr bool kNeedRec = (MODE != MODE_LOW_COST);
  constexpr bool kNeedAux = (MODE == MODE_AVG_VIEW_COST || MODE == MODE_LOW_COST);
  constexpr bool kFloat = (MODE == MODE_RRT_DISCOUNT || MODE == MODE_GOAL_DISTANCE || MODE == MODE_LOW_COST);
  const uint32_t tx_bytes = (kNeedRec ? (uint32_t)(a.box_w * side) : 0u) + (kNeedAux ? (uint32_t)(a.box_w16 * 2 * side) : 0u);

  // lane geometry (fixed for the whole kernel)
  const int ncr = a.box_w >> 4, rpi = 32 / ncr;
  const int lrow = lane / ncr, lcc = lane - lrow * ncr;
  const bool lane_on = lane < rpi * ncr;
  const int ncr16 = a.box_w16 >> 3, rpi16 = 32 / ncr16;
  const int lrow16 = lane / ncr16, lcc16 = lane - lrow16 * ncr16;
  const bool lane_on16 = lane < rpi16 * ncr16;
  // which record bit feeds counts[AUX]
  const uint32_t auxbit = (MODE == MODE_AVG_VIEW_COST) ? REC_OBS_FREE
                          : (MODE == MODE_COUNT && a.evaluator == SIPP_EVAL_EXPAND_REACHABLE) ? REC_UNK_REACH : REC_UNK_PATH;
  const uint32_t auxmask = auxbit * 0x01010101u;
  const int auxshift = __ffs(auxbit) - 1;
  // per-byte-lane SIMD counters take 4 * auxbit per iteration at most: iterations between flushes
  const int flush_iters = 255 / (4 * (int)auxbit);
  const int my_iters = (lane_on && lrow < side) ? (side - lrow + rpi - 1) / rpi : 0;
  // column masks of the unaligned window, tabulated once per CTA: entry [r0 * ncr + chunk] holds the four byte masks of
  // 16-byte chunk `chunk` when the window starts at byte r0 of the tile row (r0 = (cx - h) & 15)
  uint32_t* mask32 = reinterpret_cast<uint32_t*>(mask_tbl);
  if (PACKED) {
    // entry [r0 * wpr + word]: window bits of 32-bit word `word` of a box row (16 cells each) when the window starts at cell r0
    // of the box (0..63)
    const int wpr_ = a.box_w >> 2;
    for (int e = threadIdx.x; e < 64 * wpr_; e += blockDim.x) {
      const int r0 = e / wpr_, wd = e - r0 * wpr_;
      mask32[e] = bit_range_mask(2 * r0 - 32 * wd, 2 * (r0 + side) - 32 * wd);
    }
  } else
  for (int e = threadIdx.x; e < 16 * ncr; e += blockDim.x) {
    const int r0 = e / ncr, cc = e - r0 * ncr;
    const int lo = r0 - 16 * cc, hi = r0 + side - 16 * cc;   // valid bytes of the chunk: [lo, hi) clipped to [0,16)
    mask_tbl[e] = make_uint4(byte_range_mask(lo, hi), byte_range_mask(lo - 4, hi - 4), byte_range_mask(lo - 8, hi - 8),
                             byte_range_mask(lo - 12, hi - 12));
  }
  __syncthreads();

  // packed scan: lane = (row slot g, word column k) of the box; constants of the whole kernel
  const int pk_wpr = PACKED ? (a.box_w >> 2) : 8, pk_rpp = 32 / pk_wpr;
  const int pk_g = lane / pk_wpr, pk_k = lane - pk_g * pk_wpr;
  const int pk_nrows = (PACKED && pk_g < pk_rpp) ? (side - pk_g + pk_rpp - 1) / pk_rpp : 0;     // rows pk_g, pk_g + rpp, ... < side
  const int pk_span = pk_rpp * (pk_nrows > 0 ? pk_nrows - 1 : 0);                              // last row offset of this lane
  const uint32_t pk_pitch = (uint32_t)a.rec2_pitch;                    // the plane is far below 4 GiB: 32-bit offsets, one IMAD.WIDE per address
  const uint32_t pk_stride = (uint32_t)pk_rpp * pk_pitch;

  // ---- per-batch lane state: the window of candidate (batch base + lane)
  Window cur, nxt;
  auto load_window = [&](int idx) {   // idx = index inside this warp's range
    Window w;
    w.on_map = 0; w.cx = w.cy = -1000000; w.vcx = w.vcy = 0.0; w.cid = 0; w.crec = 0;
    if (idx < total) {
      const int cid = a.sel ? a.sel[c0 + idx] : c0 + idx;
      const double2 p = *reinterpret_cast<const double2*>(a.xy + 2 * (size_t)cid);
      w = make_window(a.g, p.x, p.y);
      w.cid = cid;
      w.crec = 0;
      if (PACKED && w.on_map && w.cx >= 0 && w.cx < a.g.W && w.cy >= 0 && w.cy < a.g.H)
        w.crec = __ldg(a.rec_plane + (size_t)w.cy * a.rec_pitch + w.cx);      // needed only in the epilogue, a batch later
      if (MODE == MODE_AVG_VIEW_COST && w.on_map && w.cx >= 0 && w.cx < a.g.W && w.cy >= 0 && w.cy < a.g.H)
        w.crec = __ldg(a.label_plane + (size_t)w.cy * a.label_pitch16 + w.cx);   // raw label of the centre cell (crec is otherwise unused on the tiled path)
    }
    return w;
  };
  auto issue = [&](const Window& w, int s) {   // executed by the one lane that owns the candidate
    const uint32_t bar = bar_base + 8u * s;
    const uint32_t dst = slot_base + (uint32_t)s * (uint32_t)a.slot_bytes;
    mbar_expect_tx(bar, tx_bytes);
    // TMA needs the box start 16-byte aligned in the contiguous dimension: start at the aligned column at or before
    // the window's first column; the box is wide enough to still cover the window (see gain_launch)
    if (kNeedRec) tma_load_2d(dst, tm_rec_p, bar, (w.cx - h) & ~15, w.cy - h);
    if (kNeedAux) tma_load_2d(dst + (uint32_t)a.aux_off, tm_aux_p, bar, (w.cx - h) & ~7, w.cy - h);
  };

  const bool fused = a.best_out != nullptr;     // every warp takes part in the CTA reduction of the fused argmax
  PeerCounters pcnt = {0, 0};                   // the exchange counters, fetched now so the tail does not wait for them
  if (fused && a.peer_mode >= 0 && threadIdx.x == 0) pcnt = peer_preload(a.peer);
  double best_v = -CUDART_INF;
  long long best_i = LLONG_MAX;
  if (total <= 0 && !fused) return;
  cur = load_window(lane);
  if (!PACKED)
    for (int k = 0; k < kStages - 1 && k < total; ++k)
      if (lane == k) issue(cur, k);   // kStages - 1 <= 7 < 32: all inside the first batch

  int s = 0, s_issue = kStages - 1;   // slot being scanned / slot refilled next (the one scanned in the previous iteration)
  uint32_t parity = 0;
  for (int base = 0; base < total; base += 32) {
    const int nb = min(32, total - base);
    nxt = load_window(base + 32 + lane);
    // per-candidate scan results, parked in the lane that owns the candidate
    uint32_t my_unk = 0, my_aux = 0, my_sum2 = 0, my_crec = 0, my_caux16 = 0;
    double my_fsum = 0.0;

    for (int j = 0; j < nb; ++j) {
      const int ahead = j + kStages - 1;
      if (!PACKED && base + ahead < total) {
        __syncwarp();   // every lane is done reading the slot that is about to be refilled
        if (ahead < 32) { if (lane == ahead) issue(cur, s_issue); }
        else            { if (lane == ahead - 32) issue(nxt, s_issue); }
      }
      const uint32_t tile = slot_base + (uint32_t)s * (uint32_t)a.slot_bytes;
      const int wcx = __shfl_sync(0xFFFFFFFFu, cur.cx, j), wcy = __shfl_sync(0xFFFFFFFFu, cur.cy, j);
      if (!PACKED)
        while (!mbar_try_wait(bar_base + 8u * s, parity)) {
        }

      // ---- window scan -------------------------------------------------------------------------------
      uint32_t n_unk = 0, n_aux = 0;       // per-lane totals
      uint32_t sum_aux2 = 0;               // AverageViewCost: sum of labels
      double fsum = 0.0;                   // floating-point modes
      if (PACKED) {
        // lane = (row slot g, word column k): the 32 lanes of one load instruction read rpp consecutive rows x wpr words (one
        // 16-byte-aligned run of 4 * wpr bytes per row). A lane sums ITS word column over its rows with
        // bit-sliced carry-save adders (Harley-Seal, blocks of eight rows): the window mask is the same for every row of a
        // column, so it is applied once to the counter planes, and eight rows cost 14 LOP3 + 2 POPC instead of 16 AND + 16 POPC.
        // All loads of a block are issued before the first use; rows / words outside the plane read as 0.
        const uint32_t m = mask32[((wcx - h) & 63) * pk_wpr + pk_k];
        const int colbyte = (((wcx - h) >> 2) & ~15) + 4 * pk_k;              // byte column of this lane's word (may be off the plane)
        const bool colok = colbyte >= 0 && colbyte + 4 <= (int)a.rec2_pitch;
        int nrows = (m != 0u && colok) ? pk_nrows : 0;
        int y = wcy - h + pk_g;                                               // this lane's rows: y, y + rpp, ...
        uint32_t ones = 0, twos = 0, fours = 0, c8b = 0, c8a = 0, sb = 0, sa = 0;
        if (y >= 0 && y + pk_span < a.g.H) {
          // every row is on the plane (all candidates except those within half_fov of the top / bottom edge): no per-load tests
          const uint8_t* ptr = a.rec2_plane + (size_t)((uint32_t)y * pk_pitch + (uint32_t)colbyte);
#define SIPP_LD(j) __ldg(reinterpret_cast<const uint32_t*>(ptr + (size_t)((uint32_t)(j) * pk_stride)))
#define SIPP_HS8(x0, x1, x2, x3, x4, x5, x6, x7)                                   \
  {                                                                                \
    uint32_t ta, tb, fa, fb, e8;                                                   \
    SIPP_CSA(ta, ones, ones, x0, x1);                                              \
    SIPP_CSA(tb, ones, ones, x2, x3);                                              \
    SIPP_CSA(fa, twos, twos, ta, tb);                                              \
    SIPP_CSA(ta, ones, ones, x4, x5);                                              \
    SIPP_CSA(tb, ones, ones, x6, x7);                                              \
    SIPP_CSA(fb, twos, twos, ta, tb);                                              \
    SIPP_CSA(e8, fours, fours, fa, fb);                                            \
    e8 &= m;                                                                       \
    c8b += __popc(e8);                                                             \
    c8a += __popc(e8 & 0xAAAAAAAAu);                                               \
  }
          // every load of the window is issued before the first use (one L2 round trip per candidate, not one per block): an odd
          // last row first, then blocks of sixteen / eight rows
          uint32_t xl = 0;
          if (nrows & 1) xl = __ldg(reinterpret_cast<const uint32_t*>(ptr + (size_t)((uint32_t)(nrows - 1) * pk_stride))) & m;
          nrows &= ~1;
          for (; nrows >= 16; nrows -= 16) {
            const uint32_t x0 = SIPP_LD(0), x1 = SIPP_LD(1), x2 = SIPP_LD(2), x3 = SIPP_LD(3), x4 = SIPP_LD(4), x5 = SIPP_LD(5), x6 = SIPP_LD(6),
                           x7 = SIPP_LD(7), x8 = SIPP_LD(8), x9 = SIPP_LD(9), xa = SIPP_LD(10), xb = SIPP_LD(11), xc = SIPP_LD(12), xd = SIPP_LD(13),
                           xe = SIPP_LD(14), xf = SIPP_LD(15);
            ptr += (size_t)(16u * pk_stride);
            SIPP_HS8(x0, x1, x2, x3, x4, x5, x6, x7)
            SIPP_HS8(x8, x9, xa, xb, xc, xd, xe, xf)
          }
          if (nrows >= 8) {
            const uint32_t x0 = SIPP_LD(0), x1 = SIPP_LD(1), x2 = SIPP_LD(2), x3 = SIPP_LD(3), x4 = SIPP_LD(4), x5 = SIPP_LD(5), x6 = SIPP_LD(6),
                           x7 = SIPP_LD(7);
            ptr += (size_t)(8u * pk_stride);
            nrows -= 8;
            SIPP_HS8(x0, x1, x2, x3, x4, x5, x6, x7)
          }
          for (; nrows > 0; --nrows) {
            const uint32_t x = SIPP_LD(0) & m;
            ptr += pk_stride;
            sb += __popc(x);
            sa += __popc(x & 0xAAAAAAAAu);
          }
          sb += __popc(xl);
          sa += __popc(xl & 0xAAAAAAAAu);
#undef SIPP_LD
#undef SIPP_HS8
        } else {
          // window cut by the top / bottom edge: rows off the plane read as 0
          const uint8_t* col = a.rec2_plane + colbyte;
          for (; nrows > 0; --nrows) {
            if (y >= 0 && y < a.g.H) {
              const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(col + (size_t)((uint32_t)y * pk_pitch))) & m;
              sb += __popc(x);
              sa += __popc(x & 0xAAAAAAAAu);
            }
            y += pk_rpp;
          }
        }
        ones &= m; twos &= m; fours &= m;
        const uint32_t both = __popc(ones) + 2u * __popc(twos) + 4u * __popc(fours) + 8u * c8b + sb;
        const uint32_t aux = __popc(ones & 0xAAAAAAAAu) + 2u * __popc(twos & 0xAAAAAAAAu) + 4u * __popc(fours & 0xAAAAAAAAu) + 8u * c8a + sa;
        n_unk = both - aux;                            // bit 1 (aux) is only ever set together with bit 0
        n_aux = aux;
      } else if (kNeedRec && lane_on) {
        const int r0 = (wcx - h) & 15;                 // the window occupies tile columns [r0, r0 + side)
        const uint4 m = mask_tbl[r0 * ncr + lcc];      // byte masks of this lane's 16-byte chunk column for that offset
        const uint32_t m0 = m.x, m1 = m.y, m2 = m.z, m3 = m.w;
        uint32_t addr = tile + (uint32_t)lane * 16u;
        const uint32_t step = (uint32_t)(rpi * ncr) * 16u;
        if (MODE == MODE_COUNT || MODE == MODE_AVG_VIEW_COST) {
          const uint32_t u0 = m0 & 0x01010101u, u1 = m1 & 0x01010101u, u2 = m2 & 0x01010101u, u3 = m3 & 0x01010101u;
          const uint32_t a0 = m0 & auxmask, a1 = m1 & auxmask, a2 = m2 & auxmask, a3 = m3 & auxmask;
          int left = my_iters;                           // rows lrow, lrow + rpi, ... < side (fixed per lane)
          while (left > 0) {
            // per-byte-lane SIMD counters, flushed before a byte lane can overflow
            uint32_t acc_unk = 0, acc_aux = 0;
            const int todo = min(left, flush_iters);
            left -= todo;
#pragma unroll 4
            for (int it = 0; it < todo; ++it) {
              const uint4 v = lds128(addr);
              addr += step;
              acc_unk += (v.x & u0) + (v.y & u1);
              acc_unk += (v.z & u2) + (v.w & u3);
              acc_aux += (v.x & a0) + (v.y & a1);
              acc_aux += (v.z & a2) + (v.w & a3);
            }
            n_unk += bytesum(acc_unk);
            n_aux += bytesum(acc_aux >> auxshift);    // every byte is a multiple of auxbit: the shift cannot mix byte lanes
          }
        } else {
          // floating-point modes: walk the selected bytes of each chunk (corner coordinates of the cells)
          const int x0 = ((wcx - h) & ~15) + lcc * 16;
          for (int row = lrow; row < side; row += rpi) {
            uint4 v = lds128(addr);
            addr += step;
            const uint32_t wds[4] = {v.x & m0, v.y & m1, v.z & m2, v.w & m3};
            const double vy = (double)(wcy - h + row) * a.g.vs;
            const double dy = a.g.goal_y - vy;
            const double dy2 = dy * dy;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t wd = wds[q];
              n_unk += __popc(wd & 0x01010101u);
              n_aux += __popc(wd & 0x02020202u);
              uint32_t sel = (MODE == MODE_RRT_DISCOUNT) ? (wd & 0x02020202u) : (wd & 0x01010101u);
              while (sel) {
                const int b = (__ffs(sel) - 1) >> 3;
                sel &= sel - 1;
                const double vx = (double)(x0 + q * 4 + b) * a.g.vs;
                const double dx = a.g.goal_x - vx;
                const double d2 = dx * dx + dy2;
                if (MODE == MODE_RRT_DISCOUNT) fsum += rsqrt(sqrt(d2) + a.discount_offset);                 // rrt_star_evaluator.cpp:82
                else fsum += (d2 > a.inv_max_gain2_hi) ? rsqrt(d2)                                           // goal_distance_evaluator.cpp:55-60
                             : ((d2 < a.inv_max_gain2_lo) ? a.max_gain : ((sqrt(d2) > a.inv_max_gain) ? 1.0 / sqrt(d2) : a.max_gain));
              }
            }
          }
        }
      }
      if (kNeedAux && lane_on16) {
        const uint32_t atile = tile + (uint32_t)a.aux_off;
        const int r0 = (wcx - h) & 7;                  // the window occupies tile columns [r0, r0 + side)
        const int lo = r0 - 8 * lcc16, hi = r0 + side - 8 * lcc16;   // valid cells of this lane's chunk: [lo, hi) clipped to [0,8)
        const uint32_t m0 = half_range_mask(lo, hi), m1 = half_range_mask(lo - 2, hi - 2);
        const uint32_t m2 = half_range_mask(lo - 4, hi - 4), m3 = half_range_mask(lo - 6, hi - 6);
        uint32_t addr = atile + (uint32_t)lane * 16u;
        const uint32_t step = (uint32_t)(rpi16 * ncr16) * 16u;
        for (int row = lrow16; row < side; row += rpi16) {
          const uint4 v = lds128(addr);
          addr += step;
          const uint32_t wds[4] = {v.x & m0, v.y & m1, v.z & m2, v.w & m3};
          if (MODE == MODE_AVG_VIEW_COST) {
#pragma unroll
            for (int q = 0; q < 4; ++q) sum_aux2 += (wds[q] & 0xFFFFu) + (wds[q] >> 16);
          } else {  // MODE_LOW_COST: bit 15 = counted (UNKNOWN, inside the bounding volume), bits 0..14 = closest cost + 1
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t l0 = wds[q] & 0xFFFFu, l1 = wds[q] >> 16;
              if (l0 & 0x8000u) { n_unk += 1; fsum += a.recip[l0 & 0x7FFFu]; }
              if (l1 & 0x8000u) { n_unk += 1; fsum += a.recip[l1 & 0x7FFFu]; }
            }
          }
        }
      }

      // ---- warp reduction + the centre cell of the tile (for the duplicated centre voxel) -------------------
      n_unk = __reduce_add_sync(0xFFFFFFFFu, n_unk);
      if (MODE != MODE_GOAL_DISTANCE && MODE != MODE_LOW_COST) n_aux = __reduce_add_sync(0xFFFFFFFFu, n_aux);
      if (MODE == MODE_AVG_VIEW_COST) sum_aux2 = __reduce_add_sync(0xFFFFFFFFu, sum_aux2);
      if (kFloat) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) fsum += __shfl_xor_sync(0xFFFFFFFFu, fsum, off);
      }
      if (lane == j) {
        my_unk = n_unk; my_aux = n_aux; my_sum2 = sum_aux2;
        if (kFloat) my_fsum = fsum;
        // the window centre cell sits at tile position (h, cx - aligned box start); only meaningful for on-map candidates
        if (PACKED) my_crec = cur.crec;
        else if (kNeedRec) my_crec = lds8(tile + (uint32_t)(h * a.box_w + (wcx - ((wcx - h) & ~15))));
        if (kNeedAux) my_caux16 = lds16(tile + (uint32_t)a.aux_off + 2u * (uint32_t)(h * a.box_w16 + (wcx - ((wcx - h) & ~7))));
      }
      s_issue = s;
      if (++s == kStages) { s = 0; parity ^= 1u; }
    }

    // ---- centre voxel duplicate + evaluator epilogue: lane j finishes candidate base + j ------------------------
    if (lane < nb) {
      const int cand = cur.cid;
      const Window& w = cur;
      uint32_t n_unk = my_unk, n_aux = my_aux, sum_aux2 = my_sum2;
      double fsum = my_fsum;
      uint32_t n_vis = 0;
      double gain = 0.0;
      uint32_t terminal = 0;
      const GainGeom& g = a.g;
      if (!w.on_map) {
        n_unk = n_aux = sum_aux2 = 0;
        fsum = 0.0;
      } else {
        n_vis = (uint32_t)(clipped_extent(w.cx, h, max(0, a.use_bv ? g.bv_kx_lo : 0), min(g.W - 1, a.use_bv ? g.bv_kx_hi : g.W - 1)) *
                           clipped_extent(w.cy, h, max(0, a.use_bv ? g.bv_ky_lo : 0), min(g.H - 1, a.use_bv ? g.bv_ky_hi : g.H - 1)));
        // centre voxel (centre coordinates): its own bounding-volume and map-server bounds tests, then the per-cell tests
        const bool c_in_bv = !a.use_bv || (w.vcx >= a.bv_x_min && w.vcx <= a.bv_x_max && w.vcy >= a.bv_y_min && w.vcy <= a.bv_y_max);
        if (c_in_bv) {
          n_vis += 1;
          // getStatusAtLocation bounds test (map_server_planexp.cpp:765). Outside -> MSR_UNKNOWN (:769), cost -1 (:762,:801),
          // not reachable (:877); the path lookup has no bounds test and lands on the window's centre cell
          // (map_planexp.cpp:226-229). The window centre cell (tile position (h,h)) is always on the map here.
          const bool c_in_bounds = 0 <= w.vcx && 0 <= w.vcy && g.map_width > w.vcx && g.map_height > w.vcy;
          const uint32_t crec = my_crec, caux16 = my_caux16;
          bool c_unk, c_free = false, c_path = false, c_reach = false;
          if (MODE == MODE_LOW_COST) {
            c_unk = c_in_bounds ? ((caux16 & 0x7FFFu) != 0) : true;      // idx != 0 <=> the cell is UNKNOWN
          } else {
            c_unk = c_in_bounds ? ((crec & REC_RAW_UNK) != 0) : true;
            c_free = c_in_bounds && (crec & REC_RAW_FREE) != 0;
            c_path = (crec & REC_PATH_RAW) != 0;
            c_reach = (crec & REC_REACH_RAW) != 0;
          }
          if (c_unk) {
            n_unk += 1;
            if (MODE == MODE_COUNT) {
              if (a.evaluator == SIPP_EVAL_EXPAND_REACHABLE) n_aux += c_reach ? 1 : 0;
              else n_aux += c_path ? 1 : 0;
            } else if (MODE == MODE_RRT_DISCOUNT) {
              n_aux += c_path ? 1 : 0;
              if (c_path) {
                const double dx = g.goal_x - w.vcx, dy = g.goal_y - w.vcy;
                fsum += 1.0 / sqrt(sqrt(dx * dx + dy * dy) + a.discount_offset);
              }
            } else if (MODE == MODE_GOAL_DISTANCE) {
              const double dx = g.goal_x - w.vcx, dy = g.goal_y - w.vcy;
              const double dist = sqrt(dx * dx + dy * dy);
              fsum += (dist > a.inv_max_gain) ? 1.0 / dist : a.max_gain;
            } else if (MODE == MODE_LOW_COST) {
              fsum += a.recip[c_in_bounds ? (caux16 & 0x7FFFu) : 1u];   // outside the map server's bounds the lookup gives -1
            }
          } else if (MODE == MODE_AVG_VIEW_COST && c_free) {
            n_aux += 1;
            sum_aux2 += w.crec;       // getVoxelWeight of the centre voxel: the cell's label, whatever the corner voxel's bounding-volume test says
          }
        }
      }
      // ---- evaluator epilogues
      if (MODE == MODE_COUNT) {
        if (a.evaluator == SIPP_EVAL_RRT_STAR) {
          if (n_unk == 0) terminal = 1;                                             // rrt_star_evaluator.cpp:53-57
          else if (a.do_binary) {
            if (n_aux > 0) {
              if (a.use_discount) {                                                 // :61-68
                // single term: round exactly like the reference (no FMA contraction) so the gain is bit-exact
                const double dx = g.goal_x - a.start_xy[2 * cand], dy = g.goal_y - a.start_xy[2 * cand + 1];
                const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                gain = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(nrm, a.discount_offset)));
              } else gain = 1.0;                                                    // :69-75
            }
          } else gain = (double)n_aux + (double)(n_unk - n_aux) * a.npg;            // :85-89 (closed form of the running sum)
        } else if (a.evaluator == SIPP_EVAL_EXPAND_REACHABLE) {
          if (n_unk == 0) terminal = 1;                                             // expand_reachable_area_evaluator.cpp:50-54
          else gain = (double)n_aux + (double)(n_unk - n_aux) * a.discount_factor;  // :65
        } else {                                                                    // NaiveEvaluator [upstream]
          gain = (double)n_unk;
          n_aux = 0;
        }
      } else if (MODE == MODE_RRT_DISCOUNT) {
        if (n_unk == 0) terminal = 1;
        else gain = fsum + (double)(n_unk - n_aux) * a.npg;                         // :78-84
      } else if (MODE == MODE_GOAL_DISTANCE) {
        gain = fsum;
        n_aux = 0;
      } else if (MODE == MODE_AVG_VIEW_COST) {
        if (n_aux > 0) {                                                            // average_view_cost_evaluator.cpp:54-56
          const double mean_view_cost = (double)sum_aux2 / (double)n_aux;
          gain = (double)n_unk * (1.0 / mean_view_cost);
        } else gain = (double)n_unk * (1.0 / g.mean_cost);                          // :57-59
      } else if (MODE == MODE_LOW_COST) {
        gain = fsum;
      }
      a.gain[cand] = gain;
      if (a.counts) *reinterpret_cast<uint4*>(a.counts + 4 * (size_t)cand) = make_uint4(n_vis, n_unk, n_aux, sum_aux2);
      if (a.fsum_out) a.fsum_out[cand] = fsum;
      if (a.terminal) a.terminal[cand] = (uint8_t)terminal;
      if (a.value) {
        // flat tree: GlobalNormalizedGain with a zeroed root: value = gain / cost when cost > 0
        const double c = 0.0 + a.cost[cand];
        const double v = c > 0 ? (0.0 + gain) / c : 0.0;
        a.value[cand] = v;
        if (a.value2) a.value2[cand] = v;
        if (better(v, (long long)cand, best_v, best_i)) { best_v = v; best_i = cand; }
      }
    }
    cur = nxt;
  }
  if (!fused) return;

  // ---- fused argmax: warp -> CTA -> last CTA done (-> peers) -------------------------------------------------------------
  __shared__ double s_bv[kMaxWarpsPerCta];
  __shared__ long long s_bi[kMaxWarpsPerCta];
  __shared__ int s_last;
  __shared__ BestRec s_mine;
  warp_argmax(best_v, best_i);
  if (lane == 0) { s_bv[warp] = best_v; s_bi[warp] = best_i; }
  __syncthreads();
  if (warp == 0) {
    best_v = lane < wpc ? s_bv[lane] : -CUDART_INF;
    best_i = lane < wpc ? s_bi[lane] : LLONG_MAX;
    warp_argmax(best_v, best_i);
    if (lane == 0) {
      a.sel_ws->partial[blockIdx.x].value = best_v;
      a.sel_ws->partial[blockIdx.x].index = best_i;
      __threadfence();                            // the partial record and this CTA's value[] stores before the ticket
      s_last = (atomicAdd(&a.sel_ws->done, 1u) == gridDim.x - 1) ? 1 : 0;
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (warp == 0) {
    best_v = -CUDART_INF;
    best_i = LLONG_MAX;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const double v = a.sel_ws->partial[b].value;
      const long long i = a.sel_ws->partial[b].index;
      if (better(v, i, best_v, best_i)) { best_v = v; best_i = i; }
    }
    warp_argmax(best_v, best_i);
    if (lane == 0) {
      s_mine = finalize_flat(n_cand, a.value, best_v, best_i, a.index_offset);
      if (a.peer_mode < 0) *a.best_out = s_mine;
      a.sel_ws->done = 0;                         // re-arm for the next launch on this stream
    }
  }
  if (a.peer_mode >= 0) {
    __syncthreads();
    peer_exchange_block(a.peer, s_mine, a.best_out, a.peer_mode, pcnt);
  }
}

// one map per launch: the tensor maps and the argument block travel in the kernel's parameter space
template <int MODE, bool PACKED = false>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, PACKED ? 2 : 1)
k_gain(const __grid_constant__ CUtensorMap tm_rec, const __grid_constant__ CUtensorMap tm_aux, const GainArgs a) {
  gain_body<MODE, PACKED>(&tm_rec, &tm_aux, a);
}
// batched (sipp_batch_cycle: an ensemble of maps scored by one launch): block (x, y) scores candidates of map y with the argument
// block as[y]. Count-type evaluators on the 2-bit plane only — that form takes no tensor map.
// One item per map: the argument block and, for the tiled forms, the map's two tensor maps IN DEVICE MEMORY (written by the host before
// the launch, 128-byte aligned; cp.async.bulk.tensor takes a tensor map from global memory as well as from the parameter space).
struct GainBatchItem {
  alignas(128) CUtensorMap tm_rec;
  alignas(128) CUtensorMap tm_aux;
  alignas(16) GainArgs a;
};
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 2)
k_gain_count_b(const GainBatchItem* __restrict__ items) {
  const GainArgs a = items[blockIdx.y].a;
  gain_body<MODE_COUNT, true>(nullptr, nullptr, a);
}
// the tiled byte path for the floating-point evaluators (and the count modes when the 2-bit plane is switched off)
template <int MODE>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 1)
k_gain_tiled_b(const GainBatchItem* __restrict__ items) {
  const GainBatchItem& it = items[blockIdx.y];
  const GainArgs a = it.a;
  // the tensor maps were written into global memory by a host copy: the tensor-map proxy has to acquire them before its first use
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(&it.tm_rec)) : "memory");
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(&it.tm_aux)) : "memory");
  gain_body<MODE, false>(&it.tm_rec, &it.tm_aux, a);
}

// ---- host side: tensor maps -----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap(sipp_ctx* ctx, CUtensorMap* out, void* base, int elem_bytes, size_t pitch_bytes, int box_w_elems, int box_h, int width_elems = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return sipp_set_err(ctx, SIPP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)(width_elems > 0 ? width_elems : ctx->W), (cuuint64_t)ctx->H};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_w_elems, (cuuint32_t)box_h};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sipp_set_err(ctx, SIPP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) box %dx%d", (int)r, box_w_elems, box_h);
  return SIPP_OK;
}

// bytes of a 2-bit-plane row that `side` cells can touch ((side + 2) / 4 + 1), plus up to 15 bytes of slack for the 16-byte
// alignment of the box start, rounded up to the 16-byte granularity of TMA
int packed_box_w(int side) { return (15 + (side + 2) / 4 + 1 + 15) & ~15; }

const CUtensorMap* cached_tmap(sipp_ctx* ctx, int half_fov, int kind, int* rc) {
  for (auto& e : ctx->tma_cache)
    if (e.half_fov == half_fov && e.kind == kind) return &e.map;
  sipp_ctx::TmaEntry e;
  e.half_fov = half_fov;
  e.kind = kind;
  const int side = 2 * half_fov + 1;
  // box width: the window plus up to 15 (7) cells of slack for the 16-byte alignment of the box start
  if (kind == 0) *rc = make_tmap(ctx, &e.map, ctx->d_rec, 1, ctx->pitch, (side + 15 + 15) & ~15, side);
  else if (kind == 1) *rc = make_tmap(ctx, &e.map, ctx->d_freelab, 2, ctx->pitch16 * 2, (side + 7 + 7) & ~7, side);
  else if (kind == 2) *rc = make_tmap(ctx, &e.map, ctx->d_closest, 2, ctx->pitch16 * 2, (side + 7 + 7) & ~7, side);
  else *rc = sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown tensor map kind %d", kind);
  if (*rc != SIPP_OK) return nullptr;
  ctx->tma_cache.push_back(e);
  return &ctx->tma_cache.back().map;
}

template <int MODE, bool PACKED = false>
int launch_mode(sipp_ctx* ctx, const CUtensorMap* tm_rec, const CUtensorMap* tm_aux, const GainArgs& a, cudaStream_t s) {
  // per device and per instantiation, once, under a process-wide lock (contexts are driven from several host threads): the SM count
  // and the dynamic shared-memory opt-in — set to the maximum once instead of per launch, so that two threads launching this kernel
  // with different tile shapes cannot interleave "set attribute" and "launch"
  int num_sms = 0;
  auto kern = k_gain<MODE, PACKED>;
  {
    static std::mutex mu;
    static struct { bool done; int num_sms; } info[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    auto& di = info[ctx->device & 63];
    if (!di.done) {
      cudaError_t e0 = cudaDeviceGetAttribute(&di.num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
      if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
      if (e0 != cudaSuccess) return sipp_set_err(ctx, SIPP_ERR_CUDA, "k_gain setup: %s", cudaGetErrorString(e0));
      di.done = true;
    }
    num_sms = di.num_sms;
  }
  // warps per CTA x tile slots per warp: as many slots as fit in ~220 KB of shared memory
  const int kStages = a.stages;
  int wpc = ctx->gain_warps > 0 ? ctx->gain_warps : kMaxWarpsPerCta;
  while (wpc > 1 && (size_t)wpc * kStages * a.slot_bytes > 220 * 1024) --wpc;
  const size_t smem = PACKED ? 0 : (size_t)wpc * kStages * a.slot_bytes;     // the packed form keeps no tiles
  if (smem > 221 * 1024) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "half_fov %d needs %zu bytes of shared memory per CTA", a.half_fov, smem);
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem);
  if (e != cudaSuccess || occ < 1) return sipp_set_err(ctx, SIPP_ERR_CUDA, "occupancy query failed: %s", cudaGetErrorString(e));
  const int want = (a.n + wpc - 1) / wpc;
  int grid = want < num_sms * occ ? want : num_sms * occ;   // persistent: a multiple of the SM count when saturated
  if (a.best_out && grid > kSelMaxBlocks) grid = kSelMaxBlocks;   // one partial record per CTA in the argmax workspace
  kern<<<grid, wpc * 32, smem, s>>>(*tm_rec, *tm_aux, a);
  ctx->launches += 1;
  e = cudaGetLastError();
  if (e != cudaSuccess) return sipp_set_err(ctx, SIPP_ERR_CUDA, "k_gain launch: %s", cudaGetErrorString(e));
  return sipp_debug_sync(ctx, s, "k_gain");
}

}  // namespace

// fills the argument block of a gain launch; *mode_out = kernel flavour, *packed_out = the 2-bit-plane form serves it
static int gain_prepare(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const double* d_xy, const double* d_start_xy,
                        const double* d_cost, double* d_gain, uint32_t* d_counts, uint8_t* d_terminal, double* d_value,
                        const int* d_sel, const int* d_n, double* d_value2, const FusedSelect* fs, GainArgs& a, int* mode_out, bool* packed_out,
                        bool* need_aux_out, double* d_fsum = nullptr) {
  const int h = p->half_fov;
  if (h < 0 || h > 112) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "half_fov %d outside [0,112] (TMA box limit 256 columns)", h);
  const int side = 2 * h + 1;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.evaluator = p->evaluator;
  a.half_fov = h;
  a.box_w = (side + 15 + 15) & ~15;
  a.box_w16 = (side + 7 + 7) & ~7;
  a.npg = p->non_path_voxel_gain;
  a.use_discount = p->use_distance_discount;
  a.do_binary = p->do_binary_check;
  a.discount_offset = p->discount_offset;
  a.max_gain = p->max_gain;
  a.inv_max_gain = 1.0 / p->max_gain;
  a.inv_max_gain2_lo = a.inv_max_gain * a.inv_max_gain * (1.0 - 1e-9);
  a.inv_max_gain2_hi = a.inv_max_gain * a.inv_max_gain * (1.0 + 1e-9);
  a.discount_factor = p->discount_factor;
  a.use_bv = p->use_bounding_volume;
  a.bv_x_min = p->bv_x_min; a.bv_x_max = p->bv_x_max; a.bv_y_min = p->bv_y_min; a.bv_y_max = p->bv_y_max;
  a.n = n;
  a.xy = d_xy;
  a.start_xy = d_start_xy;
  a.cost = d_cost;
  a.gain = d_gain;
  a.counts = d_counts;
  a.terminal = d_terminal;
  a.value = d_cost ? d_value : nullptr;
  a.value2 = d_cost ? d_value2 : nullptr;
  a.fsum_out = d_fsum;
  a.recip = ctx->d_recip;
  a.label_plane = ctx->d_label;
  a.label_pitch16 = ctx->pitch16;
  a.sel = d_sel;
  a.n_dev = d_n;
  if (fs) {
    if (!d_cost || !d_value || d_sel || d_n) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "fused argmax needs cost / value and a dense candidate list");
    a.best_out = fs->best_out;
    a.sel_ws = reinterpret_cast<SelectWs*>(ctx->d_select_ws);
    a.index_offset = fs->index_offset;
    if (fs->peer) a.peer = *fs->peer;
    a.peer_mode = fs->peer ? fs->peer_mode : -1;
  }
  int mode;
  switch (p->evaluator) {
    case SIPP_EVAL_RRT_STAR: mode = (p->use_distance_discount && !p->do_binary_check) ? MODE_RRT_DISCOUNT : MODE_COUNT; break;
    case SIPP_EVAL_EXPAND_REACHABLE:
    case SIPP_EVAL_NAIVE: mode = MODE_COUNT; break;
    case SIPP_EVAL_GOAL_DISTANCE: mode = MODE_GOAL_DISTANCE; break;
    case SIPP_EVAL_AVERAGE_VIEW_COST: mode = MODE_AVG_VIEW_COST; break;
    case SIPP_EVAL_EXPAND_LOW_COST: mode = MODE_LOW_COST; break;
    default: return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown evaluator %d", p->evaluator);
  }
  if (mode == MODE_COUNT && p->evaluator == SIPP_EVAL_RRT_STAR && p->do_binary_check && p->use_distance_discount && !d_start_xy)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "RRTStarEvaluator binary+discount mode needs start_xy (trajectory[0])");
  // count-type evaluators scan the 2-bit plane (SIPP_GAIN_PACKED=0: the byte record, as the other evaluators do)
  static const bool packed_env = !(getenv("SIPP_GAIN_PACKED") && atoi(getenv("SIPP_GAIN_PACKED")) == 0);
  const bool packed = packed_env && mode == MODE_COUNT && packed_box_w(side) <= 64;   // mask table: 64 cell offsets x <= 4 chunks
  if (packed) {
    a.box_w = packed_box_w(side);
    a.rec_plane = ctx->d_rec;
    a.rec_pitch = ctx->pitch;
    a.rec2_plane = ctx->d_rec2[p->evaluator == SIPP_EVAL_EXPAND_REACHABLE ? 1 : 0];
    a.rec2_pitch = ctx->pitch / 4;
  }
  const size_t rec_bytes = (size_t)a.box_w * side, aux_bytes = (size_t)a.box_w16 * 2 * side;
  const bool need_aux = (mode == MODE_AVG_VIEW_COST || mode == MODE_LOW_COST), need_rec = (mode != MODE_LOW_COST);
  a.aux_off = need_rec ? (int)((rec_bytes + 127) & ~(size_t)127) : 0;
  a.slot_bytes = (int)((need_aux ? (a.aux_off + aux_bytes) : rec_bytes) + 127) & ~127;
  a.stages = ctx->gain_stages > 0 ? ctx->gain_stages : (packed ? 4 : 2);
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  while (a.stages > 2 && (size_t)a.stages * a.slot_bytes > 220 * 1024) --a.stages;
  *mode_out = mode;
  *packed_out = packed;
  *need_aux_out = need_aux;
  return SIPP_OK;
}

int gain_launch(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const double* d_xy, const double* d_start_xy,
                const double* d_cost, double* d_gain, uint32_t* d_counts, uint8_t* d_terminal, double* d_value,
                cudaStream_t s, const int* d_sel, const int* d_n, double* d_value2, const FusedSelect* fs, double* d_fsum) {
  if (n <= 0) return SIPP_OK;
  GainArgs a;
  int mode = 0;
  bool packed = false, need_aux = false;
  int rc = gain_prepare(ctx, p, g, n, d_xy, d_start_xy, d_cost, d_gain, d_counts, d_terminal, d_value, d_sel, d_n, d_value2, fs, a, &mode, &packed, &need_aux, d_fsum);
  if (rc != SIPP_OK) return rc;
  const int h = p->half_fov;
  const CUtensorMap* tm_rec = cached_tmap(ctx, h, 0, &rc);      // the packed form does not use it (plain loads)
  if (!tm_rec) return rc;
  const CUtensorMap* tm_aux = tm_rec;
  if (need_aux) {
    tm_aux = cached_tmap(ctx, h, mode == MODE_LOW_COST ? 2 : 1, &rc);
    if (!tm_aux) return rc;
  }
  switch (mode) {
    case MODE_COUNT: return packed ? launch_mode<MODE_COUNT, true>(ctx, tm_rec, tm_aux, a, s) : launch_mode<MODE_COUNT>(ctx, tm_rec, tm_aux, a, s);
    case MODE_RRT_DISCOUNT: return launch_mode<MODE_RRT_DISCOUNT>(ctx, tm_rec, tm_aux, a, s);
    case MODE_GOAL_DISTANCE: return launch_mode<MODE_GOAL_DISTANCE>(ctx, tm_rec, tm_aux, a, s);
    case MODE_AVG_VIEW_COST: return launch_mode<MODE_AVG_VIEW_COST>(ctx, tm_rec, tm_aux, a, s);
    case MODE_LOW_COST: return launch_mode<MODE_LOW_COST>(ctx, tm_rec, tm_aux, a, s);
  }
  return SIPP_OK;
}

// ---- segments viewed from several sampled viewpoints (sensor_model/sampling_time > 0, sensor_model_planexp.cpp:18-90) -----------------
// The reference concatenates the voxel lists of a segment's views and removes only ADJACENT duplicates (:34,:61). A view's list is
// [centre voxel (half-integer cell coordinates), corner voxels (integer cell coordinates) ...] (map_planexp.cpp:81-101), so a seam
// between two lists never joins equal elements and the evaluator sees every view's voxels: its counts are the sums of the per-view
// counts and its running floating-point sum continues from view to view. k_gain scores every view as a candidate of its own and
// leaves the accumulators (counts + fsum_out); this kernel adds them up per segment, in view order, and applies the evaluator's
// closing formulas — the ones at the end of k_gain, restated over the sums.
namespace {
struct CombineArgs {
  int n, mode, evaluator, do_binary, use_discount;
  double npg, discount_factor, discount_offset, goal_x, goal_y, mean_cost;
  const int* view_off;          // [n + 1]
  const uint32_t* vcounts;      // [views][4]
  const double* vfsum;          // [views]
  const double* start_xy;       // [n][2] (RRTStar binary + discount: trajectory[0])
  double* gain;                 // [n]
  uint32_t* counts;             // optional [n][4]
  uint8_t* terminal;            // optional [n]
};
__global__ void k_gain_combine(const CombineArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  uint32_t n_vis = 0, n_unk = 0, n_aux = 0, sum2 = 0;
  double fsum = 0.0;
  for (int v = a.view_off[i]; v < a.view_off[i + 1]; ++v) {
    const uint4 c = *reinterpret_cast<const uint4*>(a.vcounts + 4 * (size_t)v);
    n_vis += c.x; n_unk += c.y; n_aux += c.z; sum2 += c.w;
    fsum = __dadd_rn(fsum, a.vfsum[v]);
  }
  double gain = 0.0;
  uint32_t terminal = 0;
  if (a.mode == MODE_COUNT) {
    if (a.evaluator == SIPP_EVAL_RRT_STAR) {
      if (n_unk == 0) terminal = 1;                                               // rrt_star_evaluator.cpp:53-57
      else if (a.do_binary) {
        if (n_aux > 0) {
          if (a.use_discount) {                                                   // :61-68
            const double dx = a.goal_x - a.start_xy[2 * i], dy = a.goal_y - a.start_xy[2 * i + 1];
            const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            gain = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(nrm, a.discount_offset)));
          } else gain = 1.0;                                                      // :69-75
        }
      } else gain = __dadd_rn((double)n_aux, __dmul_rn((double)(n_unk - n_aux), a.npg));      // :85-89
    } else if (a.evaluator == SIPP_EVAL_EXPAND_REACHABLE) {
      if (n_unk == 0) terminal = 1;                                               // expand_reachable_area_evaluator.cpp:50-54
      else gain = __dadd_rn((double)n_aux, __dmul_rn((double)(n_unk - n_aux), a.discount_factor));
    } else {                                                                      // NaiveEvaluator [upstream]
      gain = (double)n_unk;
      n_aux = 0;
    }
  } else if (a.mode == MODE_RRT_DISCOUNT) {
    if (n_unk == 0) terminal = 1;
    else gain = __dadd_rn(fsum, __dmul_rn((double)(n_unk - n_aux), a.npg));       // :78-84
  } else if (a.mode == MODE_GOAL_DISTANCE) {
    gain = fsum;
    n_aux = 0;
  } else if (a.mode == MODE_AVG_VIEW_COST) {
    if (n_aux > 0) {                                                              // average_view_cost_evaluator.cpp:54-56
      const double mean_view_cost = __ddiv_rn((double)sum2, (double)n_aux);
      gain = __dmul_rn((double)n_unk, __ddiv_rn(1.0, mean_view_cost));
    } else gain = __dmul_rn((double)n_unk, __ddiv_rn(1.0, a.mean_cost));          // :57-59
  } else if (a.mode == MODE_LOW_COST) {
    gain = fsum;
  }
  a.gain[i] = gain;
  if (a.counts) *reinterpret_cast<uint4*>(a.counts + 4 * (size_t)i) = make_uint4(n_vis, n_unk, n_aux, sum2);
  if (a.terminal) a.terminal[i] = (uint8_t)terminal;
}
}  // namespace

// n segments, segment i viewed from the views d_view_off[i] .. d_view_off[i + 1]) of d_view_xy (n_views in total). Scratch (device):
// d_vgain [n_views] doubles, d_vcounts [n_views][4] uint32, d_vfsum [n_views] doubles. d_start_xy: [n][2] (needed by RRTStar binary + discount).
int gain_launch_views(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const int* d_view_off, int n_views, const double* d_view_xy,
                      const double* d_start_xy, double* d_vgain, uint32_t* d_vcounts, double* d_vfsum, double* d_gain, uint32_t* d_counts,
                      uint8_t* d_terminal, cudaStream_t s) {
  if (n <= 0) return SIPP_OK;
  int rc;
  // every view as a candidate of its own; the per-view gains are not used (binary + discount reads a start point per candidate: the view itself)
  if (n_views > 0 && (rc = gain_launch(ctx, p, g, n_views, d_view_xy, d_view_xy, nullptr, d_vgain, d_vcounts, nullptr, nullptr, s, nullptr, nullptr, nullptr,
                                       nullptr, d_vfsum)) != SIPP_OK)
    return rc;
  CombineArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  switch (p->evaluator) {
    case SIPP_EVAL_RRT_STAR: a.mode = (p->use_distance_discount && !p->do_binary_check) ? MODE_RRT_DISCOUNT : MODE_COUNT; break;
    case SIPP_EVAL_EXPAND_REACHABLE:
    case SIPP_EVAL_NAIVE: a.mode = MODE_COUNT; break;
    case SIPP_EVAL_GOAL_DISTANCE: a.mode = MODE_GOAL_DISTANCE; break;
    case SIPP_EVAL_AVERAGE_VIEW_COST: a.mode = MODE_AVG_VIEW_COST; break;
    case SIPP_EVAL_EXPAND_LOW_COST: a.mode = MODE_LOW_COST; break;
    default: return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown evaluator %d", p->evaluator);
  }
  if (a.mode == MODE_COUNT && p->evaluator == SIPP_EVAL_RRT_STAR && p->do_binary_check && p->use_distance_discount && !d_start_xy)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "RRTStarEvaluator binary+discount mode needs start_xy (trajectory[0])");
  a.evaluator = p->evaluator;
  a.do_binary = p->do_binary_check;
  a.use_discount = p->use_distance_discount;
  a.npg = p->non_path_voxel_gain;
  a.discount_factor = p->discount_factor;
  a.discount_offset = p->discount_offset;
  a.goal_x = g.goal_x; a.goal_y = g.goal_y; a.mean_cost = g.mean_cost;
  a.view_off = d_view_off;
  a.vcounts = d_vcounts;
  a.vfsum = d_vfsum;
  a.start_xy = d_start_xy;
  a.gain = d_gain;
  a.counts = d_counts;
  a.terminal = d_terminal;
  k_gain_combine<<<(n + 255) / 256, 256, 0, s>>>(a);
  ctx->launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sipp_set_err(ctx, SIPP_ERR_CUDA, "k_gain_combine launch: %s", cudaGetErrorString(e));
  return sipp_debug_sync(ctx, s, "k_gain_combine");
}

// ---- batched gain + value + argmax (sipp_batch_cycle) ---------------------------------------------------------------------------
size_t gain_args_bytes() { return sizeof(GainBatchItem); }

// item of map `ctx` for a batched launch (fused value + argmax into fs->best_out). *kind_out: 0 = count-type evaluator on the 2-bit
// plane, 1 + MODE = tiled byte path of that mode; every map of a batch must come out with the same kind (same parameters: it does)
int gain_fill_batch_args(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const double* d_xy, const double* d_cost, double* d_gain,
                         uint8_t* d_terminal, double* d_value, const FusedSelect* fs, void* dst, int* kind_out, int* smem_out, int* wpc_out) {
  GainBatchItem it;
  memset(&it, 0, sizeof(it));
  int mode = 0;
  bool packed = false, need_aux = false;
  int rc = gain_prepare(ctx, p, g, n, d_xy, nullptr, d_cost, d_gain, nullptr, d_terminal, d_value, nullptr, nullptr, nullptr, fs, it.a, &mode, &packed, &need_aux);
  if (rc != SIPP_OK) return rc;
  if (mode == MODE_LOW_COST)
    return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "batched evaluation does not serve ExpandLowCostAreaEvaluator (its closest-observed planes are not maintained by the batched ingest)");
  if (mode == MODE_COUNT && p->evaluator == SIPP_EVAL_RRT_STAR && p->do_binary_check && p->use_distance_discount)
    return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "batched evaluation does not serve RRTStarEvaluator's binary + discount mode (it needs per-candidate start points)");
  *kind_out = packed ? 0 : 1 + mode;
  *smem_out = 0;
  *wpc_out = ctx->gain_warps > 0 ? ctx->gain_warps : kMaxWarpsPerCta;
  if (!packed) {
    const CUtensorMap* tm_rec = cached_tmap(ctx, p->half_fov, 0, &rc);
    if (!tm_rec) return rc;
    const CUtensorMap* tm_aux = tm_rec;
    if (need_aux) {
      tm_aux = cached_tmap(ctx, p->half_fov, 1, &rc);
      if (!tm_aux) return rc;
    }
    memcpy(&it.tm_rec, tm_rec, sizeof(CUtensorMap));
    memcpy(&it.tm_aux, tm_aux, sizeof(CUtensorMap));
    int wpc = ctx->gain_warps > 0 ? ctx->gain_warps : kMaxWarpsPerCta;
    while (wpc > 1 && (size_t)wpc * it.a.stages * it.a.slot_bytes > 220 * 1024) --wpc;
    *smem_out = wpc * it.a.stages * it.a.slot_bytes;
    *wpc_out = wpc;
    if (*smem_out > 221 * 1024) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "half_fov %d needs %d bytes of shared memory per CTA", p->half_fov, *smem_out);
  }
  memcpy(dst, &it, sizeof(it));
  return SIPP_OK;
}

template <int MODE>
static cudaError_t launch_tiled_b(sipp_ctx* c0, const GainBatchItem* items, int grid, int m, int threads, int smem, cudaStream_t s) {
  static std::mutex mu;
  static bool done[64] = {};
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!done[c0->device & 63]) {
      cudaError_t e0 = cudaFuncSetAttribute(k_gain_tiled_b<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
      if (e0 != cudaSuccess) return e0;
      done[c0->device & 63] = true;
    }
  }
  k_gain_tiled_b<MODE><<<dim3((unsigned)grid, (unsigned)m), threads, smem, s>>>(items);
  return cudaGetLastError();
}

int gain_launch_batch(sipp_ctx* c0, const void* d_args, int m, int n, int kind, int smem, int wpc, cudaStream_t s) {
  if (m <= 0 || n <= 0) return SIPP_OK;
  int num_sms = 0;
  {
    static std::mutex mu;
    static struct { bool done; int num_sms; } info[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    auto& di = info[c0->device & 63];
    if (!di.done) {
      cudaError_t e0 = cudaDeviceGetAttribute(&di.num_sms, cudaDevAttrMultiProcessorCount, c0->device);
      if (e0 != cudaSuccess) return sipp_set_err(c0, SIPP_ERR_CUDA, "batched gain setup: %s", cudaGetErrorString(e0));
      di.done = true;
    }
    num_sms = di.num_sms;
  }
  const GainBatchItem* items = reinterpret_cast<const GainBatchItem*>(d_args);
  cudaError_t e = cudaSuccess;
  if (kind == 0) {
    const int want = (n + wpc - 1) / wpc;
    // every map gets the same grid; together they should fill the GPU about twice (2 CTAs per SM)
    int per_map = (2 * num_sms + m - 1) / m;
    if (per_map < 1) per_map = 1;
    int grid = want < per_map ? want : per_map;
    if (grid > kSelMaxBlocks) grid = kSelMaxBlocks;
    k_gain_count_b<<<dim3((unsigned)grid, (unsigned)m), wpc * 32, 0, s>>>(items);
    e = cudaGetLastError();
  } else {
    // tiled form: one CTA per SM (its tile slots fill the shared memory); the maps share the SMs
    const int want = (n + wpc - 1) / wpc;
    int per_map = (num_sms + m - 1) / m;
    if (per_map < 1) per_map = 1;
    int grid = want < per_map ? want : per_map;
    if (grid > kSelMaxBlocks) grid = kSelMaxBlocks;
    switch (kind - 1) {
      case MODE_COUNT: e = launch_tiled_b<MODE_COUNT>(c0, items, grid, m, wpc * 32, smem, s); break;
      case MODE_RRT_DISCOUNT: e = launch_tiled_b<MODE_RRT_DISCOUNT>(c0, items, grid, m, wpc * 32, smem, s); break;
      case MODE_GOAL_DISTANCE: e = launch_tiled_b<MODE_GOAL_DISTANCE>(c0, items, grid, m, wpc * 32, smem, s); break;
      case MODE_AVG_VIEW_COST: e = launch_tiled_b<MODE_AVG_VIEW_COST>(c0, items, grid, m, wpc * 32, smem, s); break;
      default: return sipp_set_err(c0, SIPP_ERR_UNSUPPORTED, "batched evaluation: no kernel for mode %d", kind - 1);
    }
  }
  if (e != cudaSuccess) return sipp_set_err(c0, SIPP_ERR_CUDA, "batched gain launch: %s", cudaGetErrorString(e));
  return SIPP_OK;
}
