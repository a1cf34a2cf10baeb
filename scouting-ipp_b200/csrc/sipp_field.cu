// Follower cost-to-go field, canonical path and path mask (hot-path rows A11/A12).
//
// The reference builds an explicit Boost graph with one node per pixel and runs boost::astar_search_tree
// (planexp_base_planners/src/prm_star_planner.cpp:38-70,132-181). Here the graph is implicit:
//   * edge (a,b) exists iff ||pose_a - pose_b|| < spacing*sqrt(2.0) in the reference's double arithmetic (:30,138)
//     and neither end was first observed as an obstacle id (:113,158-168). The lattice poses X[i] = off + s*i carry
//     ulp noise, so the length of an edge depends on its column/row only through a handful of distinct
//     |dX|^2 / |dY|^2 values ("classes", 9..14 per axis); the host tabulates length and admission per class pair
//     with the reference's exact expression and the kernels look them up.
//   * weight = dist * (c_a + c_b) / 2.0 (:178-181). With half costs hc = c/2 (exact), dist * (hc_a + hc_b) has the
//     same bits: scaling by 2 commutes with IEEE rounding.
// d[v] = min_u d[u] + w(u,v) is relaxed to its fixed point. In IEEE double the least fixed point does not depend
// on the relaxation order (rounding is monotone), so a chaotic tile-parallel relaxation reproduces the bits a
// sequential Dijkstra / A* leaves in d[goal]; all adds and multiplies use _rn intrinsics so nothing is fused.
//
// Relaxation = asynchronous tile-level label correcting on 32x32 tiles: persistent warps pull active tiles from a ring
// queue, a warp iterates its tile in shared memory to the tile-local fixed point with Gauss-Seidel row sweeps, merges
// the result with atomicMin and queues the neighbours whose halo it lowered. No grid barriers (see k_relax).
#include <cooperative_groups.h>
#include <type_traits>
#include <math_constants.h>

#include <cmath>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sipp_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int TS = 32;        // tile edge (cells)
constexpr int TP = TS + 2;    // with a one-cell halo
constexpr int kMaxDd = 256;   // diagonal class table entries kept in shared memory (9..14 classes per axis in practice)

struct RelaxArgs {
  FieldGeom g;
  double* field;            // [H][W]
  const uint16_t* pcode;    // [H][pitch16]
  size_t pitch16;
  int ntx, nty;
  int* flags;               // [ntiles] 1 = tile is in the work queue
  int* queue;               // [cap] ring buffer of tile ids, -1 = empty slot
  int cap;
  int* ctl;                 // see CTL_* below
  const double* tile_w;     // [ntiles][4][TP][TP] edge-weight cache (k_weights)
  int spin_limit;           // watchdog: a warp that polls an empty queue this many times aborts the kernel
  int mode;                 // bit0: publish after every sweep (else once per run); bit1: exclusive tile ownership (owner re-runs)
  int pape;                 // 1 (shared ownership only): second ring queue + cap for tiles that ran before in this plan; it is served first
  int check;                // 1 (needs mode bit0): test whether the opposite sweep has anything to do before running it (sweep_needed)
  uint8_t* touched;         // optional [ntiles]: set for every tile that was activated (incremental re-plan: where d may have changed)
};
enum { CTL_HEAD = 0, CTL_TAIL = 1, CTL_PENDING = 2, CTL_ACTIVATIONS = 3, CTL_SWEEPS = 4, CTL_ABORT = 5, CTL_PUSHES = 6, CTL_HEAD1 = 9, CTL_TAIL1 = 10 };


// fire-and-forget minimum on the bit pattern of a non-negative double (integer order == numeric order). Must be RED (no
// return value): with the returning form (ATOMG) every store of a tile waited for its L2 round trip.
__device__ __forceinline__ void red_min_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.min.u64 [%0], %1;" ::"l"(p), "l"(__double_as_longlong(v)) : "memory");
}
// orders this thread's earlier writes before its later ones at gpu scope (lighter than __threadfence's fence.sc)
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_i32(int* p, int v) {
  asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double half_cost(uint32_t code, double init_half) {
  if (code == PCODE_UNEXPLORED) return init_half;
  if (code == PCODE_REMOVED) return CUDART_INF;
  return 0.5 * (double)code;
}

// ---- tile solver: one WARP per 32x32 tile, Gauss-Seidel row sweeps ------------------------------------------
// A warp owns a tile for one activation. Lane x owns interior column x. A sweep walks the rows top->bottom (or
// bottom->top): row r is relaxed against the already final row r-1 (N, NW, NE) and then against itself (W, E) with a
// warp-shuffle fixed-point loop, so information crosses the whole tile in one sweep in the sweep direction and along
// rows. Sweeps alternate direction until one of them changes nothing: at that point every cell satisfies all eight
// inequalities d[v] <= d[u] + w(u,v), i.e. the tile is at its local fixed point. No block barriers.
constexpr int kRelaxWarps = 5;

constexpr int kTileWDoubles = 4 * TP * TP;   // per-tile block of the weight cache: we, ws, wse, wsw planes
struct TileSmem {
  double d[TP][TP];      // tentative costs incl. one-cell halo
  // edge weights of the tile (k_weights computes them once per plan; they depend on the map, not on d), copied in as ONE
  // block by cp.async.bulk: weight of the edge from (r,c) to
  double we[TP][TP];     //   (r, c+1)
  double ws[TP][TP];     //   (r+1, c)
  double wse[TP][TP];    //   (r+1, c+1)
  double wsw[TP][TP];    //   (r+1, c-1)      +inf: edge missing (removed node, off map, diagonal not admitted)
};
static_assert(sizeof(TileSmem) == (TP * TP + kTileWDoubles) * sizeof(double), "weight planes must be contiguous");
static_assert((TP * TP * sizeof(double)) % 16 == 0 && (kTileWDoubles * sizeof(double)) % 16 == 0, "bulk copy needs 16-byte granularity");
// The same worker without the staged copy of the weights: only d lives in shared memory (9 KB instead of 46 KB per warp, so three
// times as many warps fit on an SM) and the sweeps read the weight planes straight from the cache block in global memory (read-only
// while the relaxation runs: ld.global.nc). More latency per activation, more activations in flight: the form for a batch of
// maps, where the workers are busy, not for one map, where the chain of dependent activations bounds the plan.
struct TileD {
  double d[TP][TP];
};
template <bool STAGED> struct WView;
template <> struct WView<true> {
  static constexpr int kAhead = 0;      // rows whose weights a sweep requests ahead of the row it works on: none (shared memory)
  const TileSmem* t;
  __device__ __forceinline__ double we(int r, int c) const { return t->we[r][c]; }
  __device__ __forceinline__ double ws(int r, int c) const { return t->ws[r][c]; }
  __device__ __forceinline__ double wse(int r, int c) const { return t->wse[r][c]; }
  __device__ __forceinline__ double wsw(int r, int c) const { return t->wsw[r][c]; }
};
template <> struct WView<false> {
  static constexpr int kAhead = 2;      // two rows ahead: an L2 round trip is longer than one row's in-row loop
  const double* w;
  __device__ __forceinline__ double we(int r, int c) const { return __ldg(w + r * TP + c); }
  __device__ __forceinline__ double ws(int r, int c) const { return __ldg(w + TP * TP + r * TP + c); }
  __device__ __forceinline__ double wse(int r, int c) const { return __ldg(w + 2 * TP * TP + r * TP + c); }
  __device__ __forceinline__ double wsw(int r, int c) const { return __ldg(w + 3 * TP * TP + r * TP + c); }
};

// weight of an edge of length `dist` between cells with half costs hca / hcb: dist * (hca + hcb), the reference's
// (dist * (ca + cb)) / 2.0 (prm_star_planner.cpp:178-181); +inf for a missing edge (never NaN, also for zero costs)
__device__ __forceinline__ double edge_w(double dist, double hca, double hcb) {
  return dist < CUDART_INF ? __dmul_rn(dist, __dadd_rn(hca, hcb)) : CUDART_INF;
}

// One sweep over the interior rows. DIR = +1: rows 1..32 using row r-1; DIR = -1: rows 32..1 using row r+1.
// The previous row never goes through shared memory: its final values at columns lx-1, lx, lx+1 are exactly what the last
// (change-free) iteration of its in-row loop held in (dl, d, dr), halo columns included, so they are carried in registers.
// Row r is relaxed against them (N, NW, NE) and against itself (W, E) in a shuffle loop that runs until no lane changes; a
// row nothing changes in costs one iteration. All edge weights come from the per-tile planes. Returns (warp-uniform)
// whether any cell was lowered; bit r-1 of `rows_changed` = this lane's cell in row r was lowered.
template <int DIR, class Tile, class WV>
__device__ __forceinline__ bool sweep_rows(Tile& t, const WV& w, int lane, uint32_t& rows_changed) {
  const int lx = lane + 1;
  bool chg = false;
  const int r_first = (DIR > 0) ? 0 : TP - 1;
  double pl = t.d[r_first][lx - 1], pc = t.d[r_first][lx], pr = t.d[r_first][lx + 1];
  // weights of the edges to the previous row's three cells and to the row neighbours, for the i-th row of the sweep
  struct W5 { double wn, wnl, wnr, wl, wr; };
  auto fetch = [&](int i) -> W5 {
    const int r = (DIR > 0) ? (1 + i) : (TS - i);
    W5 q;
    q.wn = (DIR > 0) ? w.ws(r - 1, lx) : w.ws(r, lx);
    q.wnl = (DIR > 0) ? w.wse(r - 1, lx - 1) : w.wsw(r, lx);
    q.wnr = (DIR > 0) ? w.wsw(r - 1, lx + 1) : w.wse(r, lx);
    q.wl = w.we(r, lx - 1);
    q.wr = w.we(r, lx);
    return q;
  };
  constexpr int AH = WV::kAhead;
  W5 ahead[AH > 0 ? AH : 1];
#pragma unroll
  for (int k = 0; k < AH; ++k) ahead[k] = fetch(k);
#pragma unroll 4
  for (int i = 0; i < TS; ++i) {
    const int r = (DIR > 0) ? (1 + i) : (TS - i);
    W5 cur;
    if constexpr (AH == 0) {
      cur = fetch(i);
    } else {      // the loads of row i + AH are in flight while rows i .. i + AH - 1 run their in-row loops
      cur = ahead[0];
#pragma unroll
      for (int k = 0; k + 1 < AH; ++k) ahead[k] = ahead[k + 1];
      if (i + AH < TS) ahead[AH - 1] = fetch(i + AH);
    }
    const double wn = cur.wn, wnl = cur.wnl, wnr = cur.wnr, wl = cur.wl, wr = cur.wr;
    const double hl = t.d[r][0], hr = t.d[r][TP - 1];      // left / right halo of this row (constant during the sweep)
    double d = t.d[r][lx];
    const double d0 = d;
    double ca = __dadd_rn(pc, wn);
    { const double c1 = __dadd_rn(pl, wnl), c2 = __dadd_rn(pr, wnr); const double c3 = c1 < c2 ? c1 : c2; ca = c3 < ca ? c3 : ca; }
    if (ca < d) d = ca;
    double dl, dr;
    bool again;
    do {
      dl = __shfl_up_sync(0xFFFFFFFFu, d, 1); dr = __shfl_down_sync(0xFFFFFFFFu, d, 1);
      if (lane == 0) dl = hl;
      if (lane == 31) dr = hr;
      const double cl = __dadd_rn(dl, wl), cr = __dadd_rn(dr, wr);
      const double c = cl < cr ? cl : cr;
      again = c < d;
      if (again) d = c;
    } while (__any_sync(0xFFFFFFFFu, again));
    pl = dl; pc = d; pr = dr;
    if (d < d0) {
      chg = true;
      rows_changed |= 1u << (r - 1);
      t.d[r][lx] = d;
    }
  }
  __syncwarp();                                    // the sweep's stores are visible to the whole warp
  return __any_sync(0xFFFFFFFFu, chg);
}

// Would a sweep in direction DIR lower anything? After a sweep in the OPPOSITE direction every cell satisfies the inequalities towards
// that sweep's previous row and towards its row neighbours (rows are final once processed), so the tile is at its local fixed point
// iff no cell can be lowered from the three cells of the row a DIR sweep would come from — exactly the first thing such a sweep
// tests for every row, here without the in-row shuffle loop and without any dependence between rows (the loads and adds of all
// rows overlap). One sweep + this test replaces the "confirming" second sweep of a run whose front came from one side.
template <int DIR, class Tile, class WV>
__device__ __forceinline__ bool sweep_needed(const Tile& t, const WV& w, int lane) {
  const int lx = lane + 1;
  bool need = false;
#pragma unroll 8
  for (int r = 1; r <= TS; ++r) {
    const int p = r - DIR;                                  // the row a DIR sweep relaxes row r against
    const double wn = (DIR > 0) ? w.ws(r - 1, lx) : w.ws(r, lx);
    const double wnl = (DIR > 0) ? w.wse(r - 1, lx - 1) : w.wsw(r, lx);
    const double wnr = (DIR > 0) ? w.wsw(r - 1, lx + 1) : w.wse(r, lx);
    double ca = __dadd_rn(t.d[p][lx], wn);
    { const double c1 = __dadd_rn(t.d[p][lx - 1], wnl), c2 = __dadd_rn(t.d[p][lx + 1], wnr); const double c3 = c1 < c2 ? c1 : c2; ca = c3 < ca ? c3 : ca; }
    need |= ca < t.d[r][lx];
  }
  return __any_sync(0xFFFFFFFFu, need);
}

// Loads d of the 34x34 window whose interior starts at (x0, y0) into shared memory. Loads are issued in batches (plain
// L2-coherent ld.global.cg: other warps update the field with L2 atomics while k_relax runs, so L1 is bypassed; an aligned
// 8-byte load cannot tear) so that their latencies overlap instead of serialising on the stores.
template <class Tile>
__device__ __forceinline__ void load_d(Tile& t, const FieldGeom& g, const double* field, int x0, int y0, int lane) {
  const int gx = x0 + lane;                  // interior column of this lane (tile column lane + 1)
  const bool colin = gx >= 0 && gx < g.W;
  double hv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {              // the two halo columns: 68 entries over the warp
    const int e = lane + 32 * j;
    const int ly = e >> 1, side_r = e & 1;
    const int hx = side_r ? x0 + TS : x0 - 1, gy = y0 - 1 + ly;
    const bool inb = e < 2 * TP && hx >= 0 && hx < g.W && gy >= 0 && gy < g.H;
    hv[j] = inb ? __ldcg(field + (size_t)gy * g.W + hx) : CUDART_INF;
  }
#pragma unroll 1
  for (int lyb = 0; lyb < TP; lyb += 17) {
    double dv[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int gy = y0 - 1 + lyb + j;
      const bool inb = colin && gy >= 0 && gy < g.H;
      dv[j] = inb ? __ldcg(field + (size_t)gy * g.W + gx) : CUDART_INF;
    }
#pragma unroll
    for (int j = 0; j < 17; ++j) t.d[lyb + j][lane + 1] = dv[j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int e = lane + 32 * j;
    if (e < 2 * TP) t.d[e >> 1][(e & 1) ? TP - 1 : 0] = hv[j];
  }
}

// Re-run of a tile by the warp that already holds it (exclusive ownership): only the border ring (rows / columns 1 and 32,
// other tiles push into it) and the halo (rows / columns 0 and 33) can have changed. Loads them, keeps the lower of (held,
// loaded) — they can only differ by a push — and returns the set of tile rows (0..33) that received a lower value.
template <class Tile>
__device__ __forceinline__ uint64_t reload_ring(Tile& t, const FieldGeom& g, const double* field, int x0, int y0, int lane) {
  const int rows4[4] = {0, 1, TS, TP - 1};
  double rv[4], cv[4], kv = CUDART_INF;
  const int gxc = x0 + lane, gyr = y0 + lane;          // this lane's interior column (tile col lane+1) / row (tile row lane+1)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int gy = y0 - 1 + rows4[k], gx = x0 - 1 + rows4[k];
    rv[k] = (gxc < g.W && gy >= 0 && gy < g.H) ? __ldcg(field + (size_t)gy * g.W + gxc) : CUDART_INF;     // row rows4[k], column lane+1
    cv[k] = (gyr < g.H && gx >= 0 && gx < g.W) ? __ldcg(field + (size_t)gyr * g.W + gx) : CUDART_INF;     // column rows4[k], row lane+1
  }
  int ky = 0, kx = 0;
  if (lane < 4) {                                       // the four halo corners
    ky = (lane < 2) ? 0 : TP - 1; kx = (lane & 1) ? TP - 1 : 0;
    const int gy = y0 - 1 + ky, gx = x0 - 1 + kx;
    if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H) kv = __ldcg(field + (size_t)gy * g.W + gx);
  }
  uint64_t E = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool lo = rv[k] < t.d[rows4[k]][lane + 1];
    if (lo) t.d[rows4[k]][lane + 1] = rv[k];
    if (__any_sync(0xFFFFFFFFu, lo)) E |= 1ull << rows4[k];
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool lo = cv[k] < t.d[lane + 1][rows4[k]];
    if (lo) t.d[lane + 1][rows4[k]] = cv[k];
    E |= (uint64_t)__ballot_sync(0xFFFFFFFFu, lo) << 1;     // lane l holds tile row l+1
  }
  {
    const bool lo = lane < 4 && kv < t.d[ky][kx];
    if (lo) t.d[ky][kx] = kv;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, lo);
    if (m & 3u) E |= 1ull;
    if (m & 12u) E |= 1ull << (TP - 1);
  }
  __syncwarp();
  return E;
}

// ---- edge-weight cache: one block of four 34x34 planes per tile (halo edges included, so neighbouring blocks overlap) -----
struct WeightArgs {
  FieldGeom g;
  const uint16_t* pcode;
  size_t pitch16;
  double* tile_w;    // [ntiles][4][TP][TP]
  int ntx;
  const int* tile_list;   // optional: block b rebuilds tile tile_list[b] (incremental re-plan: only the tiles a patch touched)
  int n_items;            // blocks with work (a batched launch is as wide as the widest map's list)
};
constexpr int kWeightThreads = 256;
__device__ __forceinline__ void weights_body(const WeightArgs& a) {
  if ((int)blockIdx.x >= a.n_items) return;
  __shared__ float hcf[TP][TP + 1];        // half costs (multiples of 0.5 <= 1799.5 or +inf: exact in fp32)
  __shared__ int xc[TP], yc[TP];           // class of the edge between tile columns c, c+1 / rows r, r+1 (-1: off map)
  __shared__ double hd[TP], vd[TP];
  __shared__ double dd[kMaxDd];
  const FieldGeom& g = a.g;
  const int tid = threadIdx.x;
  const int tile = a.tile_list ? a.tile_list[blockIdx.x] : (int)blockIdx.x;
  const int tx = tile % a.ntx, ty = tile / a.ntx;
  const int x0 = tx * TS, y0 = ty * TS;
  const bool dd_in_smem = g.Kx * g.Ky <= kMaxDd;
  if (dd_in_smem)
    for (int i = tid; i < g.Kx * g.Ky; i += kWeightThreads) dd[i] = g.ddist[i];
  const double* ddp = dd_in_smem ? dd : g.ddist;
  for (int idx = tid; idx < TP * TP; idx += kWeightThreads) {
    const int r = idx / TP, c = idx - r * TP;
    const int gx = x0 - 1 + c, gy = y0 - 1 + r;
    const bool inb = gx >= 0 && gx < g.W && gy >= 0 && gy < g.H;
    hcf[r][c] = (float)half_cost(inb ? (uint32_t)a.pcode[(size_t)gy * a.pitch16 + gx] : (uint32_t)PCODE_REMOVED, g.init_half_cost);
  }
  if (tid < TP) {
    const int gy = y0 - 1 + tid, gx = x0 - 1 + tid;     // edges (gy, gy+1) and (gx, gx+1)
    const bool oky = gy >= 0 && gy + 1 < g.H, okx = gx >= 0 && gx + 1 < g.W;
    const int cy = oky ? g.ycls[gy] : -1, cx = okx ? g.xcls[gx] : -1;
    yc[tid] = cy;
    vd[tid] = oky ? g.vdist[cy] : CUDART_INF;
    xc[tid] = cx;
    hd[tid] = okx ? g.hdist[cx] : CUDART_INF;
  }
  __syncthreads();
  double* out = a.tile_w + (size_t)tile * kTileWDoubles;
  for (int idx = tid; idx < TP * TP; idx += kWeightThreads) {
    const int r = idx / TP, c = idx - r * TP;
    const double h0 = (double)hcf[r][c];
    const bool rn = r + 1 < TP, cn = c + 1 < TP;
    auto ddc = [&](int xcl_, int ycl_) { return (xcl_ >= 0 && ycl_ >= 0) ? ddp[xcl_ * g.Ky + ycl_] : CUDART_INF; };
    out[idx] = cn ? edge_w(hd[c], h0, (double)hcf[r][c + 1]) : CUDART_INF;
    out[TP * TP + idx] = rn ? edge_w(vd[r], h0, (double)hcf[r + 1][c]) : CUDART_INF;
    out[2 * TP * TP + idx] = (rn && cn) ? edge_w(ddc(xc[c], yc[r]), h0, (double)hcf[r + 1][c + 1]) : CUDART_INF;
    out[3 * TP * TP + idx] = (rn && c > 0) ? edge_w(ddc(xc[c - 1], yc[r]), h0, (double)hcf[r + 1][c - 1]) : CUDART_INF;
  }
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Persistent, barrier-free worker kernel. Warps pull tile ids from a global ring queue, relax the tile in shared memory with
// alternating row sweeps and, after EVERY sweep that lowered something, publish: store the lowered cells, relax across the
// tile border into the neighbours' cells, queue the neighbours that received a lower value. Why this is exact:
//  (1) every value ever written is the rounded length of a real path (an upper bound of the fixed point) and the field only
//      ever decreases: cells other tiles may write too (our border ring, the neighbours' border rings) are merged with
//      atomicMin on the bit pattern (for non-negative doubles the integer order is the numeric order); interior cells are
//      written by the tile's owner only — a tile is run by at most one warp at a time (state bit RUN);
//  (2) whoever lowers a cell of tile T afterwards sets T's QUEUED bit; if T is idle that queues it, if T is running its owner
//      sees the bit when it finishes and runs T again (after clearing the bit and BEFORE re-loading the tile), so no update
//      is missed; every activation offers all its border values to the neighbours at least once, so a border cell lowered by
//      one neighbour is carried on to the others;
//  (3) the kernel ends when the number of queued + running tiles drops to zero, i.e. at a global fixed point, which is unique.
constexpr int ST_QUEUED = 1, ST_RUN = 2, ST_VISITED = 4;

// Batched form (ROAM): ONE work queue for all maps of a batch. An entry is (map << kRoamShift) | tile; a worker takes whatever
// comes next and reads that map's argument block, so the workers go where the work is — a map with a long re-plan gets all the
// warps the others no longer need, instead of a fixed share of CTAs that hold their SMs until their own map is drained. Tile
// flags, statistics and the field stay per map; HEAD / TAIL / PENDING / ABORT are the batch's (pending == 0: every map is at its
// fixed point). k_roam_collect moves what the per-map seeding kernels queued into it; k_gather_plan_out rewinds its counters.
constexpr int kRoamShift = 20;
struct RoamQ {
  int* queue; int cap; int* ctl;
  const char* args; size_t stride;     // map y's RelaxArgs at args + y * stride
  int spin_limit;                      // the watchdog's (map 0's argument block may be one that does not run this step)
};

template <bool ROAM, bool STAGED>
__device__ __forceinline__ void relax_body_t(const RelaxArgs& a0, const RoamQ& rq) {
  extern __shared__ __align__(16) uint8_t relax_smem[];
  __shared__ __align__(8) uint64_t s_bar[kRelaxWarps];      // STAGED only (kRelaxWarps warps per CTA)
  using Tile = typename std::conditional<STAGED, TileSmem, TileD>::type;

  int* const q_queue = ROAM ? rq.queue : a0.queue;
  int* const q_ctl = ROAM ? rq.ctl : a0.ctl;
  const int q_cap = ROAM ? rq.cap : a0.cap;
  const int pape = ROAM ? 0 : a0.pape;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Tile& t = reinterpret_cast<Tile*>(relax_smem)[warp];
  uint32_t bar = 0, bar_parity = 0;
  if constexpr (STAGED) {
    bar = smem_addr(&s_bar[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }

  int sweeps_local = 0, act_local = 0, push_local = 0;
#ifdef SIPP_RELAX_PROFILE
  long long tp[6] = {0, 0, 0, 0, 0, 0};   // pop, load d, sweeps, store+cross, push, weights
  long long tk = clock64();
#define SIPP_TICK(i) { const long long now_ = clock64(); tp[i] += now_ - tk; tk = now_; }
#else
#define SIPP_TICK(i)
#endif
  for (;;) {
    // ---- pop: take a ticket, wait for its slot to be filled (or for global termination)
    int ti = -2;
    if (lane == 0) {
      // two-ring form: the ring of re-activated tiles first when it holds entries (a snapshot: a ticket taken in vain just waits there)
      const bool ring1 = pape && (int)((unsigned)ld_volatile_i32(&q_ctl[CTL_TAIL1]) - (unsigned)ld_volatile_i32(&q_ctl[CTL_HEAD1])) > 0;
      const unsigned ticket = (unsigned)atomicAdd(&q_ctl[ring1 ? CTL_HEAD1 : CTL_HEAD], 1);
      int* slot = q_queue + (ring1 ? q_cap : 0) + (ticket % (unsigned)q_cap);
      int spins = 0;
      for (;;) {
        const int v = ld_volatile_i32(slot);
        if (v >= 0) {
          st_volatile_i32(slot, -1);
          ti = v;
          break;
        }
        if (ld_volatile_i32(&q_ctl[CTL_PENDING]) == 0 || ld_volatile_i32(&q_ctl[CTL_ABORT]) != 0) break;
        if (++spins > (ROAM ? rq.spin_limit : a0.spin_limit)) {
          atomicExch(&q_ctl[CTL_ABORT], 1);
          break;
        }
        __nanosleep(64);
      }
      if (ti >= 0) {
        const RelaxArgs& am = ROAM ? *reinterpret_cast<const RelaxArgs*>(rq.args + (size_t)(ti >> kRoamShift) * rq.stride) : a0;
        const int tm = ROAM ? (ti & ((1 << kRoamShift) - 1)) : ti;
        atomicExch(&am.flags[tm], (am.mode & 2) ? ST_RUN : (pape ? ST_VISITED : 0));   // queued -> running; a lowering from here on sets QUEUED again
        fence_gpu();
      }
    }
    ti = __shfl_sync(0xFFFFFFFFu, ti, 0);
    if (ti < 0) break;
    SIPP_TICK(0)
    const int tag = ROAM ? (ti >> kRoamShift) << kRoamShift : 0;      // this entry's map, as it goes into the entries pushed below
    if (ROAM) ti &= (1 << kRoamShift) - 1;
    const RelaxArgs& a = ROAM ? *reinterpret_cast<const RelaxArgs*>(rq.args + (size_t)(tag >> kRoamShift) * rq.stride) : a0;
    const FieldGeom& g = a.g;
    const int act0 = act_local, sweeps0 = sweeps_local, push0 = push_local;

    const int tx = ti % a.ntx, ty = ti / a.ntx;
    const int x0 = tx * TS, y0 = ty * TS;
    if (a.touched && lane == 0) a.touched[ti] = 1;
    // the tile's weight block: one bulk async copy global -> shared, completion on this warp's mbarrier; it overlaps with the
    // loads of d below
    __syncwarp();           // everyone is done reading the previous tile's planes
    WView<STAGED> wv;
    bool weights_pending = STAGED;
    if constexpr (STAGED) {
      wv.t = &t;
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kTileWDoubles * sizeof(double))) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(&t.we[0][0])),
                     "l"(a.tile_w + (size_t)ti * kTileWDoubles), "r"((uint32_t)(kTileWDoubles * sizeof(double))), "r"(bar)
                     : "memory");
      }
    } else {
      wv.w = a.tile_w + (size_t)ti * kTileWDoubles;
    }

    bool first_run = true;
    for (;;) {   // runs of this tile: again whenever somebody lowered one of its cells while it was running
      // first run: the whole window; re-run: only the border ring and the halo can have changed (the interior is ours alone
      // and still in shared memory)
      if (first_run) load_d(t, g, a.field, x0, y0, lane);
      else reload_ring(t, g, a.field, x0, y0, lane);
      first_run = false;
      if (weights_pending) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar), "r"(bar_parity) : "memory");
        bar_parity ^= 1u;
        weights_pending = false;
      }
      __syncwarp();
      ++act_local;
      SIPP_TICK(1)

      // alternate downward / upward sweeps until one of them changes nothing (its opposite was completed right before it);
      // publish after every sweep that lowered something, and at least once per run
      bool published = false, any_change = false;
      for (int it = 0;; ++it) {
        uint32_t rows_changed = 0;
        const bool c = (it & 1) ? sweep_rows<-1>(t, wv, lane, rows_changed) : sweep_rows<+1>(t, wv, lane, rows_changed);
        ++sweeps_local;
        SIPP_TICK(2)
        const bool last = !c && it >= 1;
        if (c) any_change = true;
        // with the one-sided test below a run may end after its first sweep: the publish every run owes (offers of values that were
        // pushed into this tile to its other neighbours) then happens here instead of after a second sweep
        const bool owe = !published && (last || (a.check && it == 0));
        if ((a.mode & 1) ? (c || owe) : last) {
          if (!(a.mode & 1)) rows_changed = any_change ? 0xFFFFFFFFu : 0u;   // one publish per run: store every row
          published = true;
          // ---- store what this lane lowered in this sweep. The border ring first: it is what the neighbours read as their halo, and it
          // has to be visible before they are told to look; the interior (nobody else reads it) follows AFTER the pushes below, off the
          // chain of dependent activations that bounds the plan
          const bool edge_col = lane == 0 || lane == 31;
          const uint32_t ring_rows = edge_col ? 0xFFFFFFFFu : 0x80000001u;
          auto store_rows = [&](uint32_t rc) {
            const int gx = x0 + lane;
            while (rc) {
              const int r = __ffs(rc) - 1;
              rc &= rc - 1;
              const int gy = y0 + r;
              double* dst = a.field + (size_t)gy * g.W + gx;
              const double v = t.d[r + 1][lane + 1];
              if (edge_col || r == 0 || r == TS - 1 || !(a.mode & 2))   // border ring: neighbours push into it concurrently
                red_min_f64(dst, v);
              else
                *dst = v;                                 // interior: this warp is the only writer
            }
          };
          store_rows(rows_changed & ring_rows);
          // ---- relax ACROSS the tile border: every halo cell n (it belongs to a neighbour tile) against its up to three
          // neighbours c on our side, cand = d[c] + w(c,n) — bit for bit what the neighbour would compute from its own halo.
          // A neighbour is queued only if cand beats the value we hold for n (then our copy is updated too), so a tile is
          // not re-activated by the echo of the values it produced itself. Lane l handles column l+1 of the top / bottom
          // halo rows and row l+1 of the left / right halo columns; lanes 0..3 the four corners.
          int m = 0;   // bit k: neighbour k (N S W E NW NE SW SE) received a lower value
          {
            const int lx = lane + 1, ly = lane + 1;
            auto offer = [&](int hy, int hx, double cand) -> bool {   // (hy,hx): halo cell in tile coordinates
              if (!(cand < t.d[hy][hx])) return false;               // +inf (edge missing / off map) never passes
              t.d[hy][hx] = cand;
              const int gx = x0 - 1 + hx, gy = y0 - 1 + hy;
              red_min_f64(a.field + (size_t)gy * g.W + gx, cand);
              return true;
            };
            auto min3 = [](double a_, double b_, double c_) { const double m_ = a_ < b_ ? a_ : b_; return m_ < c_ ? m_ : c_; };
            // top halo row (0, lx) from interior row 1; bottom halo row (TP-1, lx) from interior row TS
            {
              double cl = CUDART_INF, cr = CUDART_INF;
              if (lx > 1) cl = __dadd_rn(t.d[1][lx - 1], wv.wsw(0, lx));
              if (lx < TS) cr = __dadd_rn(t.d[1][lx + 1], wv.wse(0, lx));
              if (offer(0, lx, min3(__dadd_rn(t.d[1][lx], wv.ws(0, lx)), cl, cr))) m |= 1;
              cl = CUDART_INF; cr = CUDART_INF;
              if (lx > 1) cl = __dadd_rn(t.d[TS][lx - 1], wv.wse(TS, lx - 1));
              if (lx < TS) cr = __dadd_rn(t.d[TS][lx + 1], wv.wsw(TS, lx + 1));
              if (offer(TP - 1, lx, min3(__dadd_rn(t.d[TS][lx], wv.ws(TS, lx)), cl, cr))) m |= 2;
            }
            // left halo column (ly, 0) from interior column 1; right halo column (ly, TP-1) from interior column TS
            {
              double cu = CUDART_INF, cd = CUDART_INF;
              if (ly > 1) cu = __dadd_rn(t.d[ly - 1][1], wv.wsw(ly - 1, 1));
              if (ly < TS) cd = __dadd_rn(t.d[ly + 1][1], wv.wse(ly, 0));
              if (offer(ly, 0, min3(__dadd_rn(t.d[ly][1], wv.we(ly, 0)), cu, cd))) m |= 4;
              cu = CUDART_INF; cd = CUDART_INF;
              if (ly > 1) cu = __dadd_rn(t.d[ly - 1][TS], wv.wse(ly - 1, TS));
              if (ly < TS) cd = __dadd_rn(t.d[ly + 1][TS], wv.wsw(ly, TP - 1));
              if (offer(ly, TP - 1, min3(__dadd_rn(t.d[ly][TS], wv.we(ly, TS)), cu, cd))) m |= 8;
            }
            if (lane < 4) {                        // corners: NW NE SW SE from the diagonal interior cell
              double cand;
              int hy, hx;
              if (lane == 0) { hy = 0; hx = 0; cand = __dadd_rn(t.d[1][1], wv.wse(0, 0)); }
              else if (lane == 1) { hy = 0; hx = TP - 1; cand = __dadd_rn(t.d[1][TS], wv.wsw(0, TP - 1)); }
              else if (lane == 2) { hy = TP - 1; hx = 0; cand = __dadd_rn(t.d[TS][1], wv.wsw(TS, 1)); }
              else { hy = TP - 1; hx = TP - 1; cand = __dadd_rn(t.d[TS][TS], wv.wse(TS, TS)); }
              if (offer(hy, hx, cand)) m |= 16 << lane;
            }
          }
          m = __reduce_or_sync(0xFFFFFFFFu, (unsigned)m);
          SIPP_TICK(3)
          if (m) {
            fence_gpu();      // the lowered neighbour cells are visible before the neighbour can be told to look
            __syncwarp();
            // neighbours: N S W E NW NE SW SE. Lane k decides for neighbour k (idle -> queued: ours to enqueue; a running tile re-runs
            // itself), then lane 0 counts all of the warp's new entries into CTL_PENDING and reserves their queue positions with ONE
            // returning atomic — the same thread later releases this tile, so its additions and its subtraction are ordered without a fence
            int nt = -1;
            if (lane < 8 && ((m >> lane) & 1)) {
              const int ntx_ = tx + ((0xA8 >> lane) & 1) - ((0x54 >> lane) & 1);      // dx: 0 0 -1 1 -1 1 -1 1
              const int nty_ = ty + ((0xC2 >> lane) & 1) - ((0x31 >> lane) & 1);      // dy: -1 1 0 0 -1 -1 1 1
              if (ntx_ >= 0 && ntx_ < a.ntx && nty_ >= 0 && nty_ < a.nty) nt = nty_ * a.ntx + ntx_;
            }
            const int oldf = nt >= 0 ? atomicOr(&a.flags[nt], ST_QUEUED) : ST_QUEUED;
            const bool enq = nt >= 0 && (oldf & (ST_QUEUED | ST_RUN)) == 0;
            const unsigned em = __ballot_sync(0xFFFFFFFFu, enq);
            if (em) {
              // two-ring form (Pape / d'Esopo): a tile that already ran in this plan carries a correction — it goes to the ring that is
              // served first, so that the correction overtakes the front instead of trailing it. Only when there is a backlog: a ring
              // with waiting workers takes the entry whatever its kind.
              unsigned m1 = 0;
              if (pape) {
                m1 = __ballot_sync(0xFFFFFFFFu, enq && (oldf & ST_VISITED));
                int w0 = 0, w1 = 0;
                if (lane == 0) {
                  w0 = (int)((unsigned)ld_volatile_i32(&q_ctl[CTL_HEAD]) - (unsigned)ld_volatile_i32(&q_ctl[CTL_TAIL]));
                  w1 = (int)((unsigned)ld_volatile_i32(&q_ctl[CTL_HEAD1]) - (unsigned)ld_volatile_i32(&q_ctl[CTL_TAIL1]));
                }
                w0 = __shfl_sync(0xFFFFFFFFu, w0, 0);
                w1 = __shfl_sync(0xFFFFFFFFu, w1, 0);
                if (w1 > 0) m1 = em;                       // workers wait on the priority ring: feed them
                else if (w0 > 0) m1 = 0;                   // workers wait on the plain ring: no backlog, order does not matter
              }
              const unsigned m0 = em & ~m1;
              unsigned pos0 = 0, pos1 = 0;
              if (lane == 0) {
                atomicAdd(&q_ctl[CTL_PENDING], __popc(em));
                if (m0) pos0 = (unsigned)atomicAdd(&q_ctl[CTL_TAIL], __popc(m0));
                if (m1) pos1 = (unsigned)atomicAdd(&q_ctl[CTL_TAIL1], __popc(m1));
              }
              pos0 = __shfl_sync(0xFFFFFFFFu, pos0, 0);
              pos1 = __shfl_sync(0xFFFFFFFFu, pos1, 0);
              if (enq) {
                // the slot is free: at most ntiles <= cap entries are outstanding (a tile is queued at most once), so the entry that
                // used this slot one lap ago was popped, and the slot reset, long ago
                const bool r1 = (m1 >> lane) & 1;
                const unsigned mm = r1 ? m1 : m0;
                const unsigned pos = (r1 ? pos1 : pos0) + (unsigned)__popc(mm & ((1u << lane) - 1u));
                st_volatile_i32(q_queue + (r1 ? q_cap : 0) + (pos % (unsigned)q_cap), nt | tag);
                ++push_local;
              }
            }
            __syncwarp();
          }
          store_rows(rows_changed & ~ring_rows);         // the interior (ordered before this run's release by the fence there)
          SIPP_TICK(4)
        }
        if (last) break;
        // the opposite sweep only if it has something to do (tested on the halo as the publish above left it)
        if (a.check && !((it & 1) ? sweep_needed<+1>(t, wv, lane) : sweep_needed<-1>(t, wv, lane))) break;
      }
      // ---- done with this run: idle again, unless one of our cells was lowered meanwhile
      int again = 0;
      if (lane == 0) {
        // shared ownership: nothing to order here — this thread's own additions to CTL_PENDING precede its subtraction (same address, same
        // thread), the border ring was fenced before the neighbours were told to look, and the interior stores only have to be
        // complete when the kernel is (a warp that re-runs this tile meanwhile starts from older, higher values: still upper bounds,
        // and the stores land as minima)
        if (!(a.mode & 2)) atomicSub(&q_ctl[CTL_PENDING], 1);
        else if (fence_gpu(), atomicCAS(&a.flags[ti], ST_RUN, 0) != ST_RUN) {
          atomicExch(&a.flags[ti], ST_RUN);      // take the request; everything lowered before this point is re-loaded below
          fence_gpu();
          again = 1;
        } else {
          atomicSub(&q_ctl[CTL_PENDING], 1);     // after our pushes were counted: pending reaches 0 only at the global fixed point
        }
      }
      again = __shfl_sync(0xFFFFFFFFu, again, 0);
      SIPP_TICK(4)
      if (!again) break;
    }
    if (ROAM) {      // statistics go to the map the tile belongs to
      const int pushed = __reduce_add_sync(0xFFFFFFFFu, push_local - push0);
      if (lane == 0) {
        if (pushed) atomicAdd(&a.ctl[CTL_PUSHES], pushed);
        atomicAdd(&a.ctl[CTL_ACTIVATIONS], act_local - act0);
        atomicAdd(&a.ctl[CTL_SWEEPS], sweeps_local - sweeps0);
      }
    }
  }
  if (ROAM) return;
#ifdef SIPP_RELAX_PROFILE
  if (lane == 0)
    for (int i = 0; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a0.ctl + 16) + i, (unsigned long long)tp[i]);
#endif
  // statistics
  atomicAdd(&a0.ctl[CTL_PUSHES], push_local);
  if (lane == 0) {
    atomicAdd(&a0.ctl[CTL_ACTIVATIONS], act_local);
    atomicAdd(&a0.ctl[CTL_SWEEPS], sweeps_local);
  }
}
__device__ __forceinline__ void relax_body(const RelaxArgs& a) { relax_body_t<false, true>(a, RoamQ{}); }

// Plan from scratch, first kernel: field <- +inf, empty work queue, zeroed control block and counters; the ingest marks of an abandoned
// warm start are cleared as well (mark16 != nullptr)
struct FieldInitArgs {
  double* field; size_t n;
  int* flags; int* queue; int ntiles, qcap; int* ctl;
  uint4* mark16; size_t n_mark16;     // optional
  int* plan_out;                      // [5], [6] <- 0
};
__device__ __forceinline__ void field_init_body(const FieldInitArgs& a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < a.n; k += stride) a.field[k] = CUDART_INF;
  for (size_t k = i; k < (size_t)a.ntiles; k += stride) a.flags[k] = 0;
  for (size_t k = i; k < (size_t)a.qcap; k += stride) a.queue[k] = -1;
  if (a.mark16)
    for (size_t k = i; k < a.n_mark16; k += stride) a.mark16[k] = make_uint4(0u, 0u, 0u, 0u);
  if (i < 32) a.ctl[i] = 0;
  if (i == 0) { a.plan_out[5] = 0; a.plan_out[6] = 0; }
}
struct FieldSeedArgs {
  double* field; int W, H, sx, sy;
  int* flags; int* queue; int ntx; int* ctl;
  int unit;      // edge of the scheduling unit: a 32 x 32 tile (k_relax) or a 128 x 128 block (k_relax_blk)
};
__device__ __forceinline__ void field_seed_body(const FieldSeedArgs& a) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (a.sx >= 0 && a.sx < a.W && a.sy >= 0 && a.sy < a.H) {
    a.field[(size_t)a.sy * a.W + a.sx] = 0.0;
    const int t = (a.sy / a.unit) * a.ntx + (a.sx / a.unit);
    a.flags[t] = 1;
    a.queue[0] = t;
    a.ctl[CTL_TAIL] = 1;
    a.ctl[CTL_PENDING] = 1;
  }
}


// ---- incremental re-plan -------------------------------------------------------------------------------------------------
// Between two plans only ingested patches change the planner's map, and only cells observed for the first time change cost
// (prm_star_planner.cpp:111-115). The previous fixed point d is reused:
//  (1) cells whose cost went UP (first seen as an obstacle, or at a label above the init cost) were marked by k_ingest_patch.
//      k_inval follows the canonical predecessor plane of the last plan FORWARD from these marks — a cell whose predecessor is
//      marked is marked — and sets d = +inf on everything marked: exactly the cells whose shortest path ran through a cell or
//      edge that got more expensive. Every value that survives is still the (rounded) length of a path whose edges kept or
//      lowered their weight, i.e. an upper bound of the new fixed point;
//  (2) the relaxation restarts from the tiles the patches touched (lower weights) and the tiles in and around the invalidated
//      region; from upper bounds it converges to the same least fixed point as from +inf (the operator is monotone, also in
//      rounded arithmetic), so the field has the bits of a from-scratch plan — tests compare the two.
// k_inval is a persistent tile work-queue kernel like k_relax, one warp per 32x32 tile; the tile is held as bit masks in registers.
constexpr int kInvalWarps = 8;
struct InvalArgs {
  int W, H, ntx, nty;
  uint8_t* mark;            // [H][pitch]
  const uint8_t* pred;      // [H][pitch]
  size_t pitch;
  double* field;
  int start_x, start_y;
  int* flags; int* queue; int cap; int* ctl;
  uint8_t* tile_seed;       // [ntiles] out: tile holds invalidated cells
  int* plan_out;            // [5] <- cells invalidated
  int spin_limit;
};
enum { CTL_MARKED = 7, CTL_EXITED = 8 };

// ROAM: the batch's shared queue (see relax_body_t) — entries carry their map, rq.args + map * rq.stride is that map's InvalArgs
template <bool ROAM>
__device__ __forceinline__ void inval_body_t(const InvalArgs& a0, const RoamQ& rq) {
  int* const q_queue = ROAM ? rq.queue : a0.queue;
  int* const q_ctl = ROAM ? rq.ctl : a0.ctl;
  const int q_cap = ROAM ? rq.cap : a0.cap;
  const int lane = threadIdx.x & 31;
  int marked_local = 0;
  for (;;) {
    int ti = -2;
    if (lane == 0) {
      const unsigned ticket = (unsigned)atomicAdd(&q_ctl[CTL_HEAD], 1);
      int* slot = q_queue + (ticket % (unsigned)q_cap);
      int spins = 0;
      for (;;) {
        const int v = ld_volatile_i32(slot);
        if (v >= 0) { st_volatile_i32(slot, -1); ti = v; break; }
        if (ld_volatile_i32(&q_ctl[CTL_PENDING]) == 0 || ld_volatile_i32(&q_ctl[CTL_ABORT]) != 0) break;
        if (++spins > (ROAM ? rq.spin_limit : a0.spin_limit)) { atomicExch(&q_ctl[CTL_ABORT], 1); break; }
        __nanosleep(64);
      }
      if (ti >= 0) {
        const InvalArgs& am = ROAM ? *reinterpret_cast<const InvalArgs*>(rq.args + (size_t)(ti >> kRoamShift) * rq.stride) : a0;
        atomicExch(&am.flags[ROAM ? (ti & ((1 << kRoamShift) - 1)) : ti], 0);      // queued -> running: a mark arriving from now on queues the tile again
        fence_gpu();
      }
    }
    ti = __shfl_sync(0xFFFFFFFFu, ti, 0);
    if (ti < 0) break;
    const int tag = ROAM ? (ti >> kRoamShift) << kRoamShift : 0;
    if (ROAM) ti &= (1 << kRoamShift) - 1;
    const InvalArgs& a = ROAM ? *reinterpret_cast<const InvalArgs*>(rq.args + (size_t)(tag >> kRoamShift) * rq.stride) : a0;
    const int marked0 = marked_local;
    const int tx = ti % a.ntx, ty = ti / a.ntx;
    const int x0 = tx * TS, y0 = ty * TS;
    // ---- load: lane = tile column (+ the two halo columns by lanes 0 / 1), all rows, every load issued before the first use. The
    // tile then lives in registers as bit masks over the rows of this lane's column: Mi = marked cells (bit r = tile row r), tb = the
    // column's cells in the halo rows (bit 0: row -1, bit 1: row 32), P[k] = cells whose canonical predecessor is neighbour k.
    uint32_t Mi = 0, fresh = 0, tb = 0;      // fresh: cells carrying an unpublished ingest mark (value 1)
    uint32_t P[8];
    uint32_t HLi, HLtb, HRi, HRtb;           // the west / east halo columns, same encoding
    {
      const int gx = x0 + lane;
      const bool colin = gx < a.W;
      const int hx = (lane & 1) ? x0 + TS : x0 - 1;
      const bool hin = lane < 2 && hx >= 0 && hx < a.W;
      uint8_t mv[TP], hv[TP], pv[TS];
#pragma unroll
      for (int r = 0; r < TP; ++r) {
        const int gy = y0 - 1 + r;
        const bool rowin = gy >= 0 && gy < a.H;
        mv[r] = (rowin && colin) ? __ldcg(a.mark + (size_t)gy * a.pitch + gx) : (uint8_t)0;
        hv[r] = (rowin && hin) ? __ldcg(a.mark + (size_t)gy * a.pitch + hx) : (uint8_t)0;
      }
#pragma unroll
      for (int r = 0; r < TS; ++r) {
        const int gy = y0 + r;
        pv[r] = (gy < a.H && colin) ? a.pred[(size_t)gy * a.pitch + gx] : (uint8_t)255;
      }
      uint32_t Hi = 0, Htb = 0;
      tb = (mv[0] != 0 ? 1u : 0u) | (mv[TP - 1] != 0 ? 2u : 0u);
      Htb = (hv[0] != 0 ? 1u : 0u) | (hv[TP - 1] != 0 ? 2u : 0u);
#pragma unroll
      for (int r = 0; r < TS; ++r) {
        Mi |= (mv[r + 1] != 0 ? 1u : 0u) << r;
        fresh |= (mv[r + 1] == 1 ? 1u : 0u) << r;
        Hi |= (hv[r + 1] != 0 ? 1u : 0u) << r;
      }
      uint32_t b0 = 0, b1 = 0, b2 = 0, bv = 0;      // bit planes of the predecessor index, bv: the cell has one
#pragma unroll
      for (int r = 0; r < TS; ++r) {
        const uint32_t pp = pv[r];
        bv |= (pp < 8u ? 1u : 0u) << r;
        b0 |= (pp & 1u) << r;
        b1 |= ((pp >> 1) & 1u) << r;
        b2 |= ((pp >> 2) & 1u) << r;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) P[k] = bv & ((k & 1) ? b0 : ~b0) & ((k & 2) ? b1 : ~b1) & ((k & 4) ? b2 : ~b2);
      HLi = __shfl_sync(0xFFFFFFFFu, Hi, 0); HLtb = __shfl_sync(0xFFFFFFFFu, Htb, 0);
      HRi = __shfl_sync(0xFFFFFFFFu, Hi, 1); HRtb = __shfl_sync(0xFFFFFFFFu, Htb, 1);
    }
    // ---- propagate to the tile-local fixed point: a cell is marked when its predecessor is (k_pred's neighbour order: k = 0..7 =
    // (dx, dy) = (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)). Each round takes the neighbour columns' masks with one shuffle
    // each, ORs in the cells whose predecessor in those columns is marked, then closes the lane's own column (N / S chains) in registers;
    // rounds repeat until no lane changes. The halo (rows -1 / 32, columns -1 / 32) is input only.
    uint32_t tbl = __shfl_up_sync(0xFFFFFFFFu, tb, 1), tbr = __shfl_down_sync(0xFFFFFFFFu, tb, 1);
    if (lane == 0) tbl = HLtb;
    if (lane == 31) tbr = HRtb;
    const uint32_t M0 = Mi;
    auto up = [](uint32_t m, uint32_t halo) { return (m << 1) | (halo & 1u); };            // the column's value one row above each row
    auto dn = [](uint32_t m, uint32_t halo) { return (m >> 1) | ((halo >> 1) << 31); };    // one row below
    for (;;) {
      uint32_t L = __shfl_up_sync(0xFFFFFFFFu, Mi, 1), R = __shfl_down_sync(0xFFFFFFFFu, Mi, 1);
      if (lane == 0) L = HLi;
      if (lane == 31) R = HRi;
      uint32_t Mn = Mi | (P[0] & up(L, tbl)) | (P[1] & L) | (P[2] & dn(L, tbl)) | (P[5] & up(R, tbr)) | (P[6] & R) | (P[7] & dn(R, tbr));
      for (;;) {
        const uint32_t v = ((P[3] & up(Mn, tb)) | (P[4] & dn(Mn, tb))) & ~Mn;
        if (!v) break;
        Mn |= v;
      }
      const bool chg = Mn != Mi;
      Mi = Mn;
      if (!__any_sync(0xFFFFFFFFu, chg)) break;
    }
    const uint32_t newly = Mi & ~M0;          // bit r: this lane's cell in row r was marked here
    // ---- publish: marks (value 2) and d = +inf for every cell marked here or carrying a fresh ingest mark
    const uint32_t pub = newly | fresh;
    {
      const int gx = x0 + lane;
      uint32_t rc = pub;
      while (rc) {
        const int r = __ffs(rc) - 1;
        rc &= rc - 1;
        const int gy = y0 + r;
        a.mark[(size_t)gy * a.pitch + gx] = 2;
        if (!(gx == a.start_x && gy == a.start_y)) a.field[(size_t)gy * a.W + gx] = CUDART_INF;   // the start keeps d = 0
        ++marked_local;
      }
    }
    const uint32_t any_pub = __reduce_or_sync(0xFFFFFFFFu, pub);
    const uint32_t any_mark = __reduce_or_sync(0xFFFFFFFFu, Mi);
    if (lane == 0 && any_mark) a.tile_seed[ti] = 1;
    // neighbours whose cells may hang off a cell published here: the side / corner of the border ring it lies on
    int m = 0;   // N S W E NW NE SW SE
    if (pub & 1u) m |= 1;
    if (pub & (1u << (TS - 1))) m |= 2;
    if (lane == 0 && pub) m |= 4;
    if (lane == 31 && pub) m |= 8;
    if (lane == 0 && (pub & 1u)) m |= 16;
    if (lane == 31 && (pub & 1u)) m |= 32;
    if (lane == 0 && (pub & (1u << (TS - 1)))) m |= 64;
    if (lane == 31 && (pub & (1u << (TS - 1)))) m |= 128;
    m = __reduce_or_sync(0xFFFFFFFFu, (unsigned)m);
    if (m & 16) m |= 1 | 4;      // a corner cell also borders the two side neighbours
    if (m & 32) m |= 1 | 8;
    if (m & 64) m |= 2 | 4;
    if (m & 128) m |= 2 | 8;
    if (any_pub) fence_gpu();
    __syncwarp();
    if (m && lane < 8 && ((m >> lane) & 1)) {
      const int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
      const int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
      const int ntx_ = tx + ddx[lane], nty_ = ty + ddy[lane];
      if (ntx_ >= 0 && ntx_ < a.ntx && nty_ >= 0 && nty_ < a.nty) {
        const int nt = nty_ * a.ntx + ntx_;
        if (atomicOr(&a.flags[nt], ST_QUEUED) == 0) {
          atomicAdd(&q_ctl[CTL_PENDING], 1);
          const unsigned pos = (unsigned)atomicAdd(&q_ctl[CTL_TAIL], 1);
          int* slot = q_queue + (pos % (unsigned)q_cap);
          int spins = 0;
          while (ld_volatile_i32(slot) != -1 && ++spins < a.spin_limit) __nanosleep(32);
          st_volatile_i32(slot, nt | tag);
        }
      }
    }
    __syncwarp();
    if (ROAM) {      // the count of invalidated cells goes to the tile's map
      const int marked = __reduce_add_sync(0xFFFFFFFFu, marked_local - marked0);
      if (lane == 0 && marked) atomicAdd(&a.plan_out[5], marked);
    }
    if (lane == 0) {
      fence_gpu();
      atomicSub(&q_ctl[CTL_PENDING], 1);
    }
  }
  if (ROAM) {      // the last warp to leave rewinds the shared queue (every slot is empty again, the workers' unserved tickets are forgotten)
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      const int total = (int)gridDim.x * kInvalWarps;
      if (atomicAdd(&q_ctl[CTL_EXITED], 1) == total - 1) {
        st_volatile_i32(&q_ctl[CTL_HEAD], 0);
        st_volatile_i32(&q_ctl[CTL_TAIL], 0);
        st_volatile_i32(&q_ctl[CTL_EXITED], 0);
      }
    }
    return;
  }
  const InvalArgs& a = a0;
  if (marked_local) atomicAdd(&a.ctl[CTL_MARKED], marked_local);
  // the last warp to leave hands the queue over to the relaxation: every slot is back to -1 and every flag to 0 (whatever was
  // pushed has been popped), only the counters have to be rewound — no separate reset kernel between the two phases
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    const int total = (int)gridDim.x * kInvalWarps;
    if (atomicAdd(&a.ctl[CTL_EXITED], 1) == total - 1) {
      __threadfence();
      a.plan_out[5] = ld_volatile_i32(&a.ctl[CTL_MARKED]);
      st_volatile_i32(&a.ctl[CTL_HEAD], 0);
      st_volatile_i32(&a.ctl[CTL_TAIL], 0);
      st_volatile_i32(&a.ctl[CTL_PENDING], 0);
    }
  }
}
__device__ __forceinline__ void inval_body(const InvalArgs& a) { inval_body_t<false>(a, RoamQ{}); }

// First kernel of an incremental plan: resets the work queue and seeds it with the tiles the patches touched. `list` (nl tile
// ids, the queue order) and `seedmap` (one byte per tile, the same set) come from the host in one copy; seedmap == nullptr: none.
struct IncBeginArgs {
  const int* list; int nl; const uint8_t* seedmap;
  int ntiles, qcap; int* flags; int* queue; int* ctl;
  uint8_t* tile_seed; int* plan_out;
};
__device__ __forceinline__ void inc_begin_body(const IncBeginArgs& a) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int t = i0; t < a.ntiles; t += stride) {
    const uint8_t sd = a.seedmap ? a.seedmap[t] : (uint8_t)0;
    a.flags[t] = sd ? ST_QUEUED : 0;
    a.tile_seed[t] = sd;
  }
  for (int i = i0; i < a.qcap; i += stride) a.queue[i] = i < a.nl ? a.list[i] : -1;
  if (i0 == 0) {
    for (int k = 0; k < 32; ++k) a.ctl[k] = 0;
    a.ctl[CTL_TAIL] = a.nl;
    a.ctl[CTL_PENDING] = a.nl;
    a.plan_out[5] = 0;      // cells invalidated
    a.plan_out[6] = 0;      // tiles the relaxation starts from
    a.plan_out[7] = 0;      // "redo the path" flag of k_path_check
  }
}
// Between k_inval and k_relax: queue <- every tile that is a seed or touches one (the invalidated region and its rim, the patch
// tiles); the marks of the seed tiles are cleared for the next plan (all marks live in seed tiles).
struct SeedRelaxArgs {
  const uint8_t* tile_seed; int ntx, nty, nbx, nby;
  int* flags; int* queue; int* ctl;
  uint8_t* mark; size_t pitch; int H;
  int* plan_out;
};
__device__ __forceinline__ void seed_relax_body(const SeedRelaxArgs& a) {
  const uint8_t* tile_seed = a.tile_seed;
  const int ntx = a.ntx, nty = a.nty, H = a.H;
  int *flags = a.flags, *queue = a.queue, *ctl = a.ctl, *plan_out = a.plan_out;
  uint8_t* mark = a.mark;
  const size_t pitch = a.pitch;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntx * nty) return;
  const int tx = t % ntx, ty = t / ntx;
  bool on = false;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = tx + dx, y = ty + dy;
      if (x >= 0 && x < ntx && y >= 0 && y < nty && tile_seed[y * ntx + x]) on = true;
    }
  if (tile_seed[t]) {
    for (int r = 0; r < TS; ++r) {
      const int gy = ty * TS + r;
      if (gy >= H) break;
      uint4* row = reinterpret_cast<uint4*>(mark + (size_t)gy * pitch + (size_t)tx * TS);     // pitch is a multiple of 128: aligned, in the row
      row[0] = make_uint4(0u, 0u, 0u, 0u);
      row[1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (!on) return;
  flags[t] = ST_QUEUED;
  atomicAdd(&ctl[CTL_PENDING], 1);
  queue[atomicAdd(&ctl[CTL_TAIL], 1)] = t;
  atomicAdd(&plan_out[6], 1);
}

#include "sipp_relax_blk.cuh"

// ---- canonical backtrack: one warp, lanes 0..7 = the 8 neighbours in ascending node-id order ------------
struct PathArgs {
  FieldGeom g;
  const double* field;
  const uint16_t* pcode;
  size_t pitch16;
  const uint8_t* pred;   // canonical predecessor plane (k_pred)
  size_t pitch;
  int* nodes;       // out, goal -> start, packed (x << 16) | y
  int cap;
  int* out;         // [0] len, [1] all explored, [2] found, [3] unused, [4] walk failed on a reachable goal
  double* cost_out;
  const int* redo;  // optional: 0 = the previous plan's path is still the canonical one (nothing on or next to it changed): keep everything
};

// ---- canonical predecessor of every cell (fully parallel), then the walk over predecessor bytes -------------------------
// pred[y][x] = index k (0..7, neighbours in ascending node id x*H+y) of the canonical predecessor of (x,y): among the u with
// d[u] + w(u,v) == d[v] bit for bit, the one with the smallest f = d[u] + h(u), lowest id on equal f;
// PRED_NONE where there is none (start node, unreachable cells). One thread per cell, the 34x34 window of d / half costs of a
// 32x32 tile staged in shared memory. Bandwidth-bound: 8 B + 2 B read, 1 B written per cell.
constexpr uint8_t PRED_NONE = 255;
constexpr int kPredThreads = 256;
struct PredArgs {
  FieldGeom g;
  const double* field;
  const uint16_t* pcode;
  size_t pitch16;
  uint8_t* pred;     // [H][pitch]
  size_t pitch;
  int ntx, nty;
  const uint8_t* sel;   // optional [ntiles]: recompute a tile only if it or one of its 8 neighbours is flagged (d unchanged elsewhere)
};
struct PredSmem {
  double d[TP][TP + 1];
  double hc[TP][TP + 1];
  double dd[kMaxDd];
  double hd[TP], vd[TP];
  int xc[TP], yc[TP];
};

__device__ __forceinline__ void pred_body(const PredArgs& a) {
  __shared__ PredSmem t;
  const FieldGeom& g = a.g;
  const int tid = threadIdx.x;
  const int tx = blockIdx.x % a.ntx, ty = blockIdx.x / a.ntx;
  const int x0 = tx * TS, y0 = ty * TS;
  if (a.sel) {
    bool on = false;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = tx + dx, y = ty + dy;
        if (x >= 0 && x < a.ntx && y >= 0 && y < a.nty && a.sel[y * a.ntx + x]) on = true;
      }
    if (!on) return;
  }
  const bool dd_in_smem = g.Kx * g.Ky <= kMaxDd;
  if (dd_in_smem)
    for (int i = tid; i < g.Kx * g.Ky; i += kPredThreads) t.dd[i] = g.ddist[i];
  const double* ddp = dd_in_smem ? t.dd : g.ddist;
  {
    // 34 x 34 window: 5 rounds of 256 threads, loads first, stores after (independent loads in flight together)
    double dv[5];
    uint32_t cv[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int idx = tid + k * kPredThreads;
      const int ly = idx / TP, lx = idx - ly * TP;
      const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
      const bool inb = idx < TP * TP && gx >= 0 && gx < g.W && gy >= 0 && gy < g.H;
      dv[k] = inb ? a.field[(size_t)gy * g.W + gx] : CUDART_INF;
      cv[k] = inb ? (uint32_t)a.pcode[(size_t)gy * a.pitch16 + gx] : (uint32_t)PCODE_REMOVED;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int idx = tid + k * kPredThreads;
      if (idx < TP * TP) {
        const int ly = idx / TP, lx = idx - ly * TP;
        t.d[ly][lx] = dv[k];
        t.hc[ly][lx] = half_cost(cv[k], g.init_half_cost);
      }
    }
  }
  if (tid < TP) {
    const int gy = y0 - 1 + tid, gx = x0 - 1 + tid;     // edges (gy, gy+1) and (gx, gx+1)
    const bool oky = gy >= 0 && gy + 1 < g.H, okx = gx >= 0 && gx + 1 < g.W;
    const int cy = oky ? g.ycls[gy] : -1, cx = okx ? g.xcls[gx] : -1;
    t.yc[tid] = cy;
    t.vd[tid] = oky ? g.vdist[cy] : CUDART_INF;
    t.xc[tid] = cx;
    t.hd[tid] = okx ? g.hdist[cx] : CUDART_INF;
  }
  __syncthreads();
  const int NDX[8] = {-1, -1, -1, 0, 0, 1, 1, 1};   // ascending node id = x*H + y
  const int NDY[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
#pragma unroll 1
  for (int c = tid; c < TS * TS; c += kPredThreads) {
    const int lx = (c & (TS - 1)) + 1, ly = (c >> 5) + 1;
    const int gx = x0 + lx - 1, gy = y0 + ly - 1;
    if (gx >= g.W || gy >= g.H) continue;
    const double dv = t.d[ly][lx], hcv = t.hc[ly][lx];
    int best = PRED_NONE;
    if (dv < CUDART_INF && hcv < CUDART_INF) {
      unsigned okmask = 0;
      double du[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ux = lx + NDX[k], uy = ly + NDY[k];
        const int ex = min(lx, ux), ey = min(ly, uy);   // local edge indices (edge e joins local e and e+1)
        double dist;
        if (NDY[k] == 0) dist = t.hd[ex];
        else if (NDX[k] == 0) dist = t.vd[ey];
        else dist = (t.xc[ex] >= 0 && t.yc[ey] >= 0) ? ddp[t.xc[ex] * g.Ky + t.yc[ey]] : CUDART_INF;
        const double hcu = t.hc[uy][ux];
        du[k] = t.d[uy][ux];
        if (dist < CUDART_INF && hcu < CUDART_INF && __dadd_rn(du[k], __dmul_rn(dist, __dadd_rn(hcu, hcv))) == dv) okmask |= 1u << k;
      }
      if (okmask) {
        if ((okmask & (okmask - 1)) == 0) best = __ffs(okmask) - 1;
        else {
          double best_f = CUDART_INF;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (okmask & (1u << k)) {
              const int hx = g.goal_x - (gx + NDX[k]), hy = g.goal_y - (gy + NDY[k]);
              const double hh = __dmul_rn(__dmul_rn(__dsqrt_rn((double)(hx * hx + hy * hy)), g.min_cost), g.spacing);   // prm_star_planner.h:108-112
              const double f = __dadd_rn(du[k], hh);
              if (best == PRED_NONE || f < best_f) { best_f = f; best = k; }     // ascending k: the lowest id wins equal f
            }
        }
      }
    }
    a.pred[(size_t)gy * a.pitch + gx] = (uint8_t)best;
  }
}

// One CTA. The walk runs over a 256x256 window of predecessor bytes held in shared memory, centred on the current node: all
// threads (re)load the window whenever the walk leaves it (>= 112 steps per load), then thread 0 chases the bytes.
constexpr int PW = 256;
constexpr int kBackThreads = 256;
constexpr int kWalk = 112;      // steps per window load: the window is centred on the current node (x0 aligned down to 16)
static_assert(PW == 256, "the walk packs a window cell as (ly << 8) | lx");
__device__ __forceinline__ void backtrack_body(const PathArgs& a) {
  extern __shared__ __align__(16) uint8_t win[];     // [PW][PW]
  __shared__ int s_x, s_y, s_len, s_fail, s_n;
  __shared__ unsigned short s_buf[kWalk];
  const FieldGeom& g = a.g;
  const int tid = threadIdx.x;
  if (a.redo && *a.redo == 0) return;
  const bool goal_ok = g.goal_x >= 0 && g.goal_x < g.W && g.goal_y >= 0 && g.goal_y < g.H;
  const double dgoal = goal_ok ? a.field[(size_t)g.goal_y * g.W + g.goal_x] : CUDART_INF;
  if (!(dgoal < CUDART_INF)) {
    if (tid == 0) { a.out[0] = 0; a.out[1] = 0; a.out[2] = 0; a.out[3] = 0; a.out[4] = 0; *a.cost_out = -1.0; }
    return;
  }
  if (tid == 0) {
    s_x = g.goal_x; s_y = g.goal_y; s_len = 1; s_fail = 0;
    a.nodes[0] = (g.goal_x << 16) | g.goal_y;
  }
  __syncthreads();
  for (;;) {
    int x = s_x, y = s_y, len = s_len;
    bool fail = s_fail != 0;
    if (fail || (x == g.start_x && y == g.start_y)) break;
    // ---- (re)load: window [x0, x0+256) x [y0, y0+256) around the current node, x0 a multiple of 16 (the pred plane's
    // pitch is a multiple of 128, so every 16-byte chunk below is aligned and inside the allocation row)
    const int x0 = (x - PW / 2) & ~15, y0 = y - PW / 2;
    {
      // 16-byte chunk c = tid + 256 k of the window: sixteen consecutive threads read one 256-byte row of the plane (coalesced) and
      // the warp's shared-memory stores are consecutive (thread = window row made every store a 32-way bank conflict: 8 us per load)
      uint4 v[PW / 16];
#pragma unroll
      for (int k = 0; k < PW / 16; ++k) {
        const int c = tid + kBackThreads * k;
        const int gy = y0 + (c >> 4), gx = x0 + 16 * (c & 15);
        const bool inb = gy >= 0 && gy < g.H && gx >= 0 && gx < (int)a.pitch;
        v[k] = inb ? *reinterpret_cast<const uint4*>(a.pred + (size_t)gy * a.pitch + gx) : make_uint4(~0u, ~0u, ~0u, ~0u);
      }
#pragma unroll
      for (int k = 0; k < PW / 16; ++k) *reinterpret_cast<uint4*>(win + (size_t)(tid + kBackThreads * k) * 16) = v[k];
    }
    __syncthreads();
    if (tid == 0) {
      // kWalk steps cannot leave the window (the walk starts >= kWalk cells from every edge and moves one cell per step), so the
      // chase is: byte load -> two shifts for (dx, dy) -> add, with no bounds tests; the visited cells go to a shared buffer and
      // all threads turn them into node records afterwards.
      int li = (y - y0) * PW + (x - x0);
      const int sx = g.start_x - x0, sy = g.start_y - y0;
      const int li_start = ((unsigned)sx < (unsigned)PW && (unsigned)sy < (unsigned)PW) ? sy * PW + sx : -1;
      const int maxn = min(kWalk, a.cap - len);
      int n = 0;
      if (x < 0 || x >= g.W || y < 0 || y >= g.H || maxn <= 0) fail = true;
      while (!fail && n < maxn) {
        const int p = win[li];
        if (p > 7) { fail = true; break; }
        // NDX = {-1,-1,-1,0,0,1,1,1}, NDY = {-1,0,1,-1,1,-1,0,1} as 2-bit fields (value + 1)
        li += (int)((0x9224u >> (2 * p)) & 3u) * PW + (int)((0xA940u >> (2 * p)) & 3u) - (PW + 1);
        s_buf[n++] = (unsigned short)li;
        if (li == li_start) break;
      }
      s_n = n;
      s_fail = fail ? 1 : 0;
    }
    __syncthreads();
    {
      const int n = s_n, len0 = s_len;
      for (int i = tid; i < n; i += kBackThreads) {
        const int li = s_buf[i];
        a.nodes[len0 + i] = ((x0 + (li & (PW - 1))) << 16) | (y0 + (li >> 8));
      }
      __syncthreads();
      if (tid == 0 && n > 0) {
        const int li = s_buf[n - 1];
        s_x = x0 + (li & (PW - 1));
        s_y = y0 + (li >> 8);
        s_len = len0 + n;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const bool fail = s_fail != 0;
    a.out[0] = fail ? 0 : s_len;
    a.out[1] = 1;
    a.out[2] = fail ? 0 : 1;
    a.out[3] = 0;
    a.out[4] = fail ? 1 : 0;      // goal reachable but the predecessor walk failed
    *a.cost_out = fail ? -1.0 : dgoal;
  }
}

// ---- poses, termination test and path-mask rasterisation (map_server_planexp.cpp:902-926) -------------------
struct RasterArgs {
  FieldGeom g;
  const int* nodes;
  const int* out;     // [0] len
  int* out_rw;        // [1] all explored (atomicAnd)
  const uint16_t* pcode;
  size_t pitch16;
  double* path_xy;    // start -> goal
  uint8_t* path;      // mask plane
  size_t pitch;
  double map_width, map_height, mpp_third;
  const int* redo;    // optional: 0 = path unchanged, the mask of the previous plan stands
};

__device__ __forceinline__ void pos2pix(const RasterArgs& a, double px, double py, int* ix, int* iy) {
  *ix = (int)__ddiv_rn(__dmul_rn(px, (double)a.g.W), a.map_width);     // map_server_planexp.cpp:340-342
  *iy = (int)__ddiv_rn(__dmul_rn(py, (double)a.g.H), a.map_height);
}
__device__ __forceinline__ void mark(const RasterArgs& a, int ix, int iy) {
  if (ix >= 0 && ix < a.g.W && iy >= 0 && iy < a.g.H) a.path[(size_t)iy * a.pitch + ix] = 1;
}
__device__ __forceinline__ void node_pose(const FieldGeom& g, int packed, double* px, double* py) {
  const int x = packed >> 16, y = packed & 0xFFFF;
  *px = __dadd_rn(g.off_x, __dmul_rn(g.spacing, (double)x));   // node_id2pose prm_star_planner.h:80
  *py = __dadd_rn(g.off_y, __dmul_rn(g.spacing, (double)y));
}

// Incremental re-plan: does the previous plan's path touch a tile whose predecessor bytes were recomputed (the tile or a neighbour
// was relaxed / invalidated / hit by a patch)? If not, the walk from the goal would read exactly the bytes it read last time.
struct PathCheckArgs {
  const int* nodes; const int* out; const uint8_t* touched; int ntx, nty; int* redo;
};
__device__ __forceinline__ void path_check_body(const PathCheckArgs& a) {
  const int* nodes = a.nodes; const int* out = a.out; const uint8_t* touched = a.touched;
  const int ntx = a.ntx, nty = a.nty;
  int* redo = a.redo;
  const int len = out[0];
  if (blockIdx.x == 0 && threadIdx.x == 0 && (out[2] == 0 || len <= 0)) *redo = 1;     // no path last time: look again
  // grid-stride over the path nodes: a single plan's grid covers the node buffer, a batched one brings a few blocks per map
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
    const int nj = nodes[j];
    const int tx = (nj >> 16) / TS, ty = (nj & 0xFFFF) / TS;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = tx + dx, y = ty + dy;
        if (x >= 0 && x < ntx && y >= 0 && y < nty && touched[y * ntx + x]) { *redo = 1; return; }
      }
  }
}
struct MaskClearArgs {
  uint8_t* path; size_t n16; const int* redo;
};
__device__ __forceinline__ void mask_clear_body(const MaskClearArgs& a) {
  if (a.redo && *a.redo == 0) return;
  uint4* p = reinterpret_cast<uint4*>(a.path);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.n16; k += stride) p[k] = make_uint4(0u, 0u, 0u, 0u);
}

// one path point: j = index start -> goal
__device__ __forceinline__ void path_finish_at(const RasterArgs& a, const int len, const int j) {
  const int nj = a.nodes[len - 1 - j];
  double px, py;
  node_pose(a.g, nj, &px, &py);
  a.path_xy[2 * j] = px;
  a.path_xy[2 * j + 1] = py;
  if (a.pcode[(size_t)(nj & 0xFFFF) * a.pitch16 + (nj >> 16)] == PCODE_UNEXPLORED) atomicAnd(a.out_rw + 1, 0);   // check_termination :183-186
  int ix, iy;
  if (j == 0) {
    pos2pix(a, px, py, &ix, &iy);                                   // :905
    mark(a, ix, iy);
    return;
  }
  double qx, qy;
  node_pose(a.g, a.nodes[len - j], &qx, &qy);                       // previous point
  const double dx = __dadd_rn(px, -qx), dy = __dadd_rn(py, -qy);
  const double distance = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const int samples = (int)__ddiv_rn(distance, a.mpp_third);        // :908
  if (samples > 0) {
    const double dpx = __ddiv_rn(dx, (double)samples), dpy = __ddiv_rn(dy, (double)samples);
    for (int i = 0; i < samples; ++i) {
      pos2pix(a, __dadd_rn(qx, __dmul_rn((double)i, dpx)), __dadd_rn(qy, __dmul_rn((double)i, dpy)), &ix, &iy);
      mark(a, ix, iy);
    }
  } else {
    pos2pix(a, px, py, &ix, &iy);
    mark(a, ix, iy);
  }
}
__device__ __forceinline__ void path_finish_body(const RasterArgs& a) {
  if (a.redo && *a.redo == 0) return;
  const int len = a.out[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) path_finish_at(a, len, j);
}


// ---- kernels: one map per launch (argument block by value) and batched (gridDim.y = maps, argument blocks in device memory) ----------
// Everything one sipp_plan launches for one map, as one block: the host fills it (plan_prepare), a single plan launches the kernels
// with the members by value, a batched plan (sipp_batch_plan: an ensemble of maps that advance together) uploads one block per map
// and launches every kernel ONCE with gridDim.y = number of maps. `on_*` = this plan runs that step.
struct PlanArgsAll {
  int on_init, on_weights, on_inc, on_inval, on_check;
  int relax_grid, inval_grid;
  FieldInitArgs fi;
  FieldSeedArgs fs;
  WeightArgs wa;
  IncBeginArgs ib;
  InvalArgs ia;
  SeedRelaxArgs sr;
  RelaxArgs ra;
  PathCheckArgs pc;
  PredArgs pr;
  PathArgs pa;
  MaskClearArgs mc;
  RasterArgs rs;
};

__global__ void k_field_init(const FieldInitArgs a) { field_init_body(a); }
__global__ void k_field_seed(const FieldSeedArgs a) { field_seed_body(a); }
__global__ void __launch_bounds__(kWeightThreads) k_weights(const WeightArgs a) { weights_body(a); }
__global__ void k_inc_begin(const IncBeginArgs a) { inc_begin_body(a); }
__global__ void __launch_bounds__(kInvalWarps * 32) k_inval(const InvalArgs a) { inval_body(a); }
__global__ void k_seed_relax(const SeedRelaxArgs a) { seed_relax_body(a); }
__global__ void __launch_bounds__(kRelaxWarps * 32) k_relax(const RelaxArgs a) { relax_body(a); }
__global__ void k_path_check(const PathCheckArgs a) { path_check_body(a); }
__global__ void __launch_bounds__(kPredThreads) k_pred(const PredArgs a) { pred_body(a); }
__global__ void __launch_bounds__(kBackThreads) k_backtrack(const PathArgs a) { backtrack_body(a); }
__global__ void k_mask_clear(const MaskClearArgs a) { mask_clear_body(a); }
__global__ void k_path_finish(const RasterArgs a) { path_finish_body(a); }

// batched forms: block (x, y) works for map y. The argument block is copied out of global memory first (the bodies then read
// registers / local memory, as they read the parameter space in the single form).
#define SIPP_BATCH_KERNEL(name, bounds, member, cond, body)                                  \
  __global__ void bounds name(const PlanArgsAll* __restrict__ all) {                         \
    const PlanArgsAll& m = all[blockIdx.y];                                                  \
    if (!(cond)) return;                                                                     \
    const auto a = m.member;                                                                 \
    body(a);                                                                                 \
  }
SIPP_BATCH_KERNEL(k_field_init_b, , fi, m.on_init, field_init_body)
SIPP_BATCH_KERNEL(k_field_seed_b, , fs, m.on_init, field_seed_body)
SIPP_BATCH_KERNEL(k_weights_b, __launch_bounds__(kWeightThreads), wa, m.on_weights, weights_body)
SIPP_BATCH_KERNEL(k_inc_begin_b, , ib, m.on_inc, inc_begin_body)
SIPP_BATCH_KERNEL(k_inval_b, __launch_bounds__(kInvalWarps * 32), ia, m.on_inval && (int)blockIdx.x < m.inval_grid, inval_body)
SIPP_BATCH_KERNEL(k_seed_relax_b, , sr, m.on_inval, seed_relax_body)
SIPP_BATCH_KERNEL(k_relax_b, __launch_bounds__(kRelaxWarps * 32), ra, (int)blockIdx.x < m.relax_grid, relax_body)
SIPP_BATCH_KERNEL(k_path_check_b, , pc, m.on_check, path_check_body)
SIPP_BATCH_KERNEL(k_pred_b, __launch_bounds__(kPredThreads), pr, true, pred_body)
SIPP_BATCH_KERNEL(k_backtrack_b, __launch_bounds__(kBackThreads), pa, true, backtrack_body)
SIPP_BATCH_KERNEL(k_mask_clear_b, , mc, true, mask_clear_body)
SIPP_BATCH_KERNEL(k_path_finish_b, , rs, true, path_finish_body)
#undef SIPP_BATCH_KERNEL

// results of a batched plan: every map's plan_out[0..9] and relaxation control block [0..23] gathered into one array (one copy back)
__global__ void k_gather_plan_out(const PlanArgsAll* __restrict__ all, int* __restrict__ out, int* roam_ctl) {
  const PlanArgsAll& m = all[blockIdx.x];
  const int t = threadIdx.x;
  if (t < 10) out[blockIdx.x * 40 + t] = m.pa.out[t];
  else if (t >= 16 && t < 40) {
    int v = m.ra.ctl[t - 16];
    if (roam_ctl && t - 16 == CTL_ABORT && v == 0) v = roam_ctl[CTL_ABORT];      // the shared queue's watchdog fails every map of the batch
    out[blockIdx.x * 40 + t] = v;
  }
  // the workers left with one unserved ticket each: rewind the shared queue's counters for the next plan (every slot is empty again
  // and pending is zero after a run that was not aborted; an aborted one is reset by the host)
  if (roam_ctl && blockIdx.x == 0 && t == 63 && roam_ctl[CTL_ABORT] == 0) {
    roam_ctl[CTL_HEAD] = 0;
    roam_ctl[CTL_TAIL] = 0;
  }
}

// shared work queue of a batch (RoamQ): block y moves what map y's seeding kernels queued — entries [0, TAIL) of its own queue,
// HEAD is 0 at this point in both plan forms — into the batch's queue, tagged with the map
__global__ void k_roam_collect(const PlanArgsAll* __restrict__ all, RoamQ rq, int inval_only) {
  if (inval_only && !all[blockIdx.x].on_inval) return;      // before the invalidation: only the maps that run it (a map that starts over holds its start tile)
  const RelaxArgs& a = all[blockIdx.x].ra;
  __shared__ unsigned base;
  const int n = ld_volatile_i32(&a.ctl[CTL_TAIL]);
  if (n <= 0) return;
  if (threadIdx.x == 0) {
    base = (unsigned)atomicAdd(&rq.ctl[CTL_TAIL], n);
    atomicAdd(&rq.ctl[CTL_PENDING], n);
  }
  __syncthreads();
  if (threadIdx.x == 0) {      // handed over: a later plan that seeds nothing must not find this count again
    st_volatile_i32(&a.ctl[CTL_TAIL], 0);
    st_volatile_i32(&a.ctl[CTL_PENDING], 0);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = a.queue[i];
    a.queue[i] = -1;
    rq.queue[(base + (unsigned)i) % (unsigned)rq.cap] = v | ((int)blockIdx.x << kRoamShift);
  }
}
__global__ void __launch_bounds__(kRelaxWarps * 32) k_relax_roam(const PlanArgsAll* __restrict__ all, RoamQ rq) {
  relax_body_t<true, true>(all[0].ra, rq);
}
__global__ void __launch_bounds__(kInvalWarps * 32) k_inval_roam(const PlanArgsAll* __restrict__ all, RoamQ rq) {
  inval_body_t<true>(all[0].ia, rq);
}
// weights read from global memory: WARPS x 9 KB of shared memory per CTA, one CTA per SM (16 / 20 / 24 warps: 100 / 90 / 80 registers per thread)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_relax_roam_g(const PlanArgsAll* __restrict__ all, RoamQ rq) {
  relax_body_t<true, false>(all[0].ra, rq);
}
__global__ void k_roam_reset(RoamQ rq) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = i0; i < rq.cap; i += gridDim.x * blockDim.x) rq.queue[i] = -1;
  if (i0 < 32) rq.ctl[i0] = 0;
}

}  // namespace

// host: geometry for the kernels
static FieldGeom make_geom(const sipp_ctx* ctx) {
  FieldGeom g;
  g.W = ctx->W; g.H = ctx->H;
  g.Kx = ctx->Kx; g.Ky = ctx->Ky;
  g.xcls = ctx->d_xcls; g.ycls = ctx->d_ycls;
  g.hdist = ctx->d_hdist; g.vdist = ctx->d_vdist; g.ddist = ctx->d_ddist;
  g.init_half_cost = ctx->init_cost / 2.0;
  g.spacing = ctx->spacing; g.off_x = ctx->off_x; g.off_y = ctx->off_y;
  g.min_cost = ctx->desc.min_rov_cost;
  g.start_x = ctx->start_node[0]; g.start_y = ctx->start_node[1];
  g.goal_x = ctx->goal_node[0]; g.goal_y = ctx->goal_node[1];
  return g;
}

namespace {

constexpr int kSpinLimit = 20 * 1000 * 1000;   // ~ seconds of polling an empty queue: only a bug can get there

struct DevInfo { int num_sms, occ; };
// function attributes and occupancy belong to the device: set up once per device of this process, under a process-wide lock
// (several contexts / devices / host threads per process)
int device_setup(sipp_ctx* ctx, DevInfo* out) {
  static std::mutex dev_mu;
  static struct { bool done; int num_sms, occ; } dev_info[64] = {};
  std::lock_guard<std::mutex> lk(dev_mu);
  auto& di = dev_info[ctx->device & 63];
  if (!di.done) {
    SIPP_CUDA_CHECK(ctx, cudaDeviceGetAttribute(&di.num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRelaxWarps * sizeof(TileSmem))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRelaxWarps * sizeof(TileSmem))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_roam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRelaxWarps * sizeof(TileSmem))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_roam_g<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * sizeof(TileD))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_roam_g<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(20 * sizeof(TileD))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_roam_g<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(24 * sizeof(TileD))));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_blk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkSmem)));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_relax_blk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkSmem)));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_backtrack, cudaFuncAttributeMaxDynamicSharedMemorySize, PW * PW));
    SIPP_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_backtrack_b, cudaFuncAttributeMaxDynamicSharedMemorySize, PW * PW));
    SIPP_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&di.occ, k_relax, kRelaxWarps * 32, kRelaxWarps * sizeof(TileSmem)));
    if (di.occ < 1) return sipp_set_err(ctx, SIPP_ERR_CUDA, "k_relax does not fit on an SM");
    di.done = true;
  }
  out->num_sms = di.num_sms;
  out->occ = di.occ;
  return SIPP_OK;
}

// Host half of a plan: decides how this plan starts (from scratch / incrementally from the patch rectangles since the last plan) and
// fills the argument blocks of every kernel. The seed block of an incremental start ([seed map: one byte per tile, padded to 4][seed
// list: nl tile ids]) is left in ctx->h_seed_stage for the caller to copy to `d_seed` (the map's own buffer, or its slice of a
// batch's staging area). blk: the block-resident relaxation kernel was selected (single plans only).
int plan_prepare(sipp_ctx* ctx, PlanArgsAll& A, uint8_t* d_seed, bool blk, const DevInfo& dev, bool* incremental_out, int* nl_out, int batch_m = 0) {
  memset(&A, 0, sizeof(A));
  const FieldGeom g = make_geom(ctx);
  const size_t ncell = (size_t)ctx->W * ctx->H;
  const int ntiles = ctx->n_tiles_x * ctx->n_tiles_y;
  // ring capacity: every tile can be queued at most once, and every waiting warp must own a distinct slot (ticket % cap)
  const int qcap = ntiles > 16384 ? ntiles : 16384;
  const int nbx = (ctx->W + BS - 1) / BS, nby = (ctx->H + BS - 1) / BS;
  A.wa.g = g;
  A.wa.pcode = ctx->d_pcode;
  A.wa.pitch16 = ctx->pitch16;
  A.wa.tile_w = ctx->d_tile_w;
  A.wa.ntx = ctx->n_tiles_x;
  A.wa.tile_list = nullptr;
  A.wa.n_items = ntiles;

  // ---- incremental start? the tiles the patches since the last plan touched (one cell of margin: a tile's weight block and
  // halo reach one cell into its neighbours)
  std::vector<int> seed_tiles;
  bool incremental = ctx->inc_enabled && ctx->warm_valid && ctx->warm_init_cost == ctx->init_cost;
  if (incremental) {
    std::vector<uint8_t> seen((size_t)ntiles, 0);
    for (size_t r = 0; r + 3 < ctx->pending_rects.size(); r += 4) {
      const int xa = ctx->pending_rects[r] - 1, ya = ctx->pending_rects[r + 1] - 1, xb = ctx->pending_rects[r + 2] + 1, yb = ctx->pending_rects[r + 3] + 1;
      const int ta = (xa < 0 ? 0 : xa) / TS, tb = (xb >= ctx->W ? ctx->W - 1 : xb) / TS;
      const int ua = (ya < 0 ? 0 : ya) / TS, ub = (yb >= ctx->H ? ctx->H - 1 : yb) / TS;
      for (int ty = ua; ty <= ub; ++ty)
        for (int tx = ta; tx <= tb; ++tx) {
          const int t = ty * ctx->n_tiles_x + tx;
          if (!seen[t]) { seen[t] = 1; seed_tiles.push_back(t); }
        }
    }
    if ((int)seed_tiles.size() * 2 > ntiles) incremental = false;     // most of the map changed: start over
  }
  const int nl = incremental ? (int)seed_tiles.size() : 0;
  ctx->plan_info[0] = incremental ? 1 : 0;
  ctx->plan_info[1] = (int64_t)seed_tiles.size();
  *incremental_out = incremental;
  *nl_out = nl;
  ctx->h_seed_stage.clear();

  if (!incremental) {
    A.on_init = 1;
    A.fi.field = ctx->d_field; A.fi.n = ncell;
    A.fi.flags = ctx->d_tile_flags; A.fi.queue = ctx->d_tile_list; A.fi.ntiles = ntiles; A.fi.qcap = 2 * qcap; A.fi.ctl = ctx->d_relax_ctl;      // both rings
    // ingest marks of an abandoned warm start must not leak into a later one
    const bool clear_marks = ctx->warm_valid || !ctx->pending_rects.empty();
    A.fi.mark16 = clear_marks ? reinterpret_cast<uint4*>(ctx->d_mark) : nullptr;
    A.fi.n_mark16 = ctx->pitch * ctx->H / 16;
    A.fi.plan_out = ctx->d_plan_out;
    A.fs.field = ctx->d_field; A.fs.W = ctx->W; A.fs.H = ctx->H; A.fs.sx = g.start_x; A.fs.sy = g.start_y;
    A.fs.flags = ctx->d_tile_flags; A.fs.queue = ctx->d_tile_list; A.fs.ntx = blk ? nbx : ctx->n_tiles_x; A.fs.ctl = ctx->d_relax_ctl;
    A.fs.unit = blk ? BS : TS;
    A.on_weights = blk ? 0 : 1;      // edge weights of every tile (they depend on the planner's map and the init cost, not on d)
  } else {
    // one host->device copy: [seed map: one byte per tile, padded to 4][seed list: nl tile ids]
    const size_t map_bytes = ((size_t)ntiles + 3) & ~(size_t)3;
    uint8_t* d_seedmap = d_seed;
    int* d_list = reinterpret_cast<int*>(d_seedmap + map_bytes);
    if (nl > 0) {
      ctx->h_seed_stage.assign(map_bytes + (size_t)nl * sizeof(int), 0);
      for (int t : seed_tiles) ctx->h_seed_stage[t] = 1;
      memcpy(ctx->h_seed_stage.data() + map_bytes, seed_tiles.data(), (size_t)nl * sizeof(int));
    }
    A.on_inc = 1;
    A.ib.list = d_list; A.ib.nl = nl; A.ib.seedmap = nl > 0 ? d_seedmap : nullptr;
    A.ib.ntiles = ntiles; A.ib.qcap = 2 * qcap; A.ib.flags = ctx->d_tile_flags; A.ib.queue = ctx->d_tile_list; A.ib.ctl = ctx->d_relax_ctl;
    A.ib.tile_seed = ctx->d_tile_seed; A.ib.plan_out = ctx->d_plan_out;
    if (nl > 0) {
      A.wa.tile_list = d_list;
      A.wa.n_items = nl;
      A.on_weights = blk ? 0 : 1;
      A.on_inval = 1;
      InvalArgs& ia = A.ia;
      ia.W = ctx->W; ia.H = ctx->H; ia.ntx = ctx->n_tiles_x; ia.nty = ctx->n_tiles_y;
      ia.mark = ctx->d_mark; ia.pred = ctx->d_pred; ia.pitch = ctx->pitch; ia.field = ctx->d_field;
      ia.start_x = g.start_x; ia.start_y = g.start_y;
      ia.flags = ctx->d_tile_flags; ia.queue = ctx->d_tile_list; ia.cap = qcap; ia.ctl = ctx->d_relax_ctl;
      ia.tile_seed = ctx->d_tile_seed;
      ia.plan_out = ctx->d_plan_out;
      ia.spin_limit = kSpinLimit;
      int igrid = (ntiles + 8 * kInvalWarps - 1) / (8 * kInvalWarps);
      if (igrid > 2 * dev.num_sms) igrid = 2 * dev.num_sms;
      if (igrid * kInvalWarps > qcap) igrid = qcap / kInvalWarps;
      if (batch_m > 0) {      // batched: the maps share the SMs (several of these small CTAs fit on one) — about four CTAs per SM in total
        const int share = 4 * dev.num_sms / batch_m < 4 ? 4 : 4 * dev.num_sms / batch_m;
        if (igrid > share) igrid = share;
      }
      A.inval_grid = igrid;
      SeedRelaxArgs& sr = A.sr;
      sr.tile_seed = ctx->d_tile_seed; sr.ntx = ctx->n_tiles_x; sr.nty = ctx->n_tiles_y; sr.nbx = nbx; sr.nby = nby;
      sr.flags = ctx->d_tile_flags; sr.queue = ctx->d_tile_list; sr.ctl = ctx->d_relax_ctl;
      sr.mark = ctx->d_mark; sr.pitch = ctx->pitch; sr.H = ctx->H; sr.plan_out = ctx->d_plan_out;
    }
  }
  static const int grid_env = getenv("SIPP_RELAX_GRID") ? atoi(getenv("SIPP_RELAX_GRID")) : 0;
  if (!blk) {
    RelaxArgs& ra = A.ra;
    ra.g = g;
    ra.field = ctx->d_field;
    ra.pcode = ctx->d_pcode;
    ra.pitch16 = ctx->pitch16;
    ra.ntx = ctx->n_tiles_x; ra.nty = ctx->n_tiles_y;
    ra.flags = ctx->d_tile_flags;
    ra.queue = ctx->d_tile_list;
    ra.cap = qcap;
    ra.ctl = ctx->d_relax_ctl;
    ra.tile_w = ctx->d_tile_w;
    ra.spin_limit = kSpinLimit;
    static const int mode_env = getenv("SIPP_RELAX_MODE") ? atoi(getenv("SIPP_RELAX_MODE")) : 1;   // measured best at 4096^2 (tools/plan_bench.py)
    ra.mode = mode_env;
    static const int check_env = getenv("SIPP_RELAX_CHECK") ? atoi(getenv("SIPP_RELAX_CHECK")) : 1;
    ra.check = (check_env != 0 && (mode_env & 1)) ? 1 : 0;
    // two-ring (Pape / d'Esopo) order: pays when the queue has a backlog, i.e. on maps with many more tiles than workers (8192^2: a third
    // of the activations, 26.9 -> 16.5 ms); on a small map every entry is taken at once and the extra counter reads only sit on the
    // chain (640 x 480: +6 %). SIPP_RELAX_PAPE=0 / 1 forces it off / on.
    static const int pape_env = getenv("SIPP_RELAX_PAPE") ? atoi(getenv("SIPP_RELAX_PAPE")) : -1;
    const bool pape = pape_env < 0 ? ntiles >= 16384 : pape_env != 0;     // 4096^2 (neutral there) and larger; 2048^2 measured 4 % slower with it
    ra.pape = (pape && !(mode_env & 2)) ? 1 : 0;
    ra.touched = incremental ? ctx->d_tile_seed : nullptr;
    // persistent workers: one CTA per SM for a large map; a small map gets one warp per ~3 tiles so that several relaxations share
    // the GPU instead of queueing for all SMs
    int grid = dev.num_sms * dev.occ;
    const int want = grid_env > 0 ? grid_env : (ntiles + 15) / 16;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    if (grid * kRelaxWarps > qcap) grid = qcap / kRelaxWarps;
    if (batch_m > 0 && incremental) {
      // batched: a persistent CTA holds its SM until its map's queue is drained, so the CTAs of the maps that re-plan incrementally
      // (a handful of tiles each) should all be resident at once — about one wave in total, two CTAs per map at least. A map that
      // starts from scratch keeps the full set of workers: the step lasts as long as its slowest map.
      const int share = dev.num_sms * dev.occ / batch_m < 2 ? 2 : dev.num_sms * dev.occ / batch_m;
      if (grid > share) grid = share;
    }
    A.relax_grid = grid;
  }
  PathArgs& pa = A.pa;
  pa.g = g;
  pa.field = ctx->d_field;
  pa.pcode = ctx->d_pcode;
  pa.pitch16 = ctx->pitch16;
  pa.nodes = ctx->d_path_nodes;
  pa.cap = ctx->path_cap_nodes;
  pa.out = ctx->d_plan_out;
  pa.cost_out = reinterpret_cast<double*>(ctx->d_plan_out + 8);
  pa.pred = ctx->d_pred;
  pa.pitch = ctx->pitch;
  // incremental: predecessor bytes are recomputed around the touched tiles only, and the walk / mask are skipped when the previous
  // path runs nowhere near them (d_plan_out[7] = "redo")
  int* d_redo = ctx->d_plan_out + 7;
  static const bool reuse_env = !(getenv("SIPP_PATH_REUSE") && atoi(getenv("SIPP_PATH_REUSE")) == 0);
  const bool reuse = incremental && reuse_env && !ctx->path_mask_foreign;
  ctx->path_mask_foreign = false;
  pa.redo = reuse ? d_redo : nullptr;
  A.on_check = reuse ? 1 : 0;
  A.pc.nodes = ctx->d_path_nodes; A.pc.out = ctx->d_plan_out; A.pc.touched = ctx->d_tile_seed;
  A.pc.ntx = ctx->n_tiles_x; A.pc.nty = ctx->n_tiles_y; A.pc.redo = d_redo;
  PredArgs& pr = A.pr;
  pr.g = g;
  pr.field = ctx->d_field;
  pr.pcode = ctx->d_pcode;
  pr.pitch16 = ctx->pitch16;
  pr.pred = ctx->d_pred;
  pr.pitch = ctx->pitch;
  pr.ntx = ctx->n_tiles_x; pr.nty = ctx->n_tiles_y;
  pr.sel = incremental ? ctx->d_tile_seed : nullptr;
  // new mask: computeCurrentPathEstimate swaps in a freshly zeroed map (map_server_planexp.cpp:894-899,904)
  A.mc.path = ctx->d_path; A.mc.n16 = ctx->pitch * ctx->H / 16; A.mc.redo = pa.redo;
  RasterArgs& rs = A.rs;
  rs.redo = pa.redo;
  rs.g = g;
  rs.nodes = ctx->d_path_nodes;
  rs.out = ctx->d_plan_out;
  rs.out_rw = ctx->d_plan_out;
  rs.pcode = ctx->d_pcode;
  rs.pitch16 = ctx->pitch16;
  rs.path_xy = ctx->d_path_xy;
  rs.path = ctx->d_path;
  rs.pitch = ctx->pitch;
  rs.map_width = ctx->desc.width;
  rs.map_height = ctx->desc.height;
  rs.mpp_third = ctx->mpp / 3.0;
  return SIPP_OK;
}

}  // namespace

int field_plan(sipp_ctx* ctx, cudaStream_t s) {
  // SIPP_PLAN_TIMING=1: per-phase device times of every plan on stderr (tuning aid)
  static const bool timing = getenv("SIPP_PLAN_TIMING") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (timing) for (auto& e : ev) cudaEventCreate(&e);
  if (timing) cudaEventRecord(ev[0], s);
  if (ctx->ev_plan[2]) cudaEventRecord(ctx->ev_plan[2], s);
  DevInfo dev;
  int rc = device_setup(ctx, &dev);
  if (rc != SIPP_OK) return rc;
  const int ntiles = ctx->n_tiles_x * ctx->n_tiles_y;
  // default: the per-tile kernel (k_relax + edge-weight cache); SIPP_RELAX_KERNEL=1: the block-resident kernel (k_relax_blk), experimental
  static const int relax_kernel_env = getenv("SIPP_RELAX_KERNEL") ? atoi(getenv("SIPP_RELAX_KERNEL")) : 0;
  const bool use_blk = relax_kernel_env != 0;
  const int nbx = (ctx->W + BS - 1) / BS, nby = (ctx->H + BS - 1) / BS, nblocks = nbx * nby;
  if (!use_blk && !ctx->d_tile_w) {      // the weight cache of the per-tile kernel is only allocated when that kernel is selected
    SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
    SIPP_CUDA_CHECK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_tile_w), (size_t)ntiles * kTileWDoubles * sizeof(double)));
  }
  PlanArgsAll A;
  bool incremental = false;
  int nl = 0;
  if ((rc = plan_prepare(ctx, A, reinterpret_cast<uint8_t*>(ctx->d_seed_list), use_blk, dev, &incremental, &nl)) != SIPP_OK) return rc;

  if (!incremental) {
    k_field_init<<<592, 256, 0, s>>>(A.fi);
    k_field_seed<<<1, 32, 0, s>>>(A.fs);
    ctx->launches += 2;
    if (A.on_weights) {
      k_weights<<<ntiles, kWeightThreads, 0, s>>>(A.wa);
      ctx->launches += 1;
    }
    SIPP_CUDA_CHECK(ctx, cudaGetLastError());
  } else {
    if (nl > 0) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_seed_list, ctx->h_seed_stage.data(), ctx->h_seed_stage.size(), cudaMemcpyHostToDevice, s));
    k_inc_begin<<<64, 256, 0, s>>>(A.ib);
    ctx->launches += 1;
    if (nl > 0) {
      if (A.on_weights) {
        k_weights<<<nl, kWeightThreads, 0, s>>>(A.wa);
        ctx->launches += 1;
      }
      k_inval<<<A.inval_grid, kInvalWarps * 32, 0, s>>>(A.ia);
      if (use_blk)
        k_seed_relax_blk<<<(ntiles + 255) / 256, 256, 0, s>>>(ctx->d_tile_seed, ctx->n_tiles_x, ctx->n_tiles_y, nbx, nby, ctx->d_tile_flags, ctx->d_tile_list,
                                                              ctx->d_relax_ctl, ctx->d_mark, ctx->pitch, ctx->H, ctx->d_plan_out);
      else
        k_seed_relax<<<(ntiles + 255) / 256, 256, 0, s>>>(A.sr);
      ctx->launches += 2;
    }
    SIPP_CUDA_CHECK(ctx, cudaGetLastError());
  }
  if (ctx->ev_plan[0]) cudaEventRecord(ctx->ev_plan[0], s);      // sipp_plan_phases: the relaxation kernel alone
  if (use_blk) {
    // block-resident kernel: one persistent CTA (16 warps, the whole 128 x 128 window in shared memory) per SM; a small map (an
    // ensemble of 640x480 maps runs many plans side by side) gets at most one CTA per block
    static const int grid_env = getenv("SIPP_RELAX_GRID") ? atoi(getenv("SIPP_RELAX_GRID")) : 0;
    BlkArgs ba;
    ba.g = A.pa.g;
    ba.field = ctx->d_field;
    ba.pcode = ctx->d_pcode;
    ba.pitch16 = ctx->pitch16;
    ba.nbx = nbx; ba.nby = nby;
    ba.ntx = ctx->n_tiles_x; ba.nty = ctx->n_tiles_y;
    ba.flags = ctx->d_tile_flags;
    ba.queue = ctx->d_tile_list;
    ba.cap = (A.fi.qcap ? A.fi.qcap : A.ib.qcap) / 2;
    ba.ctl = ctx->d_relax_ctl;
    ba.spin_limit = kSpinLimit;
    ba.touched = incremental ? ctx->d_tile_seed : nullptr;
    const float hf = (float)ba.g.init_half_cost;
    ba.init_inexact = ((double)hf != ba.g.init_half_cost) ? 1 : 0;      // the shared-memory half costs are fp32
    int grid = grid_env > 0 ? grid_env : dev.num_sms;
    if (grid > nblocks) grid = nblocks;
    if (grid < 1) grid = 1;
    if (ba.init_inexact) k_relax_blk<true><<<grid, kBlkThreads, sizeof(BlkSmem), s>>>(ba);
    else k_relax_blk<false><<<grid, kBlkThreads, sizeof(BlkSmem), s>>>(ba);
  } else {
    // a plain launch: the workers never wait for each other (no grid barrier; any one warp can drain the queue alone), so CTAs that
    // only become resident later — e.g. while other contexts' kernels hold SMs — just join late. Cooperative launches of different
    // streams do not overlap, which serialised the plans of an ensemble.
    k_relax<<<A.relax_grid, kRelaxWarps * 32, kRelaxWarps * sizeof(TileSmem), s>>>(A.ra);
  }
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->launches += 1;
  if (ctx->ev_plan[1]) cudaEventRecord(ctx->ev_plan[1], s);

  if (timing) cudaEventRecord(ev[1], s);
  if (A.on_check) {
    k_path_check<<<(ctx->path_cap_nodes + 255) / 256, 256, 0, s>>>(A.pc);
    ctx->launches += 1;
  }
  k_pred<<<ntiles, kPredThreads, 0, s>>>(A.pr);
  k_backtrack<<<1, kBackThreads, PW * PW, s>>>(A.pa);
  ctx->launches += 2;
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());

  if (timing) cudaEventRecord(ev[2], s);
  k_mask_clear<<<592, 256, 0, s>>>(A.mc);
  k_path_finish<<<(ctx->path_cap_nodes + 255) / 256, 256, 0, s>>>(A.rs);
  ctx->launches += 2;
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());
  if (ctx->ev_plan[3]) cudaEventRecord(ctx->ev_plan[3], s);
  if (timing) {
    cudaEventRecord(ev[3], s);
    cudaEventSynchronize(ev[3]);
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, ev[0], ev[1]);
    cudaEventElapsedTime(&t12, ev[1], ev[2]);
    cudaEventElapsedTime(&t23, ev[2], ev[3]);
    fprintf(stderr, "[sipp] plan phases: init+relax %.3f ms, backtrack %.3f ms, mask %.3f ms\n", t01, t12, t23);
#ifdef SIPP_RELAX_PROFILE
    {      // -DSIPP_RELAX_PROFILE: the workers' clock64 accounts (k_relax: SIPP_TICK), summed over all warps
      int ctl[32];
      cudaMemcpy(ctl, ctx->d_relax_ctl, sizeof(ctl), cudaMemcpyDeviceToHost);
      unsigned long long tp[6];
      memcpy(tp, ctl + 16, sizeof(tp));
      const double act = ctl[CTL_ACTIVATIONS] > 0 ? ctl[CTL_ACTIVATIONS] : 1;
      fprintf(stderr, "[sipp] relax cycles per activation: pop+wait %.0f  load %.0f  sweeps %.0f  store+cross %.0f  push+release %.0f  (activations %d, sweeps %d; "
              "warp-cycles in total %.3g)\n", tp[0] / act, tp[1] / act, tp[2] / act, tp[3] / act, tp[4] / act, ctl[CTL_ACTIVATIONS], ctl[CTL_SWEEPS],
              (double)(tp[0] + tp[1] + tp[2] + tp[3] + tp[4]));
    }
#endif
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return SIPP_OK;
}

// Batched plan: the maps of `ctxs` (same device, same grid size; every context idle) are planned by ONE sequence of launches, every
// kernel with gridDim.y = m. d_args: m PlanArgsAll blocks; d_seed: m slices of seed_stride bytes; d_out: m x 40 ints (plan_out[0..9] at
// 0, the relaxation control block at 16). The host copies of the argument / seed blocks are built in h_args / h_seed (pinned, at
// least as large). Returns after enqueueing on `s` (the caller synchronises and reads d_out).
size_t field_plan_batch_args_bytes() { return sizeof(PlanArgsAll); }

// roam_queue / roam_cap / roam_ctl: the batch's shared relaxation queue (field_roam_*), nullptr = every map's relaxation on CTAs of its own
size_t field_roam_queue_ints(int ntiles, int m) {
  if (m >= (1 << (31 - kRoamShift)) || ntiles >= (1 << kRoamShift)) return 0;      // the entry's map / tile fields
  size_t cap = 1024;
  while (cap < (size_t)ntiles * m) cap <<= 1;      // a tile is queued at most once: never more entries than tiles
  return cap;
}
int field_roam_reset(sipp_ctx* c0, int* roam_queue, int roam_cap, int* roam_ctl, cudaStream_t s) {
  RoamQ rq{roam_queue, roam_cap, roam_ctl, nullptr, 0, 0};
  k_roam_reset<<<64, 256, 0, s>>>(rq);
  SIPP_CUDA_CHECK(c0, cudaGetLastError());
  return SIPP_OK;
}

int field_plan_batch(sipp_ctx* const* ctxs, int m, void* h_args, void* d_args, uint8_t* h_seed, uint8_t* d_seed, size_t seed_stride, int* d_out,
                     cudaStream_t s, int* launches, const std::function<void(int, const std::function<void(int)>&)>* runner,
                     int* roam_queue, int roam_cap, int* roam_ctl, int roam_share, int roam_kind) {
  if (m <= 0) return SIPP_OK;
  sipp_ctx* c0 = ctxs[0];
  DevInfo dev;
  int rc = device_setup(c0, &dev);
  if (rc != SIPP_OK) return rc;
  PlanArgsAll* A = reinterpret_cast<PlanArgsAll*>(h_args);
  const int ntiles = c0->n_tiles_x * c0->n_tiles_y;
  int any_init = 0, any_inc = 0, any_inval = 0, any_weights_all = 0, max_nl = 0, any_check = 0, max_relax = 1, max_inval = 1, sum_relax = 0, sum_inval = 0;
  for (int i = 0; i < m; ++i) {
    sipp_ctx* ctx = ctxs[i];
    if (!ctx->d_tile_w) {
      SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
      SIPP_CUDA_CHECK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_tile_w), (size_t)ntiles * kTileWDoubles * sizeof(double)));
    }
  }
  std::vector<int> rcs((size_t)m, SIPP_OK);
  const std::function<void(int)> prep = [&](int i) {      // pure host work, no dependence between maps: shared with the batch's helper threads
    sipp_ctx* ctx = ctxs[i];
    bool inc = false;
    int nl = 0;
    if ((rcs[i] = plan_prepare(ctx, A[i], d_seed + (size_t)i * seed_stride, false, dev, &inc, &nl, m)) != SIPP_OK) return;
    if (ctx->h_seed_stage.size() > seed_stride) { rcs[i] = sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "seed block larger than the batch slice"); return; }
    if (!ctx->h_seed_stage.empty()) memcpy(h_seed + (size_t)i * seed_stride, ctx->h_seed_stage.data(), ctx->h_seed_stage.size());
  };
  if (runner) (*runner)(m, prep);
  else for (int i = 0; i < m; ++i) prep(i);
  for (int i = 0; i < m; ++i) {
    if (rcs[i] != SIPP_OK) { if (i) sipp_set_err(c0, rcs[i], "%s", ctxs[i]->err.c_str()); return rcs[i]; }
    any_init |= A[i].on_init;
    any_inc |= A[i].on_inc;
    any_inval |= A[i].on_inval;
    any_check |= A[i].on_check;
    if (A[i].on_weights) {
      if (A[i].on_init) any_weights_all = 1;
      else if (A[i].wa.n_items > max_nl) max_nl = A[i].wa.n_items;
    }
    if (A[i].relax_grid > max_relax) max_relax = A[i].relax_grid;
    // shared queue: the workers are not tied to a map, so the grid follows the batch's tiles (one warp per ~3 tiles, as for a single
    // map), not the per-map shares above — idle workers only wait on their tickets
    sum_relax += roam_queue ? (ntiles + 15) / 16 : A[i].relax_grid;
    if (A[i].inval_grid > max_inval) max_inval = A[i].inval_grid;
    if (A[i].on_inval) sum_inval += A[i].inval_grid;
  }
  SIPP_CUDA_CHECK(c0, cudaMemcpyAsync(d_args, h_args, (size_t)m * sizeof(PlanArgsAll), cudaMemcpyHostToDevice, s));
  if (any_inval) SIPP_CUDA_CHECK(c0, cudaMemcpyAsync(d_seed, h_seed, (size_t)m * seed_stride, cudaMemcpyHostToDevice, s));
  const PlanArgsAll* dA = reinterpret_cast<const PlanArgsAll*>(d_args);
  const unsigned M = (unsigned)m;
  int n = 0;
  if (any_init) {
    k_field_init_b<<<dim3(64, M), 256, 0, s>>>(dA);
    k_field_seed_b<<<dim3(1, M), 32, 0, s>>>(dA);
    n += 2;
  }
  if (any_inc) { k_inc_begin_b<<<dim3(16, M), 256, 0, s>>>(dA); n += 1; }
  const int wgrid = any_weights_all ? ntiles : max_nl;
  if (wgrid > 0) { k_weights_b<<<dim3(wgrid, M), kWeightThreads, 0, s>>>(dA); n += 1; }
  if (any_inval && roam_queue) {
    RoamQ rq{roam_queue, roam_cap, roam_ctl, reinterpret_cast<const char*>(d_args) + offsetof(PlanArgsAll, ia), sizeof(PlanArgsAll), kSpinLimit};
    int grid = 4 * dev.num_sms / (roam_share > 1 ? roam_share : 1);      // small CTAs (no shared memory): several per SM
    if (grid > sum_inval) grid = sum_inval;
    if (grid < 1) grid = 1;
    k_roam_collect<<<M, 256, 0, s>>>(dA, rq, 1);
    k_inval_roam<<<grid, kInvalWarps * 32, 0, s>>>(dA, rq);
    k_seed_relax_b<<<dim3((ntiles + 255) / 256, M), 256, 0, s>>>(dA);
    n += 3;
  } else if (any_inval) {
    k_inval_b<<<dim3(max_inval, M), kInvalWarps * 32, 0, s>>>(dA);
    k_seed_relax_b<<<dim3((ntiles + 255) / 256, M), 256, 0, s>>>(dA);
    n += 2;
  }
  if (roam_queue) {
    RoamQ rq{roam_queue, roam_cap, roam_ctl, reinterpret_cast<const char*>(d_args) + offsetof(PlanArgsAll, ra), sizeof(PlanArgsAll), kSpinLimit};
    int grid = dev.num_sms * dev.occ / (roam_share > 1 ? roam_share : 1);      // batches that run side by side split the SMs
    if (grid > sum_relax) grid = sum_relax;
    if (grid < 1) grid = 1;
    k_roam_collect<<<M, 256, 0, s>>>(dA, rq, 0);
    if (roam_kind == 2) {      // weights from global memory, one CTA of 16 / 20 / 24 warps per SM (SIPP_ROAM_WARPS)
      static const int warps_env = getenv("SIPP_ROAM_WARPS") ? atoi(getenv("SIPP_ROAM_WARPS")) : 20;
      const int warps = warps_env >= 24 ? 24 : warps_env >= 20 ? 20 : 16;
      int grid_g = dev.num_sms / (roam_share > 1 ? roam_share : 1);
      const int want_g = (sum_relax * kRelaxWarps + warps - 1) / warps;
      if (grid_g > want_g) grid_g = want_g;
      if (grid_g < 1) grid_g = 1;
      if (warps == 24) k_relax_roam_g<24><<<grid_g, 24 * 32, 24 * sizeof(TileD), s>>>(dA, rq);
      else if (warps == 20) k_relax_roam_g<20><<<grid_g, 20 * 32, 20 * sizeof(TileD), s>>>(dA, rq);
      else k_relax_roam_g<16><<<grid_g, 16 * 32, 16 * sizeof(TileD), s>>>(dA, rq);
    } else {
      k_relax_roam<<<grid, kRelaxWarps * 32, kRelaxWarps * sizeof(TileSmem), s>>>(dA, rq);
    }
    n += 1;
  } else {
    k_relax_b<<<dim3(max_relax, M), kRelaxWarps * 32, kRelaxWarps * sizeof(TileSmem), s>>>(dA);
  }
  if (any_check) { k_path_check_b<<<dim3(4, M), 256, 0, s>>>(dA); n += 1; }
  k_pred_b<<<dim3(ntiles, M), kPredThreads, 0, s>>>(dA);
  k_backtrack_b<<<dim3(1, M), kBackThreads, PW * PW, s>>>(dA);
  k_mask_clear_b<<<dim3(32, M), 256, 0, s>>>(dA);
  k_path_finish_b<<<dim3(4, M), 256, 0, s>>>(dA);
  k_gather_plan_out<<<M, 64, 0, s>>>(dA, d_out, roam_queue ? roam_ctl : nullptr);
  n += 6;
  SIPP_CUDA_CHECK(c0, cudaGetLastError());
  if (launches) *launches = n;
  return SIPP_OK;
}
