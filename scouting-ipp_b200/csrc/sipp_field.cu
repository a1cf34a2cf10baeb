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
// Relaxation = block-asynchronous Bellman-Ford on 32x32 tiles: a persistent cooperative kernel keeps a
// ballot-compacted list of active tiles; a CTA iterates its tile in shared memory to a local fixed point, writes
// it back and activates the neighbours whose halo it changed; rounds are separated by grid barriers.
#include <cooperative_groups.h>
#include <math_constants.h>

#include <cmath>
#include <cstring>

#include "sipp_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int TS = 32;        // tile edge (cells)
constexpr int TP = TS + 2;    // with a one-cell halo
constexpr int kRelaxThreads = 256;
constexpr int kMaxDd = 1024;  // diagonal class table entries kept in shared memory

struct RelaxArgs {
  FieldGeom g;
  double* field;            // [H][W]
  const uint16_t* pcode;    // [H][pitch16]
  size_t pitch16;
  int ntx, nty;
  uint8_t* flags;           // [2][ntiles]
  int* list;                // [ntiles]
  int* ctl;                 // [0],[1] list counts by round parity, [2] rounds, [3] activations, [4] sweeps, [5] aborted
  int max_rounds;
};

__device__ __forceinline__ double ld_relaxed(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double half_cost(uint32_t code, double init_half) {
  if (code == PCODE_UNEXPLORED) return init_half;
  if (code == PCODE_REMOVED) return CUDART_INF;
  return 0.5 * (double)code;
}

// one relaxation of cell (lx,ly) (tile-local coordinates incl. halo) against its 8 neighbours
__device__ __forceinline__ double relax_cell(const volatile double (*sd)[TP], const double (*shc)[TP], const double (*sq)[TP],
                                             const double* s_hd, const double* s_vd, int lx, int ly, double hc0) {
  double best = sd[ly][lx];
#define RELAX(nx, ny, dist)                                                              \
  {                                                                                      \
    const double cand = __dadd_rn(sd[ny][nx], __dmul_rn((dist), __dadd_rn(hc0, shc[ny][nx]))); \
    if (cand < best) best = cand;                                                        \
  }
  RELAX(lx - 1, ly, s_hd[lx - 1]);
  RELAX(lx + 1, ly, s_hd[lx]);
  RELAX(lx, ly - 1, s_vd[ly - 1]);
  RELAX(lx, ly + 1, s_vd[ly]);
  RELAX(lx - 1, ly - 1, sq[ly - 1][lx - 1]);
  RELAX(lx + 1, ly - 1, sq[ly - 1][lx]);
  RELAX(lx - 1, ly + 1, sq[ly][lx - 1]);
  RELAX(lx + 1, ly + 1, sq[ly][lx]);
#undef RELAX
  return best;
}

__global__ void __launch_bounds__(kRelaxThreads)
k_relax(const RelaxArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sd_[TP][TP];
  __shared__ double shc[TP][TP];
  __shared__ double sq[TP][TP];      // sq[y][x]: diagonal length of the quad whose top-left cell is (x,y)
  __shared__ double s_hd[TP], s_vd[TP];
  __shared__ int s_xc[TP], s_yc[TP];
  __shared__ double s_dd[kMaxDd];
  __shared__ int s_mask;
  volatile double (*sd)[TP] = sd_;

  const FieldGeom& g = a.g;
  const int ntiles = a.ntx * a.nty;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = kRelaxThreads / 32;
  const bool dd_in_smem = g.Kx * g.Ky <= kMaxDd;
  if (dd_in_smem)
    for (int i = threadIdx.x; i < g.Kx * g.Ky; i += blockDim.x) s_dd[i] = g.ddist[i];
  const double* ddp = dd_in_smem ? s_dd : g.ddist;
  __syncthreads();

  int sweeps_local = 0, act_local = 0;
  int round = 0;
  for (;; ++round) {
    const int par = round & 1;
    uint8_t* flags_cur = a.flags + (size_t)par * ntiles;
    uint8_t* flags_next = a.flags + (size_t)(par ^ 1) * ntiles;
    // ---- (1) frontier compaction with warp ballots
    for (int base = (blockIdx.x * nwarps + warp) * 32; base < ntiles; base += gridDim.x * nwarps * 32) {
      const int t = base + lane;
      const int f = t < ntiles ? flags_cur[t] : 0;
      if (f) flags_cur[t] = 0;
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, f);
      if (ballot) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&a.ctl[par], __popc(ballot));
        pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
        if (f) a.list[pos + __popc(ballot & ((1u << lane) - 1u))] = t;
      }
    }
    grid.sync();
    const int count = *reinterpret_cast<volatile int*>(&a.ctl[par]);
    if (count == 0 || round >= a.max_rounds) {
      if (count != 0 && blockIdx.x == 0 && threadIdx.x == 0) a.ctl[5] = 1;
      break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl[par ^ 1] = 0;   // next round's counter; nobody touches it before the next barrier

    // ---- (2) relax the active tiles
    for (int li = blockIdx.x; li < count; li += gridDim.x) {
      const int t = a.list[li];
      const int tx = t % a.ntx, ty = t / a.ntx;
      const int x0 = tx * TS, y0 = ty * TS;
      if (threadIdx.x == 0) s_mask = 0;
      for (int idx = threadIdx.x; idx < TP * TP; idx += blockDim.x) {
        const int lx = idx % TP, ly = idx / TP;
        const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        const bool inb = gx >= 0 && gx < g.W && gy >= 0 && gy < g.H;
        sd[ly][lx] = inb ? ld_relaxed(a.field + (size_t)gy * g.W + gx) : CUDART_INF;
        shc[ly][lx] = inb ? half_cost(a.pcode[(size_t)gy * a.pitch16 + gx], g.init_half_cost) : CUDART_INF;
      }
      if (threadIdx.x < TP) {
        const int c = threadIdx.x;
        const int gx = x0 - 1 + c;   // edge between global column gx and gx+1
        const bool ok = gx >= 0 && gx + 1 < g.W;
        const int cls = ok ? g.xcls[gx] : -1;
        s_xc[c] = cls;
        s_hd[c] = ok ? g.hdist[cls] : CUDART_INF;
      } else if (threadIdx.x >= 64 && threadIdx.x < 64 + TP) {
        const int r = threadIdx.x - 64;
        const int gy = y0 - 1 + r;
        const bool ok = gy >= 0 && gy + 1 < g.H;
        const int cls = ok ? g.ycls[gy] : -1;
        s_yc[r] = cls;
        s_vd[r] = ok ? g.vdist[cls] : CUDART_INF;
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < TP * TP; idx += blockDim.x) {
        const int lx = idx % TP, ly = idx / TP;
        const int xc = s_xc[lx], yc = s_yc[ly];
        sq[ly][lx] = (xc >= 0 && yc >= 0) ? ddp[xc * g.Ky + yc] : CUDART_INF;
      }
      __syncthreads();

      // chaotic in-place sweeps to the tile-local fixed point; thread owns column ix, rows iy0 + 8k
      const int ix = threadIdx.x & 31, iy0 = threadIdx.x >> 5;
      int mask = 0;
      bool any = false;
      int changed;
      do {
        changed = 0;
#pragma unroll
        for (int k = 0; k < TS / 8; ++k) {
          const int iy = iy0 + 8 * k;
          const int lx = ix + 1, ly = iy + 1;
          const double hc0 = shc[ly][lx];
          if (hc0 == CUDART_INF) continue;     // removed node: no edges
          const double cur = sd[ly][lx];
          const double best = relax_cell(sd, shc, sq, s_hd, s_vd, lx, ly, hc0);
          if (best < cur) {
            sd[ly][lx] = best;
            changed = 1;
            if (iy == 0) mask |= 1;            // N
            if (iy == TS - 1) mask |= 2;       // S
            if (ix == 0) mask |= 4;            // W
            if (ix == TS - 1) mask |= 8;       // E
          }
        }
        any |= (changed != 0);
        ++sweeps_local;
        changed = __syncthreads_or(changed);
      } while (changed);
      ++act_local;

      // write back what this thread lowered (its own cells only), then publish
      if (any) {
#pragma unroll
        for (int k = 0; k < TS / 8; ++k) {
          const int iy = iy0 + 8 * k;
          const int gx = x0 + ix, gy = y0 + iy;
          if (gx < g.W && gy < g.H) st_relaxed(a.field + (size_t)gy * g.W + gx, sd[iy + 1][ix + 1]);
        }
        __threadfence();
      }
      if (mask) atomicOr(&s_mask, mask);
      __syncthreads();
      if (threadIdx.x < 8) {
        // neighbours: N S W E NW NE SW SE. A corner cell sits on two edges, so "N and W changed" may over-activate the
        // NW tile when different cells set the two bits — harmless (an activation with nothing to do converges in one sweep).
        const int m = s_mask;
        const int k = threadIdx.x;
        const int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
        const int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
        const int need[8] = {1, 2, 4, 8, 1 | 4, 1 | 8, 2 | 4, 2 | 8};
        const bool on = (k < 4) ? ((m & need[k]) != 0) : ((m & need[k]) == need[k]);
        const int ntx_ = tx + ddx[k], nty_ = ty + ddy[k];
        if (on && ntx_ >= 0 && ntx_ < a.ntx && nty_ >= 0 && nty_ < a.nty) flags_next[nty_ * a.ntx + ntx_] = 1;
      }
      __syncthreads();
    }
    grid.sync();
  }
  // statistics
  if (lane == 0 && warp == 0) {
    atomicAdd(&a.ctl[3], act_local);
    atomicAdd(&a.ctl[4], sweeps_local);
    if (blockIdx.x == 0) a.ctl[2] = round;
  }
}

__global__ void k_field_init(double* field, size_t n, int W, int sx, int sy, int H, uint8_t* flags, int ntiles2, int ntx, int* ctl) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < n; k += stride) field[k] = CUDART_INF;
  for (size_t k = i; k < (size_t)ntiles2; k += stride) flags[k] = 0;
  if (i < 8) ctl[i] = 0;
}
__global__ void k_field_seed(double* field, int W, int H, int sx, int sy, uint8_t* flags, int ntx) {
  if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
    field[(size_t)sy * W + sx] = 0.0;
    flags[(sy / TS) * ntx + (sx / TS)] = 1;
  }
}

// ---- canonical backtrack: one warp, lanes 0..7 = the 8 neighbours in ascending node-id order ------------
struct PathArgs {
  FieldGeom g;
  const double* field;
  const uint16_t* pcode;
  size_t pitch16;
  int* nodes;       // out, goal -> start, packed (x << 16) | y
  int cap;
  int* out;         // [0] len, [1] all explored, [2] found, [3] unused, [4] walk failed on a reachable goal
  double* cost_out;
};

__device__ __forceinline__ double edge_len(const FieldGeom& g, int ax, int ay, int bx, int by) {
  const int mx = min(ax, bx), my = min(ay, by);
  if (ay == by) return g.hdist[g.xcls[mx]];
  if (ax == bx) return g.vdist[g.ycls[my]];
  return g.ddist[g.xcls[mx] * g.Ky + g.ycls[my]];
}

__global__ void k_backtrack(const PathArgs a) {
  const FieldGeom& g = a.g;
  const int lane = threadIdx.x;
  int x = g.goal_x, y = g.goal_y;
  const bool goal_ok = x >= 0 && x < g.W && y >= 0 && y < g.H;
  const double dgoal = goal_ok ? a.field[(size_t)y * g.W + x] : CUDART_INF;
  if (!(dgoal < CUDART_INF)) {
    if (lane == 0) { a.out[0] = 0; a.out[1] = 0; a.out[2] = 0; a.out[3] = 0; a.out[4] = 0; *a.cost_out = -1.0; }
    return;
  }
  const int NDX[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  const int NDY[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
  int len = 0;
  if (lane == 0) a.nodes[0] = (x << 16) | y;
  len = 1;
  bool fail = false;
  while (!(x == g.start_x && y == g.start_y)) {
    const double dv = a.field[(size_t)y * g.W + x];
    const double hcv = half_cost(a.pcode[(size_t)y * a.pitch16 + x], g.init_half_cost);
    bool ok = false;
    double f = CUDART_INF;
    if (lane < 8) {
      const int ux = x + NDX[lane], uy = y + NDY[lane];
      if (ux >= 0 && ux < g.W && uy >= 0 && uy < g.H) {
        const double dist = edge_len(g, ux, uy, x, y);
        const double hcu = half_cost(a.pcode[(size_t)uy * a.pitch16 + ux], g.init_half_cost);
        if (dist < CUDART_INF && hcu < CUDART_INF && hcv < CUDART_INF) {
          const double du = a.field[(size_t)uy * g.W + ux];
          const double w = __dmul_rn(dist, __dadd_rn(hcu, hcv));
          if (__dadd_rn(du, w) == dv) {
            ok = true;
            const int gx = g.goal_x - ux, gy = g.goal_y - uy;
            const double hh = __dmul_rn(__dmul_rn(__dsqrt_rn((double)(gx * gx + gy * gy)), g.min_cost), g.spacing);   // prm_star_planner.h:108-112
            f = __dadd_rn(du, hh);
          }
        }
      }
    }
    // min f, lowest lane on ties
    int best_lane = ok ? lane : 99;
    double best_f = f;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      const double of = __shfl_xor_sync(0xFFFFFFFFu, best_f, off);
      const int ol = __shfl_xor_sync(0xFFFFFFFFu, best_lane, off);
      const bool take = (ol != 99) && (best_lane == 99 || of < best_f || (of == best_f && ol < best_lane));
      if (take) { best_f = of; best_lane = ol; }
    }
    best_lane = __shfl_sync(0xFFFFFFFFu, best_lane, 0);
    if (best_lane == 99) { fail = true; break; }
    x += NDX[best_lane];
    y += NDY[best_lane];
    if (len >= a.cap) { fail = true; break; }
    if (lane == 0) a.nodes[len] = (x << 16) | y;
    ++len;
  }
  if (lane == 0) {
    a.out[0] = fail ? 0 : len;
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

__global__ void k_path_finish(const RasterArgs a) {
  const int len = a.out[0];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;   // path point index, start -> goal
  if (j >= len) return;
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

int field_plan(sipp_ctx* ctx, cudaStream_t s) {
  const FieldGeom g = make_geom(ctx);
  const size_t ncell = (size_t)ctx->W * ctx->H;
  const int ntiles = ctx->n_tiles_x * ctx->n_tiles_y;
  k_field_init<<<592, 256, 0, s>>>(ctx->d_field, ncell, ctx->W, g.start_x, g.start_y, ctx->H, ctx->d_tile_flags, 2 * ntiles,
                                   ctx->n_tiles_x, ctx->d_relax_ctl);
  k_field_seed<<<1, 1, 0, s>>>(ctx->d_field, ctx->W, ctx->H, g.start_x, g.start_y, ctx->d_tile_flags, ctx->n_tiles_x);
  ctx->launches += 2;
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());

  static int num_sms = 0, occ = 0;
  if (!num_sms) {
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
    SIPP_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_relax, kRelaxThreads, 0));
    if (occ < 1) return sipp_set_err(ctx, SIPP_ERR_CUDA, "k_relax does not fit on an SM");
  }
  RelaxArgs ra;
  ra.g = g;
  ra.field = ctx->d_field;
  ra.pcode = ctx->d_pcode;
  ra.pitch16 = ctx->pitch16;
  ra.ntx = ctx->n_tiles_x; ra.nty = ctx->n_tiles_y;
  ra.flags = ctx->d_tile_flags;
  ra.list = ctx->d_tile_list;
  ra.ctl = ctx->d_relax_ctl;
  ra.max_rounds = 8 * ntiles + 1024;
  int grid = num_sms * occ;
  if (grid > ntiles) grid = ntiles < 1 ? 1 : ntiles;
  void* kargs[] = {(void*)&ra};
  SIPP_CUDA_CHECK(ctx, cudaLaunchCooperativeKernel((void*)k_relax, dim3(grid), dim3(kRelaxThreads), kargs, 0, s));
  ctx->launches += 1;

  PathArgs pa;
  pa.g = g;
  pa.field = ctx->d_field;
  pa.pcode = ctx->d_pcode;
  pa.pitch16 = ctx->pitch16;
  pa.nodes = ctx->d_path_nodes;
  pa.cap = ctx->path_cap_nodes;
  pa.out = ctx->d_plan_out;
  pa.cost_out = reinterpret_cast<double*>(ctx->d_plan_out + 8);
  k_backtrack<<<1, 32, 0, s>>>(pa);
  ctx->launches += 1;
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());

  // new mask: computeCurrentPathEstimate swaps in a freshly zeroed map (map_server_planexp.cpp:894-899,904)
  SIPP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_path, 0, ctx->pitch * ctx->H, s));
  RasterArgs rs;
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
  const int blocks = (ctx->path_cap_nodes + 255) / 256;
  k_path_finish<<<blocks, 256, 0, s>>>(rs);
  ctx->launches += 2;
  SIPP_CUDA_CHECK(ctx, cudaGetLastError());
  return SIPP_OK;
}
