// Trajectory cost as a line integral of the map cost: "CostPlanExpMapIntegral"
// (advanced_planner_modules/src/cost_computer/cost_planexp_map_integral.cpp:21-73). One thread per trajectory segment: the
// reference accumulates the sample position and the cost sequentially (end += direction * step), so the samples of one segment
// form a dependent chain — it is kept in order, with _rn arithmetic (no contraction), and the result has the reference's bits.
// The parallelism is across segments (thousands of tree nodes per pass); the label plane is L2 resident.
#include "sipp_internal.h"

namespace {

struct CostArgs {
  int n;
  const int32_t* off;
  const double* pts;
  const double* parent_cost;
  double* cost;
  uint8_t* valid;
  const uint16_t* label;   // map_cost_: last observed label, 0 = never observed
  size_t pitch16;
  int W, H;
  double map_width, map_height, mpp, mean_cost, msd;
  int accumulate;
};

// MapServerPlanExp::getCostAtLocation (map_server_planexp.cpp:757-762)
__device__ __forceinline__ double cost_at(const CostArgs& a, double x, double y) {
  if (0 <= x && 0 <= y && a.map_width > x && a.map_height > y) {
    int r = (int)__ddiv_rn(y, a.mpp), c = (int)__ddiv_rn(x, a.mpp);
    r = min(r, a.H - 1);      // cannot happen for the cell sizes sipp_create accepts (index math verified exact); keeps the load in range
    c = min(c, a.W - 1);
    return (double)a.label[(size_t)r * a.pitch16 + c];
  }
  return -1.0;
}

__global__ void k_cost_integral(const CostArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  const int p0 = a.off[s], np = a.off[s + 1] - p0;
  if (np < 2) {
    a.cost[s] = 0.0;
    if (a.valid) a.valid[s] = 0;
    return;
  }
  const double* P = a.pts + 3 * (size_t)p0;
  double cost = 0.0;
  double lx = P[0], ly = P[1], lz = P[2];
  for (int i = 1; i < np; ++i) {
    const double px = P[3 * i], py = P[3 * i + 1], pz = P[3 * i + 2];
    const double dx = __dadd_rn(px, -lx), dy = __dadd_rn(py, -ly), dz = __dadd_rn(pz, -lz);
    const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));   // Eigen norm(): sqrt(x2 + y2 + z2)
    double seg_cost;
    if (dist > 0.005) {
      const double ux = __ddiv_rn(dx, dist), uy = __ddiv_rn(dy, dist), uz = __ddiv_rn(dz, dist);
      const int n_sub = (int)__ddiv_rn(dist, a.msd);
      seg_cost = 0.0;
      double ex = lx, ey = ly, ez = lz;
      double ws = cost_at(a, lx, ly);                   // weight of the sub-segment's start = the previous end (same point, same lookup)
      for (int j = 0; j <= n_sub; ++j) {
        const double sub = (j == n_sub) ? __dadd_rn(dist, -__dmul_rn((double)n_sub, a.msd)) : a.msd;
        ex = __dadd_rn(ex, __dmul_rn(ux, sub));
        ey = __dadd_rn(ey, __dmul_rn(uy, sub));
        ez = __dadd_rn(ez, __dmul_rn(uz, sub));
        const double we = cost_at(a, ex, ey);
        double sc = __ddiv_rn(__dmul_rn(sub, __dadd_rn(ws, we)), 2.0);
        if (ws == 0 || we == 0) sc = __dmul_rn(sub, a.mean_cost);
        seg_cost = __dadd_rn(seg_cost, sc);
        ws = we;
      }
      (void)ez;
    } else {
      const double w = cost_at(a, px, py);
      seg_cost = (w == 0) ? __dmul_rn(dist, a.mean_cost) : __dmul_rn(dist, w);
    }
    cost = __dadd_rn(cost, seg_cost);
    lx = px; ly = py; lz = pz;
  }
  if (a.accumulate && a.parent_cost) cost = __dadd_rn(cost, a.parent_cost[s]);
  a.cost[s] = cost;
  if (a.valid) a.valid[s] = 1;
}

// ---- MavSimulator::computeCost (mav_simulator/src/mav_simulator.cpp:785-821) over the ground-truth label plane -----------------------
struct TravelArgs {
  int n;
  const double* start;   // n x 3
  const double* goal;    // n x 3
  double* cost;
  const uint16_t* truth;
  size_t pitch16;
  int W, H;
  double map_width, map_height, max_step, speed;
  int formulation;
};

// position2pixelLocation (:582-587): (int)(p * cols / width); a pixel outside the map reads label 0
__device__ __forceinline__ int truth_at(const TravelArgs& a, double x, double y) {
  const double fx = __ddiv_rn(__dmul_rn(x, (double)a.W), a.map_width), fy = __ddiv_rn(__dmul_rn(y, (double)a.H), a.map_height);
  if (!(fx > -2.0e9 && fx < 2.0e9 && fy > -2.0e9 && fy < 2.0e9)) return 0;
  const int ix = (int)fx, iy = (int)fy;
  if (ix < 0 || ix >= a.W || iy < 0 || iy >= a.H) return 0;
  return (int)a.truth[(size_t)iy * a.pitch16 + ix];
}

__global__ void k_travel_cost(const TravelArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  const double sx = a.start[3 * s], sy = a.start[3 * s + 1], sz = a.start[3 * s + 2];
  const double dx = __dadd_rn(a.goal[3 * s], -sx), dy = __dadd_rn(a.goal[3 * s + 1], -sy), dz = __dadd_rn(a.goal[3 * s + 2], -sz);
  const double distance = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));      // :788
  double cost = 0.0;
  const double q = __ddiv_rn(distance, a.max_step);
  const int n_elements = (q < 2.0e9) ? (int)ceil(q) : 0;                                                                   // :789 (NaN / absurd lengths: nothing)
  if (n_elements > 0) {
    const double step_distance = __ddiv_rn(distance, (double)n_elements);                                                  // :791
    const double stx = __ddiv_rn(dx, (double)n_elements), sty = __ddiv_rn(dy, (double)n_elements), stz = __ddiv_rn(dz, (double)n_elements);   // :792
    double cx = sx, cy = sy, cz = sz;
    int ls = truth_at(a, cx, cy);
    for (int i = 0; i < n_elements; ++i) {
      const double tx = __dadd_rn(cx, stx), ty = __dadd_rn(cy, sty), tz = __dadd_rn(cz, stz);                              // :796
      const int le = truth_at(a, tx, ty);
      const double average_segment_cost = __ddiv_rn((double)(ls + le), 2.0);                                              // :799 (int + int, then / 2.0)
      if (a.formulation == SIPP_TRAVEL_MAP_COST) cost = __dadd_rn(cost, __dmul_rn(average_segment_cost, step_distance));  // :805
      else if (a.formulation == SIPP_TRAVEL_DISTANCE_COST) cost = __dadd_rn(cost, step_distance);                          // :808
      else cost = __dadd_rn(cost, __ddiv_rn(step_distance, a.speed));                                                      // :811
      cx = tx; cy = ty; cz = tz;
      ls = le;                     // the next step starts where this one ended: same position, same lookup
    }
  }
  a.cost[s] = cost;
}

}  // namespace

cudaError_t launch_travel_cost(sipp_ctx* ctx, const sipp_travel_params& p, int n, const double* d_start, const double* d_goal, double* d_cost, cudaStream_t s) {
  TravelArgs a;
  a.n = n; a.start = d_start; a.goal = d_goal; a.cost = d_cost;
  a.truth = ctx->d_truth; a.pitch16 = ctx->pitch16; a.W = ctx->W; a.H = ctx->H;
  a.map_width = ctx->desc.width; a.map_height = ctx->desc.height; a.max_step = p.max_step_size; a.speed = p.mav_speed;
  a.formulation = p.formulation;
  k_travel_cost<<<(n + 127) / 128, 128, 0, s>>>(a);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_cost_integral(sipp_ctx* ctx, const sipp_cost_params& p, int n, const int32_t* d_off, const double* d_pts, const double* d_parent_cost,
                                 double* d_cost, uint8_t* d_valid, cudaStream_t s) {
  CostArgs a;
  a.n = n; a.off = d_off; a.pts = d_pts; a.parent_cost = d_parent_cost; a.cost = d_cost; a.valid = d_valid;
  a.label = ctx->d_label; a.pitch16 = ctx->pitch16; a.W = ctx->W; a.H = ctx->H;
  a.map_width = ctx->desc.width; a.map_height = ctx->desc.height; a.mpp = ctx->mpp; a.mean_cost = ctx->mean_cost; a.msd = p.max_sample_distance;
  a.accumulate = p.accumulate;
  k_cost_integral<<<(n + 127) / 128, 128, 0, s>>>(a);
  ctx->launches += 1;
  return cudaGetLastError();
}
