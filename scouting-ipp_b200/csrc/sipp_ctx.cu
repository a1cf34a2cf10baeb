// C ABI of libsipp.so (include/sipp.h): context, host-side sequential bookkeeping, staging and launches.
// There is no CPU compute path in here: every result comes from the kernels in sipp_{map,gain,field,select}.cu.
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>

#include "sipp_internal.h"
#include "sipp_peer.cuh"

static thread_local std::string g_create_err;

int sipp_set_err(sipp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_err = buf;
  return code;
}

int sipp_debug_sync(sipp_ctx* ctx, cudaStream_t s, const char* what) {
  static const bool on = getenv("SIPP_DEBUG_SYNC") != nullptr;
  if (!on) return SIPP_OK;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return sipp_set_err(ctx, SIPP_ERR_CUDA, "[debug sync] %s: %s", what, cudaGetErrorString(e));
  return SIPP_OK;
}

namespace {

inline bool in_ids(const int32_t* ids, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (ids[i] == v) return true;
  return false;
}

template <class T>
int dev_alloc(sipp_ctx* ctx, T** p, size_t count) {
  SIPP_CUDA_CHECK(ctx, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return SIPP_OK;
}

// Index math of MapPlanExp / MapServerPlanExp must be exact for the cell-record planes to be addressable by the
// voxel index: (int)((k*vs)/mpp) == k, (int)((k*vs)*(1/vs)) == k and likewise for the centre voxel (k+0.5)*vs
// (map_planexp.cpp:134-144, map_server_planexp.cpp:759,766). True whenever width/W is a power of two (every shipped
// map: 0.0625); otherwise the reference itself reads neighbouring cells and we refuse instead of diverging silently.
bool index_math_exact(int W, int H, double width, double height, double vs, double mpp) {
  const double inv = 1 / vs;
  const int n = W > H ? W : H;
  for (int k = 0; k < n; ++k) {
    const double v = (double)k * vs, c = ((double)k + 0.5) * vs;
    if ((int)(v / mpp) != k || (int)(v * inv) != k || (int)(c / mpp) != k || (int)(c * inv) != k) return false;
    if (k < W && !(width > v && width > c)) return false;
    if (k < H && !(height > v && height > c)) return false;
  }
  return true;
}

// Geometry of the implicit per-pixel graph — PRMStarPlanner::set_derived_params / build_graph
// (prm_star_planner.cpp:16-36,132-156; node_id2pose / closest_xy prm_star_planner.h:80-81).
int build_geometry(sipp_ctx* ctx) {
  const int W = ctx->W, H = ctx->H;
  const double width = ctx->desc.width, height = ctx->desc.height;
  ctx->spacing = width / (W + 1);
  ctx->max_dist = ctx->spacing * std::sqrt(2.0);
  ctx->off_x = (width - (W + 1) * ctx->spacing) / 2.0;
  ctx->off_y = (height - (H + 1) * ctx->spacing) / 2.0;
  ctx->X.resize(W);
  ctx->Y.resize(H);
  for (int i = 0; i < W; ++i) ctx->X[i] = ctx->off_x + ctx->spacing * i;
  for (int j = 0; j < H; ++j) ctx->Y[j] = ctx->off_y + ctx->spacing * j;
  ctx->start_node[0] = (int)std::round((ctx->desc.start_x - ctx->off_x) / ctx->spacing);
  ctx->start_node[1] = (int)std::round((ctx->desc.start_y - ctx->off_y) / ctx->spacing);
  ctx->goal_node[0] = (int)std::round((ctx->desc.goal_x - ctx->off_x) / ctx->spacing);
  ctx->goal_node[1] = (int)std::round((ctx->desc.goal_y - ctx->off_y) / ctx->spacing);

  // edge-length classes: distinct values of dX*dX (dX = X[i] - X[i+1]) and dY*dY
  std::vector<uint8_t> xcls(W > 1 ? W - 1 : 1, 0), ycls(H > 1 ? H - 1 : 1, 0);
  std::vector<double> xs, ys;
  auto classify = [](const std::vector<double>& P, std::vector<uint8_t>& cls, std::vector<double>& sq) -> bool {
    std::map<double, int> ids;
    for (size_t i = 0; i + 1 < P.size(); ++i) {
      const double d = P[i] - P[i + 1];
      const double s = d * d;
      auto it = ids.find(s);
      if (it == ids.end()) {
        if (ids.size() >= SIPP_MAX_CLASSES) return false;
        it = ids.emplace(s, (int)ids.size()).first;
        sq.push_back(s);
      }
      cls[i] = (uint8_t)it->second;
    }
    if (sq.empty()) sq.push_back(0.0);
    return true;
  };
  if (!classify(ctx->X, xcls, xs) || !classify(ctx->Y, ycls, ys))
    return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "more than %d distinct lattice edge lengths per axis", SIPP_MAX_CLASSES);
  ctx->Kx = (int)xs.size();
  ctx->Ky = (int)ys.size();
  const double inf = INFINITY;
  std::vector<double> hd(ctx->Kx), vd(ctx->Ky), dd((size_t)ctx->Kx * ctx->Ky);
  // (pose_a - pose_b).norm() = sqrt(dX*dX + dY*dY); an axial edge has dY (or dX) == 0.0 exactly
  for (int a = 0; a < ctx->Kx; ++a) { const double d = std::sqrt(xs[a] + 0.0 * 0.0); hd[a] = d < ctx->max_dist ? d : inf; }
  for (int b = 0; b < ctx->Ky; ++b) { const double d = std::sqrt(0.0 * 0.0 + ys[b]); vd[b] = d < ctx->max_dist ? d : inf; }
  for (int a = 0; a < ctx->Kx; ++a)
    for (int b = 0; b < ctx->Ky; ++b) {
      const double d = std::sqrt(xs[a] + ys[b]);
      dd[(size_t)a * ctx->Ky + b] = d < ctx->max_dist ? d : inf;   // prm_star_planner.cpp:138
    }
  if (dev_alloc(ctx, &ctx->d_xcls, xcls.size()) || dev_alloc(ctx, &ctx->d_ycls, ycls.size()) || dev_alloc(ctx, &ctx->d_hdist, hd.size()) ||
      dev_alloc(ctx, &ctx->d_vdist, vd.size()) || dev_alloc(ctx, &ctx->d_ddist, dd.size()))
    return SIPP_ERR_CUDA;
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(ctx->d_xcls, xcls.data(), xcls.size(), cudaMemcpyHostToDevice));
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(ctx->d_ycls, ycls.data(), ycls.size(), cudaMemcpyHostToDevice));
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(ctx->d_hdist, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice));
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(ctx->d_vdist, vd.data(), vd.size() * 8, cudaMemcpyHostToDevice));
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(ctx->d_ddist, dd.data(), dd.size() * 8, cudaMemcpyHostToDevice));
  return SIPP_OK;
}

int host_clear(sipp_ctx* ctx) {
  ctx->h_state.assign((size_t)ctx->W * ctx->H, SIPP_STATE_UNKNOWN);
  ctx->mean_cost = 0;
  ctx->mean_counter = 0;
  ctx->init_cost = ctx->desc.unknown_init_cost;
  ctx->path_valid = false;
  ctx->field_valid = false;
  ctx->warm_valid = false;
  ctx->pending_rects.clear();
  ctx->views_dirty = true;
  SIPP_CUDA_CHECK(ctx, launch_clear_planes(ctx, ctx->stream));
  SIPP_CUDA_CHECK(ctx, launch_closest_clear(ctx, ctx->stream));
  ctx->lowcost_view_dirty = true;
  return SIPP_OK;
}

// grows a device buffer (contents are not preserved)
int ensure_dev(sipp_ctx* ctx, void** p, size_t* cur, size_t need) {
  if (*cur >= need) return SIPP_OK;
  if (*p) { cudaStreamSynchronize(ctx->stream); cudaFree(*p); *p = nullptr; *cur = 0; }
  size_t cap = need + need / 4 + 4096;
  SIPP_CUDA_CHECK(ctx, cudaMalloc(p, cap));
  *cur = cap;
  return SIPP_OK;
}

struct Stage {   // carve-up of ctx->d_stage for a batch of n candidates
  double *xy, *start_xy, *cost, *gain, *value;
  uint32_t* counts;
  int32_t* parent;
  uint8_t* terminal;
};
size_t stage_bytes(size_t n) { return n * (16 + 16 + 8 + 8 + 8 + 16 + 4 + 1) + 1024; }
Stage carve(void* base, size_t n) {
  Stage s;
  char* p = reinterpret_cast<char*>(base);
  s.xy = reinterpret_cast<double*>(p); p += 16 * n;
  s.start_xy = reinterpret_cast<double*>(p); p += 16 * n;
  s.counts = reinterpret_cast<uint32_t*>(p); p += 16 * n;   // written with 128-bit stores: keep it 16-byte aligned
  s.cost = reinterpret_cast<double*>(p); p += 8 * n;
  s.gain = reinterpret_cast<double*>(p); p += 8 * n;
  s.value = reinterpret_cast<double*>(p); p += 8 * n;
  s.parent = reinterpret_cast<int32_t*>(p); p += 4 * n;
  s.terminal = reinterpret_cast<uint8_t*>(p);
  return s;
}

// bounding volume -> inclusive index range of the corner voxels k*vs that lie inside [lo, hi]
void bv_range(int n, double vs, double lo, double hi, int* klo, int* khi) {
  int a = n, b = -1;
  for (int k = 0; k < n; ++k) {
    const double v = (double)k * vs;
    if (v >= lo && v <= hi) { if (k < a) a = k; if (k > b) b = k; }
  }
  *klo = a;
  *khi = b;
}

GainGeom make_gain_geom(const sipp_ctx* ctx, const sipp_eval_params* p) {
  GainGeom g;
  g.W = ctx->W; g.H = ctx->H;
  g.vs = ctx->vs;
  g.inv_vs = 1 / ctx->vs;
  g.map_width = ctx->desc.width; g.map_height = ctx->desc.height;
  g.goal_x = ctx->desc.goal_x; g.goal_y = ctx->desc.goal_y;
  g.mean_cost = ctx->mean_cost;
  g.bv_kx_lo = 0; g.bv_kx_hi = ctx->W - 1; g.bv_ky_lo = 0; g.bv_ky_hi = ctx->H - 1;
  if (p->use_bounding_volume) {
    bv_range(ctx->W, ctx->vs, p->bv_x_min, p->bv_x_max, &g.bv_kx_lo, &g.bv_kx_hi);
    bv_range(ctx->H, ctx->vs, p->bv_y_min, p->bv_y_max, &g.bv_ky_lo, &g.bv_ky_hi);
  }
  return g;
}

bool same_bv(const sipp_eval_params& a, const sipp_eval_params& b) {
  if (a.use_bounding_volume != b.use_bounding_volume) return false;
  if (!a.use_bounding_volume) return true;
  return a.bv_x_min == b.bv_x_min && a.bv_x_max == b.bv_x_max && a.bv_y_min == b.bv_y_min && a.bv_y_max == b.bv_y_max;
}

// (re)build the evaluator view planes when the map, the path mask or the bounding volume changed
int ensure_views(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, cudaStream_t s) {
  if (p->evaluator == SIPP_EVAL_EXPAND_LOW_COST) {
    // ExpandLowCostAreaEvaluator: its own view plane + the reciprocal table recip[idx] (idx = closest cost + 1). The view packs idx
    // into 15 bits next to the "counted" bit: a label above 32766 cannot be represented there — refuse instead of aliasing costs.
    if (ctx->max_cost > (double)(kRecipEntries - 2))
      return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "ExpandLowCostAreaEvaluator: labels up to %d only (a label of %.0f was ingested)", kRecipEntries - 2, ctx->max_cost);
    const bool bv_changed = !(ctx->view_params_valid && same_bv(ctx->view_params, *p));
    if (ctx->lowcost_view_dirty || bv_changed) {
      SIPP_CUDA_CHECK(ctx, launch_build_lowcost_view(ctx, g, p->use_bounding_volume != 0, s));
      ctx->lowcost_view_dirty = false;
      if (bv_changed) ctx->views_dirty = true;    // the other views were built for another bounding volume
      ctx->view_params = *p;
      ctx->view_params_valid = true;
    }
    const int aug = p->augment_unknown_with_mean != 0;
    if (!ctx->recip_valid || ctx->recip_augment != aug || memcmp(&ctx->recip_mean, &ctx->mean_cost, sizeof(double)) != 0) {
      std::vector<double> r(kRecipEntries, 0.0);
      r[1] = aug ? 1.0 / ctx->mean_cost : 0.0;                  // cost <= 0: expand_low_cost_area_evaluator.cpp:55-57,64-66
      for (int k = 2; k < kRecipEntries; ++k) r[k] = 1.0 / (double)(k - 1);   // :53,62
      SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_recip, r.data(), r.size() * 8, cudaMemcpyHostToDevice, s));
      SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));         // r goes out of scope
      ctx->recip_valid = true; ctx->recip_augment = aug; ctx->recip_mean = ctx->mean_cost;
    }
    return SIPP_OK;
  }
  if (!ctx->views_dirty && ctx->view_params_valid && same_bv(ctx->view_params, *p)) return SIPP_OK;
  if (!(ctx->view_params_valid && same_bv(ctx->view_params, *p))) ctx->lowcost_view_dirty = true;   // built for another bounding volume
  SIPP_CUDA_CHECK(ctx, launch_build_views(ctx, g, p->use_bounding_volume != 0, s));
  if (int rc = sipp_debug_sync(ctx, s, "k_build_views")) return rc;
  ctx->views_dirty = false;
  ctx->view_params = *p;
  ctx->view_params_valid = true;
  return SIPP_OK;
}

int check_params(sipp_ctx* ctx, const sipp_eval_params* p) {
  if (!p) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "params == NULL");
  if (p->evaluator < 0 || p->evaluator >= SIPP_EVAL_COUNT_) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown evaluator %d", p->evaluator);
  const bool needs_path = p->evaluator == SIPP_EVAL_RRT_STAR;
  if (needs_path && !ctx->path_valid)
    return sipp_set_err(ctx, SIPP_ERR_NO_PATH_MASK, "RRTStarEvaluator needs a path mask: call sipp_plan or sipp_set_path_mask first");
  return SIPP_OK;
}

// the sequential part of MapServerPlanExp::mapDataCallback (map_server_planexp.cpp:238-260,276): running mean with its
// integer division, max cost, state mirror, goal-explored switch. Per-cell planes are updated by k_ingest_patch.
int host_ingest(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* labels, const uint8_t* observed) {
  const int W = ctx->W, H = ctx->H;
  for (size_t i = 0; i < (size_t)rows * cols; ++i)
    if ((!observed || observed[i]) && (labels[i] < 0 || labels[i] > 65533))
      return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "label %d outside [0,65533]", labels[i]);
  const size_t gc = (size_t)ctx->goal_loc[1] * W + ctx->goal_loc[0];
  const bool goal_inside = ctx->goal_loc[0] >= 0 && ctx->goal_loc[0] < W && ctx->goal_loc[1] >= 0 && ctx->goal_loc[1] < H;
  const bool pre_goal = goal_inside && ctx->h_state[gc] != SIPP_STATE_UNKNOWN;
  for (int i = 0; i < rows; ++i) {
    const int y = y0 + i;
    if (y < 0 || y >= H) continue;
    for (int j = 0; j < cols; ++j) {
      const int x = x0 + j;
      if (x < 0 || x >= W) continue;
      if (observed && !observed[(size_t)i * cols + j]) continue;
      const size_t c = (size_t)y * W + x;
      const int data_point = labels[(size_t)i * cols + j];
      if (ctx->h_state[c] == SIPP_STATE_UNKNOWN) {   // :246-248, int / int on the second term
        ctx->mean_cost = ctx->mean_cost * ((double)ctx->mean_counter / (ctx->mean_counter + 1)) + data_point / (ctx->mean_counter + 1);
        ++ctx->mean_counter;
      } else if (ctx->h_state[c] == SIPP_STATE_FREE) {   // :249-251: cost was just overwritten, so this adds 0.0 / n
        ctx->mean_cost = ctx->mean_cost + (data_point - (double)data_point) / ctx->mean_counter;
      }
      ctx->h_state[c] = in_ids(ctx->desc.obstacle_ids_mav, ctx->desc.n_obstacle_ids_mav, data_point) ? SIPP_STATE_OCCUPIED : SIPP_STATE_FREE;
      if (data_point > ctx->max_cost) ctx->max_cost = data_point;
    }
  }
  const bool post_goal = goal_inside && ctx->h_state[gc] != SIPP_STATE_UNKNOWN;
  if (post_goal && !pre_goal && ctx->init_cost != ctx->desc.min_rov_cost) ctx->init_cost = ctx->desc.min_rov_cost;   // :276 -> update_unknown_cost
  return SIPP_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int sipp_abi_version(void) { return SIPP_ABI_VERSION; }

const char* sipp_last_error(const sipp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void* sipp_stream(sipp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int64_t sipp_launch_count(const sipp_ctx* ctx) { return ctx ? ctx->launches : 0; }

void sipp_destroy(sipp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  void* ptrs[] = {ctx->d_state, ctx->d_label, ctx->d_pcode, ctx->d_path, ctx->d_rec, ctx->d_freelab, ctx->d_closest, ctx->d_cclab, ctx->d_closest_dist,
                  ctx->d_recip, ctx->d_field, ctx->d_pred, ctx->d_tile_w, ctx->d_xcls, ctx->d_ycls, ctx->d_hdist, ctx->d_vdist, ctx->d_ddist, ctx->d_tile_flags,
                  ctx->d_tile_list, ctx->d_relax_ctl, ctx->d_path_nodes, ctx->d_path_xy, ctx->d_plan_out, ctx->d_stage, ctx->d_tree_ws,
                  ctx->d_best, ctx->d_select_ws, ctx->d_mark, ctx->d_tile_seed, ctx->d_seed_list, ctx->d_rec2[0], ctx->d_rec2[1], ctx->d_truth};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  peer_release(ctx);
  if (ctx->d_peer_box) cudaFree(ctx->d_peer_box);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->cycle_graph) cudaGraphExecDestroy(ctx->cycle_graph);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  for (int i = 0; i < 8; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
  }
  for (int i = 0; i < 4; ++i)
    if (ctx->ev_plan[i]) cudaEventDestroy(ctx->ev_plan[i]);
  delete ctx;
}

int sipp_create(const sipp_map_desc* d, int device, sipp_ctx** out) {
  if (!out) return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "out == NULL");
  *out = nullptr;
  if (!d) return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "desc == NULL");
  if (d->W <= 0 || d->H <= 0 || d->W > 32767 || d->H > 32767 || !(d->width > 0) || !(d->height > 0))
    return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "bad map size %d x %d (%.3f x %.3f m)", d->W, d->H, d->width, d->height);
  if (d->n_obstacle_ids_mav < 0 || d->n_obstacle_ids_mav > SIPP_MAX_IDS || d->n_obstacle_ids_rov < 0 || d->n_obstacle_ids_rov > SIPP_MAX_IDS)
    return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "too many obstacle ids (max %d)", SIPP_MAX_IDS);
  // map cfg validity as the reference checks it (map_server_planexp.cpp:95-107)
  if (std::abs(1 - ((d->width / d->W) / (d->height / d->H))) >= 0.01)
    return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "cells are not square (map_server_planexp.cpp:97)");
  if (!(d->goal_x >= 0 && d->goal_y >= 0 && d->goal_x < d->width && d->goal_y < d->height && d->start_x >= 0 && d->start_y >= 0 &&
        d->start_x < d->width && d->start_y < d->height))
    return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "start / goal outside the map (map_server_planexp.cpp:104-105)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return sipp_set_err(nullptr, SIPP_ERR_CUDA, "no CUDA device (%s); libsipp has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return sipp_set_err(nullptr, SIPP_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, ndev);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return sipp_set_err(nullptr, SIPP_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10) return sipp_set_err(nullptr, SIPP_ERR_CUDA, "device %d is sm_%d%d; libsipp is built for sm_100a only", device, prop.major, prop.minor);

  sipp_ctx* ctx = new sipp_ctx();
  ctx->desc = *d;
  ctx->device = device;
  ctx->launches = 0;
  ctx->gain_warps = getenv("SIPP_GAIN_WARPS") ? atoi(getenv("SIPP_GAIN_WARPS")) : 0;
  ctx->gain_stages = getenv("SIPP_GAIN_STAGES") ? atoi(getenv("SIPP_GAIN_STAGES")) : 0;
  ctx->W = d->W;
  ctx->H = d->H;
  ctx->mpp = d->width / (unsigned int)d->W;   // m_per_pixel_ map_server_planexp.cpp:116
  ctx->vs = d->width / (unsigned int)d->W;    // c_voxel_size_ map_planexp.cpp:42
  ctx->goal_loc[0] = (int)(d->goal_x / ctx->mpp);
  ctx->goal_loc[1] = (int)(d->goal_y / ctx->mpp);
  ctx->start_loc[0] = (int)(d->start_x / ctx->mpp);
  ctx->start_loc[1] = (int)(d->start_y / ctx->mpp);
  ctx->max_cost = 0.0;
  ctx->path_generation = 0;
  ctx->view_params_valid = false;
  memset(ctx->plan_stats, 0, sizeof(ctx->plan_stats));
  // zero all device pointers (sipp_ctx has non-trivial members, so no memset of the whole struct)
  ctx->d_state = ctx->d_path = ctx->d_rec = nullptr;
  ctx->d_label = ctx->d_pcode = ctx->d_freelab = ctx->d_closest = ctx->d_cclab = nullptr;
  ctx->lowcost_view_dirty = true; ctx->recip_valid = false;
  ctx->d_closest_dist = nullptr; ctx->d_recip = nullptr; ctx->d_field = nullptr; ctx->d_pred = nullptr; ctx->d_tile_w = nullptr; ctx->d_truth = nullptr;
  ctx->d_xcls = ctx->d_ycls = nullptr; ctx->d_hdist = ctx->d_vdist = ctx->d_ddist = nullptr;
  ctx->d_tile_flags = nullptr; ctx->d_tile_list = nullptr; ctx->d_relax_ctl = nullptr;
  ctx->d_path_nodes = nullptr; ctx->d_path_xy = nullptr; ctx->d_plan_out = nullptr;
  ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0; ctx->d_stage = nullptr; ctx->d_stage_bytes = 0;
  ctx->d_tree_ws = nullptr; ctx->d_tree_ws_bytes = 0; ctx->d_best = nullptr; ctx->d_select_ws = nullptr;
  ctx->stream = nullptr;
  ctx->s_in = ctx->s_out = nullptr;
  ctx->cycle_graph = nullptr;
  ctx->cycle_graph_launches = 0;
  for (int i = 0; i < 8; ++i) ctx->ev_in[i] = ctx->ev_k[i] = nullptr;
  for (int i = 0; i < 4; ++i) ctx->ev_plan[i] = nullptr;

  auto fail = [&](int rc) {
    g_create_err = ctx->err;
    sipp_destroy(ctx);
    return rc;
  };
  if (!index_math_exact(ctx->W, ctx->H, d->width, d->height, ctx->vs, ctx->mpp)) {
    sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "voxel size %.17g makes the reference's position->cell index math inexact; use a power-of-two m_per_pixel", ctx->vs);
    return fail(SIPP_ERR_UNSUPPORTED);
  }
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); return fail(SIPP_ERR_CUDA); }
  e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev_plan[i]);
  if (e != cudaSuccess) { sipp_set_err(ctx, SIPP_ERR_CUDA, "copy streams / events: %s", cudaGetErrorString(e)); return fail(SIPP_ERR_CUDA); }
  ctx->cycle_chunks = getenv("SIPP_CYCLE_CHUNKS") ? atoi(getenv("SIPP_CYCLE_CHUNKS")) : 0;
  ctx->inc_enabled = !(getenv("SIPP_INCREMENTAL") && atoi(getenv("SIPP_INCREMENTAL")) == 0);
  ctx->cycle_use_graph = !(getenv("SIPP_CYCLE_GRAPH") && atoi(getenv("SIPP_CYCLE_GRAPH")) == 0);
  // opt-in: measured on B200 (tools/zerocopy_probe.py) it does not beat the copy-engine pipeline — every warp stores its results at
  // the end of its one 32-candidate batch, so the 1.1 MB of PCIe writes land in a burst after the scan instead of under it
  ctx->cycle_zero_copy = getenv("SIPP_CYCLE_ZEROCOPY") && atoi(getenv("SIPP_CYCLE_ZEROCOPY")) != 0;
  e = cudaMallocHost(&ctx->h_pinned, 256);
  if (e != cudaSuccess) { sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaMallocHost: %s", cudaGetErrorString(e)); return fail(SIPP_ERR_CUDA); }
  ctx->h_pinned_bytes = 256;

  ctx->pitch = ((size_t)ctx->W + 127) & ~(size_t)127;
  ctx->pitch16 = ctx->pitch;
  const size_t n8 = ctx->pitch * ctx->H;
  ctx->n_tiles_x = (ctx->W + 31) / 32;
  ctx->n_tiles_y = (ctx->H + 31) / 32;
  const size_t ntiles = (size_t)ctx->n_tiles_x * ctx->n_tiles_y;
  size_t cap = (size_t)16 * (ctx->W + ctx->H);
  if (cap < (1u << 20)) cap = 1u << 20;
  if (cap > (size_t)ctx->W * ctx->H) cap = (size_t)ctx->W * ctx->H;
  ctx->path_cap_nodes = (int)cap;
  int rc = SIPP_OK;
  rc |= dev_alloc(ctx, &ctx->d_state, n8);
  rc |= dev_alloc(ctx, &ctx->d_path, n8);
  rc |= dev_alloc(ctx, &ctx->d_rec, n8);
  rc |= dev_alloc(ctx, &ctx->d_rec2[0], n8 / 4);
  rc |= dev_alloc(ctx, &ctx->d_rec2[1], n8 / 4);
  rc |= dev_alloc(ctx, &ctx->d_label, n8);
  rc |= dev_alloc(ctx, &ctx->d_pcode, n8);
  rc |= dev_alloc(ctx, &ctx->d_freelab, n8);
  rc |= dev_alloc(ctx, &ctx->d_closest, n8);
  rc |= dev_alloc(ctx, &ctx->d_recip, (size_t)kRecipEntries);
  if (ctx->desc.compute_closest_observed) {
    rc |= dev_alloc(ctx, &ctx->d_cclab, n8);
    rc |= dev_alloc(ctx, &ctx->d_closest_dist, n8);
  }
  rc |= dev_alloc(ctx, &ctx->d_field, (size_t)ctx->W * ctx->H);
  rc |= dev_alloc(ctx, &ctx->d_pred, n8);
  rc |= dev_alloc(ctx, &ctx->d_mark, n8);
  rc |= dev_alloc(ctx, &ctx->d_tile_seed, ntiles);
  rc |= dev_alloc(ctx, &ctx->d_seed_list, ntiles + ntiles / 4 + 2);    // seed map bytes + seed list ints
  // d_tile_w (edge-weight cache of the per-tile relaxation kernel) is allocated by field_plan when SIPP_RELAX_KERNEL=0 selects that kernel
  rc |= dev_alloc(ctx, &ctx->d_tile_flags, ntiles);
  rc |= dev_alloc(ctx, &ctx->d_tile_list, 2 * (ntiles > 16384 ? ntiles : (size_t)16384));   // two ring queues (plain + re-activated tiles), each >= tiles and >= resident warps
  rc |= dev_alloc(ctx, &ctx->d_relax_ctl, 32);
  rc |= dev_alloc(ctx, &ctx->d_path_nodes, cap);
  rc |= dev_alloc(ctx, &ctx->d_path_xy, 2 * cap);
  rc |= dev_alloc(ctx, &ctx->d_plan_out, 16);
  rc |= dev_alloc(ctx, &ctx->d_best, 4);
  if (rc == SIPP_OK) {
    e = cudaMalloc(&ctx->d_select_ws, select_workspace_bytes());
    if (e != cudaSuccess) rc = sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    else cudaMemset(ctx->d_select_ws, 0, select_workspace_bytes());
  }
  if (rc != SIPP_OK) return fail(SIPP_ERR_CUDA);
  if ((rc = build_geometry(ctx)) != SIPP_OK) return fail(rc);
  if ((rc = host_clear(ctx)) != SIPP_OK) return fail(rc);
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { sipp_set_err(ctx, SIPP_ERR_CUDA, "init sync: %s", cudaGetErrorString(e)); return fail(SIPP_ERR_CUDA); }
  *out = ctx;
  return SIPP_OK;
}

#define SIPP_ENTER(ctx)                                               \
  if (!(ctx)) return SIPP_ERR_INVALID_ARG;                            \
  std::lock_guard<std::mutex> lock__((ctx)->mu);                      \
  {                                                                   \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                   \
    if (e__ != cudaSuccess) return sipp_set_err((ctx), SIPP_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e__)); \
  }

int sipp_map_clear(sipp_ctx* ctx) {
  SIPP_ENTER(ctx);
  // max_cost_ survives clearMap() in the reference (map_server_planexp.cpp:461-487 never touches it)
  ctx->path_generation = 0;
  int rc = host_clear(ctx);
  if (rc != SIPP_OK) return rc;
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return SIPP_OK;
}

int sipp_patch_origin(const sipp_ctx* ctx, double px, double py, int32_t rows, int32_t cols, int32_t* x0, int32_t* y0) {
  if (!ctx || !x0 || !y0) return SIPP_ERR_INVALID_ARG;
  const int ix = (int)(px * (long)ctx->W / ctx->desc.width);     // position2pixelLocation map_server_planexp.cpp:340-342
  const int iy = (int)(py * (long)ctx->H / ctx->desc.height);
  *x0 = ix - (int)(cols / 2);                                     // :227-228
  *y0 = iy - (int)(rows / 2);
  return SIPP_OK;
}

// host half of a patch ingest: the sequential scalars + the bookkeeping of the incremental re-plan; the caller launches k_ingest_patch
static int ingest_host_part(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* labels, const uint8_t* observed) {
  if (!labels || rows <= 0 || cols <= 0) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad patch %d x %d", rows, cols);
  int rc = host_ingest(ctx, x0, y0, rows, cols, labels, observed);
  if (rc != SIPP_OK) return rc;
  // incremental re-plan bookkeeping: the patch rectangle joins the list the next plan starts from; a change of the init cost
  // (goal explored, map_server_planexp.cpp:276) changes every unexplored cell and ends the warm start
  if (ctx->warm_valid && ctx->init_cost != ctx->warm_init_cost) ctx->warm_valid = false;
  if (ctx->warm_valid) {
    const int xa = x0 < 0 ? 0 : x0, ya = y0 < 0 ? 0 : y0, xb = x0 + cols - 1 >= ctx->W ? ctx->W - 1 : x0 + cols - 1,
              yb = y0 + rows - 1 >= ctx->H ? ctx->H - 1 : y0 + rows - 1;
    if (xa <= xb && ya <= yb) {
      if (ctx->pending_rects.size() >= 4 * 64) ctx->warm_valid = false;      // too many patches since the last plan: start over
      else { ctx->pending_rects.push_back(xa); ctx->pending_rects.push_back(ya); ctx->pending_rects.push_back(xb); ctx->pending_rects.push_back(yb); }
    }
  }
  ctx->lowcost_view_dirty = true;
  ctx->views_dirty = true;
  ctx->field_valid = false;
  return SIPP_OK;
}

static int ingest_common(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* labels, const uint8_t* observed) {
  int rc = ingest_host_part(ctx, x0, y0, rows, cols, labels, observed);
  if (rc != SIPP_OK) return rc;
  const size_t n = (size_t)rows * cols;
  const size_t need = n * 4 + (observed ? n : 0) + 256;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, need)) != SIPP_OK) return rc;
  int32_t* d_lab = reinterpret_cast<int32_t*>(ctx->d_stage);
  uint8_t* d_obs = observed ? reinterpret_cast<uint8_t*>(ctx->d_stage) + n * 4 : nullptr;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_lab, labels, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (observed) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_obs, observed, n, cudaMemcpyHostToDevice, ctx->stream));
  SIPP_CUDA_CHECK(ctx, launch_ingest_patch(ctx, x0, y0, rows, cols, d_lab, d_obs, ctx->stream));
  // setClosestObservedMaps runs per MapData patch (map_server_planexp.cpp:264); the masked whole-map upload is not one
  if (ctx->desc.compute_closest_observed && !observed) SIPP_CUDA_CHECK(ctx, launch_closest_update(ctx, x0, y0, rows, cols, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return SIPP_OK;
}

int sipp_map_update_patch(sipp_ctx* ctx, int32_t x0, int32_t y0, int32_t rows, int32_t cols, const int32_t* labels) {
  SIPP_ENTER(ctx);
  return ingest_common(ctx, x0, y0, rows, cols, labels, nullptr);
}

int sipp_map_upload_full(sipp_ctx* ctx, const int32_t* labels, const uint8_t* observed) {
  SIPP_ENTER(ctx);
  return ingest_common(ctx, 0, 0, ctx->H, ctx->W, labels, observed);
}

int sipp_update_unknown_cost(sipp_ctx* ctx, double new_cost) {
  SIPP_ENTER(ctx);
  if (new_cost != ctx->init_cost) ctx->warm_valid = false;
  ctx->init_cost = new_cost;   // unexplored cells carry PCODE_UNEXPLORED; their cost is resolved at plan time
  ctx->field_valid = false;
  return SIPP_OK;
}

int sipp_map_stats(const sipp_ctx* ctx, double* mean_cost, double* max_cost, int64_t* mean_counter) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  if (mean_cost) *mean_cost = ctx->mean_cost;
  if (max_cost) *max_cost = ctx->max_cost;
  if (mean_counter) *mean_counter = ctx->mean_counter;
  return SIPP_OK;
}

int sipp_map_download(sipp_ctx* ctx, uint8_t* state, int32_t* cost, uint8_t* reach, uint8_t* path) {
  SIPP_ENTER(ctx);
  const size_t n = (size_t)ctx->W * ctx->H;
  int rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, n * 7 + 256);
  if (rc != SIPP_OK) return rc;
  char* base = reinterpret_cast<char*>(ctx->d_stage);
  int32_t* d_cost = reinterpret_cast<int32_t*>(base);
  uint8_t* d_state = reinterpret_cast<uint8_t*>(base + 4 * n);
  uint8_t* d_reach = d_state + n;
  uint8_t* d_path = d_reach + n;
  SIPP_CUDA_CHECK(ctx, launch_download_planes(ctx, d_state, d_cost, d_reach, d_path, ctx->stream));
  if (state) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(state, d_state, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (cost) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(cost, d_cost, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (reach) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(reach, d_reach, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (path) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(path, d_path, n, cudaMemcpyDeviceToHost, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return SIPP_OK;
}

int sipp_set_path_mask(sipp_ctx* ctx, const uint8_t* mask) {
  SIPP_ENTER(ctx);
  ctx->path_mask_foreign = true;
  if (!mask) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "mask == NULL");
  SIPP_CUDA_CHECK(ctx, cudaMemcpy2DAsync(ctx->d_path, ctx->pitch, mask, ctx->W, ctx->W, ctx->H, cudaMemcpyHostToDevice, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->path_valid = true;
  ctx->views_dirty = true;
  return SIPP_OK;
}

int sipp_path_mask_generation(const sipp_ctx* ctx) { return ctx ? ctx->path_generation : -1; }

static int plan_collect(sipp_ctx* ctx, const int* out, const int* ctl, double* cost, int32_t* status, int* len_out, bool* found_out);

static int plan_impl(sipp_ctx* ctx, double* cost, int32_t* status, double* path_xy, int32_t path_cap, int32_t* path_len) {
  int rc = field_plan(ctx, ctx->stream);
  if (rc != SIPP_OK) { ctx->warm_valid = false; return rc; }
  // results come back through the context's pinned block (a copy into pageable memory is staged by the driver under a global lock)
  int* out = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->h_pinned) + 64);   // [10]
  int* ctl = out + 16;   // [24] CTL_* of sipp_field.cu: head, tail, pending, activations, sweeps, abort, pushes; [16..23] watchdog diagnostics
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->d_plan_out, 10 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctl, ctx->d_relax_ctl, 24 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  int len = 0;
  bool found = false;
  if ((rc = plan_collect(ctx, out, ctl, cost, status, &len, &found)) != SIPP_OK) return rc;
  if (path_len) *path_len = found ? len : 0;
  if (path_xy && found && path_cap > 0) {
    const int m = len < path_cap ? len : path_cap;
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(path_xy, ctx->d_path_xy, (size_t)m * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SIPP_OK;
}

// host half of a plan's end: out = plan_out[0..9], ctl = the relaxation control block [0..23] as the kernels left them
static int plan_collect(sipp_ctx* ctx, const int* out, const int* ctl, double* cost, int32_t* status, int* len_out, bool* found_out) {
  ctx->plan_stats[0] = ctl[6];
  ctx->plan_stats[1] = ctl[3];
  ctx->plan_stats[2] = ctl[4];
  ctx->plan_stats[3] = ctl[5];
  ctx->pending_rects.clear();
  ctx->plan_info[2] = out[6];
  ctx->plan_info[3] = out[5];
  if (ctl[5]) {
    ctx->warm_valid = false;
    return sipp_set_err(ctx, SIPP_ERR_CUDA, "cost-to-go relaxation aborted by its watchdog (code %d: 1 = work queue starved, 2 = sub-tile sweeps, 3 = block drain; "
                        "head %d tail %d pending %d; diag %d %d %d %d %d)", ctl[5], ctl[0], ctl[1], ctl[2], ctl[16], ctl[17], ctl[18], ctl[19], ctl[20]);
  }
  // the converged field + predecessor plane are the warm start of the next plan
  ctx->warm_valid = ctx->inc_enabled;
  ctx->warm_init_cost = ctx->init_cost;
  double c;
  memcpy(&c, &out[8], sizeof(double));
  const int len = out[0];
  const bool found = out[2] != 0;
  // PRMStarPlanner::plan (prm_star_planner.cpp:38-70): cost -1 / TERMINAL_INVALID when the search exhausts the graph.
  // A reachable goal whose predecessor walk does not terminate (zero-weight plateaus can make the canonical rule cycle;
  // never the case with the shipped label sets, all traversable labels >= 30) is reported like the reference's
  // catch-all: INVALID, cost -1, no path.
  int st = SIPP_PLAN_TERMINAL_INVALID;
  if (found) st = out[1] ? SIPP_PLAN_TERMINAL_VALID : SIPP_PLAN_VALID;
  else if (out[4]) st = SIPP_PLAN_INVALID;
  if (cost) *cost = found ? c : -1.0;
  if (status) *status = st;
  *len_out = len;
  *found_out = found;
  ctx->path_generation += 1;     // path_counter_ map_server_planexp.cpp:896
  ctx->path_valid = true;
  ctx->field_valid = true;
  ctx->views_dirty = true;
  return SIPP_OK;
}

int sipp_plan(sipp_ctx* ctx, double* cost, int32_t* status, double* path_xy, int32_t path_cap, int32_t* path_len) {
  SIPP_ENTER(ctx);
  return plan_impl(ctx, cost, status, path_xy, path_cap, path_len);
}

int sipp_path_download(sipp_ctx* ctx, double* path_xy, int32_t path_cap, int32_t* path_len) {
  SIPP_ENTER(ctx);
  if (!ctx->field_valid) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "no path: call sipp_plan first");
  int* out = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->h_pinned) + 64);
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->d_plan_out, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int len = out[2] != 0 ? out[0] : 0;
  if (path_len) *path_len = len;
  if (path_xy && len > 0 && path_cap > 0) {
    const int m = len < path_cap ? len : path_cap;
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(path_xy, ctx->d_path_xy, (size_t)m * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SIPP_OK;
}

int sipp_plan_phases(sipp_ctx* ctx, double ms[2]) {
  SIPP_ENTER(ctx);
  if (!ms) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "ms == NULL");
  if (!ctx->field_valid) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "no plan yet");
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  float relax = 0.f, total = 0.f;
  SIPP_CUDA_CHECK(ctx, cudaEventElapsedTime(&relax, ctx->ev_plan[0], ctx->ev_plan[1]));
  SIPP_CUDA_CHECK(ctx, cudaEventElapsedTime(&total, ctx->ev_plan[2], ctx->ev_plan[3]));
  ms[0] = relax;
  ms[1] = total;
  return SIPP_OK;
}

int sipp_plan_info(const sipp_ctx* ctx, int64_t info[4]) {
  if (!ctx || !info) return SIPP_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) info[i] = ctx->plan_info[i];
  return SIPP_OK;
}

int sipp_set_incremental(sipp_ctx* ctx, int32_t enable) {
  SIPP_ENTER(ctx);
  ctx->inc_enabled = enable != 0;
  if (!enable) ctx->warm_valid = false;
  return SIPP_OK;
}

int sipp_plan_stats(const sipp_ctx* ctx, int64_t stats[4]) {
  if (!ctx || !stats) return SIPP_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) stats[i] = ctx->plan_stats[i];
  return SIPP_OK;
}

int sipp_field_download(sipp_ctx* ctx, double* field) {
  SIPP_ENTER(ctx);
  if (!field) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "field == NULL");
  if (!ctx->field_valid) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "no field: call sipp_plan first");
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(field, ctx->d_field, (size_t)ctx->W * ctx->H * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return SIPP_OK;
}

// ---- gain ----------------------------------------------------------------------------------------------
int sipp_gain_batch_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy, const double* d_start_xy,
                        double* d_gain, uint32_t* d_counts, uint8_t* d_terminal, void* stream) {
  SIPP_ENTER(ctx);
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (n < 0 || (n > 0 && (!d_xy || !d_gain))) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad batch arguments");
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  return gain_launch(ctx, params, g, n, d_xy, d_start_xy, nullptr, d_gain, d_counts, d_terminal, nullptr, s);
}

int sipp_gain_batch(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy, const double* start_xy, double* gain,
                    uint32_t* counts, uint8_t* terminal) {
  SIPP_ENTER(ctx);
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (n < 0 || (n > 0 && (!xy || !gain))) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad batch arguments");
  if (n == 0) return SIPP_OK;
  cudaStream_t s = ctx->stream;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, stage_bytes(n))) != SIPP_OK) return rc;
  const Stage st = carve(ctx->d_stage, n);
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.xy, xy, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  if (start_xy) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.start_xy, start_xy, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  if ((rc = gain_launch(ctx, params, g, n, st.xy, start_xy ? st.start_xy : nullptr, nullptr, st.gain, counts ? st.counts : nullptr,
                        terminal ? st.terminal : nullptr, nullptr, s)) != SIPP_OK)
    return rc;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(gain, st.gain, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (counts) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, st.counts, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
  if (terminal) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(terminal, st.terminal, (size_t)n, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  return SIPP_OK;
}

// ---- value + selection -----------------------------------------------------------------------------------
// device carve-up for segments with several views: [view_xy nv x 16][vcounts nv x 16][vgain nv x 8][vfsum nv x 8][counts n x 16][start n x 16][gain n x 8]
// [view_off (n + 1) x 4, padded][terminal n]
struct ViewStage {
  double *view_xy, *vgain, *vfsum, *start_xy, *gain;
  uint32_t *vcounts, *counts;
  int* view_off;
  uint8_t* terminal;
};
static size_t view_stage_bytes(size_t n, size_t nv) { return nv * 48 + n * 41 + 4 * (n + 1) + 256; }
static ViewStage carve_views(void* base, size_t n, size_t nv) {
  ViewStage v;
  char* p = reinterpret_cast<char*>(base);
  v.view_xy = reinterpret_cast<double*>(p); p += 16 * nv;
  v.vcounts = reinterpret_cast<uint32_t*>(p); p += 16 * nv;
  v.counts = reinterpret_cast<uint32_t*>(p); p += 16 * n;
  v.start_xy = reinterpret_cast<double*>(p); p += 16 * n;
  v.vgain = reinterpret_cast<double*>(p); p += 8 * nv;
  v.vfsum = reinterpret_cast<double*>(p); p += 8 * nv;
  v.gain = reinterpret_cast<double*>(p); p += 8 * n;
  v.view_off = reinterpret_cast<int*>(p); p += 4 * ((n + 1 + 3) & ~(size_t)3);
  v.terminal = reinterpret_cast<uint8_t*>(p);
  return v;
}
static int check_views(sipp_ctx* ctx, int32_t n, const int32_t* view_offset, const double* view_xy, int64_t* nv_out) {
  if (!view_offset || view_offset[0] != 0) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "view_offset must start at 0");
  for (int i = 0; i < n; ++i)
    if (view_offset[i + 1] < view_offset[i]) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "view_offset decreases at segment %d", i);
  *nv_out = view_offset[n];
  if (*nv_out > 0 && !view_xy) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "view_xy == NULL");
  return SIPP_OK;
}

int sipp_gain_batch_views(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const int32_t* view_offset, const double* view_xy,
                          const double* start_xy, double* gain, uint32_t* counts, uint8_t* terminal) {
  SIPP_ENTER(ctx);
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (n < 0 || (n > 0 && !gain)) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad batch arguments");
  if (n == 0) return SIPP_OK;
  int64_t nv = 0;
  if ((rc = check_views(ctx, n, view_offset, view_xy, &nv)) != SIPP_OK) return rc;
  cudaStream_t s = ctx->stream;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, view_stage_bytes((size_t)n, (size_t)nv))) != SIPP_OK) return rc;
  const ViewStage st = carve_views(ctx->d_stage, (size_t)n, (size_t)nv);
  if (nv > 0) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.view_xy, view_xy, (size_t)nv * 16, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.view_off, view_offset, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s));
  if (start_xy) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.start_xy, start_xy, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  if ((rc = gain_launch_views(ctx, params, g, n, st.view_off, (int)nv, st.view_xy, start_xy ? st.start_xy : nullptr, st.vgain, st.vcounts, st.vfsum, st.gain,
                              counts ? st.counts : nullptr, terminal ? st.terminal : nullptr, s)) != SIPP_OK)
    return rc;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(gain, st.gain, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (counts) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, st.counts, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
  if (terminal) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(terminal, st.terminal, (size_t)n, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  return SIPP_OK;
}

int sipp_value_select_dev(sipp_ctx* ctx, int32_t n, const double* d_gain, const double* d_cost, const int32_t* d_parent, double* d_value,
                          int32_t* d_best_child, double* d_best_value, void* stream) {
  SIPP_ENTER(ctx);
  if (n < 1 || !d_gain || !d_cost || !d_value) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad value_select arguments");
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  if (d_parent) {
    int rc = ensure_dev(ctx, &ctx->d_tree_ws, &ctx->d_tree_ws_bytes, (size_t)n * 16 + 512);
    if (rc != SIPP_OK) return rc;
    SIPP_CUDA_CHECK(ctx, launch_value_tree(ctx, n, d_gain, nullptr, d_cost, d_parent, d_value, ctx->d_best, ctx->d_tree_ws, s));
  } else {
    SIPP_CUDA_CHECK(ctx, launch_value_flat(ctx, n, d_gain, d_cost, d_value, s));
    // children are nodes 1..n-1: scan value+1, the kernel shifts the index back by one
    SIPP_CUDA_CHECK(ctx, launch_select_flat(ctx, n - 1, d_value + 1, 1, ctx->d_best, d_value, s));
  }
  // unpack the 16-byte record into the caller's device scalars
  if (d_best_value) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_best_value, &ctx->d_best->value, 8, cudaMemcpyDeviceToDevice, s));
  if (d_best_child) {
    // index is int64 in the record; the low 32 bits (little endian) are the int32 the caller asked for
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_best_child, &ctx->d_best->index, 4, cudaMemcpyDeviceToDevice, s));
  }
  return SIPP_OK;
}

int sipp_value_select(sipp_ctx* ctx, int32_t n, const double* gain, const double* cost, const int32_t* parent, double* value,
                      int32_t* best_child, double* best_value) {
  SIPP_ENTER(ctx);
  if (n < 1 || !gain || !cost) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad value_select arguments");
  if (parent)
    for (int i = 1; i < n; ++i)
      if (parent[i] < 0 || parent[i] >= i) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "parent[%d] = %d violates 0 <= parent[i] < i", i, parent[i]);
  cudaStream_t s = ctx->stream;
  int rc;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, stage_bytes(n))) != SIPP_OK) return rc;
  const Stage st = carve(ctx->d_stage, n);
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.gain, gain, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.cost, cost, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  if (parent) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.parent, parent, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  if (parent) {
    if ((rc = ensure_dev(ctx, &ctx->d_tree_ws, &ctx->d_tree_ws_bytes, (size_t)n * 16 + 512)) != SIPP_OK) return rc;
    SIPP_CUDA_CHECK(ctx, launch_value_tree(ctx, n, st.gain, nullptr, st.cost, st.parent, st.value, ctx->d_best, ctx->d_tree_ws, s));
  } else {
    SIPP_CUDA_CHECK(ctx, launch_value_flat(ctx, n, st.gain, st.cost, st.value, s));
    if (n > 1) SIPP_CUDA_CHECK(ctx, launch_select_flat(ctx, n - 1, st.value + 1, 1, ctx->d_best, st.value, s));
  }
  BestRec br;
  br.value = 0.0;
  br.index = -1;
  if (parent || n > 1) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(&br, ctx->d_best, sizeof(br), cudaMemcpyDeviceToHost, s));
  if (value) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(value, st.value, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  if (br.index == -2) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "tree deeper than 512 levels");
  if (best_child) *best_child = (int32_t)br.index;
  if (best_value) *best_value = br.value;
  return SIPP_OK;
}

int sipp_value_preorder(sipp_ctx* ctx, int32_t n, const double* gain_new, const double* gain_old, const double* cost, const int32_t* parent,
                        double* value) {
  SIPP_ENTER(ctx);
  if (n < 1 || !gain_new || !gain_old || !cost || !parent || !value) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad value_preorder arguments");
  for (int i = 1; i < n; ++i)
    if (parent[i] < 0 || parent[i] >= i) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "parent[%d] = %d violates 0 <= parent[i] < i", i, parent[i]);
  cudaStream_t s = ctx->stream;
  int rc;
  // stage: gain_new | gain_old | cost | value (doubles) | parent (int32)
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, (size_t)n * 36 + 256)) != SIPP_OK) return rc;
  double* d_new = reinterpret_cast<double*>(ctx->d_stage);
  double *d_old = d_new + n, *d_cost = d_old + n, *d_value = d_cost + n;
  int32_t* d_parent = reinterpret_cast<int32_t*>(d_value + n);
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_new, gain_new, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_old, gain_old, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cost, cost, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_parent, parent, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  if ((rc = ensure_dev(ctx, &ctx->d_tree_ws, &ctx->d_tree_ws_bytes, (size_t)n * 16 + 512)) != SIPP_OK) return rc;
  SIPP_CUDA_CHECK(ctx, launch_value_tree(ctx, n, d_new, d_old, d_cost, d_parent, d_value, ctx->d_best, ctx->d_tree_ws, s));
  BestRec br;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(&br, ctx->d_best, sizeof(br), cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(value, d_value, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  if (br.index == -2) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "tree deeper than 512 levels");
  return SIPP_OK;
}

// ---- fused cycle -----------------------------------------------------------------------------------------
static int cycle_dev_impl(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy, const double* d_start_xy, const double* d_cost,
                          double* d_gain, uint8_t* d_terminal, double* d_value, void* d_best, void* stream, bool shard, int64_t index_offset,
                          int peer_mode = 0) {
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (shard && ctx->peer_world < 1) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "peers are not attached (sipp_peer_attach)");
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  if (shard && n == 0 && d_best) {
    // an empty shard (fewer candidates than ranks, or everything pruned on this rank) still takes part in the lockstep exchange:
    // it publishes {0.0, -1} — "never wins" — so the step counter advances on every rank and nobody waits for a record that never comes
    const BestRec none = {0.0, -1};
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_best + 1, &none, sizeof(none), cudaMemcpyHostToDevice, s));
    SIPP_CUDA_CHECK(ctx, launch_peer_exchange(ctx, ctx->d_best + 1, reinterpret_cast<BestRec*>(d_best), peer_mode == 1 ? PEER_PIPELINED : PEER_SYNC, s));
    return SIPP_OK;
  }
  if (n < 1 || !d_xy || !d_cost || !d_gain || !d_value || !d_best) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad cycle arguments");
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  // one launch: the argmax (and, sharded, the exchange of the shard winners) is the tail of the gain kernel
  static const bool fuse = !(getenv("SIPP_FUSED_SELECT") && atoi(getenv("SIPP_FUSED_SELECT")) == 0);
  FusedSelect fs;
  fs.best_out = reinterpret_cast<BestRec*>(d_best);
  fs.index_offset = shard ? (long long)index_offset : 0;
  fs.peer = shard ? &ctx->peer : nullptr;
  fs.peer_mode = peer_mode;
  if ((rc = gain_launch(ctx, params, g, n, d_xy, d_start_xy, d_cost, d_gain, nullptr, d_terminal, d_value, s, nullptr, nullptr, nullptr,
                        fuse ? &fs : nullptr)) != SIPP_OK)
    return rc;
  if (!fuse)
    SIPP_CUDA_CHECK(ctx, launch_select_flat(ctx, n, d_value, fs.index_offset, fs.best_out, nullptr, s, fs.peer, fs.peer_mode));
  return SIPP_OK;
}

int sipp_cycle_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy, const double* d_start_xy, const double* d_cost,
                   double* d_gain, uint8_t* d_terminal, double* d_value, void* d_best, void* stream) {
  SIPP_ENTER(ctx);
  return cycle_dev_impl(ctx, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, d_best, stream, false, 0);
}

int sipp_cycle_shard_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy, const double* d_start_xy,
                         const double* d_cost, double* d_gain, uint8_t* d_terminal, double* d_value, int64_t index_offset,
                         void* d_best_global, void* stream) {
  SIPP_ENTER(ctx);
  return cycle_dev_impl(ctx, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, d_best_global, stream, true, index_offset);
}

// device-side alias of a pinned (page-locked, mapped) host pointer; nullptr for pageable memory
static void* pinned_alias(const void* p) {
  if (!p) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
  cudaGetLastError();
  return nullptr;
}

int sipp_cycle_shard_pipelined_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy, const double* d_start_xy,
                                   const double* d_cost, double* d_gain, uint8_t* d_terminal, double* d_value, int64_t index_offset,
                                   void* d_prev_best_global, void* stream) {
  SIPP_ENTER(ctx);
  return cycle_dev_impl(ctx, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, d_prev_best_global, stream, true, index_offset, 1);
}

static int cycle_host_impl(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy, const double* start_xy, const double* cost,
                           double* gain, uint8_t* terminal, double* value, int64_t* best_index, double* best_value, bool shard, int64_t index_offset) {
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (shard && ctx->peer_world < 1) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "peers are not attached (sipp_peer_attach)");
  cudaStream_t s = ctx->stream;
  if (shard && n == 0) {       // empty shard: exchange only (see cycle_dev_impl)
    const BestRec none = {0.0, -1};
    BestRec* h_best0 = reinterpret_cast<BestRec*>(ctx->h_pinned);
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_best + 1, &none, sizeof(none), cudaMemcpyHostToDevice, s));
    SIPP_CUDA_CHECK(ctx, launch_peer_exchange(ctx, ctx->d_best + 1, ctx->d_best, PEER_SYNC, s));
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(h_best0, ctx->d_best, sizeof(BestRec), cudaMemcpyDeviceToHost, s));
    SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
    if (best_index) *best_index = h_best0->index;
    if (best_value) *best_value = h_best0->value;
    if (h_best0->index == -3) return sipp_set_err(ctx, SIPP_ERR_CUDA, "peer exchange timed out (a rank died or the ranks are not calling in lockstep)");
    return SIPP_OK;
  }
  if (n < 1 || !xy || !cost) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad cycle arguments");
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, stage_bytes(n))) != SIPP_OK) return rc;
  const Stage st = carve(ctx->d_stage, n);
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  // Chunked three-stage pipeline: host->device copies of chunk c+1 (stream s_in) and device->host copies of chunk c-1 (s_out)
  // run under the gain kernel of chunk c (s); the argmax runs once over all values. Small batches go through in one piece.
  // The whole pipeline (copies, kernels, cross-stream events) is captured once as a CUDA graph and replayed with a single
  // launch while the caller keeps passing the same buffers and parameters — the per-call cost of ~25 driver calls was as
  // large as the transfers themselves.
  int chunks = ctx->cycle_chunks > 0 ? ctx->cycle_chunks : (n >= 16384 ? 2 : 1);
  if (chunks > 8) chunks = 8;
  const int per = (n + chunks - 1) / chunks;
  BestRec* h_best = reinterpret_cast<BestRec*>(ctx->h_pinned);
  // Zero-copy results (SIPP_CYCLE_ZEROCOPY=1): when every result array the caller passed is pinned host memory, the gain kernel stores gain / terminal /
  // value straight into it over PCIe (posted writes, coalesced 256-byte runs) and the argmax kernel writes the winner into the
  // context's pinned record: no device->host copies, no output stream. Inputs still travel by copy engine, chunk c+1 under the
  // gain kernel of chunk c — a kernel that reads its poses over PCIe cannot start a candidate before its pose has arrived.
  double* a_gain = reinterpret_cast<double*>(pinned_alias(gain));
  uint8_t* a_term = reinterpret_cast<uint8_t*>(pinned_alias(terminal));
  double* a_value = reinterpret_cast<double*>(pinned_alias(value));
  BestRec* a_best = reinterpret_cast<BestRec*>(pinned_alias(h_best));
  const bool zc = ctx->cycle_zero_copy && a_best && (!gain || a_gain) && (!terminal || a_term) && (!value || a_value) && (gain || terminal || value);
  FusedSelect fs;
  fs.best_out = zc ? a_best : ctx->d_best;
  fs.index_offset = shard ? (long long)index_offset : 0;
  fs.peer = shard ? &ctx->peer : nullptr;
  fs.peer_mode = 0;
  auto enqueue = [&]() -> int {
    int launched = 0;
    for (int c = 0; c < chunks; ++c) {
      const int lo = c * per, m = (lo + per <= n ? per : n - lo);
      if (m <= 0) break;
      cudaStream_t si = chunks > 1 ? ctx->s_in : s, so = chunks > 1 ? ctx->s_out : s;
      if (chunks > 1 && c == 0) {   // fork the copy streams off s (required inside a capture, harmless outside)
        SIPP_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_k[7], s));
        SIPP_CUDA_CHECK(ctx, cudaStreamWaitEvent(si, ctx->ev_k[7], 0));
      }
      SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.xy + 2 * (size_t)lo, xy + 2 * (size_t)lo, (size_t)m * 16, cudaMemcpyHostToDevice, si));
      if (start_xy) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.start_xy + 2 * (size_t)lo, start_xy + 2 * (size_t)lo, (size_t)m * 16, cudaMemcpyHostToDevice, si));
      SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(st.cost + lo, cost + lo, (size_t)m * 8, cudaMemcpyHostToDevice, si));
      if (chunks > 1) {
        SIPP_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_in[c], si));
        SIPP_CUDA_CHECK(ctx, cudaStreamWaitEvent(s, ctx->ev_in[c], 0));
      }
      const int grc = gain_launch(ctx, params, g, m, st.xy + 2 * (size_t)lo, start_xy ? st.start_xy + 2 * (size_t)lo : nullptr, st.cost + lo,
                                  (zc && gain) ? a_gain + lo : st.gain + lo, nullptr, (zc && terminal) ? a_term + lo : st.terminal + lo, st.value + lo, s,
                                  nullptr, nullptr, (zc && value) ? a_value + lo : nullptr, chunks == 1 ? &fs : nullptr);
      if (grc != SIPP_OK) return grc;
      ++launched;
      if (zc) continue;
      if (chunks > 1) {
        SIPP_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_k[c], s));
        SIPP_CUDA_CHECK(ctx, cudaStreamWaitEvent(so, ctx->ev_k[c], 0));
      }
      if (gain) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(gain + lo, st.gain + lo, (size_t)m * 8, cudaMemcpyDeviceToHost, so));
      if (terminal) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(terminal + lo, st.terminal + lo, (size_t)m, cudaMemcpyDeviceToHost, so));
      if (value) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(value + lo, st.value + lo, (size_t)m * 8, cudaMemcpyDeviceToHost, so));
    }
    // N GPUs: the argmax kernel also exchanges the shard winners through peer memory and leaves the global winner in d_best
    if (chunks > 1) {           // several gain launches: one argmax over all values (a single launch carries it in its tail)
      SIPP_CUDA_CHECK(ctx, launch_select_flat(ctx, n, st.value, fs.index_offset, fs.best_out, nullptr, s, fs.peer));
      ++launched;
    }
    if (!zc) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(h_best, ctx->d_best, sizeof(BestRec), cudaMemcpyDeviceToHost, s));
    if (chunks > 1 && !zc) {      // join the output stream back into s
      SIPP_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_in[7], ctx->s_out));
      SIPP_CUDA_CHECK(ctx, cudaStreamWaitEvent(s, ctx->ev_in[7], 0));
    }
    ctx->cycle_graph_launches = launched;
    return SIPP_OK;
  };
  if (ctx->cycle_use_graph && chunks > 1) {   // a single chunk is seven driver calls: capturing / instantiating a graph costs more
    // key: everything the captured nodes bake in
    std::vector<unsigned char> key;
    auto put = [&](const void* p, size_t nb) { const unsigned char* b = reinterpret_cast<const unsigned char*>(p); key.insert(key.end(), b, b + nb); };
    const void* ptrs[8] = {xy, start_xy, cost, gain, terminal, value, ctx->d_stage, ctx->d_rec};
    put(&n, sizeof(n)); put(&chunks, sizeof(chunks)); put(ptrs, sizeof(ptrs)); put(params, sizeof(*params)); put(&g, sizeof(g));
    const int64_t shard_key[3] = {shard ? 1 : 0, shard ? index_offset : 0, zc ? 1 : 0};
    put(shard_key, sizeof(shard_key));
    if (!ctx->cycle_graph || key != ctx->cycle_key) {
      if (ctx->cycle_graph) { cudaGraphExecDestroy(ctx->cycle_graph); ctx->cycle_graph = nullptr; }
      const int64_t launches_before = ctx->launches;
      SIPP_CUDA_CHECK(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
      rc = enqueue();
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(s, &graph);
      ctx->launches = launches_before;       // nothing ran yet
      if (rc != SIPP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return sipp_set_err(ctx, SIPP_ERR_CUDA, "graph capture of the cycle pipeline failed: %s", cudaGetErrorString(ce));
      const cudaError_t ie = cudaGraphInstantiate(&ctx->cycle_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { ctx->cycle_graph = nullptr; return sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
      ctx->cycle_key = key;
    }
    SIPP_CUDA_CHECK(ctx, cudaGraphLaunch(ctx->cycle_graph, s));
    ctx->launches += ctx->cycle_graph_launches;
  } else {
    if ((rc = enqueue()) != SIPP_OK) return rc;
  }
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  const BestRec br = *h_best;
  if (best_index) *best_index = br.index;
  if (best_value) *best_value = br.value;
  if (shard && br.index == -3) return sipp_set_err(ctx, SIPP_ERR_CUDA, "peer exchange timed out (a rank died or the ranks are not calling in lockstep)");
  return SIPP_OK;
}

int sipp_cycle(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy, const double* start_xy, const double* cost,
               double* gain, uint8_t* terminal, double* value, int32_t* best_index, double* best_value) {
  SIPP_ENTER(ctx);
  int64_t bi = -1;
  const int rc = cycle_host_impl(ctx, params, n, xy, start_xy, cost, gain, terminal, value, &bi, best_value, false, 0);
  if (rc == SIPP_OK && best_index) *best_index = (int32_t)bi;
  return rc;
}

int sipp_cycle_shard(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy, const double* start_xy, const double* cost,
                     int64_t index_offset, double* gain, uint8_t* terminal, double* value, int64_t* best_index_global, double* best_value) {
  SIPP_ENTER(ctx);
  return cycle_host_impl(ctx, params, n, xy, start_xy, cost, gain, terminal, value, best_index_global, best_value, true, index_offset);
}

// ---- trajectory cost integral --------------------------------------------------------------------------
int sipp_cost_integral(sipp_ctx* ctx, const sipp_cost_params* params, int32_t n, const int32_t* point_offset, const double* points_xyz,
                       const double* parent_cost, double* cost, uint8_t* valid) {
  SIPP_ENTER(ctx);
  if (!params || n < 0 || (n > 0 && (!point_offset || !cost))) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad cost_integral arguments");
  if (!(params->max_sample_distance > 0)) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "max_sample_distance must be positive");
  if (n == 0) return SIPP_OK;
  if (point_offset[0] != 0) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "point_offset[0] must be 0");
  for (int i = 0; i < n; ++i)
    if (point_offset[i + 1] < point_offset[i]) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "point_offset is not non-decreasing at %d", i);
  const size_t np = (size_t)point_offset[n];
  if (np > 0 && !points_xyz) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "points_xyz == NULL");
  cudaStream_t s = ctx->stream;
  int rc;
  const size_t N = (size_t)n, need = np * 24 + N * 8 * 2 + (N + 1) * 4 + N + 512;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, need)) != SIPP_OK) return rc;
  double* d_pts = reinterpret_cast<double*>(ctx->d_stage);
  double *d_parent = d_pts + 3 * np, *d_cost = d_parent + N;
  int32_t* d_off = reinterpret_cast<int32_t*>(d_cost + N);
  uint8_t* d_valid = reinterpret_cast<uint8_t*>(d_off + N + 1);
  if (np) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pts, points_xyz, np * 24, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_off, point_offset, (N + 1) * 4, cudaMemcpyHostToDevice, s));
  const bool acc = params->accumulate && parent_cost;
  if (acc) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_parent, parent_cost, N * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, launch_cost_integral(ctx, *params, n, d_off, d_pts, acc ? d_parent : nullptr, d_cost, d_valid, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(cost, d_cost, N * 8, cudaMemcpyDeviceToHost, s));
  if (valid) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(valid, d_valid, N, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  return SIPP_OK;
}

// ---- simulator travel cost --------------------------------------------------------------------------------
int sipp_truth_upload(sipp_ctx* ctx, const int32_t* labels) {
  SIPP_ENTER(ctx);
  if (!labels) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "labels == NULL");
  const size_t n = (size_t)ctx->W * ctx->H;
  std::vector<uint16_t> packed(ctx->pitch16 * (size_t)ctx->H, 0);
  for (int y = 0; y < ctx->H; ++y)
    for (int x = 0; x < ctx->W; ++x) {
      const int32_t v = labels[(size_t)y * ctx->W + x];
      if (v < 0 || v > 65533) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "label %d outside [0,65533]", v);
      packed[(size_t)y * ctx->pitch16 + x] = (uint16_t)v;
    }
  (void)n;
  if (!ctx->d_truth) SIPP_CUDA_CHECK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_truth), packed.size() * sizeof(uint16_t)));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_truth, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return SIPP_OK;
}

int sipp_travel_cost(sipp_ctx* ctx, const sipp_travel_params* params, int32_t n, const double* start_xyz, const double* goal_xyz, double* cost) {
  SIPP_ENTER(ctx);
  if (!params || n < 0 || (n > 0 && (!start_xyz || !goal_xyz || !cost))) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad travel_cost arguments");
  if (params->formulation < SIPP_TRAVEL_MAP_COST || params->formulation > SIPP_TRAVEL_TIME_COST)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown cost formulation %d", params->formulation);      // mav_simulator.cpp:813-815
  if (!(params->max_step_size > 0) || !(params->mav_speed > 0)) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "max_step_size and mav_speed must be positive");
  if (!ctx->d_truth) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "no ground-truth map: call sipp_truth_upload first");
  if (n == 0) return SIPP_OK;
  cudaStream_t s = ctx->stream;
  const size_t N = (size_t)n;
  int rc;
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, N * 56 + 256)) != SIPP_OK) return rc;
  double* d_start = reinterpret_cast<double*>(ctx->d_stage);
  double *d_goal = d_start + 3 * N, *d_cost = d_goal + 3 * N;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_start, start_xyz, N * 24, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_goal, goal_xyz, N * 24, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, launch_travel_cost(ctx, *params, n, d_start, d_goal, d_cost, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(cost, d_cost, N * 8, cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  return SIPP_OK;
}

// ---- episode replay --------------------------------------------------------------------------------------
int sipp_episode_run(sipp_ctx* ctx, const int32_t* truth, const sipp_eval_params* params, int32_t cycles, int32_t n, const double* cand_xy,
                     const double* cand_cost, int32_t fov, int32_t init, double start_x, double start_y, double* path_cost,
                     int32_t* plan_status, int32_t* path_len, int32_t* best_index, double* best_value) {
  SIPP_ENTER(ctx);
  if (!truth || !params || cycles < 0 || n < 1 || !cand_xy || !cand_cost || fov < 1 || fov > 4096 || init < 0)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad episode arguments");
  const int W = ctx->W, H = ctx->H;
  int rc;
  std::vector<int32_t> patch;
  auto ingest_window = [&](int x0, int y0, int rows, int cols) -> int {   // the simulator's MapData patch: truth labels, 0 off the image
    patch.assign((size_t)rows * cols, 0);
    for (int i = 0; i < rows; ++i) {
      const int y = y0 + i;
      if (y < 0 || y >= H) continue;
      for (int j = 0; j < cols; ++j) {
        const int x = x0 + j;
        if (x >= 0 && x < W) patch[(size_t)i * cols + j] = truth[(size_t)y * W + x];
      }
    }
    return ingest_common(ctx, x0, y0, rows, cols, patch.data(), nullptr);
  };
  if (init > 0 && (rc = ingest_window(0, 0, init < H ? init : H, init < W ? init : W)) != SIPP_OK) return rc;   // mav_simulator.cpp:156-159
  double px = start_x, py = start_y;
  for (int c = 0; c < cycles; ++c) {
    const int ix = (int)(px * (long)W / ctx->desc.width), iy = (int)(py * (long)H / ctx->desc.height);   // sipp_patch_origin
    if ((rc = ingest_window(ix - fov / 2, iy - fov / 2, fov, fov)) != SIPP_OK) return rc;
    double pc = -1.0;
    int32_t st = 0, len = 0;
    if ((rc = plan_impl(ctx, &pc, &st, nullptr, 0, &len)) != SIPP_OK) return rc;
    int64_t bi = -1;
    double bv = 0.0;
    const double* xy = cand_xy + (size_t)c * n * 2;
    if ((rc = cycle_host_impl(ctx, params, n, xy, nullptr, cand_cost + (size_t)c * n, nullptr, nullptr, nullptr, &bi, &bv, false, 0)) != SIPP_OK) return rc;
    if (path_cost) path_cost[c] = pc;
    if (plan_status) plan_status[c] = st;
    if (path_len) path_len[c] = len;
    if (best_index) best_index[c] = (int32_t)bi;
    if (best_value) best_value[c] = bv;
    if (bi >= 0) { px = xy[2 * bi]; py = xy[2 * bi + 1]; }
  }
  return SIPP_OK;
}

static int update_tree_impl(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n, const int32_t* parent,
                            const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                            const double* cur_xy, int32_t path_changed, int32_t* n_rescored, const double* cost_below, const int32_t* view_offset,
                            const double* view_xy);

int sipp_update_tree(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n, const int32_t* parent,
                     const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                     const double* cur_xy, int32_t path_changed, int32_t* n_rescored, const double* cost_below) {
  SIPP_ENTER(ctx);
  return update_tree_impl(ctx, params, updater, n, parent, end_xy, start_xy, gain, cost, value, terminal, cur_xy, path_changed, n_rescored, cost_below, nullptr,
                          nullptr);
}

int sipp_update_tree_views(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n, const int32_t* parent,
                           const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                           const double* cur_xy, int32_t path_changed, int32_t* n_rescored, const double* cost_below, const int32_t* view_offset,
                           const double* view_xy) {
  SIPP_ENTER(ctx);
  if (!view_offset) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "view_offset == NULL");
  return update_tree_impl(ctx, params, updater, n, parent, end_xy, start_xy, gain, cost, value, terminal, cur_xy, path_changed, n_rescored, cost_below,
                          view_offset, view_xy);
}

static int update_tree_impl(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n, const int32_t* parent,
                            const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                            const double* cur_xy, int32_t path_changed, int32_t* n_rescored, const double* cost_below, const int32_t* view_offset,
                            const double* view_xy) {
  int rc = check_params(ctx, params);
  if (rc != SIPP_OK) return rc;
  if (!updater || n < 1 || !parent || !end_xy || !gain || !cost || !value || !terminal || !cur_xy)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad update_tree arguments");
  if (updater->updater < 0 || updater->updater > SIPP_UPD_DIRECT || updater->following < 0 || updater->following > SIPP_FOLLOW_UPDATE_ALL_NON_TERMINAL)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "unknown updater %d / following updater %d", updater->updater, updater->following);
  for (int i = 1; i < n; ++i)
    if (parent[i] < 0 || parent[i] >= i) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "parent[%d] = %d violates 0 <= parent[i] < i", i, parent[i]);
  cudaStream_t s = ctx->stream;
  int64_t nv = 0;
  if (view_offset && (rc = check_views(ctx, n, view_offset, view_xy, &nv)) != SIPP_OK) return rc;
  // device staging: doubles end_xy[2n] start_xy[2n] gain_old gain_new gain_tmp cost value value_tmp cost_below | int32 parent sel n_sel | u8 flags x5
  // | with views: a ViewStage behind it (256-byte aligned)
  const size_t N = (size_t)n, need0 = (N * (16 + 16 + 8 * 7) + N * 8 + 64 + N * 5 + 256 + 255) & ~(size_t)255;
  const size_t need = need0 + (view_offset ? view_stage_bytes(N, (size_t)nv) : 0);
  if ((rc = ensure_dev(ctx, &ctx->d_stage, &ctx->d_stage_bytes, need)) != SIPP_OK) return rc;
  double* d_end = reinterpret_cast<double*>(ctx->d_stage);
  double *d_start = d_end + 2 * N, *d_gold = d_start + 2 * N, *d_gnew = d_gold + N, *d_gtmp = d_gnew + N, *d_cost = d_gtmp + N, *d_val = d_cost + N,
         *d_vtmp = d_val + N, *d_cbelow = d_vtmp + N;
  int32_t* d_parent = reinterpret_cast<int32_t*>(d_cbelow + N);
  int* d_sel = d_parent + N;
  int* d_nsel = d_sel + N;                       // 16 ints reserved
  uint8_t* d_term = reinterpret_cast<uint8_t*>(d_nsel + 16);
  uint8_t *d_ttmp = d_term + N, *d_dog = d_ttmp + N, *d_fol = d_dog + N, *d_dov = d_fol + N;
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_end, end_xy, N * 16, cudaMemcpyHostToDevice, s));
  if (start_xy) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_start, start_xy, N * 16, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_gold, gain, N * 8, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemsetAsync(d_gold, 0, 8, s));       // the root just became the current segment: gain = cost = 0 [upstream]
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cost, cost, N * 8, cudaMemcpyHostToDevice, s));
  if (!cost_below) SIPP_CUDA_CHECK(ctx, cudaMemsetAsync(d_cost, 0, 8, s));     // with cost_below the caller states the root's cost of this pass
  if (cost_below) {     // update_cost: `cost` holds this pass's costs, cost_below last pass's (what computeValue still finds below the visited node)
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cbelow, cost_below, N * 8, cudaMemcpyHostToDevice, s));
    SIPP_CUDA_CHECK(ctx, cudaMemsetAsync(d_cbelow, 0, 8, s));
  }
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_val, value, N * 8, cudaMemcpyHostToDevice, s));
  if (updater->updater != SIPP_UPD_RRT_STAR) SIPP_CUDA_CHECK(ctx, cudaMemsetAsync(d_val, 0, 8, s));   // ... and value = 0
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_parent, parent, N * 4, cudaMemcpyHostToDevice, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(d_term, terminal, N, cudaMemcpyHostToDevice, s));
  // 1. the updater predicate for every node + compaction of the nodes whose gain is recomputed
  SIPP_CUDA_CHECK(ctx, launch_update_select(ctx, *updater, n, d_end, d_gold, d_term, cur_xy[0], cur_xy[1], path_changed != 0, d_dog, d_fol, d_sel, d_nsel, s));
  // 2. one batched gain launch over the compacted list (count read on the device), results scattered by node index
  const GainGeom g = make_gain_geom(ctx, params);
  if ((rc = ensure_views(ctx, params, g, s)) != SIPP_OK) return rc;
  if (view_offset) {
    // several sampled viewpoints per node: every view of every node is scored (the trees are small next to the GPU), the per-node sums
    // land in gain_tmp / terminal_tmp and k_update_apply takes them for the selected nodes only
    const ViewStage vs = carve_views(reinterpret_cast<char*>(ctx->d_stage) + need0, N, (size_t)nv);
    if (nv > 0) SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(vs.view_xy, view_xy, (size_t)nv * 16, cudaMemcpyHostToDevice, s));
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(vs.view_off, view_offset, (N + 1) * 4, cudaMemcpyHostToDevice, s));
    if ((rc = gain_launch_views(ctx, params, g, n, vs.view_off, (int)nv, vs.view_xy, start_xy ? d_start : d_end, vs.vgain, vs.vcounts, vs.vfsum, d_gtmp, nullptr,
                                d_ttmp, s)) != SIPP_OK)
      return rc;
  } else if ((rc = gain_launch(ctx, params, g, n, d_end, start_xy ? d_start : nullptr, nullptr, d_gtmp, nullptr, d_ttmp, nullptr, s, d_sel, d_nsel)) != SIPP_OK) return rc;
  // 3. new gains / terminal flags, which nodes get a new value
  const int sets_terminal = params->evaluator == SIPP_EVAL_RRT_STAR || params->evaluator == SIPP_EVAL_EXPAND_REACHABLE;
  SIPP_CUDA_CHECK(ctx, launch_update_apply(ctx, *updater, n, sets_terminal, d_dog, d_fol, d_gold, d_gtmp, d_ttmp, d_gnew, d_term, d_dov, s));
  // 4. values with the pre-order semantics (new gains on the chain root..i, old gains below i)
  if ((rc = ensure_dev(ctx, &ctx->d_tree_ws, &ctx->d_tree_ws_bytes, N * 16 + 512)) != SIPP_OK) return rc;
  // the root-first call of the patched planner re-scores the root unless the chain starts with an RRTStarUpdater; its new gain then
  // enters every value below it (pinned against the reference's own updater sources: tests/test_ref_pin.py)
  const int root_live = updater->updater != SIPP_UPD_RRT_STAR;
  SIPP_CUDA_CHECK(ctx, launch_value_tree(ctx, n, d_gnew, d_gold, d_cost, d_parent, d_vtmp, ctx->d_best, ctx->d_tree_ws, s, root_live, cost_below ? d_cbelow : nullptr));
  SIPP_CUDA_CHECK(ctx, launch_update_value(ctx, n, d_dov, d_vtmp, d_val, s));
  BestRec* h_best = reinterpret_cast<BestRec*>(ctx->h_pinned);
  int* h_nsel = reinterpret_cast<int*>(h_best + 1);
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(h_best, ctx->d_best, sizeof(BestRec), cudaMemcpyDeviceToHost, s));
  SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(h_nsel, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, s));
  const size_t first = root_live ? 0 : 1;     // an RRTStarUpdater chain never touches the root's numbers
  if (N > first) {
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(gain + first, d_gnew + first, (N - first) * 8, cudaMemcpyDeviceToHost, s));
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(value + first, d_val + first, (N - first) * 8, cudaMemcpyDeviceToHost, s));
    SIPP_CUDA_CHECK(ctx, cudaMemcpyAsync(terminal + first, d_term + first, N - first, cudaMemcpyDeviceToHost, s));
  }
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(s));
  if (h_best->index == -2) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "tree deeper than 512 levels");
  if (n_rescored) *n_rescored = *h_nsel;
  return SIPP_OK;
}

int sipp_merge_best_dev(sipp_ctx* ctx, const void* d_recs, int32_t world, int64_t shard_stride, void* d_out, void* stream) {
  SIPP_ENTER(ctx);
  if (!d_recs || !d_out || world < 1) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad merge_best arguments");
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  SIPP_CUDA_CHECK(ctx, launch_merge_best(ctx, reinterpret_cast<const BestRec*>(d_recs), world, (long long)shard_stride, reinterpret_cast<BestRec*>(d_out), s));
  return SIPP_OK;
}

// =====================================================================================================
// Batched ensembles (BASELINE config 5): m maps of one size on one device advance TOGETHER — every kernel of a
// step (patch ingest, re-plan, view planes, gain + value + argmax) is launched once with gridDim.y = m and reads its per-map argument
// block from device memory, instead of m contexts x ~20 stream operations per map-cycle fed by m host threads. One host->device
// copy per step carries all argument blocks and inputs, one device->host copy all results, one synchronisation per map-cycle.
// The contexts keep all state (sipp_map_download, sipp_field_download, ... keep working between batch calls); they must be idle
// when a batch call starts and must not be driven from other threads while it runs.
// Replaces, for evaluation runs over many maps, data_analysis/evaluation.py:39-62,139-144 + planexp_eval_planner.cpp:201-303
// (one planner process per saved map) and advanced_planner_ros/scripts/run_experiment.sh:190-212 (one experiment per map).
// host work of a batch step is a loop over maps with no dependence between them (the sequential scalars of an ingest, the seed
// tiles of a plan, argument blocks): a few helper threads share it with the calling thread. No CUDA calls are made from the helpers.
struct BatchPool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  const std::function<void(int)>* fn = nullptr;
  int n = 0, active = 0;
  std::atomic<int> next{0};
  uint64_t gen = 0;
  bool stop = false;
  void start(int helpers) {
    for (int t = 0; t < helpers; ++t)
      th.emplace_back([this] {
        uint64_t seen = 0;
        for (;;) {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || gen != seen; });
          if (stop) return;
          seen = gen;
          lk.unlock();
          for (int i; (i = next.fetch_add(1)) < n;) (*fn)(i);
          lk.lock();
          if (--active == 0) cv_done.notify_one();
        }
      });
  }
  void run(int count, const std::function<void(int)>& f) {
    if (th.empty() || count < 8) {
      for (int i = 0; i < count; ++i) f(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      fn = &f; n = count; next.store(0); active = (int)th.size(); ++gen;
    }
    cv.notify_all();
    for (int i; (i = next.fetch_add(1)) < count;) f(i);
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return active == 0; });
  }
  ~BatchPool() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
};

struct sipp_batch {
  int m, device;
  BatchPool pool;
  std::vector<sipp_ctx*> ctx;
  cudaStream_t stream;
  std::mutex mu;
  size_t plan_args, seed_stride;      // bytes per map
  // pinned host / device staging; one region per step of a map-cycle so that no step overwrites what an earlier one still copies
  uint8_t *h_plan, *d_plan;           // m x PlanArgsAll
  uint8_t *h_seed, *d_seed;           // m x seed_stride
  int *h_out, *d_out;                 // m x 40: plan_out[0..9], relaxation control block at [16..39]
  uint8_t *h_ing, *d_ing;  size_t ing_bytes;     // ingest: m argument blocks + m patches
  uint8_t *h_cyc, *d_cyc;  size_t cyc_bytes;     // cycle: m view blocks + m gain blocks + candidates (xy, cost)
  uint8_t *d_res;  size_t res_bytes;  // cycle scratch: gain, value, terminal per candidate; best records at the front
  BestRec* h_best;                    // m records
  int* d_roam;  int roam_cap;  bool roam_dirty;  int roam_share, roam_kind;     // shared relaxation queue (32 control ints, then roam_cap slots); cap 0 = off
  int64_t launches;
  char err[320];
};

static int batch_err(sipp_batch* b, int rc, sipp_ctx* from) {
  if (from) snprintf(b->err, sizeof(b->err), "%s", from->err.c_str());
  return rc;
}
static int batch_fail(sipp_batch* b, int rc, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b->err, sizeof(b->err), fmt, ap);
  va_end(ap);
  return rc;
}
#define SIPP_BATCH_CUDA(b, expr)                                                                                          \
  do {                                                                                                                    \
    cudaError_t e__ = (expr);                                                                                             \
    if (e__ != cudaSuccess) return batch_fail((b), SIPP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static int batch_grow(sipp_batch* b, uint8_t** h, uint8_t** d, size_t* cur, size_t need) {
  if (*cur >= need) return SIPP_OK;
  SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
  if (h && *h) cudaFreeHost(*h);
  if (*d) cudaFree(*d);
  if (h) *h = nullptr;
  *d = nullptr; *cur = 0;
  need = (need + (need >> 2) + 4095) & ~(size_t)4095;
  if (h) SIPP_BATCH_CUDA(b, cudaMallocHost(reinterpret_cast<void**>(h), need));
  SIPP_BATCH_CUDA(b, cudaMalloc(reinterpret_cast<void**>(d), need));
  *cur = need;
  return SIPP_OK;
}

const char* sipp_batch_last_error(const sipp_batch* b) { return b ? b->err : "batch == NULL"; }
int64_t sipp_batch_launch_count(const sipp_batch* b) { return b ? b->launches : 0; }

void sipp_batch_destroy(sipp_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
  void* hp[] = {b->h_plan, b->h_seed, b->h_out, b->h_ing, b->h_cyc, b->h_best};
  void* dp[] = {b->d_plan, b->d_seed, b->d_out, b->d_ing, b->d_cyc, b->d_res, b->d_roam};
  for (void* p : hp) if (p) cudaFreeHost(p);
  for (void* p : dp) if (p) cudaFree(p);
  delete b;
}

int sipp_batch_create(sipp_ctx* const* ctxs, int32_t m, sipp_batch** out) {
  if (!ctxs || m < 1 || !out) return SIPP_ERR_INVALID_ARG;
  *out = nullptr;
  for (int i = 0; i < m; ++i) {
    if (!ctxs[i]) return SIPP_ERR_INVALID_ARG;
    if (ctxs[i]->device != ctxs[0]->device || ctxs[i]->W != ctxs[0]->W || ctxs[i]->H != ctxs[0]->H)
      return sipp_set_err(ctxs[i], SIPP_ERR_INVALID_ARG, "a batch takes maps of one size on one device (map %d differs from map 0)", i);
    if (ctxs[i]->desc.compute_closest_observed)
      return sipp_set_err(ctxs[i], SIPP_ERR_UNSUPPORTED, "batched ingest does not maintain the closest-observed planes (compute_closest_observed)");
    for (int j = 0; j < i; ++j)
      if (ctxs[j] == ctxs[i]) return sipp_set_err(ctxs[i], SIPP_ERR_INVALID_ARG, "context %d appears twice in the batch", i);
  }
  sipp_batch* b = new sipp_batch();
  b->m = m;
  b->device = ctxs[0]->device;
  b->ctx.assign(ctxs, ctxs + m);
  b->stream = nullptr;
  b->h_plan = b->d_plan = b->h_seed = b->d_seed = b->h_ing = b->d_ing = b->h_cyc = b->d_cyc = b->d_res = nullptr;
  b->h_out = b->d_out = nullptr;
  b->h_best = nullptr;
  b->ing_bytes = b->cyc_bytes = b->res_bytes = 0;
  b->launches = 0;
  b->err[0] = 0;
  auto fail = [&](int rc) { sipp_set_err(ctxs[0], rc, "%s", b->err); sipp_batch_destroy(b); return rc; };
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(batch_fail(b, SIPP_ERR_CUDA, "cudaSetDevice(%d) failed", b->device));
  const int ntiles = ctxs[0]->n_tiles_x * ctxs[0]->n_tiles_y;
  b->plan_args = field_plan_batch_args_bytes();
  b->seed_stride = ((((size_t)ntiles + 3) & ~(size_t)3) + (size_t)ntiles * sizeof(int) + 15) & ~(size_t)15;
  cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&b->h_plan), b->plan_args * m);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->d_plan), b->plan_args * m);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&b->h_seed), b->seed_stride * m);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->d_seed), b->seed_stride * m);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&b->h_out), sizeof(int) * 40 * m);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->d_out), sizeof(int) * 40 * m);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&b->h_best), sizeof(BestRec) * m);
  // one relaxation work queue for all maps (SIPP_BATCH_ROAM=0: every map's relaxation on worker CTAs of its own); SIPP_BATCH_SHARE = how
  // many batches the caller runs side by side on this GPU (each takes that fraction of the SMs for its persistent workers)
  // SIPP_BATCH_ROAM: 0 = off, 1 = workers with the staged weight copy (5 warps per SM), 2 = weights read from global memory (16 warps per SM)
  b->roam_kind = getenv("SIPP_BATCH_ROAM") ? atoi(getenv("SIPP_BATCH_ROAM")) : 2;
  b->roam_cap = b->roam_kind == 0 ? 0 : (int)field_roam_queue_ints(ntiles, m);
  b->roam_share = getenv("SIPP_BATCH_SHARE") ? atoi(getenv("SIPP_BATCH_SHARE")) : 1;
  b->roam_dirty = true;
  if (e == cudaSuccess && b->roam_cap > 0) e = cudaMalloc(reinterpret_cast<void**>(&b->d_roam), sizeof(int) * (32 + (size_t)b->roam_cap));
  if (e != cudaSuccess) return fail(batch_fail(b, SIPP_ERR_CUDA, "batch staging: %s", cudaGetErrorString(e)));
  // helper threads for the per-map host loops: SIPP_BATCH_THREADS (total, including the caller), default 4
  int threads = getenv("SIPP_BATCH_THREADS") ? atoi(getenv("SIPP_BATCH_THREADS")) : 4;
  if (threads > 64) threads = 64;
  if (threads > 1 && m >= 8) b->pool.start(threads - 1);
  *out = b;
  return SIPP_OK;
}

#define SIPP_BATCH_ENTER(b)                                                                                       \
  if (!(b)) return SIPP_ERR_INVALID_ARG;                                                                          \
  std::lock_guard<std::mutex> lock__((b)->mu);                                                                    \
  if (cudaSetDevice((b)->device) != cudaSuccess) return batch_fail((b), SIPP_ERR_CUDA, "cudaSetDevice failed")

int sipp_batch_set_sharing(sipp_batch* b, int32_t batches_on_this_gpu) {
  SIPP_BATCH_ENTER(b);
  if (batches_on_this_gpu < 1 || batches_on_this_gpu > 64) return batch_fail(b, SIPP_ERR_INVALID_ARG, "batches_on_this_gpu %d out of range [1, 64]", batches_on_this_gpu);
  b->roam_share = batches_on_this_gpu;
  return SIPP_OK;
}

// ---- steps: enqueue on the batch stream; nothing here synchronises ------------------------------------------------------------
// patches: labels + (size_t)i * label_stride = the rows x cols patch of map i (host memory)
static int batch_ingest_enqueue(sipp_batch* b, const int32_t* x0, const int32_t* y0, int rows, int cols, const int32_t* const* labels) {
  const int m = b->m;
  if (!x0 || !y0 || !labels || rows <= 0 || cols <= 0) return batch_fail(b, SIPP_ERR_INVALID_ARG, "bad batch patch %d x %d", rows, cols);
  // [m argument blocks, packed][pad to 256][m patches, 16-byte aligned each]
  const size_t ab = ingest_args_bytes(), pb = ((size_t)rows * cols * sizeof(int32_t) + 15) & ~(size_t)15;
  const size_t poff = (ab * m + 255) & ~(size_t)255;
  int rc = batch_grow(b, &b->h_ing, &b->d_ing, &b->ing_bytes, poff + pb * m);
  if (rc != SIPP_OK) return rc;
  for (int i = 0; i < m; ++i)
    if (!labels[i]) return batch_fail(b, SIPP_ERR_INVALID_ARG, "patch %d == NULL", i);
  std::vector<int> rcs((size_t)m, SIPP_OK);
  b->pool.run(m, [&](int i) {
    sipp_ctx* ctx = b->ctx[i];
    if ((rcs[i] = ingest_host_part(ctx, x0[i], y0[i], rows, cols, labels[i], nullptr)) != SIPP_OK) return;
    memcpy(b->h_ing + poff + pb * i, labels[i], (size_t)rows * cols * sizeof(int32_t));
    fill_ingest_args(ctx, x0[i], y0[i], rows, cols, reinterpret_cast<const int32_t*>(b->d_ing + poff + pb * i), b->h_ing + ab * i);
  });
  for (int i = 0; i < m; ++i)
    if (rcs[i] != SIPP_OK) return batch_err(b, rcs[i], b->ctx[i]);
  SIPP_BATCH_CUDA(b, cudaMemcpyAsync(b->d_ing, b->h_ing, poff + pb * m, cudaMemcpyHostToDevice, b->stream));
  SIPP_BATCH_CUDA(b, launch_ingest_patch_batch(b->d_ing, m, rows, cols, b->stream));
  b->launches += 1;
  return SIPP_OK;
}

static int batch_plan_enqueue(sipp_batch* b) {
  int n = 0;
  const std::function<void(int, const std::function<void(int)>&)> runner = [b](int count, const std::function<void(int)>& f) { b->pool.run(count, f); };
  const bool roam = b->roam_cap > 0;
  if (roam && b->roam_dirty) {
    const int rr = field_roam_reset(b->ctx[0], b->d_roam + 32, b->roam_cap, b->d_roam, b->stream);
    if (rr != SIPP_OK) return batch_err(b, rr, b->ctx[0]);
    b->roam_dirty = false;
    b->launches += 1;
  }
  int rc = field_plan_batch(b->ctx.data(), b->m, b->h_plan, b->d_plan, b->h_seed, b->d_seed, b->seed_stride, b->d_out, b->stream, &n, &runner,
                            roam ? b->d_roam + 32 : nullptr, b->roam_cap, roam ? b->d_roam : nullptr, b->roam_share, b->roam_kind);
  if (rc != SIPP_OK) {
    for (sipp_ctx* c : b->ctx) c->warm_valid = false;
    return batch_err(b, rc, b->ctx[0]);
  }
  b->launches += n;
  SIPP_BATCH_CUDA(b, cudaMemcpyAsync(b->h_out, b->d_out, sizeof(int) * 40 * b->m, cudaMemcpyDeviceToHost, b->stream));
  for (sipp_ctx* c : b->ctx) c->path_valid = true;      // the masks exist in stream order; plan_collect confirms after the synchronisation
  return SIPP_OK;
}
// after the stream was synchronised
static int batch_plan_collect(sipp_batch* b, double* cost, int32_t* status, int32_t* path_len, size_t stride) {
  int first_rc = SIPP_OK;
  for (int i = 0; i < b->m; ++i) {
    int len = 0;
    bool found = false;
    double c = -1.0;
    int32_t st = 0;
    const int rc = plan_collect(b->ctx[i], b->h_out + 40 * i, b->h_out + 40 * i + 16, &c, &st, &len, &found);
    if (rc != SIPP_OK && first_rc == SIPP_OK) first_rc = batch_err(b, rc, b->ctx[i]);
    if (rc != SIPP_OK) b->roam_dirty = true;      // a watchdog abort leaves entries in the shared queue
    if (cost) cost[i * stride] = c;
    if (status) status[i * stride] = st;
    if (path_len) path_len[i * stride] = found ? len : 0;
  }
  return first_rc;
}

// candidates of map i: xy[i] (n x 2 doubles), cost[i] (n doubles), host memory
static int batch_cycle_enqueue(sipp_batch* b, const sipp_eval_params* params, int n, const double* const* xy, const double* const* cost) {
  const int m = b->m;
  if (!params || n < 1 || !xy || !cost) return batch_fail(b, SIPP_ERR_INVALID_ARG, "bad batch cycle arguments");
  // [m view blocks, packed][pad][m gain blocks, packed][pad][m x (xy, cost)]
  const size_t vb = build_views_args_bytes(), gb = gain_args_bytes();
  const size_t goff = (vb * m + 255) & ~(size_t)255, coff = (goff + gb * m + 255) & ~(size_t)255;
  const size_t cb = ((size_t)n * 24 + 15) & ~(size_t)15;      // xy + cost of one map (the kernel reads xy as 16-byte pairs)
  int rc = batch_grow(b, &b->h_cyc, &b->d_cyc, &b->cyc_bytes, coff + cb * m);
  if (rc != SIPP_OK) return rc;
  // scratch per map: gain, value (doubles), terminal (bytes, padded); the m best records sit at the front
  const size_t rb = ((size_t)n * 17 + 15) & ~(size_t)15, best_bytes = ((size_t)m * sizeof(BestRec) + 255) & ~(size_t)255;
  if ((rc = batch_grow(b, nullptr, &b->d_res, &b->res_bytes, best_bytes + rb * m)) != SIPP_OK) return rc;
  uint8_t* h_views = b->h_cyc;
  uint8_t* h_gain = b->h_cyc + goff;
  uint8_t* h_cand = b->h_cyc + coff;
  uint8_t* d_cand = b->d_cyc + coff;
  BestRec* d_best = reinterpret_cast<BestRec*>(b->d_res);
  bool any_views = false;
  std::vector<uint8_t> stale_views((size_t)m, 0);
  if (params->evaluator == SIPP_EVAL_EXPAND_LOW_COST) return batch_fail(b, SIPP_ERR_UNSUPPORTED, "batched evaluation does not serve ExpandLowCostAreaEvaluator");
  std::vector<int> kinds((size_t)m, 0), smems((size_t)m, 0), wpcs((size_t)m, 0);
  for (int i = 0; i < m; ++i)
    if (!xy[i] || !cost[i]) return batch_fail(b, SIPP_ERR_INVALID_ARG, "candidates of map %d == NULL", i);
  std::vector<int> rcs((size_t)m, SIPP_OK);
  b->pool.run(m, [&](int i) {
    sipp_ctx* ctx = b->ctx[i];
    if ((rcs[i] = check_params(ctx, params)) != SIPP_OK) return;
    const GainGeom g = make_gain_geom(ctx, params);
    const bool stale = !(!ctx->views_dirty && ctx->view_params_valid && same_bv(ctx->view_params, *params));
    stale_views[i] = stale ? 1 : 0;
    fill_build_views_args(ctx, g, params->use_bounding_volume != 0, stale ? 1 : 0, h_views + vb * i);
    memcpy(h_cand + cb * i, xy[i], (size_t)n * 16);
    memcpy(h_cand + cb * i + (size_t)n * 16, cost[i], (size_t)n * 8);
    const double* d_xy = reinterpret_cast<const double*>(d_cand + cb * i);
    const double* d_cost = d_xy + (size_t)n * 2;
    double* d_gain = reinterpret_cast<double*>(b->d_res + best_bytes + rb * i);
    double* d_value = d_gain + n;
    uint8_t* d_term = reinterpret_cast<uint8_t*>(d_value + n);
    FusedSelect fs;
    fs.best_out = d_best + i;
    fs.index_offset = 0;
    fs.peer = nullptr;
    fs.peer_mode = 0;
    rcs[i] = gain_fill_batch_args(ctx, params, g, n, d_xy, d_cost, d_gain, d_term, d_value, &fs, h_gain + gb * i, &kinds[i], &smems[i], &wpcs[i]);
  });
  for (int i = 0; i < m; ++i) {
    if (rcs[i] != SIPP_OK) return batch_err(b, rcs[i], b->ctx[i]);
    any_views |= stale_views[i] != 0;
    if (kinds[i] != kinds[0] || smems[i] != smems[0] || wpcs[i] != wpcs[0]) return batch_fail(b, SIPP_ERR_INVALID_ARG, "map %d needs another kernel shape than map 0", i);
  }
  SIPP_BATCH_CUDA(b, cudaMemcpyAsync(b->d_cyc, b->h_cyc, coff + cb * m, cudaMemcpyHostToDevice, b->stream));
  for (int i = 0; i < m; ++i)          // every argument block is in place: only now do the contexts learn that their views get rebuilt
    if (stale_views[i]) {
      sipp_ctx* ctx = b->ctx[i];
      if (!(ctx->view_params_valid && same_bv(ctx->view_params, *params))) ctx->lowcost_view_dirty = true;
      ctx->views_dirty = false;
      ctx->view_params = *params;
      ctx->view_params_valid = true;
    }
  if (any_views) {
    SIPP_BATCH_CUDA(b, launch_build_views_batch(b->d_cyc, m, b->ctx[0]->pitch, b->ctx[0]->H, b->stream));
    b->launches += 1;
  }
  if ((rc = gain_launch_batch(b->ctx[0], b->d_cyc + goff, m, n, kinds[0], smems[0], wpcs[0], b->stream)) != SIPP_OK) return batch_err(b, rc, b->ctx[0]);
  b->launches += 1;
  SIPP_BATCH_CUDA(b, cudaMemcpyAsync(b->h_best, d_best, sizeof(BestRec) * m, cudaMemcpyDeviceToHost, b->stream));
  return SIPP_OK;
}

// ---- public calls --------------------------------------------------------------------------------------------------------------
int sipp_batch_ingest(sipp_batch* b, const int32_t* x0, const int32_t* y0, int32_t rows, int32_t cols, const int32_t* const* labels) {
  SIPP_BATCH_ENTER(b);
  int rc = batch_ingest_enqueue(b, x0, y0, rows, cols, labels);
  if (rc != SIPP_OK) return rc;
  SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
  return SIPP_OK;
}

int sipp_batch_plan(sipp_batch* b, double* cost, int32_t* status, int32_t* path_len) {
  SIPP_BATCH_ENTER(b);
  int rc = batch_plan_enqueue(b);
  if (rc != SIPP_OK) return rc;
  SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
  return batch_plan_collect(b, cost, status, path_len, 1);
}

int sipp_batch_cycle(sipp_batch* b, const sipp_eval_params* params, int32_t n, const double* const* xy, const double* const* cost, int64_t* best_index,
                     double* best_value, double* const* gain) {
  SIPP_BATCH_ENTER(b);
  int rc = batch_cycle_enqueue(b, params, n, xy, cost);
  if (rc != SIPP_OK) return rc;
  if (gain) {
    const size_t rb = ((size_t)n * 17 + 15) & ~(size_t)15, best_bytes = ((size_t)b->m * sizeof(BestRec) + 255) & ~(size_t)255;
    for (int i = 0; i < b->m; ++i)
      if (gain[i]) SIPP_BATCH_CUDA(b, cudaMemcpyAsync(gain[i], b->d_res + best_bytes + rb * i, (size_t)n * 8, cudaMemcpyDeviceToHost, b->stream));
  }
  SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
  for (int i = 0; i < b->m; ++i) {
    if (best_index) best_index[i] = b->h_best[i].index;
    if (best_value) best_value[i] = b->h_best[i].value;
  }
  return SIPP_OK;
}

// m episodes in lockstep (sipp_episode_run for every map of the batch): per cycle ONE ingest launch, one plan sequence, one scoring
// launch and one synchronisation for all maps. truth[i]: W x H labels of map i; cand_xy[i] / cand_cost[i]: cycles x n candidates of map i.
// Outputs (any may be NULL): m x cycles, map-major.
int sipp_batch_episode_run(sipp_batch* b, const int32_t* const* truth, const sipp_eval_params* params, int32_t cycles, int32_t n,
                           const double* const* cand_xy, const double* const* cand_cost, int32_t fov, int32_t init, double start_x, double start_y,
                           double* path_cost, int32_t* plan_status, int32_t* path_len, int32_t* best_index, double* best_value) {
  SIPP_BATCH_ENTER(b);
  if (!truth || !params || cycles < 0 || n < 1 || !cand_xy || !cand_cost || fov < 1 || fov > 4096 || init < 0)
    return batch_fail(b, SIPP_ERR_INVALID_ARG, "bad episode arguments");
  const int m = b->m, W = b->ctx[0]->W, H = b->ctx[0]->H;
  for (int i = 0; i < m; ++i)
    if (!truth[i] || !cand_xy[i] || !cand_cost[i]) return batch_fail(b, SIPP_ERR_INVALID_ARG, "inputs of map %d == NULL", i);
  std::vector<std::vector<int32_t>> patch((size_t)m);
  std::vector<const int32_t*> pp((size_t)m);
  std::vector<int32_t> x0((size_t)m), y0((size_t)m);
  std::vector<double> px((size_t)m, start_x), py((size_t)m, start_y);
  std::vector<const double*> xy((size_t)m), cc((size_t)m);
  auto cut = [&](int rows, int cols) {       // the simulator's MapData patches: truth labels, 0 off the image
    b->pool.run(m, [&](int k) {
      patch[k].assign((size_t)rows * cols, 0);
      for (int i = 0; i < rows; ++i) {
        const int y = y0[k] + i;
        if (y < 0 || y >= H) continue;
        for (int j = 0; j < cols; ++j) {
          const int x = x0[k] + j;
          if (x >= 0 && x < W) patch[k][(size_t)i * cols + j] = truth[k][(size_t)y * W + x];
        }
      }
      pp[k] = patch[k].data();
    });
  };
  int rc;
  if (init > 0) {       // mav_simulator.cpp:156-159
    for (int k = 0; k < m; ++k) x0[k] = y0[k] = 0;
    const int rows = init < H ? init : H, cols = init < W ? init : W;
    cut(rows, cols);
    if ((rc = batch_ingest_enqueue(b, x0.data(), y0.data(), rows, cols, pp.data())) != SIPP_OK) return rc;
    SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
  }
  // SIPP_BATCH_TIMING=1: where a batch-cycle's wall time goes (host preparation per step, the wait for the stream) and how long the
  // stream was busy (events around the enqueued work), printed once per episode
  const bool timing = getenv("SIPP_BATCH_TIMING") && atoi(getenv("SIPP_BATCH_TIMING")) != 0;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};      // cut, ingest, plan, cycle (host, enqueue included), wait, collect, stream busy (ms)
  cudaEvent_t tev[2] = {nullptr, nullptr};
  if (timing) { cudaEventCreate(&tev[0]); cudaEventCreate(&tev[1]); }
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  for (int c = 0; c < cycles; ++c) {
    auto tp = now();
    auto lap = [&](int i) { if (timing) { acc[i] += ms_since(tp); tp = now(); } };
    if (timing) cudaEventRecord(tev[0], b->stream);
    for (int k = 0; k < m; ++k) {
      const sipp_ctx* ctx = b->ctx[k];
      const int ix = (int)(px[k] * (long)W / ctx->desc.width), iy = (int)(py[k] * (long)H / ctx->desc.height);   // sipp_patch_origin
      x0[k] = ix - fov / 2;
      y0[k] = iy - fov / 2;
      xy[k] = cand_xy[k] + (size_t)c * n * 2;
      cc[k] = cand_cost[k] + (size_t)c * n;
    }
    cut(fov, fov);
    lap(0);
    if ((rc = batch_ingest_enqueue(b, x0.data(), y0.data(), fov, fov, pp.data())) != SIPP_OK) return rc;
    lap(1);
    if ((rc = batch_plan_enqueue(b)) != SIPP_OK) return rc;
    lap(2);
    if ((rc = batch_cycle_enqueue(b, params, n, xy.data(), cc.data())) != SIPP_OK) return rc;
    if (timing) cudaEventRecord(tev[1], b->stream);
    lap(3);
    SIPP_BATCH_CUDA(b, cudaStreamSynchronize(b->stream));
    lap(4);
    if ((rc = batch_plan_collect(b, path_cost ? path_cost + c : nullptr, plan_status ? plan_status + c : nullptr, path_len ? path_len + c : nullptr,
                                 (size_t)cycles)) != SIPP_OK)
      return rc;
    for (int k = 0; k < m; ++k) {
      const int64_t bi = b->h_best[k].index;
      if (best_index) best_index[(size_t)k * cycles + c] = (int32_t)bi;
      if (best_value) best_value[(size_t)k * cycles + c] = b->h_best[k].value;
      if (bi >= 0) { px[k] = xy[k][2 * bi]; py[k] = xy[k][2 * bi + 1]; }
    }
    lap(5);
    if (timing) {
      float t = 0;
      if (cudaEventElapsedTime(&t, tev[0], tev[1]) == cudaSuccess) acc[6] += t;
    }
  }
  if (timing) {
    const double k = cycles > 0 ? 1.0 / cycles : 0.0;
    fprintf(stderr, "[sipp] batch of %d maps, ms per batch-cycle: host cut %.3f, ingest %.3f, plan %.3f, cycle %.3f, wait for the stream %.3f, collect %.3f; "
            "first launch to last kernel on the stream %.3f\n", m, acc[0] * k, acc[1] * k, acc[2] * k, acc[3] * k, acc[4] * k, acc[5] * k, acc[6] * k);
    cudaEventDestroy(tev[0]); cudaEventDestroy(tev[1]);
  }
  return SIPP_OK;
}

}  // extern "C"
