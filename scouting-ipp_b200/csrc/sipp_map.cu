// Map-store kernels: patch ingest (A10), evaluator view planes, debug read-back.
//
// Reference behaviour restated here (cell loop of MapServerPlanExp::mapDataCallback,
// map_server_planexp/src/map_server_planexp.cpp:238-260, and PRMStarPlanner::update_states,
// planexp_base_planners/src/prm_star_planner.cpp:104-119). The sequential scalars of that loop
// (mean_cost_, max_cost_, the goal-explored switch) are kept on the host in sipp_ctx.cu; everything
// per-cell runs here.
#include <cmath>

#include "sipp_internal.h"

namespace {

struct IdList {
  int n;
  int ids[SIPP_MAX_IDS];
};

__device__ __forceinline__ bool in_ids(const IdList& l, int v) {
  bool hit = false;
#pragma unroll 4
  for (int i = 0; i < l.n; ++i) hit |= (l.ids[i] == v);
  return hit;
}

// One thread per patch cell. labels: rows x cols int32 row-major (MapData.data); observed: optional mask.
struct IngestArgs {
  uint8_t* state; uint16_t* label; uint16_t* pcode;
  size_t pitch, pitch16;
  int W, H, x0, y0, rows, cols;
  const int32_t* labels; const uint8_t* observed;
  IdList mav, rov;
  uint8_t* mark; int init_cost_ceil;
};
__device__ __forceinline__ void ingest_patch_body(const IngestArgs& a) {
  uint8_t* __restrict__ state = a.state; uint16_t* __restrict__ label = a.label; uint16_t* __restrict__ pcode = a.pcode;
  const size_t pitch = a.pitch, pitch16 = a.pitch16;
  const int W = a.W, H = a.H, x0 = a.x0, y0 = a.y0, rows = a.rows, cols = a.cols;
  const int32_t* __restrict__ labels = a.labels; const uint8_t* __restrict__ observed = a.observed;
  const IdList& mav = a.mav; const IdList& rov = a.rov;
  uint8_t* __restrict__ mark = a.mark; const int init_cost_ceil = a.init_cost_ceil;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * cols) return;
  const int i = (int)(t / cols), j = (int)(t % cols);
  const int y = y0 + i, x = x0 + j;
  if (!(0 <= y && 0 <= x && y < H && x < W)) return;             // map_server_planexp.cpp:240
  if (observed && !observed[t]) return;
  const int lab = labels[t];
  label[(size_t)y * pitch16 + x] = (uint16_t)lab;                  // map_cost_ = label  :242
  state[(size_t)y * pitch + x] = in_ids(mav, lab) ? SIPP_STATE_OCCUPIED : SIPP_STATE_FREE;  // :252-256
  uint16_t* pc = pcode + (size_t)y * pitch16 + x;
  if (*pc == PCODE_UNEXPLORED) {                                   // is_new  prm_star_planner.cpp:111
    const bool removed = in_ids(rov, lab);
    *pc = removed ? (uint16_t)PCODE_REMOVED : (uint16_t)lab;       // :112-115
    // incremental re-plan: the planner believed init_cost here; a cell whose cost went UP (or whose edges are gone) invalidates
    // every cost-to-go value whose path runs through it (k_inval follows the predecessor plane from these marks)
    if (mark && (removed || lab > init_cost_ceil)) mark[(size_t)y * pitch + x] = 1;
  }
}

// MapServerPlanExp::setClosestObservedMaps (map_server_planexp.cpp:803-823) for one ingested patch [tlx,brx) x [tly,bry):
// every UNKNOWN cell within 15 columns / 10 rows of the patch takes over the cost of the nearest cell of the closed box
// [tlx..brx] x [tly..bry] if that is closer than what it had. One thread per window cell; a cell only writes its own
// entries and only reads map_cost_, so the reference's sequential double loop is order independent.
// dist2 holds the SQUARED pixel distance (an exact small integer in fp32): sqrt is monotone on distinct integers, so
// `stored <= distance` of the reference is `stored2 <= distance2`; the initial value sqrt(width^2 + height^2) [metres!] is
// represented by the sentinel qmax + 0.5, qmax = largest integer whose root is below it. The reference admits i == W /
// j == H and a box that reaches one cell past the patch (and the map): those out-of-range accesses are skipped.
__global__ void k_closest_update(const uint8_t* __restrict__ state, const uint16_t* __restrict__ label, uint16_t* __restrict__ cclab,
                                 float* __restrict__ dist2, size_t pitch, size_t pitch16, int W, int H, int tlx, int tly, int brx, int bry) {
  const int ww = brx - tlx + 31, wh = bry - tly + 21;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ww * wh) return;
  const int i = tlx - 15 + (int)(t % ww), j = tly - 10 + (int)(t / ww);
  if (i < 0 || i >= W || j < 0 || j >= H) return;
  if (state[(size_t)j * pitch + i] != SIPP_STATE_UNKNOWN) return;
  const int cx = min(brx, max(tlx, i)), cy = min(bry, max(tly, j));
  const float q = (float)((i - cx) * (i - cx) + (j - cy) * (j - cy));
  float* dq = dist2 + (size_t)j * pitch16 + i;
  if (*dq <= q) return;
  if (cx < 0 || cx >= W || cy < 0 || cy >= H) return;
  *dq = q;
  cclab[(size_t)j * pitch16 + i] = label[(size_t)cy * pitch16 + cx];
}

__global__ void k_fill_f32(float* p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// ExpandLowCostAreaEvaluator view (expand_low_cost_area_evaluator.cpp:50-69): per cell a 16-bit entry
//   bits 0..14  idx = closest-observed cost + 1 when the cell is UNKNOWN (labels up to 32766: ensure_views refuses larger ones) (1 when there is no closest cost: plane not computed
//               -> the lookup returns -1, or never set -> 0), 0 when the cell is observed
//   bit 15      the corner voxel of the cell is counted: UNKNOWN and inside the bounding volume
// The kernel adds recip[idx] per counted cell: recip[1] = augment ? 1/mean : 0, recip[k] = 1/(k-1).
__global__ void k_build_lowcost_view(const uint8_t* __restrict__ state, const uint16_t* __restrict__ cclab, uint16_t* __restrict__ view,
                                     size_t pitch, size_t pitch16, int W, int H, int bv_kx_lo, int bv_kx_hi, int bv_ky_lo, int bv_ky_hi) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)pitch16 * H) return;
  const int y = (int)(t / (long long)pitch16), x = (int)(t % (long long)pitch16);
  uint32_t e = 0;
  if (x < W && state[(size_t)y * pitch + x] == SIPP_STATE_UNKNOWN) {
    e = (cclab ? (uint32_t)cclab[(size_t)y * pitch16 + x] : 0u) + 1u;
    if (x >= bv_kx_lo && x <= bv_kx_hi && y >= bv_ky_lo && y <= bv_ky_hi) e |= 0x8000u;
  }
  view[(size_t)y * pitch16 + x] = (uint16_t)e;
}

// Builds the evaluator view planes. 16 cells per thread, 128-bit loads/stores.
//   rec     : packed gain record (REC_* bits)
//   freelab : label where the cell is observed FREE and inside the bounding volume, else 0
struct BuildViewsArgs {
  const uint8_t* state; const uint8_t* path; const uint16_t* label;
  uint8_t* rec; uint16_t* freelab;
  size_t pitch, pitch16;
  int W, H, bv_kx_lo, bv_kx_hi, bv_ky_lo, bv_ky_hi, reach_x, reach_y, use_reach;
  uint8_t* rec2_path; uint8_t* rec2_reach;
  int on;       // batched launch: this map's views are stale
};
__device__ __forceinline__ void build_views_body(const BuildViewsArgs& a) {
  const uint8_t* __restrict__ state = a.state; const uint8_t* __restrict__ path = a.path; const uint16_t* __restrict__ label = a.label;
  uint8_t* __restrict__ rec = a.rec; uint16_t* __restrict__ freelab = a.freelab;
  const size_t pitch = a.pitch, pitch16 = a.pitch16;
  const int W = a.W, H = a.H, bv_kx_lo = a.bv_kx_lo, bv_kx_hi = a.bv_kx_hi, bv_ky_lo = a.bv_ky_lo, bv_ky_hi = a.bv_ky_hi;
  const int reach_x = a.reach_x, reach_y = a.reach_y, use_reach = a.use_reach;
  uint8_t* __restrict__ rec2_path = a.rec2_path; uint8_t* __restrict__ rec2_reach = a.rec2_reach;
  const int chunks_per_row = (int)(pitch / 16);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)chunks_per_row * H) return;
  const int y = (int)(t / chunks_per_row);
  const int xc = (int)(t % chunks_per_row) * 16;
  const uint4 sv = *reinterpret_cast<const uint4*>(state + (size_t)y * pitch + xc);
  const uint4 pv = *reinterpret_cast<const uint4*>(path + (size_t)y * pitch + xc);
  const uint4 l0 = *reinterpret_cast<const uint4*>(label + (size_t)y * pitch16 + xc);
  const uint4 l1 = *reinterpret_cast<const uint4*>(label + (size_t)y * pitch16 + xc + 8);
  const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
  const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
  const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
  uint32_t rw[4] = {0, 0, 0, 0};
  uint32_t fw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t w2p = 0, w2r = 0;     // 2 bits per cell: bit 0 = REC_UNK, bit 1 = REC_UNK_PATH / REC_UNK_REACH (the count evaluators' planes)
  const bool row_in_bv = (y >= bv_ky_lo && y <= bv_ky_hi);
  const bool row_reach = use_reach && (y >= reach_y - 1 && y <= reach_y + 1);
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int x = xc + b;
    const uint32_t st = (sw[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
    const uint32_t pm = (pw[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
    const uint32_t lb = (lw[b >> 1] >> ((b & 1) * 16)) & 0xFFFFu;
    const bool inside = x < W;
    const bool inbv = inside && row_in_bv && x >= bv_kx_lo && x <= bv_kx_hi;
    const bool r3 = inside && row_reach && (x >= reach_x - 1 && x <= reach_x + 1);
    const bool unk = inside && st == SIPP_STATE_UNKNOWN;
    uint32_t r = 0;
    if (inbv && unk) r |= REC_UNK;
    if (inbv && unk && pm) r |= REC_UNK_PATH;
    if (inbv && unk && r3) r |= REC_UNK_REACH;
    if (inbv && st == SIPP_STATE_FREE) r |= REC_OBS_FREE;
    if (unk) r |= REC_RAW_UNK;
    if (inside && st == SIPP_STATE_FREE) r |= REC_RAW_FREE;
    if (inside && pm) r |= REC_PATH_RAW;
    if (r3) r |= REC_REACH_RAW;
    rw[b >> 2] |= r << ((b & 3) * 8);
    w2p |= ((r & REC_UNK) | (r & REC_UNK_PATH)) << (2 * b);                 // REC_UNK = 1, REC_UNK_PATH = 2: already the two bits
    w2r |= ((r & REC_UNK) | ((r & REC_UNK_REACH) >> 1)) << (2 * b);
    if (inbv && st == SIPP_STATE_FREE) fw[b >> 1] |= lb << ((b & 1) * 16);
  }
  const size_t o2 = (size_t)y * (pitch / 4) + (size_t)(xc / 4);
  *reinterpret_cast<uint32_t*>(rec2_path + o2) = w2p;
  *reinterpret_cast<uint32_t*>(rec2_reach + o2) = w2r;
  *reinterpret_cast<uint4*>(rec + (size_t)y * pitch + xc) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
  *reinterpret_cast<uint4*>(freelab + (size_t)y * pitch16 + xc) = make_uint4(fw[0], fw[1], fw[2], fw[3]);
  *reinterpret_cast<uint4*>(freelab + (size_t)y * pitch16 + xc + 8) = make_uint4(fw[4], fw[5], fw[6], fw[7]);
}

// debug read-back into compact W*H arrays
__global__ void k_download(const uint8_t* __restrict__ state, const uint16_t* __restrict__ label,
                           const uint8_t* __restrict__ path, size_t pitch, size_t pitch16, int W, int H,
                           uint8_t* __restrict__ o_state, int32_t* __restrict__ o_cost, uint8_t* __restrict__ o_reach,
                           uint8_t* __restrict__ o_path, int reach_x, int reach_y, int use_reach) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)W * H) return;
  const int y = (int)(t / W), x = (int)(t % W);
  if (o_state) o_state[t] = state[(size_t)y * pitch + x];
  if (o_cost) o_cost[t] = label[(size_t)y * pitch16 + x];
  if (o_reach) o_reach[t] = (use_reach && x == reach_x && y == reach_y) ? 1 : 0;   // only the start cell is MSR_REACHABLE
  if (o_path) o_path[t] = path[(size_t)y * pitch + x];
}

IdList make_ids(const int32_t* ids, int n) {
  IdList l;
  l.n = n;
  for (int i = 0; i < SIPP_MAX_IDS; ++i) l.ids[i] = i < n ? ids[i] : -1;
  return l;
}

// one map per launch (arguments by value) / batched (gridDim.y = maps, argument blocks in device memory: sipp_batch_*)
__global__ void k_ingest_patch(const IngestArgs a) { ingest_patch_body(a); }
__global__ void k_build_views(const BuildViewsArgs a) { build_views_body(a); }
__global__ void k_ingest_patch_b(const IngestArgs* __restrict__ as) {
  const IngestArgs a = as[blockIdx.y];
  ingest_patch_body(a);
}
__global__ void k_build_views_b(const BuildViewsArgs* __restrict__ as) {
  if (!as[blockIdx.y].on) return;
  const BuildViewsArgs a = as[blockIdx.y];
  build_views_body(a);
}

IngestArgs make_ingest_args(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* d_labels, const uint8_t* d_observed) {
  IngestArgs a;
  a.state = ctx->d_state; a.label = ctx->d_label; a.pcode = ctx->d_pcode;
  a.pitch = ctx->pitch; a.pitch16 = ctx->pitch16;
  a.W = ctx->W; a.H = ctx->H; a.x0 = x0; a.y0 = y0; a.rows = rows; a.cols = cols;
  a.labels = d_labels; a.observed = d_observed;
  a.mav = make_ids(ctx->desc.obstacle_ids_mav, ctx->desc.n_obstacle_ids_mav);
  a.rov = make_ids(ctx->desc.obstacle_ids_rov, ctx->desc.n_obstacle_ids_rov);
  a.mark = ctx->warm_valid ? ctx->d_mark : nullptr;
  a.init_cost_ceil = (int)floor(ctx->warm_init_cost);
  return a;
}
BuildViewsArgs make_build_views_args(sipp_ctx* ctx, const GainGeom& g, bool use_bv, int on) {
  BuildViewsArgs a;
  a.state = ctx->d_state; a.path = ctx->d_path; a.label = ctx->d_label; a.rec = ctx->d_rec; a.freelab = ctx->d_freelab;
  a.pitch = ctx->pitch; a.pitch16 = ctx->pitch16; a.W = ctx->W; a.H = ctx->H;
  a.bv_kx_lo = use_bv ? g.bv_kx_lo : 0; a.bv_kx_hi = use_bv ? g.bv_kx_hi : ctx->W - 1;
  a.bv_ky_lo = use_bv ? g.bv_ky_lo : 0; a.bv_ky_hi = use_bv ? g.bv_ky_hi : ctx->H - 1;
  a.reach_x = ctx->start_loc[0]; a.reach_y = ctx->start_loc[1]; a.use_reach = ctx->desc.compute_reachability;
  a.rec2_path = ctx->d_rec2[0]; a.rec2_reach = ctx->d_rec2[1];
  a.on = on;
  return a;
}

}  // namespace

size_t ingest_args_bytes() { return sizeof(IngestArgs); }
size_t build_views_args_bytes() { return sizeof(BuildViewsArgs); }
void fill_ingest_args(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* d_labels, void* dst) {
  const IngestArgs a = make_ingest_args(ctx, x0, y0, rows, cols, d_labels, nullptr);
  memcpy(dst, &a, sizeof(a));
}
void fill_build_views_args(sipp_ctx* ctx, const GainGeom& g, bool use_bv, int on, void* dst) {
  const BuildViewsArgs a = make_build_views_args(ctx, g, use_bv, on);
  memcpy(dst, &a, sizeof(a));
}
cudaError_t launch_ingest_patch_batch(const void* d_args, int m, int rows, int cols, cudaStream_t s) {
  const long long n = (long long)rows * cols;
  if (n <= 0 || m <= 0) return cudaSuccess;
  k_ingest_patch_b<<<dim3((unsigned)((n + 255) / 256), (unsigned)m), 256, 0, s>>>(reinterpret_cast<const IngestArgs*>(d_args));
  return cudaGetLastError();
}
cudaError_t launch_build_views_batch(const void* d_args, int m, size_t pitch, int H, cudaStream_t s) {
  const long long n = (long long)(pitch / 16) * H;
  if (n <= 0 || m <= 0) return cudaSuccess;
  k_build_views_b<<<dim3((unsigned)((n + 255) / 256), (unsigned)m), 256, 0, s>>>(reinterpret_cast<const BuildViewsArgs*>(d_args));
  return cudaGetLastError();
}

cudaError_t launch_clear_planes(sipp_ctx* ctx, cudaStream_t s) {
  // MapServerPlanExp::clearMap (map_server_planexp.cpp:461-487) + PRMStarPlanner::set_derived_params (:24-27)
  cudaError_t e;
  const size_t n8 = ctx->pitch * ctx->H, n16 = ctx->pitch16 * ctx->H;
  if ((e = cudaMemsetAsync(ctx->d_state, SIPP_STATE_UNKNOWN, n8, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_label, 0, n16 * 2, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_pcode, 0xFF, n16 * 2, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_path, 0, n8, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_mark, 0, n8, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_rec, 0, n8, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_rec2[0], 0, n8 / 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_rec2[1], 0, n8 / 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_freelab, 0, n16 * 2, s)) != cudaSuccess) return e;
  ctx->launches += 6;
  return cudaSuccess;
}

cudaError_t launch_ingest_patch(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* d_labels,
                                const uint8_t* d_observed, cudaStream_t s) {
  const long long n = (long long)rows * cols;
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  k_ingest_patch<<<blocks, threads, 0, s>>>(make_ingest_args(ctx, x0, y0, rows, cols, d_labels, d_observed));
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_build_views(sipp_ctx* ctx, const GainGeom& g, bool use_bv, cudaStream_t s) {
  const long long n = (long long)(ctx->pitch / 16) * ctx->H;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  k_build_views<<<blocks, threads, 0, s>>>(make_build_views_args(ctx, g, use_bv, 1));
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_closest_update(sipp_ctx* ctx, int x0, int y0, int rows, int cols, cudaStream_t s) {
  const long long n = (long long)(cols + 31) * (rows + 21);
  const int threads = 256;
  k_closest_update<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(ctx->d_state, ctx->d_label, ctx->d_cclab, ctx->d_closest_dist, ctx->pitch,
                                                                             ctx->pitch16, ctx->W, ctx->H, x0, y0, x0 + cols, y0 + rows);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_closest_clear(sipp_ctx* ctx, cudaStream_t s) {
  if (!ctx->d_cclab) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(ctx->d_cclab, 0, ctx->pitch16 * ctx->H * 2, s);
  if (e != cudaSuccess) return e;
  // sentinel for the initial distance sqrt(width^2 + height^2): qmax + 0.5 with qmax = max{q : sqrt(q) < that}
  const double d0 = std::sqrt(ctx->desc.width * ctx->desc.width + ctx->desc.height * ctx->desc.height);
  long long q = (long long)std::floor(d0 * d0);
  while (q >= 0 && !(std::sqrt((double)q) < d0)) --q;
  while (std::sqrt((double)(q + 1)) < d0) ++q;
  const float sentinel = q > 8000000 ? 8.0e6f : (float)q + 0.5f;     // squared window distances never exceed 15^2 + 10^2
  k_fill_f32<<<592, 256, 0, s>>>(ctx->d_closest_dist, ctx->pitch16 * ctx->H, sentinel);
  ctx->launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_build_lowcost_view(sipp_ctx* ctx, const GainGeom& g, bool use_bv, cudaStream_t s) {
  const long long n = (long long)ctx->pitch16 * ctx->H;
  const int threads = 256;
  const int klo_x = use_bv ? g.bv_kx_lo : 0, khi_x = use_bv ? g.bv_kx_hi : ctx->W - 1;
  const int klo_y = use_bv ? g.bv_ky_lo : 0, khi_y = use_bv ? g.bv_ky_hi : ctx->H - 1;
  k_build_lowcost_view<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(ctx->d_state, ctx->d_cclab, ctx->d_closest, ctx->pitch, ctx->pitch16,
                                                                                 ctx->W, ctx->H, klo_x, khi_x, klo_y, khi_y);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_download_planes(sipp_ctx* ctx, uint8_t* o_state, int32_t* o_cost, uint8_t* o_reach, uint8_t* o_path,
                                   cudaStream_t s) {
  const long long n = (long long)ctx->W * ctx->H;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  k_download<<<blocks, threads, 0, s>>>(ctx->d_state, ctx->d_label, ctx->d_path, ctx->pitch, ctx->pitch16, ctx->W, ctx->H,
                                        o_state, o_cost, o_reach, o_path, ctx->start_loc[0], ctx->start_loc[1],
                                        ctx->desc.compute_reachability);
  ctx->launches += 1;
  return cudaGetLastError();
}
