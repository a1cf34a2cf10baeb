// C++ host driving several contexts of ONE process, one std::thread per rank (rank r on device r % devices): the shape of a
// single-process multi-GPU planner host. Every thread keeps the replicated map, scores its shard of the candidates with
// sipp_cycle_shard (host pointers in, host results + the GLOBAL winner out) and the ranks meet only inside the kernel, through
// their peer-memory mailboxes (same-process attach: plain peer pointers, cudaDeviceEnablePeerAccess between devices).
// Checked against the oracle scanning all candidates on the CPU (linked as the CHECKER only).
//
//   test_shard_threads <ranks> <devices>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../../include/sipp.h"

extern "C" {
struct orc_ctx;
orc_ctx* orc_create(const sipp_map_desc* d);
void orc_destroy(orc_ctx* c);
int orc_map_upload_full(orc_ctx* c, const int32_t* labels, const uint8_t* observed);
int orc_plan(orc_ctx* c, double* cost, int32_t* status, double* path_xy, int32_t path_cap, int32_t* path_len, int update_mask);
int orc_cycle(orc_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy, const double* cost, double* gain,
              uint8_t* terminal, double* value, int32_t* best_index, double* best_value, int variant, int threads);
}

namespace {
struct Barrier {   // std::barrier is C++20; a generation-counting one is enough here
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const int gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen != gen_; });
  }
  std::mutex m_; std::condition_variable cv_; int n_, count_ = 0, gen_ = 0;
};

std::vector<int32_t> makeLabels(int W, int H, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<int32_t> L((size_t)W * H, 30);
  const int labels[5] = {45, 60, 90, 120, 0};
  for (int r = 0; r < 24; ++r) {
    const int lab = labels[r % 5];
    const int w = (lab ? W / 10 : W / 24) + (int)(rng() % (W / 8)), h = (lab ? H / 10 : H / 24) + (int)(rng() % (H / 8));
    const int x = (int)(rng() % W), y = (int)(rng() % H);
    for (int j = y; j < std::min(H, y + h); ++j)
      for (int i = x; i < std::min(W, x + w); ++i) L[(size_t)j * W + i] = lab;
  }
  for (int j = 0; j < 24; ++j)
    for (int i = 0; i < 24; ++i) { L[(size_t)j * W + i] = 30; L[(size_t)(H - 1 - j) * W + (W - 1 - i)] = 30; }
  return L;
}
}  // namespace

int main(int argc, char** argv) {
  const int ranks = argc > 1 ? atoi(argv[1]) : 2, devices = argc > 2 ? atoi(argv[2]) : 1;
  if (ranks < 1 || ranks > SIPP_PEER_MAX_WORLD || devices < 1) { std::fprintf(stderr, "usage: %s <ranks 1..16> <devices>\n", argv[0]); return 2; }
  if (ranks > devices) {
    // The host-pointer calls block (staged copies from pageable memory, stream synchronisation); two ranks that share a device
    // share its driver context, and the rank that blocks while its kernel waits for the peer keeps the peer's thread from
    // launching. One rank per device is the supported shape for these calls; contexts that share a device use the asynchronous
    // _dev entry points from one thread (tests/test_gpu_peer.py).
    std::fprintf(stderr, "%d ranks need %d devices (one rank per device for the blocking host-pointer calls)\n", ranks, ranks);
    return 2;
  }
  const int W = 320, H = 240, T = 6000, steps = 4;
  sipp_map_desc d;
  std::memset(&d, 0, sizeof(d));
  d.W = W; d.H = H; d.width = W * 0.0625; d.height = H * 0.0625;
  d.start_x = d.start_y = 0.5; d.goal_x = d.width - 0.5; d.goal_y = d.height - 0.5;
  d.min_rov_cost = 30; d.max_rov_cost = 120; d.unknown_init_cost = 120; d.unknown_id = 3599;
  d.n_obstacle_ids_mav = 1; d.obstacle_ids_mav[0] = 0; d.n_obstacle_ids_rov = 1; d.obstacle_ids_rov[0] = 0;
  d.compute_reachability = 1;
  const std::vector<int32_t> L = makeLabels(W, H, 5);
  std::vector<uint8_t> observed((size_t)W * H, 0);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) observed[(size_t)y * W + x] = (x < 200 && y < 160) ? 1 : 0;
  sipp_eval_params P;
  std::memset(&P, 0, sizeof(P));
  P.evaluator = SIPP_EVAL_RRT_STAR; P.half_fov = 32; P.non_path_voxel_gain = 0.001; P.discount_offset = 0.1; P.max_gain = 100.0; P.discount_factor = 0.1;

  // candidates of every step (all ranks see the same arrays; rank r owns [lo_r, hi_r))
  std::mt19937_64 rng(99);
  std::uniform_real_distribution<double> ux(0.0, d.width), uy(0.0, d.height), uc(0.2, 5.0);
  std::vector<std::vector<double>> xy(steps, std::vector<double>(2 * (size_t)T)), cost(steps, std::vector<double>(T));
  for (int s = 0; s < steps; ++s)
    for (int i = 0; i < T; ++i) { xy[s][2 * i] = ux(rng); xy[s][2 * i + 1] = uy(rng); cost[s][i] = uc(rng); }
  for (int i = 0; i < 2; ++i) { xy[1][2 * (T - 1) + i] = xy[1][2 * 3 + i]; }   // an exact cross-shard tie candidate pair in step 1
  cost[1][T - 1] = cost[1][3];

  // the oracle's answer: one scan over all candidates
  orc_ctx* o = orc_create(&d);
  orc_map_upload_full(o, L.data(), observed.data());
  double oc; int32_t ost, oln;
  orc_plan(o, &oc, &ost, nullptr, 0, &oln, 1);
  std::vector<int32_t> want_i(steps);
  std::vector<double> want_v(steps);
  {
    std::vector<double> g(T), v(T);
    std::vector<uint8_t> t(T);
    for (int s = 0; s < steps; ++s) orc_cycle(o, &P, T, xy[s].data(), nullptr, cost[s].data(), g.data(), t.data(), v.data(), &want_i[s], &want_v[s], 1, 4);
  }
  orc_destroy(o);

  std::vector<unsigned char> handles((size_t)ranks * SIPP_PEER_HANDLE_BYTES);
  Barrier bar(ranks);
  std::atomic<int> failures{0};
  std::vector<std::thread> th;
  for (int r = 0; r < ranks; ++r)
    th.emplace_back([&, r] {
      auto fail = [&](const char* what, sipp_ctx* c) { std::fprintf(stderr, "rank %d: %s: %s\n", r, what, sipp_last_error(c)); ++failures; };
      sipp_ctx* c = nullptr;
      bool ok = sipp_create(&d, r % devices, &c) == SIPP_OK;
      if (!ok) fail("sipp_create", nullptr);
      double pc; int32_t st, ln;
      if (ok && (sipp_map_upload_full(c, L.data(), observed.data()) != SIPP_OK || sipp_plan(c, &pc, &st, nullptr, 0, &ln) != SIPP_OK)) { fail("map / plan", c); ok = false; }
      if (ok && (pc != oc || ln != oln)) { std::fprintf(stderr, "rank %d: plan %.17g (%d nodes) vs oracle %.17g (%d)\n", r, pc, ln, oc, oln); ++failures; }
      if (ok && sipp_peer_export(c, handles.data() + (size_t)r * SIPP_PEER_HANDLE_BYTES) != SIPP_OK) { fail("peer_export", c); ok = false; }
      bar.wait();                                   // every handle is in the table
      if (ok && sipp_peer_attach(c, r, ranks, handles.data()) != SIPP_OK) { fail("peer_attach", c); ok = false; }
      bar.wait();
      const int base = T / ranks, rem = T % ranks;
      const int lo = r * base + std::min(r, rem), n = base + (r < rem ? 1 : 0);
      std::vector<double> gain(n), value(n);
      std::vector<uint8_t> term(n);
      for (int s = 0; s < steps && failures == 0; ++s) {
        int64_t bi = -9;
        double bv = 0;
        const int rc = sipp_cycle_shard(c, &P, n, xy[s].data() + 2 * (size_t)lo, nullptr, cost[s].data() + lo, lo, gain.data(), term.data(), value.data(), &bi, &bv);
        if (rc != SIPP_OK) { fail("sipp_cycle_shard", c); break; }
        const bool vok = bv == want_v[s] || std::fabs(bv - want_v[s]) <= 1e-5 * std::fabs(want_v[s]);
        if (bi != want_i[s] || !vok) { std::fprintf(stderr, "rank %d step %d: winner %lld (%.17g) vs oracle %d (%.17g)\n", r, s, (long long)bi, bv, want_i[s], want_v[s]); ++failures; }
      }
      bar.wait();                                   // nobody frees a mailbox a peer may still write into
      if (c) { sipp_peer_detach(c); sipp_destroy(c); }
    });
  for (auto& t : th) t.join();
  std::printf("shard_threads: %d rank(s) on %d device(s), %d step(s): %d failure(s)\n", ranks, devices, steps, failures.load());
  return failures ? 1 : 0;
}
