// Block-resident relaxation of the follower cost-to-go field (row A11; included by sipp_field.cu inside its anonymous namespace).
//
// Same fixed point as k_relax — d[v] = min_u d[u] + w(u,v), w = dist * (c_a + c_b) / 2 (prm_star_planner.cpp:178-181), which in IEEE
// double does not depend on the relaxation order — but organised for the B200's 227 KB of shared memory per SM:
//
//   * unit of global scheduling = a BLOCK of 128 x 128 cells. One CTA (16 warps) takes a block from the global ring queue and keeps
//     its whole window in shared memory: 130 x 130 doubles of d (one-cell halo) + 130 x 130 fp32 half costs = 203 KB. Edge weights are
//     not cached anywhere: w = dist * (hc_a + hc_b) is evaluated in the sweep from the half costs (the edge lengths come from the
//     per-column / per-row class tables of the noisy lattice), so there is no 37 KB weight block per 32 x 32 tile to fetch and no
//     k_weights pass;
//   * inside the block every warp OWNS one 32 x 32 sub-tile (4 x 4 of them). A warp runs Gauss-Seidel row sweeps over its sub-tile
//     (lane = column, previous row carried in registers, in-row shuffle fixed point — the sweep of k_relax) directly on the shared
//     array: the halo of a sub-tile is simply the neighbouring warp's cells. A warp that lowers a cell on its border ring sets the
//     neighbouring sub-tile's dirty flag (shared memory) — a hand-off between sub-tiles costs a flag instead of the publish / queue /
//     pop / load round trip through L2 that a hop between k_relax tiles costs (8 of its 19 us). The sub-tiles run in rounds separated
//     by a block barrier at which the block votes on "any flag set?";
//   * when no sub-tile is dirty, the CTA publishes: the cells it lowered go to global memory with plain stores (a block is run by at
//     most one CTA at a time, nobody else writes its cells), and a neighbouring block is queued only if relaxing across the border
//     would lower the copy of its ring we hold (no echo activations). A block whose halo was lowered while it ran goes back to the
//     queue.
//
// STATUS: bit-identical to k_relax (tools/relax_ab.py, tests/test_gpu_parity.py::test_block_resident_relaxation_is_bit_identical) but
// SLOWER on B200 — 14.0 ms against 8.0 ms at 4096^2, 62.7 against 31.1 ms at 8192^2, 2.8 against 1.35 ms at 640 x 480 (r2_j12 logs):
// 8.6 runs per block and 3.6 x the sub-tile sweeps of k_relax, because a block run restarts its sub-tiles for every small improvement
// that crosses it and only 148 x 16 warps are resident. It stays selectable (SIPP_RELAX_KERNEL=1) as the A/B partner of k_relax; the
// default is k_relax.
//
// Exactness: every value ever stored is the rounded length of a real path (an upper bound of the fixed point), values only decrease,
// and the kernel ends when no block is queued or running, i.e. at the fixed point — the argument of k_relax, unchanged.
// Concurrent readers of a shared-memory cell see an old or a new value of an aligned 8-byte word, both upper bounds.

constexpr int BS = 128;            // block edge (cells)
constexpr int BPD = BS + 2;        // with the one-cell halo
constexpr int kBlkSub = BS / TS;   // 4 sub-tiles per block edge
constexpr int kBlkWarps = kBlkSub * kBlkSub;
constexpr int kBlkThreads = kBlkWarps * 32;

struct BlkSmem {
  double d[BPD][BPD];        // tentative costs incl. halo
  float hc[BPD][BPD];        // half costs (multiples of 0.5 or +inf: exact in fp32); NaN = unexplored cell whose init cost is not
  double hd[BPD], vd[BPD];   // length of the edge between block columns c, c+1 / rows r, r+1 (1.0 where one end is off the map:
                             // the off-map cell's half cost is +inf, so the weight is +inf anyway)
  double dd[kMaxDd];         // diagonal length per class pair (+inf: not admitted, prm_star_planner.cpp:138)
  int xc[BPD], yc[BPD];      // class of those edges, premultiplied: xc[c] = class * Ky
  int dirty[kBlkWarps];      // sub-tile has unseen input
  int blk;                   // block being run (-1: the queue is drained)
  int ring_m;                // neighbour blocks (N S W E NW NE SW SE) whose side of our ring was lowered in this run
  int notify_m;              // neighbour blocks that have to look at it
  int vote[3];               // "another round" votes, slot = round % 3
};
static_assert(sizeof(BlkSmem) <= 227 * 1024, "block window must fit the opt-in shared memory of one SM");

struct BlkArgs {
  FieldGeom g;
  double* field;            // [H][W]
  const uint16_t* pcode;    // [H][pitch16]
  size_t pitch16;
  int nbx, nby;             // blocks per row / column
  int ntx, nty;             // 32 x 32 tiles per row / column (touched[] granularity)
  int* flags;               // [nblocks] ST_QUEUED / ST_RUN
  int* queue;               // [cap] ring of block ids, -1 = empty slot
  int cap;
  int* ctl;                 // CTL_* of k_relax
  int spin_limit;
  int init_inexact;         // init_cost / 2 is not representable in fp32: unexplored cells are patched in the sweep
  uint8_t* touched;         // optional [ntiles]: set for every 32 x 32 tile in which a cell was lowered
};

__device__ __forceinline__ double ldv(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
__device__ __forceinline__ void stv(double* p, double v) { *reinterpret_cast<volatile double*>(p) = v; }
__device__ __forceinline__ int ldvi(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

template <bool PATCH>
__device__ __forceinline__ double hc_at(const BlkSmem& s, int r, int c, double init_half) {
  const float f = s.hc[r][c];
  if (PATCH && f != f) return init_half;
  return (double)f;
}

// weight of the edge between two neighbouring cells given in block coordinates (publish phase; the sweep has its own copy)
template <bool PATCH>
__device__ __forceinline__ double blk_edge_w(const BlkSmem& s, const double* ddp, int ra, int ca, int rb, int cb, double init_half) {
  const int ey = min(ra, rb), ex = min(ca, cb);
  double dist;
  if (ra == rb) dist = s.hd[ex];
  else if (ca == cb) dist = s.vd[ey];
  else dist = ddp[s.xc[ex] + s.yc[ey]];
  return dist < CUDART_INF ? __dmul_rn(dist, __dadd_rn(hc_at<PATCH>(s, ra, ca, init_half), hc_at<PATCH>(s, rb, cb, init_half))) : CUDART_INF;
}

// One sweep of the warp's sub-tile whose first cell is (r0, c0) in block coordinates. Row r is relaxed against the three cells of the
// previous row (carried in registers from that row's last shuffle iteration), then against itself with a shuffle loop that runs
// until no lane changes. fmin() drops a NaN candidate (an absent diagonal of length +inf between two zero-cost cells gives inf * 0).
template <int DIR, bool PATCH>
__device__ __forceinline__ bool blk_sweep(BlkSmem& s, const double* __restrict__ ddp, const double init_half, const int r0, const int c0,
                                          const int lane, uint32_t& rows_changed) {
  const int c = c0 + lane;
  const double hdl = s.hd[c - 1], hdr = s.hd[c];
  const int xl = s.xc[c - 1], xr = s.xc[c];
  const int pfirst = DIR > 0 ? r0 - 1 : r0 + TS;
  double pl = ldv(&s.d[pfirst][c - 1]), pc = ldv(&s.d[pfirst][c]), pr = ldv(&s.d[pfirst][c + 1]);
  double hpl = hc_at<PATCH>(s, pfirst, c - 1, init_half), hpc = hc_at<PATCH>(s, pfirst, c, init_half), hpr = hc_at<PATCH>(s, pfirst, c + 1, init_half);
  bool chg = false;
#pragma unroll 2
  for (int i = 0; i < TS; ++i) {
    const int r = DIR > 0 ? r0 + i : r0 + TS - 1 - i;
    const int ey = DIR > 0 ? r - 1 : r;                      // the vertical edges of this step join rows ey, ey + 1
    const double hrl = hc_at<PATCH>(s, r, c - 1, init_half), hrc = hc_at<PATCH>(s, r, c, init_half), hrr = hc_at<PATCH>(s, r, c + 1, init_half);
    const double vdv = s.vd[ey];
    const int ycv = s.yc[ey];
    const double ddl = ddp[xl + ycv], ddr = ddp[xr + ycv];
    const double wn = __dmul_rn(vdv, __dadd_rn(hpc, hrc));
    const double wnl = __dmul_rn(ddl, __dadd_rn(hpl, hrc));
    const double wnr = __dmul_rn(ddr, __dadd_rn(hpr, hrc));
    const double wl = __dmul_rn(hdl, __dadd_rn(hrl, hrc));
    const double wr = __dmul_rn(hdr, __dadd_rn(hrc, hrr));
    const double hl = ldv(&s.d[r][c0 - 1]), hr = ldv(&s.d[r][c0 + TS]);     // the neighbours' cells (or the block halo), as they are now
    double d = ldv(&s.d[r][c]);
    const double d0 = d;
    d = fmin(d, fmin(__dadd_rn(pc, wn), fmin(__dadd_rn(pl, wnl), __dadd_rn(pr, wnr))));
    double dl, dr;
    bool again;
    do {
      dl = __shfl_up_sync(0xFFFFFFFFu, d, 1); dr = __shfl_down_sync(0xFFFFFFFFu, d, 1);
      if (lane == 0) dl = hl;
      if (lane == 31) dr = hr;
      const double cnd = fmin(__dadd_rn(dl, wl), __dadd_rn(dr, wr));
      again = cnd < d;
      if (again) d = cnd;
    } while (__any_sync(0xFFFFFFFFu, again));
    pl = dl; pc = d; pr = dr;
    hpl = hrl; hpc = hrc; hpr = hrr;
    if (d < d0) {
      chg = true;
      rows_changed |= 1u << (r - r0);
      stv(&s.d[r][c], d);
    }
  }
  __syncwarp();
  return __any_sync(0xFFFFFFFFu, chg);
}

template <bool PATCH>
__global__ void __launch_bounds__(kBlkThreads, 1)
k_relax_blk(const BlkArgs a) {
  extern __shared__ __align__(16) uint8_t blk_smem_raw[];
  BlkSmem& s = *reinterpret_cast<BlkSmem*>(blk_smem_raw);
  const FieldGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sx = warp & (kBlkSub - 1), sy = warp / kBlkSub;
  const int r0 = 1 + sy * TS, c0 = 1 + sx * TS;            // this warp's sub-tile in block coordinates
  const double init_half = g.init_half_cost;
  const bool dd_in_smem = g.Kx * g.Ky <= kMaxDd;
  if (dd_in_smem)
    for (int i = tid; i < g.Kx * g.Ky; i += kBlkThreads) s.dd[i] = g.ddist[i];
  const double* ddp = dd_in_smem ? s.dd : g.ddist;
  int sweeps_local = 0, act_local = 0, push_local = 0;

  for (;;) {
    // ---- pop a block: one thread takes a ticket and waits for its slot (or for global termination)
    if (tid == 0) {
      int bi = -2;
      const unsigned ticket = (unsigned)atomicAdd(&a.ctl[CTL_HEAD], 1);
      int* slot = a.queue + (ticket % (unsigned)a.cap);
      int spins = 0;
      for (;;) {
        if (ld_volatile_i32(&a.ctl[CTL_ABORT]) != 0) break;       // an aborted plan: do not start another block
        const int v = ld_volatile_i32(slot);
        if (v >= 0) {
          st_volatile_i32(slot, -1);
          bi = v;
          break;
        }
        if (ld_volatile_i32(&a.ctl[CTL_PENDING]) == 0) break;
        if (++spins > a.spin_limit) {
          atomicExch(&a.ctl[CTL_ABORT], 1);
          break;
        }
        __nanosleep(64);
      }
      if (bi >= 0) {
        atomicExch(&a.flags[bi], ST_RUN);      // queued -> running: whoever lowers our halo from now on sets QUEUED, and we re-run
        fence_gpu();
      }
      s.blk = bi;
    }
    __syncthreads();
    const int bi = s.blk;
    if (bi < 0) break;
    const int bx = bi % a.nbx, by = bi / a.nbx;
    const int gx0 = bx * BS - 1, gy0 = by * BS - 1;          // global coordinates of block cell (0, 0) (the halo corner)

    // ---- load the window: d (L2-coherent loads: other CTAs publish while we run) and the half costs; loads are issued in batches
    // of 11 per thread before the first shared-memory store so that their latencies overlap
    constexpr int kLoadRounds = (BPD * BPD + kBlkThreads - 1) / kBlkThreads;      // 34
    constexpr int kLoadBatch = 17;
    static_assert(kLoadRounds % kLoadBatch == 0, "load batches must tile the window");
#pragma unroll 1
    for (int b0 = 0; b0 < kLoadRounds; b0 += kLoadBatch) {
      double dv[kLoadBatch];
      uint32_t cv[kLoadBatch];
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) {
        const int idx = tid + (b0 + j) * kBlkThreads;
        const int r = idx / BPD, c = idx - r * BPD;
        const int gx = gx0 + c, gy = gy0 + r;
        const bool inb = idx < BPD * BPD && gx >= 0 && gx < g.W && gy >= 0 && gy < g.H;
        dv[j] = inb ? __ldcg(a.field + (size_t)gy * g.W + gx) : CUDART_INF;
        cv[j] = inb ? (uint32_t)__ldg(a.pcode + (size_t)gy * a.pitch16 + gx) : (uint32_t)PCODE_REMOVED;
      }
#pragma unroll
      for (int j = 0; j < kLoadBatch; ++j) {
        const int idx = tid + (b0 + j) * kBlkThreads;
        if (idx < BPD * BPD) {
          const int r = idx / BPD, c = idx - r * BPD;
          float h;
          if (cv[j] == PCODE_UNEXPLORED) h = PATCH ? __int_as_float(0x7FC00000) : (float)init_half;
          else if (cv[j] == PCODE_REMOVED) h = __int_as_float(0x7F800000);
          else h = 0.5f * (float)cv[j];
          s.d[r][c] = dv[j];
          s.hc[r][c] = h;
        }
      }
    }
    if (tid < BPD) {
      const int gy = gy0 + tid, gx = gx0 + tid;               // edges (gy, gy + 1) and (gx, gx + 1)
      const bool oky = gy >= 0 && gy + 1 < g.H, okx = gx >= 0 && gx + 1 < g.W;
      const int cy = oky ? g.ycls[gy] : 0, cx = okx ? g.xcls[gx] : 0;
      s.yc[tid] = cy;
      s.vd[tid] = oky ? g.vdist[cy] : 1.0;
      s.xc[tid] = cx * g.Ky;
      s.hd[tid] = okx ? g.hdist[cx] : 1.0;
    }
    if (tid < kBlkWarps) {
      const int tx_ = tid & (kBlkSub - 1), ty_ = tid / kBlkSub;
      s.dirty[tid] = (bx * BS + tx_ * TS < g.W && by * BS + ty_ * TS < g.H) ? 1 : 0;     // sub-tiles off the map have nothing to do
    }
    __syncthreads();
    if (tid == 0) {
      s.ring_m = 0;
      s.notify_m = 0;
      s.vote[0] = s.vote[1] = s.vote[2] = 0;
    }
    __syncthreads();

    uint32_t acc = 0;            // rows of this lane's column lowered since the last publish
    {                            // one run of this block
      ++act_local;
      // ---- the sub-tiles relax in ROUNDS until none is dirty. In a round every warp whose flag is set takes it and runs its sub-tile to
      // the local fixed point, reading its neighbours' cells as they are and flagging the neighbours whose ring it lowered; then the
      // block votes at a barrier: each warp arrives only once it has stopped, and a flag is always set before its setter arrives, so the
      // flags the last arriving warp reads are final — if any is set there is another round. (An earlier form, a shared counter of dirty +
      // running sub-tiles polled by the idle warps, lost count on the GPU and was dropped; waiting warps now sit in the barrier instead
      // of competing for issue slots.)
      for (int round = 0;; ++round) {
        int mine = 0;
        if (lane == 0) mine = ldvi(&s.dirty[warp]) ? atomicExch(&s.dirty[warp], 0) : 0;
        mine = __shfl_sync(0xFFFFFFFFu, mine, 0);
        if (mine) {
          __threadfence_block();
          for (int it = 0;; ++it) {
            uint32_t rows_changed = 0;
            const bool ch = (it & 1) ? blk_sweep<-1, PATCH>(s, ddp, init_half, r0, c0, lane, rows_changed)
                                     : blk_sweep<+1, PATCH>(s, ddp, init_half, r0, c0, lane, rows_changed);
            ++sweeps_local;
            if (ch) {
              acc |= rows_changed;
              // who reads the cells we lowered: N S W E NW NE SW SE
              const unsigned top = __ballot_sync(0xFFFFFFFFu, (rows_changed & 1u) != 0), bot = __ballot_sync(0xFFFFFFFFu, (rows_changed >> 31) != 0);
              const uint32_t lft = __shfl_sync(0xFFFFFFFFu, rows_changed, 0), rgt = __shfl_sync(0xFFFFFFFFu, rows_changed, 31);
              const int m = (top ? 1 : 0) | (bot ? 2 : 0) | (lft ? 4 : 0) | (rgt ? 8 : 0) | ((lft & 1u) ? 16 : 0) | ((rgt & 1u) ? 32 : 0) |
                            ((lft >> 31) ? 64 : 0) | ((rgt >> 31) ? 128 : 0);
              __threadfence_block();       // the lowered cells before the flags
              if (lane < 8 && ((m >> lane) & 1)) {
                const int nsx = sx + ((0xA8 >> lane) & 1) - ((0x54 >> lane) & 1);      // dx of N S W E NW NE SW SE: 0 0 -1 1 -1 1 -1 1
                const int nsy = sy + ((0xC2 >> lane) & 1) - ((0x31 >> lane) & 1);      // dy:                        -1 1 0 0 -1 -1 1 1
                const int ox = nsx < 0 ? -1 : (nsx >= kBlkSub ? 1 : 0), oy = nsy < 0 ? -1 : (nsy >= kBlkSub ? 1 : 0);
                if (ox == 0 && oy == 0) {
                  atomicExch(&s.dirty[nsy * kBlkSub + nsx], 1);
                } else {
                  // a cell on the block's ring: the neighbouring block in direction (ox, oy) reads it
                  const int nb = oy == 0 ? (ox < 0 ? 2 : 3) : (ox == 0 ? (oy < 0 ? 0 : 1) : ((oy < 0 ? 4 : 6) + (ox < 0 ? 0 : 1)));
                  atomicOr(&s.ring_m, 1 << nb);
                }
              }
              __syncwarp();
            }
            if (!ch && it >= 1) break;
            if (it > 100000) {                 // watchdog: a sub-tile cannot need this many sweeps (values only ever decrease)
              if (lane == 0) { atomicExch(&a.ctl[CTL_ABORT], 2); a.ctl[16] = bi; a.ctl[17] = warp; }
              break;
            }
          }
          __threadfence_block();             // flags and cells before this warp's arrival at the vote
        }
        // the vote (plain barrier + a slot per round: barrier.red raised an illegal-instruction fault on this part). Slot r % 3 collects
        // round r; thread 0 clears the slot of round r + 2 after barrier r, which nobody writes before barrier r + 1.
        const int slot = round % 3;
        const int flagged = (lane < kBlkWarps) ? ldvi(&s.dirty[lane]) : 0;
        if (__any_sync(0xFFFFFFFFu, flagged != 0) && lane == 0) *reinterpret_cast<volatile int*>(&s.vote[slot]) = 1;
        __syncthreads();
        const int more = ldvi(&s.vote[slot]);
        if (tid == 0) *reinterpret_cast<volatile int*>(&s.vote[(round + 2) % 3]) = 0;
        if (!more) break;
        if (round > (1 << 20)) {             // watchdog: the block's sub-tiles never drained
          if (tid == 0) { atomicExch(&a.ctl[CTL_ABORT], 3); a.ctl[18] = bi; }
          break;
        }
      }
      // every flag is clear and nobody is sweeping

      // ---- publish: the cells this warp lowered (the block is ours alone: plain stores)
      {
        const int gx = gx0 + c0 + lane;
        uint32_t rc = acc;
        while (rc) {
          const int r = __ffs(rc) - 1;
          rc &= rc - 1;
          a.field[(size_t)(gy0 + r0 + r) * g.W + gx] = s.d[r0 + r][c0 + lane];
        }
        if (a.touched && __any_sync(0xFFFFFFFFu, acc != 0) && lane == 0) {
          const int ttx = bx * kBlkSub + sx, tty = by * kBlkSub + sy;
          if (ttx < a.ntx && tty < a.nty) a.touched[tty * a.ntx + ttx] = 1;
        }
        acc = 0;
      }
      // ---- which neighbouring blocks have to look: relax ACROSS the border into our copy of their ring (what they would compute
      // from their halo, bit for bit); queue a neighbour only if that beats the value we hold for one of its cells
      {
        const int ring = s.ring_m;
        int m = 0;
        if (ring) {
          const int q = tid >> 7, k = (tid & 127) + 1;          // side N S W E, position 1..128 along it
          if ((ring >> q) & 1) {
            double cand = CUDART_INF, cur;
            if (q < 2) {
              const int hr = q == 0 ? 0 : BPD - 1, ir = q == 0 ? 1 : BS;
              cur = s.d[hr][k];
              for (int dc = -1; dc <= 1; ++dc) {
                const int cc = k + dc;
                if (cc < 1 || cc > BS) continue;
                cand = fmin(cand, __dadd_rn(s.d[ir][cc], blk_edge_w<PATCH>(s, ddp, ir, cc, hr, k, init_half)));
              }
            } else {
              const int hcx = q == 2 ? 0 : BPD - 1, ic = q == 2 ? 1 : BS;
              cur = s.d[k][hcx];
              for (int dr_ = -1; dr_ <= 1; ++dr_) {
                const int rr = k + dr_;
                if (rr < 1 || rr > BS) continue;
                cand = fmin(cand, __dadd_rn(s.d[rr][ic], blk_edge_w<PATCH>(s, ddp, rr, ic, k, hcx, init_half)));
              }
            }
            if (cand < cur) m |= 1 << q;
          }
          if (tid < 4 && ((ring >> (4 + tid)) & 1)) {          // corners NW NE SW SE: one diagonal edge each
            const int hr = (tid & 2) ? BPD - 1 : 0, hcx = (tid & 1) ? BPD - 1 : 0;
            const int ir = (tid & 2) ? BS : 1, ic = (tid & 1) ? BS : 1;
            if (__dadd_rn(s.d[ir][ic], blk_edge_w<PATCH>(s, ddp, ir, ic, hr, hcx, init_half)) < s.d[hr][hcx]) m |= 16 << tid;
          }
          m = __reduce_or_sync(0xFFFFFFFFu, (unsigned)m);
          if (lane == 0 && m) atomicOr(&s.notify_m, m);
        }
      }
      __threadfence();       // our stores are visible at gpu scope before a neighbour can be told to look
      __syncthreads();
      if (warp == 0) {
        const int m = s.notify_m;
        if (lane < 8 && ((m >> lane) & 1)) {
          const int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
          const int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
          const int nbx_ = bx + ddx[lane], nby_ = by + ddy[lane];
          if (nbx_ >= 0 && nbx_ < a.nbx && nby_ >= 0 && nby_ < a.nby) {
            const int nb = nby_ * a.nbx + nbx_;
            if (atomicOr(&a.flags[nb], ST_QUEUED) == 0) {      // idle -> queued: ours to enqueue (a running block re-runs itself)
              atomicAdd(&a.ctl[CTL_PENDING], 1);
              const unsigned pos = (unsigned)atomicAdd(&a.ctl[CTL_TAIL], 1);
              int* slot = a.queue + (pos % (unsigned)a.cap);
              int spins = 0;
              while (ld_volatile_i32(slot) != -1 && ++spins < a.spin_limit) __nanosleep(32);   // never expected: cap >= number of blocks
              st_volatile_i32(slot, nb);
              ++push_local;
            }
          }
        }
        __syncwarp();
        // ---- done with this run: idle again, unless our halo was lowered meanwhile
        if (lane == 0) {
          fence_gpu();
          if (atomicCAS(&a.flags[bi], ST_RUN, 0) != ST_RUN) {
            // somebody lowered our halo while we ran: the block goes back to the queue (still counted in CTL_PENDING) and is loaded
            // afresh by whoever pops it. (Re-running it in place on the window we hold, re-reading only the halo, produced wrong
            // fields whenever two CTAs were involved and was dropped; the reload is 170 KB out of L2.)
            atomicExch(&a.flags[bi], ST_QUEUED);
            const unsigned pos = (unsigned)atomicAdd(&a.ctl[CTL_TAIL], 1);
            int* slot = a.queue + (pos % (unsigned)a.cap);
            int spins = 0;
            while (ld_volatile_i32(slot) != -1 && ++spins < a.spin_limit) __nanosleep(32);
            st_volatile_i32(slot, bi);
            ++push_local;
          } else {
            atomicSub(&a.ctl[CTL_PENDING], 1);     // after our pushes were counted: pending reaches 0 only at the global fixed point
          }
          s.ring_m = 0;
          s.notify_m = 0;
        }
      }
      __syncthreads();
    }
  }
  // statistics (plan_stats: pushes, block runs, sub-tile sweeps)
  if (tid == 0) atomicAdd(&a.ctl[CTL_ACTIVATIONS], act_local);
  if (lane == 0) atomicAdd(&a.ctl[CTL_SWEEPS], sweeps_local);
  if (push_local) atomicAdd(&a.ctl[CTL_PUSHES], push_local);
}

// Incremental re-plan, between k_inval and k_relax_blk: queue <- every block that holds a seed tile or touches one; the marks of the seed
// tiles are cleared for the next plan (all marks live in seed tiles). One thread per 32 x 32 tile; the threads with an id below the
// number of blocks also decide for one block each.
__global__ void k_seed_relax_blk(const uint8_t* tile_seed, int ntx, int nty, int nbx, int nby, int* flags, int* queue, int* ctl, uint8_t* mark,
                                 size_t pitch, int H, int* plan_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntx * nty && tile_seed[t]) {
    const int tx = t % ntx, ty = t / ntx;
    for (int r = 0; r < TS; ++r) {
      const int gy = ty * TS + r;
      if (gy >= H) break;
      uint4* row = reinterpret_cast<uint4*>(mark + (size_t)gy * pitch + (size_t)tx * TS);     // pitch is a multiple of 128: aligned, in the row
      row[0] = make_uint4(0u, 0u, 0u, 0u);
      row[1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (t >= nbx * nby) return;
  const int bx = t % nbx, by = t / nbx;
  bool on = false;
  for (int ty = by * kBlkSub - 1; ty <= by * kBlkSub + kBlkSub && !on; ++ty)
    for (int tx = bx * kBlkSub - 1; tx <= bx * kBlkSub + kBlkSub; ++tx)
      if (tx >= 0 && tx < ntx && ty >= 0 && ty < nty && tile_seed[ty * ntx + tx]) { on = true; break; }
  if (!on) return;
  flags[t] = ST_QUEUED;
  atomicAdd(&ctl[CTL_PENDING], 1);
  queue[atomicAdd(&ctl[CTL_TAIL], 1)] = t;
  atomicAdd(&plan_out[6], 1);
}
