// Multi-GPU exchange of the shard winners through peer memory: host side (mailbox export / attach) and the standalone
// exchange kernel. The fused form is the tail of k_select (sipp_select.cu); the protocol is described in sipp_peer.cuh.
#include <unistd.h>

#include <cstdlib>
#include <cstring>

#include "sipp_peer.cuh"

namespace {

struct HandleBlob {                 // what sipp_peer_export writes (<= SIPP_PEER_HANDLE_BYTES)
  uint32_t magic;
  int32_t device;
  int32_t has_ipc;                  // 0: cudaIpcGetMemHandle failed here — only contexts of the same process can attach
  int32_t pad;
  int64_t pid;
  uint64_t ptr;                     // the mailbox address in the exporting process
  cudaIpcMemHandle_t ipc;           // 64 bytes
};
static_assert(sizeof(HandleBlob) <= SIPP_PEER_HANDLE_BYTES, "handle blob too large");
constexpr uint32_t kMagic = 0x53495050u;   // "SIPP"

__global__ void __launch_bounds__(32) k_peer_exchange(const PeerDev pd, const BestRec* __restrict__ mine, BestRec* out, int mode) {
  PeerCounters cnt = {0, 0};
  if (threadIdx.x == 0) cnt = peer_preload(pd);
  BestRec m;
  m.value = 0.0;
  m.index = -1;
  if (mine) m = *mine;
  peer_exchange_block(pd, m, out, mode, cnt);
}

}  // namespace

cudaError_t launch_peer_exchange(sipp_ctx* ctx, const BestRec* d_mine, BestRec* d_out, int mode, cudaStream_t s) {
  k_peer_exchange<<<1, 32, 0, s>>>(ctx->peer, d_mine, d_out, mode);
  ctx->launches += 1;
  return cudaGetLastError();
}

void peer_release(sipp_ctx* ctx) {
  for (int r = 0; r < SIPP_PEER_MAX_WORLD; ++r)
    if (ctx->peer_mapped[r]) {
      cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
      ctx->peer_mapped[r] = nullptr;
    }
  ctx->peer_world = 0;
}

extern "C" {

int sipp_peer_export(sipp_ctx* ctx, void* handle) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!handle) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "handle == NULL");
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->d_peer_box) {
    // a mailbox of its own allocation: cudaIpcGetMemHandle shares the whole allocation the pointer belongs to
    SIPP_CUDA_CHECK(ctx, cudaMalloc(&ctx->d_peer_box, sizeof(PeerBox)));
  }
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  SIPP_CUDA_CHECK(ctx, cudaMemset(ctx->d_peer_box, 0, sizeof(PeerBox)));
  SIPP_CUDA_CHECK(ctx, cudaDeviceSynchronize());
  HandleBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = kMagic;
  b.device = ctx->device;
  b.pid = (int64_t)getpid();
  b.ptr = (uint64_t)(uintptr_t)ctx->d_peer_box;
  b.has_ipc = cudaIpcGetMemHandle(&b.ipc, ctx->d_peer_box) == cudaSuccess ? 1 : 0;
  cudaGetLastError();
  memset(handle, 0, SIPP_PEER_HANDLE_BYTES);
  memcpy(handle, &b, sizeof(b));
  return SIPP_OK;
}

int sipp_peer_attach(sipp_ctx* ctx, int32_t rank, int32_t world, const void* handles) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!handles || world < 1 || world > SIPP_PEER_MAX_WORLD || rank < 0 || rank >= world)
    return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad peer_attach arguments (rank %d, world %d, max %d)", rank, world, SIPP_PEER_MAX_WORLD);
  if (!ctx->d_peer_box) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "sipp_peer_export must be called before sipp_peer_attach");
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  peer_release(ctx);
  PeerDev pd;
  memset(&pd, 0, sizeof(pd));
  pd.rank = rank;
  pd.world = world;
  const char* to = getenv("SIPP_PEER_TIMEOUT_MS");
  pd.timeout_ns = (unsigned long long)(to ? atoll(to) : 20000) * 1000000ull;
  const int64_t me = (int64_t)getpid();
  for (int r = 0; r < world; ++r) {
    HandleBlob b;
    memcpy(&b, reinterpret_cast<const char*>(handles) + (size_t)r * SIPP_PEER_HANDLE_BYTES, sizeof(b));
    if (b.magic != kMagic) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "handle %d is not a sipp peer handle", r);
    if (r == rank) {
      if (b.pid != me || b.ptr != (uint64_t)(uintptr_t)ctx->d_peer_box)
        return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "handle %d is not the one this context exported", r);
      pd.box[r] = ctx->d_peer_box;
    } else if (b.pid == me) {
      // same process (several contexts driven by one host program): the pointer is valid as is; other device -> peer access
      if (b.device != ctx->device) {
        int can = 0;
        SIPP_CUDA_CHECK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, b.device));
        if (!can) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", ctx->device, b.device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      pd.box[r] = reinterpret_cast<PeerBox*>((uintptr_t)b.ptr);
    } else {
      if (!b.has_ipc) return sipp_set_err(ctx, SIPP_ERR_UNSUPPORTED, "rank %d could not export a CUDA IPC handle for its mailbox", r);
      void* p = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        peer_release(ctx);
        return sipp_set_err(ctx, SIPP_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, device %d): %s", r, b.device, cudaGetErrorString(e));
      }
      ctx->peer_mapped[r] = p;
      pd.box[r] = reinterpret_cast<PeerBox*>(p);
    }
  }
  ctx->peer = pd;
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  if (ctx->cycle_graph) { cudaGraphExecDestroy(ctx->cycle_graph); ctx->cycle_graph = nullptr; }   // a captured pipeline may hold the old table
  return SIPP_OK;
}

int sipp_peer_detach(sipp_ctx* ctx) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  peer_release(ctx);
  if (ctx->cycle_graph) { cudaGraphExecDestroy(ctx->cycle_graph); ctx->cycle_graph = nullptr; }
  return SIPP_OK;
}

int sipp_peer_status(sipp_ctx* ctx, int32_t* status, int64_t* steps_done) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!ctx->d_peer_box) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "no peer mailbox (sipp_peer_export was never called)");
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SIPP_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long v[3];
  SIPP_CUDA_CHECK(ctx, cudaMemcpy(v, &ctx->d_peer_box->published, sizeof(v), cudaMemcpyDeviceToHost));
  if (steps_done) *steps_done = (int64_t)v[1];     // collected
  if (status) *status = (int32_t)v[2];
  return SIPP_OK;
}

int sipp_peer_exchange_dev(sipp_ctx* ctx, const void* d_mine, void* d_out, void* stream) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!d_mine || !d_out) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad peer_exchange arguments");
  if (ctx->peer_world < 1) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "peers are not attached (sipp_peer_attach)");
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  SIPP_CUDA_CHECK(ctx, launch_peer_exchange(ctx, reinterpret_cast<const BestRec*>(d_mine), reinterpret_cast<BestRec*>(d_out), PEER_SYNC, s));
  return SIPP_OK;
}

int sipp_peer_flush_dev(sipp_ctx* ctx, void* d_out, void* stream) {
  if (!ctx) return SIPP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!d_out) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "bad peer_flush arguments");
  if (ctx->peer_world < 1) return sipp_set_err(ctx, SIPP_ERR_INVALID_ARG, "peers are not attached (sipp_peer_attach)");
  SIPP_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  SIPP_CUDA_CHECK(ctx, launch_peer_exchange(ctx, nullptr, reinterpret_cast<BestRec*>(d_out), PEER_COLLECT_ONLY, s));
  return SIPP_OK;
}

}  // extern "C"
