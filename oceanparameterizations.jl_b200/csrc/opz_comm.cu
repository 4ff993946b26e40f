// opz_comm.cu — the one collective of the path, inside the boundary: the sum over ranks of the NN-parameter gradient
// and the loss components of a training step (the loss is a mean over simulations, wind_mixing/src/NDE_training.jl:290-301,
// so with the cases sharded over GPUs every rank normalises by the global batch and the exchange is a plain sum).
//
// One process (or host thread) per GPU, one opz_ctx per GPU.  NCCL is bound at run time (dlopen of libnccl.so.2 on first
// use): a single-GPU user of libopz.so needs no NCCL at all, and inside a process that already carries a NCCL (PyTorch's)
// the same library instance is shared.  The all-reduce is enqueued on the context's stream right behind the BPTT kernels,
// so the `_dev` entry points stay asynchronous and the host-buffer entry points copy back already-reduced values.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: nothing here links against libnccl
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "opz_internal.h"
#include "../../include/opz.h"

namespace opz {

namespace {
struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string why;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  // 1. a NCCL this process already carries (PyTorch's): share that instance; 2. OPZ_NCCL_LIB; 3. the system library.
  //    RTLD_LOCAL: nothing else may bind to our choice through the global namespace.  (The loader still de-duplicates by
  //    SONAME: whoever loads "libnccl.so.2" first decides for the whole process.  A host that will import PyTorch later
  //    should point OPZ_NCCL_LIB at PyTorch's copy — the Python package does so on import.)
  g_nccl.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
  const char* names[4] = {getenv("OPZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 3 && !g_nccl.handle; ++i)
    if (names[i] && *names[i]) g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
  if (!g_nccl.handle) { g_nccl.why = std::string("cannot load libnccl.so.2 (set OPZ_NCCL_LIB): ") + (dlerror() ? dlerror() : "?"); return; }
#define OPZ_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name)); \
  if (!g_nccl.field) { g_nccl.why = std::string("libnccl lacks ") + name; g_nccl.handle = nullptr; return; }
  OPZ_SYM(GetUniqueId, "ncclGetUniqueId")
  OPZ_SYM(CommInitRank, "ncclCommInitRank")
  OPZ_SYM(CommDestroy, "ncclCommDestroy")
  OPZ_SYM(AllReduce, "ncclAllReduce")
  OPZ_SYM(GroupStart, "ncclGroupStart")
  OPZ_SYM(GroupEnd, "ncclGroupEnd")
  OPZ_SYM(GetErrorString, "ncclGetErrorString")
  OPZ_SYM(GetVersion, "ncclGetVersion")
#undef OPZ_SYM
}

int nccl_ready() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.handle) { set_error(g_nccl.why); return OPZ_ENCCL; }
  return OPZ_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
  set_error(std::string("NCCL error: ") + g_nccl.GetErrorString(r) + " in " + what);
  return OPZ_ENCCL;
}
#define OPZ_NCCL(expr)                                   \
  do {                                                   \
    ncclResult_t _r = (expr);                            \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr);  \
  } while (0)
}  // namespace

// Sum over the ranks of the context's communicator of up to three device buffers (gradient, loss components, mPP
// gradient), in place, enqueued on the context's stream as ONE NCCL group (one launch-latency, not three).
int comm_allreduce_sum(opz_ctx* ctx, float* const* bufs, const size_t* counts, int n) {
  if (!ctx->nccl_comm || ctx->comm_ranks <= 1) return OPZ_OK;   // single rank: the sum over ranks is the value itself
  int rc = nccl_ready();
  if (rc) return rc;
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  OPZ_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; ++i)
    if (bufs[i] && counts[i]) OPZ_NCCL(g_nccl.AllReduce(bufs[i], bufs[i], counts[i], ncclFloat32, ncclSum, comm, ctx->stream));
  OPZ_NCCL(g_nccl.GroupEnd());
  ctx->collectives += 1;
  return OPZ_OK;
}

void comm_release(opz_ctx* ctx) {
  if (ctx->nccl_comm && ctx->owns_comm && g_nccl.handle) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->owns_comm = false;
  ctx->comm_ranks = 1;
  ctx->comm_rank = 0;
}

}  // namespace opz

using namespace opz;

extern "C" {

int opz_comm_get_unique_id(void* id_out) {
  if (!id_out) { set_error("null id_out"); return OPZ_EINVAL; }
  static_assert(sizeof(ncclUniqueId) == OPZ_COMM_ID_BYTES, "opz.h: OPZ_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
  int rc = nccl_ready();
  if (rc) return rc;
  ncclUniqueId id;
  OPZ_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return OPZ_OK;
}

int opz_ctx_comm_init(opz_ctx* ctx, const void* id, int n_ranks, int rank) {
  if (!ctx || !id) { set_error("null ctx/id"); return OPZ_EINVAL; }
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("opz_ctx_comm_init: need 0 <= rank < n_ranks"); return OPZ_EINVAL; }
  int rc = nccl_ready();
  if (rc) return rc;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  comm_release(ctx);
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  OPZ_NCCL(g_nccl.CommInitRank(&comm, n_ranks, uid, rank));
  ctx->nccl_comm = comm;
  ctx->owns_comm = true;
  ctx->comm_ranks = n_ranks;
  ctx->comm_rank = rank;
  return OPZ_OK;
}

int opz_ctx_comm_adopt(opz_ctx* ctx, void* nccl_comm, int n_ranks, int rank) {
  if (!ctx) { set_error("null ctx"); return OPZ_EINVAL; }
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("opz_ctx_comm_adopt: need 0 <= rank < n_ranks"); return OPZ_EINVAL; }
  if (nccl_comm) {
    int rc = nccl_ready();
    if (rc) return rc;
  } else if (n_ranks > 1) {
    set_error("opz_ctx_comm_adopt: null communicator with n_ranks > 1");
    return OPZ_EINVAL;
  }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  comm_release(ctx);
  ctx->nccl_comm = nccl_comm;
  ctx->owns_comm = false;
  ctx->comm_ranks = n_ranks;
  ctx->comm_rank = rank;
  return OPZ_OK;
}

int opz_ctx_comm_destroy(opz_ctx* ctx) {
  if (!ctx) return OPZ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  return OPZ_OK;
}

int opz_ctx_comm_info(const opz_ctx* ctx, int* n_ranks, int* rank, int64_t* collectives) {
  if (!ctx) { set_error("null ctx"); return OPZ_EINVAL; }
  if (n_ranks) *n_ranks = ctx->comm_ranks;
  if (rank) *rank = ctx->comm_rank;
  if (collectives) *collectives = ctx->collectives;
  return OPZ_OK;
}

int opz_comm_nccl_version(int* version) {
  if (!version) { set_error("null version"); return OPZ_EINVAL; }
  int rc = nccl_ready();
  if (rc) return rc;
  OPZ_NCCL(g_nccl.GetVersion(version));
  return OPZ_OK;
}

/* contiguous shard [lo, hi) of B cases owned by `rank` of `n_ranks`: the first B % n_ranks ranks own one more */
int opz_shard_range(int64_t B, int rank, int n_ranks, int64_t* lo, int64_t* hi) {
  if (B < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !lo || !hi) { set_error("opz_shard_range: bad argument"); return OPZ_EINVAL; }
  const int64_t base = B / n_ranks, extra = B % n_ranks;
  *lo = rank * base + (rank < extra ? rank : extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
  return OPZ_OK;
}

}  // extern "C"
