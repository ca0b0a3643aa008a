// fdx_comm.cu -- the ONE exchange step of data-parallel training behind the C-ABI:
//   jax.lax.pmean(grads, "data") / pmean(loss)      trainer/general_diffusion_trainer.py:325,334
//   (mesh / shard_map plumbing :340-345; simple_trainer.py:175-180)
// as NCCL all-reduce(avg) calls on caller-chosen ranges ("buckets") of the flat f32 gradient buffer, issued
// on a caller-supplied stream so that the trainer can launch a bucket as soon as the backward pass has
// finished the parameters it covers (NVLink 5 / NVSwitch; NVLS in-switch reduction when NCCL enables it).
//
// libfdx.so keeps NO link-time dependency on NCCL (it must load on the CPU-only build box): the library is
// bound at run time - first the copy already mapped into the process (PyTorch's libnccl.so.2), then
// $FDX_NCCL_LIB, then the system search path.  One communicator per process, one process per GPU.
#include "fdx_common.cuh"
#include "../../include/fdx.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0, ncclAvg = 4 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int bind_nccl() {
  if (g_nccl.AllReduce) return FDX_OK;
  const char* env = getenv("FDX_NCCL_LIB");
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by PyTorch?
  if (!h && env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    fdx_set_error("comm: libnccl.so.2 not found (set FDX_NCCL_LIB to its path): %s", dlerror());
    return FDX_ERR_UNSUPPORTED;
  }
  g_nccl.handle = h;
#define FDX_BIND(field, sym)                                                   \
  *(void**)(&g_nccl.field) = dlsym(h, sym);                                    \
  if (!g_nccl.field) {                                                         \
    fdx_set_error("comm: symbol %s missing from libnccl", sym);                \
    g_nccl.AllReduce = nullptr;                                                \
    return FDX_ERR_UNSUPPORTED;                                                \
  }
  FDX_BIND(GetUniqueId, "ncclGetUniqueId")
  FDX_BIND(CommInitRank, "ncclCommInitRank")
  FDX_BIND(CommDestroy, "ncclCommDestroy")
  FDX_BIND(GroupStart, "ncclGroupStart")
  FDX_BIND(GroupEnd, "ncclGroupEnd")
  FDX_BIND(GetVersion, "ncclGetVersion")
  FDX_BIND(GetErrorString, "ncclGetErrorString")
  FDX_BIND(AllReduce, "ncclAllReduce")
#undef FDX_BIND
  return FDX_OK;
}

int nccl_check(int r, const char* what) {
  if (r == ncclSuccess) return FDX_OK;
  fdx_set_error("NCCL error %d (%s) at %s", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
  return FDX_ERR_CUDA;
}

}  // namespace

struct fdx_comm {
  ncclComm_t comm;
  int rank, world;
};

extern "C" {

int fdx_comm_unique_id(void* id128) {
  FDX_REQUIRE(id128, "comm_unique_id: null buffer");
  int s = bind_nccl();
  if (s != FDX_OK) return s;
  ncclUniqueId id;
  if ((s = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) != FDX_OK) return s;
  memcpy(id128, &id, sizeof(id));
  return FDX_OK;
}

int fdx_comm_init(fdx_comm** out, int rank, int world, const void* id128) {
  FDX_REQUIRE(out && id128, "comm_init: null pointer");
  FDX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank %d of %d", rank, world);
  int s = bind_nccl();
  if (s != FDX_OK) return s;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  if ((s = nccl_check(g_nccl.CommInitRank(&c, world, id, rank), "ncclCommInitRank")) != FDX_OK) return s;
  fdx_comm* h = new fdx_comm{c, rank, world};
  *out = h;
  return FDX_OK;
}

int fdx_comm_allreduce_avg(fdx_comm* c, float* buf, long long n, void* stream) {
  FDX_REQUIRE(c && c->comm && buf && n > 0, "comm_allreduce_avg: bad arguments");
  return nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclAvg, c->comm, (cudaStream_t)stream),
                    "ncclAllReduce");
}

int fdx_comm_world(const fdx_comm* c, int* rank, int* world) {
  FDX_REQUIRE(c, "comm_world: null communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return FDX_OK;
}

int fdx_comm_nccl_version(void) {
  if (bind_nccl() != FDX_OK) return 0;
  int v = 0;
  g_nccl.GetVersion(&v);
  return v;
}

int fdx_comm_destroy(fdx_comm* c) {
  if (!c) return FDX_OK;
  int s = FDX_OK;
  if (c->comm) s = nccl_check(g_nccl.CommDestroy(c->comm), "ncclCommDestroy");
  delete c;
  return s;
}

}  // extern "C"
