// comm.cu — the one exchange step of the hot path: all-gather of the per-rank (target_prob, top1) score blocks.
//
// Masks are independent work units (generate_gp_training_data_imagenet.py:221-266 has no cross-mask state); rank r of W
// scores a contiguous slice and every rank needs the full table afterwards (the GP rank fits on all of it).  The
// reference declares `--dist-backend`, `--world-size` (generate_gp_training_data_imagenet.py:72-77) but never
// initialises a process group (:572 only computes `args.distributed`); this is the backend it stops short of.
//
// NCCL is bound at run time with dlopen (the copy PyTorch already loaded is reused when present), so libnib.so has no
// link-time dependency on a particular NCCL build.  Bootstrap (who is rank 0, how the 128-byte unique id travels) is
// the caller's business — the host mirror broadcasts it over torch.distributed; the data path is ncclAllGather on the
// caller's stream, fed directly by the score kernel's (prob, top1) table (score.cu), no packing kernels in between.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace nib {

// minimal NCCL surface (nccl.h 2.x ABI: ncclUniqueId is 128 opaque bytes, ncclFloat32 == 7, ncclSuccess == 0)
typedef struct ncclComm* nccl_comm_t;
typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*pfn_ncclGetUniqueId)(nccl_unique_id*);
typedef int (*pfn_ncclCommInitRank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*pfn_ncclAllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*pfn_ncclCommDestroy)(nccl_comm_t);
typedef const char* (*pfn_ncclGetErrorString)(int);
static constexpr int kNcclFloat32 = 7;

struct NcclApi {
  void* handle;
  pfn_ncclGetUniqueId GetUniqueId;
  pfn_ncclCommInitRank CommInitRank;
  pfn_ncclAllGather AllGather;
  pfn_ncclCommDestroy CommDestroy;
  pfn_ncclGetErrorString GetErrorString;
};
static NcclApi g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static int load_nccl() {
  if (g_nccl.handle) return NIB_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // already in the process (PyTorch's copy)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("nib_comm: libnccl.so.2 not found (%s); load PyTorch first or put NCCL on LD_LIBRARY_PATH", dlerror());
    return NIB_ESTATE;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (pfn_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (pfn_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
  a.AllGather = (pfn_ncclAllGather)dlsym(h, "ncclAllGather");
  a.CommDestroy = (pfn_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (pfn_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
    set_error("nib_comm: libnccl.so.2 lacks a required symbol");
    return NIB_ESTATE;
  }
  g_nccl = a;
  return NIB_OK;
}

#define NIB_NCCL(expr)                                                                          \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != 0) {                                                                              \
      nib::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nib::g_nccl.GetErrorString(_r)); \
      return NIB_ECUDA;                                                                         \
    }                                                                                           \
  } while (0)

}  // namespace nib

struct nib_comm {
  nib::nccl_comm_t comm;
  int rank, world;
};

extern "C" {

int nib_comm_unique_id(void* h_id128) {
  NIB_REQUIRE(h_id128 != nullptr, "nib_comm_unique_id: null buffer");
  int rc = nib::load_nccl();
  if (rc != NIB_OK) return rc;
  nib::nccl_unique_id id;
  NIB_NCCL(nib::g_nccl.GetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return NIB_OK;
}

int nib_comm_init(const void* h_id128, int rank, int world, nib_comm** out) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(h_id128 && out, "nib_comm_init: null pointer");
  NIB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "nib_comm_init: rank %d outside world %d", rank, world);
  int rc = nib::load_nccl();
  if (rc != NIB_OK) return rc;
  nib::nccl_unique_id id;
  memcpy(&id, h_id128, sizeof(id));
  nib_comm* c = new nib_comm();
  c->rank = rank;
  c->world = world;
  c->comm = nullptr;
  int r = nib::g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    nib::set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, nib::g_nccl.GetErrorString(r));
    delete c;
    return NIB_ECUDA;
  }
  *out = c;
  return NIB_OK;
}

int nib_comm_destroy(nib_comm* comm) {
  if (!comm) return NIB_OK;
  if (comm->comm && nib::g_nccl.CommDestroy) nib::g_nccl.CommDestroy(comm->comm);
  delete comm;
  return NIB_OK;
}

int nib_allgather_scores(nib_comm* comm, const float* d_local, int rows_per_rank, float* d_table, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(comm && comm->comm, "nib_allgather_scores: communicator not initialised");
  NIB_REQUIRE(d_local && d_table && rows_per_rank >= 0, "nib_allgather_scores: bad arguments");
  if (rows_per_rank == 0) return NIB_OK;
  NIB_NCCL(nib::g_nccl.AllGather(d_local, d_table, (size_t)rows_per_rank * 2, nib::kNcclFloat32, comm->comm,
                                 (cudaStream_t)stream));
  return NIB_OK;
}

}  // extern "C"
