// NCCL entry points of the C ABI: the two exchange steps of the data-parallel path --
// gradient averaging before each optimizer step (the reference's nn.DataParallel reduce,
// train.py:145-152,497) and the int64 confusion-matrix sum of the evaluation (train.py:47 accumulated
// over a sharded loader) -- for callers that bind libb200seg.so without torch.distributed.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") (inside a PyTorch process that is the copy
// torch already loaded), so the library has no link-time dependency on it and still loads where NCCL
// is absent; every b200_nccl_* call then returns B200_EDRIVER.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "status.h"
#include "b200seg.h"

namespace b200 {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    memset(&api, 0, sizeof(api));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

static int nccl_fail(NcclApi* a, const char* what, ncclResult_t r) {
  return set_error(B200_ECUDA, "%s: %s", what, a->GetErrorString(r));
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_nccl_unique_id(void* id128) {
  NcclApi* a = nccl_api();
  if (!a) return set_error(B200_EDRIVER, "libnccl.so.2 not found (dlopen)");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(a, "ncclGetUniqueId", r);
  memcpy(id128, &id, sizeof(id));
  return B200_OK;
}

int b200_nccl_init(const void* id128, int world, int rank, void** comm_out) {
  NcclApi* a = nccl_api();
  if (!a) return set_error(B200_EDRIVER, "libnccl.so.2 not found (dlopen)");
  if (world < 1 || rank < 0 || rank >= world || comm_out == nullptr)
    return set_error(B200_EINVAL, "nccl_init: bad world / rank (%d / %d)", world, rank);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = a->CommInitRank(&comm, world, id, rank);   // on the calling thread's current device
  if (r != ncclSuccess) return nccl_fail(a, "ncclCommInitRank", r);
  *comm_out = comm;
  return B200_OK;
}

int b200_nccl_allreduce_grads(void* comm, float* flat, int64_t count, int average, cudaStream_t stream) {
  NcclApi* a = nccl_api();
  if (!a) return set_error(B200_EDRIVER, "libnccl.so.2 not found (dlopen)");
  if (comm == nullptr || count < 0) return set_error(B200_EINVAL, "nccl_allreduce_grads: bad arguments");
  if (count == 0) return B200_OK;
  ncclResult_t r = a->AllReduce(flat, flat, (size_t)count, ncclFloat32, average ? ncclAvg : ncclSum,
                                static_cast<ncclComm_t>(comm), stream);
  if (r != ncclSuccess) return nccl_fail(a, "ncclAllReduce(grads)", r);
  return B200_OK;
}

int b200_nccl_allreduce_hist(void* comm, int64_t* hist, int count, cudaStream_t stream) {
  NcclApi* a = nccl_api();
  if (!a) return set_error(B200_EDRIVER, "libnccl.so.2 not found (dlopen)");
  if (comm == nullptr || count < 0) return set_error(B200_EINVAL, "nccl_allreduce_hist: bad arguments");
  if (count == 0) return B200_OK;
  ncclResult_t r = a->AllReduce(hist, hist, (size_t)count, ncclInt64, ncclSum, static_cast<ncclComm_t>(comm), stream);
  if (r != ncclSuccess) return nccl_fail(a, "ncclAllReduce(hist)", r);
  return B200_OK;
}

int b200_nccl_destroy(void* comm) {
  NcclApi* a = nccl_api();
  if (!a) return set_error(B200_EDRIVER, "libnccl.so.2 not found (dlopen)");
  if (comm == nullptr) return B200_OK;
  ncclResult_t r = a->CommDestroy(static_cast<ncclComm_t>(comm));
  if (r != ncclSuccess) return nccl_fail(a, "ncclCommDestroy", r);
  return B200_OK;
}

}  // extern "C"
