// shard.cu -- see shard.cuh.
#include "shard.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the entry points are resolved with dlsym
#include <stdio.h>
#include <string.h>

namespace cfb {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return GetUniqueId && CommInitRank && CommDestroy && Broadcast && GetErrorString; }
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // RTLD_NOLOAD first: inside a PyTorch process this is the NCCL torch already runs on
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.CommAbort = (decltype(api.CommAbort))dlsym(api.handle, "ncclCommAbort");
      api.Broadcast = (decltype(api.Broadcast))dlsym(api.handle, "ncclBroadcast");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    }
  }
  return api;
}

thread_local char g_err[256];
const char* fail(const char* what, ncclResult_t r) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL unavailable");
  return g_err;
}

}  // namespace

int FrameShard::uniqueId(unsigned char id[128], const char** err) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi& n = nccl();
  if (!n.ok()) {
    *err = "libnccl.so.2 could not be loaded";
    return 1;
  }
  ncclUniqueId u;
  ncclResult_t r = n.GetUniqueId(&u);
  if (r != ncclSuccess) {
    *err = fail("ncclGetUniqueId", r);
    return 1;
  }
  memcpy(id, &u, 128);
  return 0;
}

int FrameShard::init(int rank, int world, const unsigned char id[128], size_t packedBytes, const char** err) {
  NcclApi& n = nccl();
  if (!n.ok()) {
    *err = "libnccl.so.2 could not be loaded";
    return 1;
  }
  if (comm_ || world < 1 || rank < 0 || rank >= world) {
    *err = "shard_init: bad rank / world, or already initialised";
    return 1;
  }
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = n.CommInitRank(&c, world, u, rank);
  if (r != ncclSuccess) {
    *err = fail("ncclCommInitRank", r);
    return 1;
  }
  comm_ = c;
  rank_ = rank;
  world_ = world;
  bytes_ = packedBytes;
  bool good = cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) == cudaSuccess;
  for (int k = 0; k < 2 && good; ++k)
    good = cudaMalloc((void**)&buf_[k], packedBytes) == cudaSuccess && cudaMemset(buf_[k], 0, packedBytes) == cudaSuccess &&
           cudaEventCreateWithFlags(&evDone_[k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&evFree_[k], cudaEventDisableTiming) == cudaSuccess;
  if (!good) {
    *err = "shard_init: device allocation failed";
    return 1;
  }
  return 0;
}

FrameShard::~FrameShard() {
  // every broadcast this rank issued has completed: release the communicator without waiting for the peers
  // (ncclCommDestroy finalises collectively and was seen to hang at the end of a 4-rank run; abort is local)
  if (stream_) cudaStreamSynchronize(stream_);
  if (comm_ && nccl().CommAbort)
    nccl().CommAbort((ncclComm_t)comm_);
  else if (comm_ && nccl().CommDestroy)
    nccl().CommDestroy((ncclComm_t)comm_);
  for (int k = 0; k < 2; ++k) {
    cudaFree(buf_[k]);
    if (evDone_[k]) cudaEventDestroy(evDone_[k]);
    if (evFree_[k]) cudaEventDestroy(evFree_[k]);
  }
  if (stream_) cudaStreamDestroy(stream_);
}

cudaError_t FrameShard::acquire(cudaStream_t consumer, uint8_t** buf, cudaEvent_t* free_evt) {
  cudaError_t e = cudaEventRecord(evFree_[cur_], consumer);
  if (e != cudaSuccess) return e;
  cur_ ^= 1;
  *buf = buf_[cur_];
  *free_evt = evFree_[cur_];  // recorded two frames ago (or never: then the wait is a no-op)
  return cudaSuccess;
}

cudaError_t FrameShard::broadcast(cudaEvent_t ready, cudaStream_t consumer, const char** err) {
  cudaError_t e;
  if ((e = cudaStreamWaitEvent(stream_, evFree_[cur_], 0)) != cudaSuccess) return e;
  if (ready && (e = cudaStreamWaitEvent(stream_, ready, 0)) != cudaSuccess) return e;
  ncclResult_t r = nccl().Broadcast(buf_[cur_], buf_[cur_], bytes_, ncclUint8, 0, (ncclComm_t)comm_, stream_);
  if (r != ncclSuccess) {
    *err = fail("ncclBroadcast", r);
    return cudaErrorUnknown;
  }
  if ((e = cudaEventRecord(evDone_[cur_], stream_)) != cudaSuccess) return e;
  return cudaStreamWaitEvent(consumer, evDone_[cur_], 0);
}

}  // namespace cfb
