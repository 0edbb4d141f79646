// shard.cuh -- object sharding across the GPUs of a node: one process per GPU, rank r
// owns the models assigned to it, ONE NCCL broadcast of the packed frame per time step (root = rank 0), every
// rank rebuilds the filtered depth and the pyramids locally.  NCCL is resolved at run time (dlopen of
// libnccl.so.2): single-GPU users carry no dependency on it, and inside a PyTorch process the library shares
// the NCCL build torch has already loaded.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cfb {

class FrameShard {
 public:
  ~FrameShard();
  // unique id for ncclCommInitRank, produced on rank 0 and distributed by the application (128 bytes)
  static int uniqueId(unsigned char id[128], const char** err);
  // collective over all ranks: creates the communicator on the current device
  int init(int rank, int world, const unsigned char id[128], size_t packedBytes, const char** err);
  bool active() const { return comm_ != nullptr; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  // Step 1 (all ranks): take the buffer of the next frame.  Everything enqueued so far on `consumer` (the previous
  // frame) is what still reads the other buffer; *free_evt is the event a stream that is about to WRITE the returned
  // buffer (the root's uploads) has to wait for.
  cudaError_t acquire(cudaStream_t consumer, uint8_t** buf, cudaEvent_t* free_evt);
  // Step 2 (all ranks): ONE ncclBroadcast of that buffer from rank 0 on the communication stream, after `ready` (the
  // root's uploads; null on the other ranks); `consumer` is made to wait for the data.
  cudaError_t broadcast(cudaEvent_t ready, cudaStream_t consumer, const char** err);
  uint8_t* current() { return buf_[cur_]; }

 private:
  void* comm_ = nullptr;  // ncclComm_t
  int rank_ = 0, world_ = 1;
  size_t bytes_ = 0;
  uint8_t* buf_[2] = {nullptr, nullptr};
  int cur_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t evDone_[2] = {nullptr, nullptr}, evFree_[2] = {nullptr, nullptr};
};

// which rank owns the model at list position `index` (0 = camera / background model) of a scene sharded over `world`
// ranks: round robin, so that rank r of a world of N owns model r of an N-model scene
inline int shard_owner(int index, int world) { return world > 0 ? index % world : 0; }

}  // namespace cfb
