// Multi-GPU plumbing of libplonk_b200.so: one process per GPU, one NCCL communicator per context, created from a
// unique id that the caller distributes with its own rendezvous (torch.distributed in plonkathon_b200/parallel.py).
// The data-path collectives -- the allgather at the join of the slab-sharded NTT and the allgather of the 256-byte
// MSM partial sums -- are issued by the library itself on the context's stream, so a sharded proof never bounces
// through the host between a kernel and its exchange step.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has, e.g. torch's): the library keeps
// loading, and every single-GPU entry point keeps working, on a machine without NCCL.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pb200 {

struct Comm {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  int log_world = 0;
  uint64_t collectives = 0;     // data-path collectives issued so far (bench / tests)
  uint64_t bytes_gathered = 0;  // bytes received by this rank in them
};

// 128 opaque bytes (ncclUniqueId) for pb200_comm_init on every rank; call on one rank only
void comm_unique_id(uint8_t out[128]);
Comm* comm_create(const uint8_t id[128], int rank, int world);
void comm_destroy(Comm* c);
// every rank contributes `bytes` at recv + rank * bytes (in place) and ends up with all world * bytes
void comm_allgather_inplace(Comm* c, void* recv, size_t bytes, cudaStream_t stream);
inline int comm_rank(const Comm* c) { return c->rank; }
inline int comm_world(const Comm* c) { return c->world; }
inline int comm_log_world(const Comm* c) { return c->log_world; }

}  // namespace pb200
