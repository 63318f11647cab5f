// NCCL binding of comm.cuh (run-time dlopen; see the header).
#include "comm.cuh"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace pb200 {

namespace {
// the handful of NCCL entry points used, with the ABI of nccl.h (2.x): ncclUniqueId is 128 bytes passed by value,
// ncclDataType_t ncclUint8 == 1, ncclResult_t ncclSuccess == 0
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_get_error_string error_string = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // prefer the copy the process has already mapped (torch's bundled libnccl.so.2), so both users share one NCCL
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
    api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    api.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    api.error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
  });
  PB_CHECK(api.handle && api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather,
           "multi-GPU entry points need NCCL (libnccl.so.2 was not found in the process or on the loader path)");
  return api;
}

void nccl_check(int rc, const char* what) {
  if (rc == 0) return;
  char b[512];
  NcclApi& api = nccl();
  snprintf(b, sizeof b, "%s failed: %s", what, api.error_string ? api.error_string(rc) : "NCCL error");
  throw Error(b);
}
}  // namespace

Comm* ctx_comm(Context* ctx) {
  PB_CHECK(ctx->comm, "this call needs a communicator on the context (pb200_comm_init)");
  return ctx->comm;
}

void comm_unique_id(uint8_t out[128]) {
  NcclUniqueId id;
  nccl_check(nccl().get_unique_id(&id), "ncclGetUniqueId");
  memcpy(out, id.internal, 128);
}

Comm* comm_create(const uint8_t id_bytes[128], int rank, int world) {
  PB_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
  int lg = 0;
  while ((1 << lg) < world) lg++;
  PB_CHECK((1 << lg) == world && world <= 8, "the sharded path needs 1, 2, 4 or 8 ranks (one box)");
  auto c = std::make_unique<Comm>();
  c->rank = rank;
  c->world = world;
  c->log_world = lg;
  NcclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  nccl_check(nccl().comm_init_rank(&c->nccl_comm, world, id, rank), "ncclCommInitRank");
  return c.release();
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->nccl_comm) nccl().comm_destroy(c->nccl_comm);
  delete c;
}

void comm_allgather_inplace(Comm* c, void* recv, size_t bytes, cudaStream_t stream) {
  PB_CHECK(c && c->nccl_comm, "no communicator on this context (pb200_comm_init)");
  const char* send = static_cast<const char*>(recv) + (size_t)c->rank * bytes;
  nccl_check(nccl().all_gather(send, recv, bytes, /*ncclUint8*/ 1, c->nccl_comm, stream), "ncclAllGather");
  c->collectives++;
  c->bytes_gathered += (uint64_t)bytes * (c->world - 1);
}

}  // namespace pb200
