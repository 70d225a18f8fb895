// pm_nccl.hpp — the one collective pattern of the slab decomposition:
// halo rows by ncclSend/ncclRecv between j-neighbours, scalars by ncclAllReduce.
// libnccl is resolved at run time (dlopen "libnccl.so.2"): a single-GPU handle never needs it, and
// inside a torch process the already-loaded NCCL of torch.distributed is the one that answers.
#pragma once
#include <dlfcn.h>
#include <cstdint>
#include <cstring>
#include <string>

#include <cuda_runtime.h>
#include <nccl.h>

struct PmNccl {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static inline bool pm_nccl_load(PmNccl* n, std::string* err) {
  if (n->lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n->lib) break;
  }
  if (!n->lib) { *err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define PM_SYM(field, name)                                                    \
  *(void**)(&n->field) = dlsym(n->lib, name);                                  \
  if (!n->field) { *err = std::string("libnccl lacks ") + name; return false; }
  PM_SYM(GetUniqueId, "ncclGetUniqueId")
  PM_SYM(CommInitRank, "ncclCommInitRank")
  PM_SYM(CommDestroy, "ncclCommDestroy")
  PM_SYM(Send, "ncclSend")
  PM_SYM(Recv, "ncclRecv")
  PM_SYM(AllReduce, "ncclAllReduce")
  PM_SYM(Broadcast, "ncclBroadcast")
  PM_SYM(GroupStart, "ncclGroupStart")
  PM_SYM(GroupEnd, "ncclGroupEnd")
  PM_SYM(GetErrorString, "ncclGetErrorString")
#undef PM_SYM
  return true;
}
#define PM_NCCL_CK(call)                                                                      \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) { *err = std::string(#call ": ") + n->GetErrorString(r_); return false; } \
  } while (0)

static inline bool pm_nccl_get_unique_id(uint8_t out[128], std::string* err) {
  static PmNccl loader;
  PmNccl* n = &loader;
  if (!pm_nccl_load(n, err)) return false;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  PM_NCCL_CK(n->GetUniqueId(&id));
  std::memcpy(out, &id, 128);
  return true;
}
static inline bool pm_nccl_init(PmNccl* n, const uint8_t idbytes[128], int nranks, int rank, std::string* err) {
  if (!pm_nccl_load(n, err)) return false;
  ncclUniqueId id;
  std::memcpy(&id, idbytes, 128);
  n->rank = rank; n->nranks = nranks;
  PM_NCCL_CK(n->CommInitRank(&n->comm, nranks, id, rank));
  return true;
}
static inline void pm_nccl_destroy(PmNccl* n) {
  if (n->comm && n->CommDestroy) n->CommDestroy(n->comm);
  n->comm = nullptr;
}
// Grouped neighbour exchange along the slab chain; null pointers skip a direction.
//   send_up -> rank+1, recv_up <- rank+1, send_dn -> rank-1, recv_dn <- rank-1.   n doubles each.
static inline bool pm_nccl_exchange(PmNccl* n, cudaStream_t st, const double* send_up, double* recv_up,
                                    const double* send_dn, double* recv_dn, size_t cnt, std::string* err) {
  PM_NCCL_CK(n->GroupStart());
  if (send_up) PM_NCCL_CK(n->Send(send_up, cnt, ncclDouble, n->rank + 1, n->comm, st));
  if (recv_up) PM_NCCL_CK(n->Recv(recv_up, cnt, ncclDouble, n->rank + 1, n->comm, st));
  if (send_dn) PM_NCCL_CK(n->Send(send_dn, cnt, ncclDouble, n->rank - 1, n->comm, st));
  if (recv_dn) PM_NCCL_CK(n->Recv(recv_dn, cnt, ncclDouble, n->rank - 1, n->comm, st));
  PM_NCCL_CK(n->GroupEnd());
  return true;
}
// max over non-negative doubles == max over their bit patterns as uint64
static inline bool pm_nccl_allreduce_max_u64(PmNccl* n, cudaStream_t st, unsigned long long* buf, size_t cnt, std::string* err) {
  PM_NCCL_CK(n->AllReduce(buf, buf, cnt, ncclUint64, ncclMax, n->comm, st));
  return true;
}
static inline bool pm_nccl_allreduce_sum_f64(PmNccl* n, cudaStream_t st, double* buf, size_t cnt, std::string* err) {
  PM_NCCL_CK(n->AllReduce(buf, buf, cnt, ncclDouble, ncclSum, n->comm, st));
  return true;
}
// A few 64-bit words along the slab chain (the exact source mean): from rank-1, then on to rank+1.
static inline bool pm_nccl_recv_words(PmNccl* n, cudaStream_t st, unsigned long long* buf, size_t cnt, int peer, std::string* err) {
  PM_NCCL_CK(n->Recv(buf, cnt, ncclUint64, peer, n->comm, st));
  return true;
}
static inline bool pm_nccl_send_words(PmNccl* n, cudaStream_t st, const unsigned long long* buf, size_t cnt, int peer, std::string* err) {
  PM_NCCL_CK(n->Send(buf, cnt, ncclUint64, peer, n->comm, st));
  return true;
}
static inline bool pm_nccl_bcast_words(PmNccl* n, cudaStream_t st, unsigned long long* buf, size_t cnt, int root, std::string* err) {
  PM_NCCL_CK(n->Broadcast(buf, buf, cnt, ncclUint64, root, n->comm, st));
  return true;
}
