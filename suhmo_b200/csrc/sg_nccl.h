// sg_nccl.h -- thin run-time binding of NCCL (dlopen "libnccl.so.2"): the single-GPU path has no NCCL
// dependency at all; with nranks > 1 the halo rows travel by ncclSend/ncclRecv over NVLink and residual norms
// by ncclAllReduce.  When torch is already imported its bundled libnccl.so.2 is the one that gets bound.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#ifndef SG_OK
#define SG_OK 0
#endif
#define SG_ERR_NCCL_ 5

struct SgNcclApi {
  void* h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SGN_SYM(name) name = (decltype(name))dlsym(h, "nccl" #name); if (!name) { err = "missing symbol nccl" #name; return false; }
    SGN_SYM(GetUniqueId) SGN_SYM(CommInitRank) SGN_SYM(CommDestroy) SGN_SYM(Send) SGN_SYM(Recv)
    SGN_SYM(AllReduce) SGN_SYM(GroupStart) SGN_SYM(GroupEnd) SGN_SYM(GetErrorString)
#undef SGN_SYM
    return true;
  }
};
static SgNcclApi g_nccl_api;

static inline int sgnccl_unique_id(void* out128, std::string& err) {
  if (!g_nccl_api.load(err)) return SG_ERR_NCCL_;
  ncclUniqueId id;
  ncclResult_t r = g_nccl_api.GetUniqueId(&id);
  if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + g_nccl_api.GetErrorString(r); return SG_ERR_NCCL_; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return SG_OK;
}

struct SgNccl {
  ncclComm_t comm = nullptr;
  int check(ncclResult_t r, const char* what, std::string& err) {
    if (r == ncclSuccess) return SG_OK;
    err = std::string(what) + ": " + g_nccl_api.GetErrorString(r);
    return SG_ERR_NCCL_;
  }
  int init(const void* uid128, int rank, int nranks, std::string& err) {
    if (!g_nccl_api.load(err)) return SG_ERR_NCCL_;
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    return check(g_nccl_api.CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank", err);
  }
  void destroy() {
    if (comm) g_nccl_api.CommDestroy(comm);
    comm = nullptr;
  }
  int group_start(std::string& err) { return check(g_nccl_api.GroupStart(), "ncclGroupStart", err); }
  int group_end(std::string& err) { return check(g_nccl_api.GroupEnd(), "ncclGroupEnd", err); }
  int send(const double* p, size_t n, int peer, cudaStream_t s, std::string& err) {
    return check(g_nccl_api.Send(p, n, ncclFloat64, peer, comm, s), "ncclSend", err);
  }
  int recv(double* p, size_t n, int peer, cudaStream_t s, std::string& err) {
    return check(g_nccl_api.Recv(p, n, ncclFloat64, peer, comm, s), "ncclRecv", err);
  }
  int allreduce_max_u8(unsigned char* p, size_t n, cudaStream_t s, std::string& err) {
    return check(g_nccl_api.AllReduce(p, p, n, ncclUint8, ncclMax, comm, s), "ncclAllReduce", err);
  }
  int allreduce(double* p, size_t n, bool is_max, cudaStream_t s, std::string& err) {
    return check(g_nccl_api.AllReduce(p, p, n, ncclFloat64, is_max ? ncclMax : ncclSum, comm, s), "ncclAllReduce", err);
  }
};
