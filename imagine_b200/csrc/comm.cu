// Collectives behind the C ABI: a host that is not Python / torch.distributed (the reference's own MPI
// mode, a C++ sampler driver) can run the sharded likelihood with nothing but this library.  NCCL is
// loaded at run time (dlopen of libnccl.so.2 -- the copy the process already holds, if any), so the
// library has no link-time dependency on it and single-GPU users never touch it.
//   imb_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host ships it to the other ranks
//                        (MPI_Bcast in the reference's world: tools/mpi_helper.py; a file; a socket)
//   imb_comm_init        every rank joins; one communicator per process (= per GPU)
//   imb_allreduce_sum    SUM all-reduce of doubles in place -- the reference's `mpi_mean` /
//                        `Allreduce` building block (tools/mpi_helper.py:132)
//   imb_alltoall         block r of `send` goes to rank r, block q of `recv` comes from rank q (grouped
//                        ncclSend / ncclRecv on the caller's stream): the exchange of the simulated maps
//                        [P(dest), K, n_local, m/P] -> [P(source), K, n_local, m/P]
//   imb_loglike_ensemble_sharded
//                        EnsembleLikelihood.call (likelihoods/ensemble_likelihood.py:68-107) for K keys
//                        whose PIXELS are sharded over the ranks: phase 1 -> all-reduce G -> phase 2 ->
//                        all-reduce P2 -> phase 3, one read-back; the value is replicated on all ranks
//                        ("copied to all nodes", ensemble_likelihood.py:79-80)
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace imb {

typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(ncclComm_t*, int, NcclId, int);
typedef int (*fn_destroy)(ncclComm_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_sendrecv)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);   // ncclSend (ncclRecv: void*)
typedef int (*fn_group)(void);

struct Nccl {
  void* handle = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_errstr errstr = nullptr;
  fn_sendrecv send = nullptr, recv = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = -1;
};

static Nccl& nccl() {
  static Nccl n;
  return n;
}

static int nccl_error(int rc, const char* where) {
  State& s = state();
  snprintf(s.cuda_err, sizeof(s.cuda_err), "%s: NCCL error %d (%s)", where, rc,
           nccl().errstr ? nccl().errstr(rc) : "?");
  return IMB_ECUDA;
}

static int load_nccl() {
  Nccl& n = nccl();
  if (n.handle) return IMB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.handle) break;
  }
  if (!n.handle) {
    snprintf(state().cuda_err, sizeof(state().cuda_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
    return IMB_EUNSUPPORTED;
  }
  n.get_id = (fn_get_id)dlsym(n.handle, "ncclGetUniqueId");
  n.init_rank = (fn_init_rank)dlsym(n.handle, "ncclCommInitRank");
  n.destroy = (fn_destroy)dlsym(n.handle, "ncclCommDestroy");
  n.allreduce = (fn_allreduce)dlsym(n.handle, "ncclAllReduce");
  n.errstr = (fn_errstr)dlsym(n.handle, "ncclGetErrorString");
  n.send = (fn_sendrecv)dlsym(n.handle, "ncclSend");
  n.recv = (fn_sendrecv)dlsym(n.handle, "ncclRecv");
  n.group_start = (fn_group)dlsym(n.handle, "ncclGroupStart");
  n.group_end = (fn_group)dlsym(n.handle, "ncclGroupEnd");
  if (!n.get_id || !n.init_rank || !n.destroy || !n.allreduce || !n.send || !n.recv || !n.group_start ||
      !n.group_end) {
    snprintf(state().cuda_err, sizeof(state().cuda_err), "libnccl lacks a required symbol");
    return IMB_EUNSUPPORTED;
  }
  return IMB_OK;
}

// NCCL enum values (nccl.h): ncclFloat64 = 8, ncclSum = 0
static int allreduce_sum(double* buf, size_t count, cudaStream_t s) {
  Nccl& n = nccl();
  if (!n.comm) return IMB_ENOTINIT;
  const int rc = n.allreduce(buf, buf, count, 8, 0, n.comm, s);
  return rc == 0 ? IMB_OK : nccl_error(rc, "ncclAllReduce");
}

}  // namespace imb

using namespace imb;

extern "C" {

int imb_comm_unique_id(void* out128) {
  if (!out128) return IMB_EINVAL;
  IMB_TRY(load_nccl());
  NcclId id;
  const int rc = nccl().get_id(&id);
  if (rc != 0) return nccl_error(rc, "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
  return IMB_OK;
}

int imb_comm_init(int nranks, int rank, const void* unique_id128) {
  IMB_TRY(check_init());
  if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id128) return IMB_EINVAL;
  IMB_TRY(load_nccl());
  Nccl& n = nccl();
  if (n.comm) {
    n.destroy(n.comm);
    n.comm = nullptr;
  }
  NcclId id;
  memcpy(&id, unique_id128, sizeof(id));
  const int rc = n.init_rank(&n.comm, nranks, id, rank);
  if (rc != 0) return nccl_error(rc, "ncclCommInitRank");
  n.nranks = nranks;
  n.rank = rank;
  return IMB_OK;
}

int imb_comm_finalize(void) {
  Nccl& n = nccl();
  if (n.comm) {
    n.destroy(n.comm);
    n.comm = nullptr;
  }
  n.nranks = 0;
  n.rank = -1;
  return IMB_OK;
}

int imb_comm_size(void) { return nccl().comm ? nccl().nranks : 0; }
int imb_comm_rank(void) { return nccl().comm ? nccl().rank : -1; }

int imb_allreduce_sum(double* device_buf, int64_t count, void* stream) {
  IMB_TRY(check_init());
  if (!device_buf || count < 0) return IMB_EINVAL;
  if (count == 0) return IMB_OK;
  return allreduce_sum(device_buf, (size_t)count, (cudaStream_t)stream);
}

int imb_alltoall(const double* send, double* recv, int64_t count_per_rank, void* stream) {
  IMB_TRY(check_init());
  Nccl& n = nccl();
  if (!n.comm) return IMB_ENOTINIT;
  if (!send || !recv || count_per_rank < 0) return IMB_EINVAL;
  if (count_per_rank == 0) return IMB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = n.group_start();
  if (rc != 0) return nccl_error(rc, "ncclGroupStart");
  for (int r = 0; r < n.nranks && rc == 0; ++r) {
    rc = n.send(send + (size_t)r * count_per_rank, (size_t)count_per_rank, 8, r, n.comm, s);
    if (rc == 0) rc = n.recv(recv + (size_t)r * count_per_rank, (size_t)count_per_rank, 8, r, n.comm, s);
  }
  const int rc_end = n.group_end();
  if (rc != 0) return nccl_error(rc, "ncclSend/ncclRecv");
  return rc_end == 0 ? IMB_OK : nccl_error(rc_end, "ncclGroupEnd");
}

int imb_loglike_ensemble_sharded(int K, const double* const* sims, int N, int m_local,
                                 int rows_per_block, int64_t block_stride, int64_t row_stride,
                                 const double* const* meas, const double* const* var, int64_t m_total,
                                 int estimator, double* G, double* P2, double* out_loglike,
                                 int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  if (!nccl().comm) return IMB_ENOTINIT;
  if (K < 1 || N < 1 || !G || !P2) return IMB_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  IMB_TRY(imb_gram_phase1(K, sims, N, m_local, rows_per_block, block_stride, row_stride, meas, G, stream));
  IMB_TRY(allreduce_sum(G, (size_t)K * N * N, s));
  IMB_TRY(imb_gram_phase2(K, sims, N, m_local, rows_per_block, block_stride, row_stride, meas, var, m_total,
                          estimator, G, P2, stream));
  IMB_TRY(allreduce_sum(P2, (size_t)K * ((size_t)N * N + N + 4), s));
  return imb_gram_phase3(K, N, m_local, G, P2, out_loglike, out_info, stream);
}

}  // extern "C"
