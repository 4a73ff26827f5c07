// npgp communicator: the ONE collective of the data-parallel ELBO step (sum all-reduce of the flat fp64 gradient buffer,
// SURVEY.md section 8(e)) behind the C ABI.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy PyTorch has already
// loaded when the caller is a torch process, the system library otherwise), so libnpgp.so itself has no link-time
// dependency on it and a single-GPU caller never touches it.  The reference has no distributed code at all.
#include <dlfcn.h>
#include <cstring>

#include "common.cuh"
#include "../../include/npgp.h"

namespace {

struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_group)(void);

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
};

// resolved once per process (read-only afterwards)
const NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (PyTorch's bundled copy)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.handle = h;
      api.get_unique_id = reinterpret_cast<fn_get_unique_id>(dlsym(h, "ncclGetUniqueId"));
      api.comm_init_rank = reinterpret_cast<fn_comm_init_rank>(dlsym(h, "ncclCommInitRank"));
      api.comm_destroy = reinterpret_cast<fn_comm_destroy>(dlsym(h, "ncclCommDestroy"));
      api.all_reduce = reinterpret_cast<fn_all_reduce>(dlsym(h, "ncclAllReduce"));
      api.group_start = reinterpret_cast<fn_group>(dlsym(h, "ncclGroupStart"));
      api.group_end = reinterpret_cast<fn_group>(dlsym(h, "ncclGroupEnd"));
    }
  }
  return (api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce && api.group_start && api.group_end)
             ? &api
             : nullptr;
}

constexpr int kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace

/* 128-byte rendezvous token: create on rank 0, ship to every rank by any means (file, socket, torch.distributed) */
extern "C" int npgp_comm_unique_id(void* id128) {
  if (!id128) return NPGP_EINVAL;
  const NcclApi* api = nccl_api();
  if (!api) return NPGP_EUNSUPPORTED;
  return api->get_unique_id(static_cast<NcclId*>(id128)) == 0 ? NPGP_OK : NPGP_EINVAL;
}

/* collective over all ranks; binds the communicator to the calling thread's current CUDA device */
extern "C" int npgp_comm_create(void** comm, const void* id128, int nranks, int rank) {
  if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return NPGP_EINVAL;
  const NcclApi* api = nccl_api();
  if (!api) return NPGP_EUNSUPPORTED;
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  if (api->comm_init_rank(&c, nranks, id, rank) != 0) return NPGP_EINVAL;
  *comm = c;
  return NPGP_OK;
}

extern "C" int npgp_comm_destroy(void* comm) {
  if (!comm) return NPGP_OK;
  const NcclApi* api = nccl_api();
  if (!api) return NPGP_EUNSUPPORTED;
  return api->comm_destroy(comm) == 0 ? NPGP_OK : NPGP_EINVAL;
}

/* buf[0 .. n) <- sum over ranks, in place, on `stream` (capturable) */
extern "C" int npgp_allreduce_f64(void* comm, double* buf, long n, cudaStream_t stream) {
  if (!comm || n < 0 || (n > 0 && !buf)) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  const NcclApi* api = nccl_api();
  if (!api) return NPGP_EUNSUPPORTED;
  return api->all_reduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, comm, stream) == 0 ? NPGP_OK : NPGP_EINVAL;
}

/* two disjoint pieces in ONE grouped launch (the head and the tail of the flat gradient buffer around the part that was
 * reduced early, csrc/svgp_step.cu); either piece may be empty */
extern "C" int npgp_allreduce_f64_pair(void* comm, double* buf1, long n1, double* buf2, long n2, cudaStream_t stream) {
  if (!comm || n1 < 0 || n2 < 0 || (n1 > 0 && !buf1) || (n2 > 0 && !buf2)) return NPGP_EINVAL;
  if (n1 == 0) return npgp_allreduce_f64(comm, buf2, n2, stream);
  if (n2 == 0) return npgp_allreduce_f64(comm, buf1, n1, stream);
  const NcclApi* api = nccl_api();
  if (!api) return NPGP_EUNSUPPORTED;
  int rc = api->group_start();
  if (rc == 0) rc = api->all_reduce(buf1, buf1, (size_t)n1, kNcclFloat64, kNcclSum, comm, stream);
  if (rc == 0) rc = api->all_reduce(buf2, buf2, (size_t)n2, kNcclFloat64, kNcclSum, comm, stream);
  const int rc2 = api->group_end();
  return (rc == 0 && rc2 == 0) ? NPGP_OK : NPGP_EINVAL;
}
