// Row-block collectives of the row-partitioned propagation behind the C ABI:
//   fr_allgather_rows       X_full[world * R, d] <- all ranks' X_local[R, d]      (the one exchange of a layer)
//   fr_reduce_scatter_rows  dX_local[R, d]      <- sum over ranks of G_full block  (its adjoint, for the batch gathers)
// NCCL is bound at RUN time: the process that loads this library has normally loaded `libnccl.so.2` already (PyTorch
// links it), so the entry points are looked up with dlopen/dlsym and the shared object carries no link-time NCCL
// dependency -- a single-GPU user never touches NCCL.  Host code only; the collectives run on the caller's stream.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

// The part of NCCL's public C API used here (nccl.h; ABI-stable across 2.x).
struct UniqueId { char internal[128]; };            // ncclUniqueId
typedef struct ncclComm *Comm;                      // ncclComm_t
constexpr int kNcclFloat = 7;                       // ncclFloat32
constexpr int kNcclSum = 0;                         // ncclSum

struct Api {
    void *handle = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Api &api() {
    static Api a;
    if (a.ok || a.handle) return a;
    const char *override_path = getenv("FR_NCCL_LIBRARY");
    if (override_path && *override_path) a.handle = dlopen(override_path, RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy already in the process
    if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
#define FR_SYM(field, name) *reinterpret_cast<void **>(&a.field) = dlsym(a.handle, name)
    FR_SYM(GetVersion, "ncclGetVersion");
    FR_SYM(GetUniqueId, "ncclGetUniqueId");
    FR_SYM(CommInitRank, "ncclCommInitRank");
    FR_SYM(CommDestroy, "ncclCommDestroy");
    FR_SYM(AllGather, "ncclAllGather");
    FR_SYM(ReduceScatter, "ncclReduceScatter");
    FR_SYM(GetErrorString, "ncclGetErrorString");
#undef FR_SYM
    a.ok = a.GetVersion && a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.ReduceScatter &&
           a.GetErrorString;
    return a;
}

int need_api(const char *what) {
    if (api().ok) return FR_OK;
    fr::set_error("%s: NCCL is not available (libnccl.so.2 not found; set FR_NCCL_LIBRARY): %s", what,
                  api().handle ? "missing symbols" : dlerror());
    return FR_EINVAL;
}

int check_nccl(int rc, const char *what) {
    if (rc == 0) return FR_OK;
    fr::set_error("%s: NCCL error %d: %s", what, rc, api().GetErrorString(rc));
    return FR_ECUDA;
}

}  // namespace

extern "C" int fr_comm_version(void) {
    int v = 0;
    if (!api().ok || api().GetVersion(&v) != 0) return 0;
    return v;
}

extern "C" int fr_comm_unique_id(void *id128) {
    FR_REQUIRE(id128, "fr_comm_unique_id: null pointer");
    if (int rc = need_api("fr_comm_unique_id")) return rc;
    UniqueId id;
    if (int rc = check_nccl(api().GetUniqueId(&id), "fr_comm_unique_id")) return rc;
    memcpy(id128, &id, sizeof id);
    return FR_OK;
}

extern "C" int fr_comm_init(const void *id128, int32_t rank, int32_t world, void **comm) {
    FR_REQUIRE(id128 && comm, "fr_comm_init: null pointer");
    FR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "fr_comm_init: rank=%d world=%d", rank, world);
    if (int rc = need_api("fr_comm_init")) return rc;
    UniqueId id;
    memcpy(&id, id128, sizeof id);
    Comm c = nullptr;
    if (int rc = check_nccl(api().CommInitRank(&c, world, id, rank), "fr_comm_init")) return rc;
    *comm = c;
    return FR_OK;
}

extern "C" int fr_comm_destroy(void *comm) {
    if (!comm) return FR_OK;
    if (int rc = need_api("fr_comm_destroy")) return rc;
    return check_nccl(api().CommDestroy((Comm)comm), "fr_comm_destroy");
}

extern "C" int fr_allgather_rows(void *comm, const float *x_local, int64_t rows_per_rank, int32_t d, float *x_full,
                                 void *stream) {
    FR_REQUIRE(comm && x_local && x_full, "fr_allgather_rows: null pointer");
    FR_REQUIRE(rows_per_rank > 0 && d > 0, "fr_allgather_rows: rows_per_rank=%lld d=%d", (long long)rows_per_rank, d);
    if (int rc = need_api("fr_allgather_rows")) return rc;
    return check_nccl(api().AllGather(x_local, x_full, (size_t)rows_per_rank * d, kNcclFloat, (Comm)comm,
                                      (cudaStream_t)stream), "fr_allgather_rows");
}

extern "C" int fr_reduce_scatter_rows(void *comm, const float *g_full, int64_t rows_per_rank, int32_t d, float *g_local,
                                      void *stream) {
    FR_REQUIRE(comm && g_full && g_local, "fr_reduce_scatter_rows: null pointer");
    FR_REQUIRE(rows_per_rank > 0 && d > 0, "fr_reduce_scatter_rows: rows_per_rank=%lld d=%d", (long long)rows_per_rank, d);
    if (int rc = need_api("fr_reduce_scatter_rows")) return rc;
    return check_nccl(api().ReduceScatter(g_full, g_local, (size_t)rows_per_rank * d, kNcclFloat, kNcclSum, (Comm)comm,
                                          (cudaStream_t)stream), "fr_reduce_scatter_rows");
}
