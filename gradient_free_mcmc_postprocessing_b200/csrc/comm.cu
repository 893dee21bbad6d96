// comm.cu - NCCL plumbing for the two exchanges the path has:
//   thinning : per greedy step, all-gather of R x {f64 value, u64 index} = 16 B per rank
//   KSD      : one sum all-reduce of the row-sum vectors
// NCCL is resolved with dlopen at gfk_comm_init time, so the library loads (and the single-GPU
// path works) on machines without NCCL, and a process that already loaded the torch-bundled
// libnccl.so.2 shares that copy.  Only the handful of prototypes used are declared here.
#include <dlfcn.h>
#include <string.h>

#include "comm.h"
#include "common.cuh"

namespace gfk {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_api;
static ncclComm_t g_comm = nullptr;
static int g_size = 1, g_rank = 0;

static int load_nccl(const char* path) {
    if (g_api.handle) return GFK_OK;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        return GFK_ENCCL;
    }
    g_api.handle = h;
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_api.AllGather = reinterpret_cast<decltype(g_api.AllGather)>(dlsym(h, "ncclAllGather"));
    g_api.AllReduce = reinterpret_cast<decltype(g_api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.CommDestroy || !g_api.AllGather ||
        !g_api.AllReduce || !g_api.GetErrorString) {
        set_error("NCCL library is missing required symbols");
        g_api = NcclApi();
        return GFK_ENCCL;
    }
    return GFK_OK;
}

static int nccl_fail(int rc, const char* what) {
    set_error("%s failed: %s", what, g_api.GetErrorString ? g_api.GetErrorString(rc) : "?");
    return GFK_ENCCL;
}

// ---- peer mailboxes ---------------------------------------------------------------------------------
static unsigned long long* g_mail_local = nullptr;
static unsigned long long* g_mail_peer[kMailRanks] = {nullptr};
static bool g_mail_open = false;
static unsigned g_tag_epoch = 1;
static unsigned g_job_count = 0;

void p2p_next_job(MailArgs* out, int steps) {
    memset(out, 0, sizeof(*out));
    if (!g_mail_open) return;
    out->enabled = 1;
    out->local = g_mail_local;
    for (int r = 0; r < kMailRanks; ++r) out->peer[r] = g_mail_peer[r];
    // every rank makes the same sequence of jobs, so the epochs agree without communication
    if (g_tag_epoch > 0xf0000000u - (unsigned)steps - 2u) g_tag_epoch = 1;   // wrap long before 2^32
    out->tag_base = g_tag_epoch;
    g_tag_epoch += (unsigned)steps + 2u;
    out->job_parity = (int)(g_job_count++ & 1u);
    out->rank = g_rank;
}

int comm_size() { return g_size; }
int comm_rank() { return g_rank; }

int comm_allgather16(const void* send, void* recv, cudaStream_t st) {
    if (!g_comm) { set_error("communicator not initialised"); return GFK_ENCCL; }
    const int rc = g_api.AllGather(send, recv, 16, ncclChar, g_comm, st);
    return rc == ncclSuccess ? GFK_OK : nccl_fail(rc, "ncclAllGather");
}

}  // namespace gfk

using namespace gfk;

extern "C" int gfk_comm_unique_id(void* id128_host, const char* nccl_path) {
    GFK_REQUIRE(id128_host, "gfk_comm_unique_id: NULL output");
    int rc = load_nccl(nccl_path);
    if (rc != GFK_OK) return rc;
    ncclUniqueId id;
    rc = g_api.GetUniqueId(&id);
    if (rc != ncclSuccess) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id128_host, &id, sizeof(id));
    return GFK_OK;
}

extern "C" int gfk_comm_init(int n_ranks, int rank, const void* id128_host, const char* nccl_path) {
    GFK_REQUIRE(n_ranks >= 1 && n_ranks <= 16, "gfk_comm_init: n_ranks = %d out of range [1, 16]", n_ranks);
    GFK_REQUIRE(rank >= 0 && rank < n_ranks, "gfk_comm_init: rank = %d out of range", rank);
    GFK_REQUIRE(!g_comm, "gfk_comm_init: communicator already initialised");
    if (n_ranks == 1) { g_size = 1; g_rank = 0; return GFK_OK; }
    GFK_REQUIRE(id128_host, "gfk_comm_init: NULL unique id");
    int rc = load_nccl(nccl_path);
    if (rc != GFK_OK) return rc;
    ncclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    rc = g_api.CommInitRank(&g_comm, n_ranks, id, rank);
    if (rc != ncclSuccess) { g_comm = nullptr; return nccl_fail(rc, "ncclCommInitRank"); }
    g_size = n_ranks;
    g_rank = rank;
    return GFK_OK;
}

extern "C" int gfk_comm_size(void) { return g_size; }
extern "C" int gfk_comm_rank(void) { return g_rank; }

extern "C" int gfk_allreduce_sum(double* buf, int64_t count, void* stream) {
    GFK_REQUIRE(buf && count >= 0, "gfk_allreduce_sum: bad arguments");
    if (g_size == 1 || count == 0) return GFK_OK;
    if (!g_comm) { set_error("communicator not initialised"); return GFK_ENCCL; }
    const int rc = g_api.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, g_comm,
                                   static_cast<cudaStream_t>(stream));
    return rc == ncclSuccess ? GFK_OK : nccl_fail(rc, "ncclAllReduce");
}

extern "C" int gfk_p2p_local_handle(void* handle64_host) {
    GFK_REQUIRE(handle64_host, "gfk_p2p_local_handle: NULL output");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!g_mail_local) {
        GFK_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_mail_local), sizeof(unsigned long long) * kMailWords));
        GFK_CUDA(cudaMemset(g_mail_local, 0, sizeof(unsigned long long) * kMailWords));
        GFK_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    GFK_CUDA(cudaIpcGetMemHandle(&h, g_mail_local));
    memcpy(handle64_host, &h, sizeof(h));
    return GFK_OK;
}

extern "C" int gfk_p2p_open(const void* handles_host, int n_ranks, int rank) {
    GFK_REQUIRE(handles_host && n_ranks >= 2 && n_ranks <= kMailRanks && rank >= 0 && rank < n_ranks,
                "gfk_p2p_open: bad arguments");
    GFK_REQUIRE(g_mail_local, "gfk_p2p_open: call gfk_p2p_local_handle first");
    GFK_REQUIRE(n_ranks == g_size && rank == g_rank, "gfk_p2p_open: communicator mismatch (call gfk_comm_init first)");
    const char* base = static_cast<const char*>(handles_host);
    for (int r = 0; r < n_ranks; ++r) {
        if (r == rank) { g_mail_peer[r] = g_mail_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, base + 64 * r, sizeof(h));
        void* ptr = nullptr;
        GFK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        g_mail_peer[r] = static_cast<unsigned long long*>(ptr);
    }
    g_mail_open = true;
    g_tag_epoch = 1;
    g_job_count = 0;
    return GFK_OK;
}

extern "C" int gfk_p2p_enabled(void) { return g_mail_open ? 1 : 0; }

#if defined(GFK_SIMT_HOST)
// emulator build of the test suite only: this rank's mailbox and the tag base / parity the NEXT job will use, so a
// test can plant a forged or stale record (tests/test_simt_fault_protocol.py)
extern "C" unsigned long long* gfk_test_mailbox(unsigned* next_tag_base, int* next_job_parity) {
    if (next_tag_base) *next_tag_base = g_tag_epoch;
    if (next_job_parity) *next_job_parity = (int)(g_job_count & 1u);
    return g_mail_local;
}
#endif

extern "C" int gfk_comm_finalize(void) {
    if (g_mail_open) {
        for (int r = 0; r < kMailRanks; ++r) {
            if (g_mail_peer[r] && g_mail_peer[r] != g_mail_local) cudaIpcCloseMemHandle(g_mail_peer[r]);
            g_mail_peer[r] = nullptr;
        }
        g_mail_open = false;
    }
    if (g_mail_local) {
        cudaFree(g_mail_local);
        g_mail_local = nullptr;
    }
    if (g_comm) {
        g_api.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_size = 1;
    g_rank = 0;
    return GFK_OK;
}
