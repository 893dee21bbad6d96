// fake_nccl.cpp - the six NCCL entry points libgfk resolves with dlopen (csrc/comm.cu), implemented over POSIX
// shared memory between emulated ranks (one process per rank).  TEST INFRASTRUCTURE ONLY: lets the CPU suite run the
// multi-rank code paths of the emulator build (tests/host/libgfk_sim.so) - communicator setup, the all-gather
// exchange fallback and the KSD all-reduce - without GPUs.
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

namespace {
constexpr int kMaxRanks = 16;
constexpr size_t kSlotBytes = 1u << 20;
struct Shared {
    volatile unsigned arrived;
    volatile unsigned generation;
    char pad[56];
    char slot[kMaxRanks][kSlotBytes];
};
struct Comm {
    Shared* sh;
    int n, rank;
    char name[128];
};

void barrier(Comm* c) {
    const unsigned gen = c->sh->generation;
    if (__atomic_add_fetch(&c->sh->arrived, 1u, __ATOMIC_SEQ_CST) == (unsigned)c->n) {
        c->sh->arrived = 0;
        __atomic_add_fetch(&c->sh->generation, 1u, __ATOMIC_SEQ_CST);
    } else {
        while (__atomic_load_n(&c->sh->generation, __ATOMIC_SEQ_CST) == gen) sched_yield();
    }
}
size_t elem_size(int dtype) { return dtype == 8 ? 8 : 1; }   // ncclFloat64 = 8, ncclChar = 0
}  // namespace

extern "C" {

struct ncclUniqueId { char internal[128]; };

int ncclGetUniqueId(ncclUniqueId* id) {
    memset(id, 0, sizeof(*id));
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(id->internal, sizeof(id->internal), "/fake_nccl_%d_%ld", (int)getpid(), (long)ts.tv_nsec);
    return 0;
}

int ncclCommInitRank(Comm** out, int n, ncclUniqueId id, int rank) {
    if (n < 1 || n > kMaxRanks) return 4;
    Comm* c = new Comm;
    c->n = n;
    c->rank = rank;
    snprintf(c->name, sizeof(c->name), "%s", id.internal);
    int fd = -1;
    for (int tries = 0; tries < 20000 && fd < 0; ++tries) {
        fd = rank == 0 ? shm_open(c->name, O_CREAT | O_RDWR, 0600) : shm_open(c->name, O_RDWR, 0600);
        if (fd < 0) usleep(500);
    }
    if (fd < 0) { delete c; return 2; }
    if (rank == 0 && ftruncate(fd, (off_t)sizeof(Shared)) != 0) { close(fd); delete c; return 2; }
    struct stat st;
    for (int tries = 0; tries < 20000; ++tries) {          // wait until rank 0 has sized the segment
        if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(Shared)) break;
        usleep(500);
    }
    void* p = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { delete c; return 2; }
    c->sh = static_cast<Shared*>(p);
    barrier(c);
    if (rank == 0) shm_unlink(c->name);                     // everyone has it mapped
    *out = c;
    return 0;
}

int ncclCommDestroy(Comm* c) {
    if (c) {
        munmap(c->sh, sizeof(Shared));
        delete c;
    }
    return 0;
}

int ncclAllGather(const void* send, void* recv, size_t count, int dtype, Comm* c, void*) {
    const size_t bytes = count * elem_size(dtype);
    if (bytes > kSlotBytes) return 4;
    memcpy(c->sh->slot[c->rank], send, bytes);
    barrier(c);
    for (int r = 0; r < c->n; ++r) memcpy(static_cast<char*>(recv) + (size_t)r * bytes, c->sh->slot[r], bytes);
    barrier(c);
    return 0;
}

int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, Comm* c, void*) {
    if (dtype != 8 || op != 0 || count * 8 > kSlotBytes) return 4;   // float64 sum only
    memcpy(c->sh->slot[c->rank], send, count * 8);
    barrier(c);
    double* out = static_cast<double*>(recv);
    for (size_t e = 0; e < count; ++e) {
        double acc = 0.0;
        for (int r = 0; r < c->n; ++r) acc += reinterpret_cast<const double*>(c->sh->slot[r])[e];
        out[e] = acc;
    }
    barrier(c);
    return 0;
}

const char* ncclGetErrorString(int rc) { return rc == 0 ? "no error" : "fake NCCL error"; }
}
