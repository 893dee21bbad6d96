// simt.cpp - fiber scheduler and fake CUDA runtime of the SIMT emulator (see simt.h).  TEST INFRASTRUCTURE ONLY.
#include <time.h>
#include "simt.h"

#include <fcntl.h>
#include <sched.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "the SIMT emulator's context switch is written for x86-64"
#endif

// callee-saved registers + stack pointer; everything else is dead across a call by the ABI
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.hidden simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

namespace simt {

constexpr size_t kStackBytes = 64 * 1024;
enum Wait { W_NONE = 0, W_BLOCK = 1, W_WARP = 2, W_NAMED = 3 };

struct NamedBarrier {
    unsigned gen = 0;
    int arrived = 0;
    int count = 0;
};

struct WarpState {
    unsigned long long buf[2][32];
    unsigned mask[2] = {0, 0};   // lanes that took part in the collective held in buf[.]
    unsigned gen = 0;
    int arrived = 0;
    int alive = 0;
    unsigned alive_mask = 0;
};

struct Fiber {
    ThreadCtx ctx;
    void* sp = nullptr;
    int wait = W_NONE;
    unsigned wait_gen = 0;
    bool done = false;
    int warp = 0, lane = 0;
    unsigned last_mask = 0;
    int wait_id = 0;
    struct Copy { void* dst; const void* src; int bytes; };
    std::vector<std::vector<Copy>> groups;    // committed cp.async groups, oldest first
    std::vector<Copy> open_group;             // copies issued since the last commit
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    NamedBarrier named[16];
    unsigned gen = 0;
    int arrived = 0;
    int alive = 0;
    void* sched_sp = nullptr;
    Fiber* running = nullptr;
    char* smem = nullptr;
    char* smem_window = nullptr;        // what shared-window addresses (descriptors) are relative to
    const std::function<void()>* fn = nullptr;
    bool failed = false;
    unsigned polls = 0;
    struct PendingTma { void* dst; TensorMap2D map; int c0, c1; unsigned long long* bar; };
    std::vector<PendingTma> tma;        // bulk tensor copies issued and not yet completed
    bool fault_injected = false;        // tests lost a copy on purpose: leftovers at exit are expected
    unsigned long long mbar_polls = 0;  // consecutive unsuccessful mbarrier tests (a wait that can never complete)
    std::vector<unsigned> tmem;         // 128 lanes x 512 columns once allocated
    int tmem_cols = 0;                  // allocated columns (0: none)
    const char* site_file = nullptr;   // call site of the __syncthreads being collected
    int site_line = 0;
};

static thread_local Block* t_blk = nullptr;

// ---- errors ------------------------------------------------------------------------------------------------
static std::mutex g_err_mu;
static char g_err_msg[1024] = "";
static std::atomic<int> g_last_error{0};
static std::atomic<bool> g_abort{false};
constexpr int kSimtLaunchFailure = 719;

static void record_error(const char* msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    if (!g_last_error.load()) {
        snprintf(g_err_msg, sizeof(g_err_msg), "%s", msg);
        g_last_error.store(kSimtLaunchFailure);
    }
    if (getenv("SIMT_VERBOSE")) fprintf(stderr, "[simt] %s\n", msg);
}

[[noreturn]] void fail(const char* fmt, ...) {
    char buf[900];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    Block* b = t_blk;
    if (!b || !b->running) {
        fprintf(stderr, "[simt] fatal outside a kernel: %s\n", buf);
        abort();
    }
    char where[1024];
    const ThreadCtx& c = b->running->ctx;
    snprintf(where, sizeof(where), "%s (block %u, thread %u)", buf, c.bid.x, c.tid.x);
    record_error(where);
    g_abort.store(true);
    b->failed = true;
    simt_switch(&b->running->sp, b->sched_sp);   // the scheduler abandons this block
    abort();
}

void check_align(const void* p, size_t a, const char* what) {
    if (a > 1 && (reinterpret_cast<uintptr_t>(p) % a) != 0) fail("misaligned address in %s: %p needs %zu-byte alignment", what, p, a);
}

ThreadCtx* cur() { return &t_blk->running->ctx; }
void* dyn_smem() { return t_blk->smem; }

static inline void yield_to_scheduler(Block* b, Fiber* f) { simt_switch(&f->sp, b->sched_sp); }

static int g_random_sched = -1;
static unsigned g_seed = 1;
static bool random_sched() {
    if (g_random_sched < 0) {
        const char* e = getenv("SIMT_SCHED");
        g_random_sched = (e && strncmp(e, "random", 6) == 0) ? 1 : 0;
        if (g_random_sched && e[6] == ':') g_seed = (unsigned)strtoul(e + 7, nullptr, 10);
    }
    return g_random_sched == 1;
}

void syncthreads(const char* file, int line) {
    Block* b = t_blk;
    Fiber* f = b->running;
    const unsigned g = b->gen;
    if (b->arrived == 0) {
        b->site_file = file;
        b->site_line = line;
    } else if (b->site_line != line || b->site_file != file) {
        fail("divergent __syncthreads: threads of one block wait at %s:%d and at %s:%d", b->site_file, b->site_line,
             file, line);
    }
    if (++b->arrived == b->alive) {
        b->arrived = 0;
        b->gen++;
        if (random_sched()) { f->wait = W_NONE; yield_to_scheduler(b, f); }
        return;
    }
    f->wait = W_BLOCK;
    f->wait_gen = g;
    yield_to_scheduler(b, f);
}

void named_barrier(int id, int count, bool wait) {
    Block* b = t_blk;
    Fiber* f = b->running;
    if (id < 1 || id > 15 || count <= 0 || count % 32 != 0) fail("named barrier %d with thread count %d", id, count);
    NamedBarrier& nb = b->named[id];
    if (nb.arrived == 0) nb.count = count;
    else if (nb.count != count) fail("named barrier %d used with thread counts %d and %d", id, nb.count, count);
    const unsigned g = nb.gen;
    if (++nb.arrived == nb.count) {
        nb.arrived = 0;
        nb.gen++;
        if (random_sched()) { f->wait = W_NONE; yield_to_scheduler(b, f); }
        return;
    }
    if (nb.arrived > nb.count) fail("named barrier %d: more arrivals than its thread count %d", id, count);
    if (!wait) {
        if (random_sched()) { f->wait = W_NONE; yield_to_scheduler(b, f); }
        return;
    }
    f->wait = W_NAMED;
    f->wait_id = id;
    f->wait_gen = g;
    yield_to_scheduler(b, f);
}

const unsigned long long* warp_gather(unsigned long long bits) {
    Block* b = t_blk;
    Fiber* f = b->running;
    WarpState& w = b->warps[f->warp];
    const unsigned g = w.gen;
    w.buf[g & 1][f->lane] = bits;
    if (++w.arrived == w.alive) {
        w.arrived = 0;
        w.mask[g & 1] = w.alive_mask;
        w.gen++;
        if (random_sched()) { f->wait = W_NONE; yield_to_scheduler(b, f); }
    } else {
        f->wait = W_WARP;
        f->wait_gen = g;
        yield_to_scheduler(b, f);
    }
    f->last_mask = w.mask[g & 1];
    return w.buf[g & 1];
}

// the lanes that took part in this thread's most recent collective (lanes that exit later must still count)
unsigned warp_alive_mask() { return t_blk->running->last_mask; }

// all 32 lanes of the warp must execute the instruction together with the same operand (.sync.aligned)
static void warp_converged(unsigned operand, const char* what) {
    const unsigned long long* all = warp_gather(operand);
    Block* b = t_blk;
    if (warp_alive_mask() != 0xffffffffu) fail("%s by a warp with exited lanes", what);
    for (int l = 0; l < 32; ++l)
        if ((unsigned)all[l] != operand) fail("%s with lane-dependent operand (%u vs %u)", what, (unsigned)all[l], operand);
    (void)b;
}

void tmem_alloc(unsigned* dst_smem, int cols) {
    warp_converged((unsigned)cols, "tcgen05.alloc");
    Block* b = t_blk;
    if (cols < 32 || cols > 512 || (cols & (cols - 1))) fail("tcgen05.alloc of %d columns", cols);
    if (b->running->lane == 0) {
        if (b->tmem_cols) fail("tcgen05.alloc while tensor memory is already allocated");
        b->tmem.assign(128 * 512, 0xdeadbeefu);
        b->tmem_cols = cols;
        *dst_smem = 0u;                                    // base address: lane 0, column 0
    }
}

void tmem_dealloc(unsigned base, int cols) {
    warp_converged(base, "tcgen05.dealloc");
    Block* b = t_blk;
    if (b->running->lane == 0) {
        if (!b->tmem_cols || cols != b->tmem_cols || base != 0u) fail("tcgen05.dealloc does not match the allocation");
        b->tmem_cols = 0;
    }
}

static std::atomic<long long> g_tmem_accesses{0};

static unsigned* tmem_row(unsigned taddr, int n, const char* what) {
    warp_converged(taddr, what);
    g_tmem_accesses.fetch_add(1, std::memory_order_relaxed);
    Block* b = t_blk;
    Fiber* f = b->running;
    const unsigned lane_base = taddr >> 16, col = taddr & 0xffffu;
    if (!b->tmem_cols) fail("%s without a tensor-memory allocation", what);
    if (lane_base != (unsigned)(f->warp % 4) * 32u) fail("%s: warp %d may not access lanes from %u", what, f->warp, lane_base);
    if (col + (unsigned)n > (unsigned)b->tmem_cols) fail("%s: columns %u..%u beyond the %d allocated", what, col, col + n, b->tmem_cols);
    return b->tmem.data() + (size_t)(lane_base + (unsigned)f->lane) * 512 + col;
}

void tmem_st(unsigned taddr, const unsigned* regs, int n) {
    unsigned* row = tmem_row(taddr, n, "tcgen05.st");
    for (int c = 0; c < n; ++c) row[c] = regs[c];
}

void tmem_ld(unsigned taddr, unsigned* regs, int n) {
    const unsigned* row = tmem_row(taddr, n, "tcgen05.ld");
    for (int c = 0; c < n; ++c) regs[c] = row[c];
}

void cp_async(void* dst, const void* src, int bytes) {
    Fiber* f = t_blk->running;
    check_align(dst, (size_t)bytes, "cp.async destination");
    check_align(src, (size_t)bytes, "cp.async source");
    f->open_group.push_back(Fiber::Copy{dst, src, bytes});
}

void cp_async_commit() {
    Fiber* f = t_blk->running;
    f->groups.push_back(std::move(f->open_group));
    f->open_group.clear();
}

static void run_copies(std::vector<Fiber::Copy>& g) {
    for (const Fiber::Copy& c : g) memcpy(c.dst, c.src, (size_t)c.bytes);
    g.clear();
}

void cp_async_wait(int allow_pending_groups) {
    Fiber* f = t_blk->running;
    while ((int)f->groups.size() > allow_pending_groups) {
        run_copies(f->groups.front());
        f->groups.erase(f->groups.begin());
    }
}

static std::atomic<long long> g_tma_loads{0};
static std::atomic<long long> g_umma_issued{0};
static std::atomic<long long> g_tma_drop{-1};     // index (in issue order) of a bulk copy to lose, tests only

// shared-window address of a pointer into dynamic shared memory (what cvta.to.shared gives on the device)
unsigned smem_u32(const void* p) {
    Block* b = t_blk;
    const char* c = static_cast<const char*>(p);
    if (!b->smem_window || c < b->smem || c >= b->smem_window + 256 * 1024) fail("shared-window address of a pointer outside dynamic shared memory");
    return (unsigned)(c - b->smem_window);
}

// ---- tcgen05.mma.cta_group::1.kind::tf32, both operands from shared memory, K-major, SWIZZLE_128B ---------------
// D[lane r][column c0 + j] (+)= sum_k A[r][k] * B[j][k],  r < M, j < N, k < 8 (one instruction = 32 bytes of K).
// The shared-memory descriptors are decoded as cute::UMMA::SmemDescriptor lays them out: start address and stride
// byte offset in 16-byte units, version 1, layout type 2 = SWIZZLE_128B; an element's address is
// start + (row / 8) * SBO + (row % 8) * 128 + 4 k with address bits [4,7) XOR-ed by bits [7,10).  Operands are
// read as tf32 (the low 13 mantissa bits do not take part); products are accumulated in float in k order.
static float tf32_of(unsigned bits) {
    bits &= 0xffffe000u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
void umma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b, unsigned idesc, bool accumulate) {
    Block* b = t_blk;
    g_umma_issued.fetch_add(1, std::memory_order_relaxed);
    if (!b->tmem_cols) fail("tcgen05.mma without a tensor-memory allocation");
    const unsigned c_format = (idesc >> 4) & 3u, a_format = (idesc >> 7) & 7u, b_format = (idesc >> 10) & 7u;
    const unsigned a_major = (idesc >> 15) & 1u, b_major = (idesc >> 16) & 1u;
    const unsigned N = ((idesc >> 17) & 0x3fu) << 3, M = ((idesc >> 24) & 0x1fu) << 4;
    if (c_format != 1 || a_format != 2 || b_format != 2) fail("tcgen05.mma kind::tf32: instruction descriptor formats %u/%u/%u", c_format, a_format, b_format);
    if (a_major != 0 || b_major != 0) fail("tcgen05.mma: the emulator models K-major operands only");
    if (M != 128 || N < 8 || N > 256 || N % 8) fail("tcgen05.mma: shape M = %u, N = %u", M, N);
    if ((idesc & 0xfu) || (idesc >> 30)) fail("tcgen05.mma: sparse / saturate / shift bits set in the instruction descriptor");
    const unsigned lane0 = tmem_d >> 16, col0 = tmem_d & 0xffffu;
    if (lane0 != 0 || col0 + N > (unsigned)b->tmem_cols) fail("tcgen05.mma: accumulator columns %u..%u beyond the allocation", col0, col0 + N);
    auto operand = [&](unsigned long long desc, unsigned rows, const char* which, std::vector<float>& out) {
        const unsigned start = (unsigned)(desc & 0x3fffu) << 4, sbo = (unsigned)((desc >> 32) & 0x3fffu) << 4;
        const unsigned version = (unsigned)((desc >> 46) & 3u), layout = (unsigned)((desc >> 61) & 7u);
        const unsigned base_offset = (unsigned)((desc >> 49) & 7u);
        if (version != 1) fail("tcgen05.mma: %s descriptor version %u (Blackwell: 1)", which, version);
        if (layout != 2) fail("tcgen05.mma: %s descriptor layout type %u (the emulator models SWIZZLE_128B = 2)", which, layout);
        if (base_offset != 0) fail("tcgen05.mma: %s descriptor base offset %u", which, base_offset);
        if (sbo != 1024) fail("tcgen05.mma: %s descriptor stride byte offset %u (8 rows of 128 bytes = 1024)", which, sbo);
        if ((start & 0x7fu) % 32 != 0 || (start & 0x7fu) + 32 > 128) fail("tcgen05.mma: %s K offset inside the swizzle atom", which);
        if (((start & ~0x7fu) % 1024) != 0) fail("tcgen05.mma: %s tile is not 1024-byte aligned (SWIZZLE_128B)", which);
        out.resize((size_t)rows * 8);
        for (unsigned r = 0; r < rows; ++r)
            for (unsigned k = 0; k < 8; ++k) {
                unsigned addr = start + (r / 8) * sbo + (r % 8) * 128 + 4 * k;
                addr ^= ((addr >> 7) & 7u) << 4;
                const char* p = b->smem_window + addr;
                if (p < b->smem) fail("tcgen05.mma: %s operand outside dynamic shared memory", which);
                unsigned bits;
                memcpy(&bits, p, 4);
                out[(size_t)r * 8 + k] = tf32_of(bits);
            }
    };
    std::vector<float> A, B;
    operand(desc_a, M, "A", A);
    operand(desc_b, N, "B", B);
    for (unsigned r = 0; r < M; ++r)
        for (unsigned j = 0; j < N; ++j) {
            unsigned* cell = b->tmem.data() + (size_t)r * 512 + col0 + j;
            float acc = 0.f;
            if (accumulate) memcpy(&acc, cell, 4);
            for (unsigned k = 0; k < 8; ++k) acc = fmaf(A[(size_t)r * 8 + k], B[(size_t)j * 8 + k], acc);
            memcpy(cell, &acc, 4);
        }
}
// tcgen05.commit.mbarrier::arrive::one: the MMAs issued so far have completed (the emulator runs them at issue)
void umma_commit(unsigned long long* bar) { mbar_arrive(bar); }

// ---- mbarrier: [63] phase, [62:48] pending arrivals, [47:32] arrival count of a phase, [31:0] pending bytes (signed) ----
namespace {
struct Mbar {
    unsigned phase, pending, count;
    int tx;
};
Mbar mbar_unpack(unsigned long long v) {
    return Mbar{(unsigned)(v >> 63), (unsigned)((v >> 48) & 0x7fffu), (unsigned)((v >> 32) & 0xffffu), (int)(unsigned)(v & 0xffffffffu)};
}
unsigned long long mbar_pack(const Mbar& m) {
    return ((unsigned long long)m.phase << 63) | ((unsigned long long)(m.pending & 0x7fffu) << 48) |
           ((unsigned long long)(m.count & 0xffffu) << 32) | (unsigned long long)(unsigned)m.tx;
}
void mbar_settle(unsigned long long* bar, Mbar m) {
    if (m.pending == 0 && m.tx == 0) { m.phase ^= 1u; m.pending = m.count; }     // the phase completes
    *bar = mbar_pack(m);
}
}  // namespace

void mbar_init(unsigned long long* bar, int count) {
    check_align(bar, 8, "mbarrier.init");
    if (count < 1 || count > 0x7fff) fail("mbarrier.init with count %d", count);
    *bar = mbar_pack(Mbar{0u, (unsigned)count, (unsigned)count, 0});
}

void mbar_arrive(unsigned long long* bar) {
    Mbar m = mbar_unpack(*bar);
    if (m.count == 0) fail("mbarrier.arrive on an uninitialised barrier");
    if (m.pending == 0) fail("mbarrier.arrive: more arrivals than the barrier's count in one phase");
    m.pending--;
    mbar_settle(bar, m);
    if (random_sched()) { Block* b = t_blk; b->running->wait = W_NONE; yield_to_scheduler(b, b->running); }
}

void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    Mbar m = mbar_unpack(*bar);
    if (m.count == 0) fail("mbarrier.arrive.expect_tx on an uninitialised barrier");
    if (m.pending == 0) fail("mbarrier.arrive.expect_tx: more arrivals than the barrier's count in one phase");
    m.tx += (int)bytes;
    m.pending--;
    mbar_settle(bar, m);
}

static void tma_complete_all(Block* b) {
    std::vector<Block::PendingTma> todo;
    todo.swap(b->tma);
    for (const Block::PendingTma& t : todo) {
        const TensorMap2D& mp = t.map;
        const size_t row_bytes = (size_t)mp.box0 * mp.elem;
        for (unsigned j = 0; j < mp.box1; ++j) {
            for (unsigned i = 0; i < mp.box0; ++i) {
                const long long g0 = (long long)t.c0 + i, g1 = (long long)t.c1 + j;
                size_t off = (size_t)j * row_bytes + (size_t)i * mp.elem;
                if (mp.swizzle == 128) off ^= ((off >> 7) & 7u) << 4;
                char* d = static_cast<char*>(t.dst) + off;
                if (g0 >= 0 && g1 >= 0 && (unsigned long long)g0 < mp.dim0 && (unsigned long long)g1 < mp.dim1)
                    memcpy(d, static_cast<const char*>(mp.base) + (size_t)g1 * mp.stride1 + (size_t)g0 * mp.elem, mp.elem);
                else
                    memset(d, 0, mp.elem);                       // out-of-bounds elements are zero-filled
            }
        }
        Mbar m = mbar_unpack(*t.bar);
        m.tx -= (int)(row_bytes * mp.box1);                      // the whole box counts, in bounds or not
        mbar_settle(t.bar, m);
    }
}

bool mbar_test(unsigned long long* bar, unsigned parity) {
    Block* b = t_blk;
    tma_complete_all(b);
    const Mbar m = mbar_unpack(*bar);
    if (m.count == 0) fail("mbarrier.try_wait on an uninitialised barrier");
    const bool done = m.phase != (parity & 1u);       // the phase with this parity has completed
    if (done) b->mbar_polls = 0;
    else if (++b->mbar_polls > 400000000ull) fail("mbarrier wait never completes (missing arrive / expect_tx?)");
    return done;
}

void tma_load_2d(void* dst, const TensorMap2D* map, int c0, int c1, unsigned long long* bar) {
    Block* b = t_blk;
    if (!map || map->magic != 0x7a3du) fail("cp.async.bulk.tensor with something that is not a tensor map");
    if (map->elem == 0 || (map->box0 * map->elem) % 16 != 0 || map->box0 > 256 || map->box1 > 256 || map->stride1 % 16 != 0 ||
        reinterpret_cast<uintptr_t>(map->base) % 16 != 0)
        fail("cp.async.bulk.tensor: tensor map violates the TMA alignment rules");
    if (map->swizzle == 128) {
        if (map->box0 * map->elem > 128) fail("cp.async.bulk.tensor: SWIZZLE_128B needs an inner box of at most 128 bytes");
        check_align(dst, 1024, "cp.async.bulk.tensor destination (SWIZZLE_128B)");
    } else {
        check_align(dst, 128, "cp.async.bulk.tensor destination");
    }
    const size_t bytes = (size_t)map->box0 * map->elem * map->box1;
    if (static_cast<char*>(dst) < b->smem) fail("cp.async.bulk.tensor destination outside dynamic shared memory");
    const long long nth = g_tma_loads.fetch_add(1, std::memory_order_relaxed);
    memset(dst, 0xff, bytes);
    if (g_tma_drop.load() == nth) { b->fault_injected = true; return; }   // fault injection: this copy never completes                                    // not there yet: a reader that does not wait sees NaNs
    b->tma.push_back(Block::PendingTma{dst, *map, c0, c1, bar});
}

}  // namespace simt
extern "C" long long simt_tma_loads() { return simt::g_tma_loads.load(); }
extern "C" long long simt_umma_issued() { return simt::g_umma_issued.load(); }
// the bulk tensor copy issued `after` copies from now is silently lost (its bytes never arrive); -1 switches it off
extern "C" void simt_drop_tma(long long after) { simt::g_tma_drop.store(after < 0 ? -1 : simt::g_tma_loads.load() + after); }   // tests: did the TMA kernel run at all?
namespace simt {
unsigned long long now_ns() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

void poll_yield() {
    Block* b = t_blk;
    if (!b || !b->running) return;
    if (g_abort.load(std::memory_order_relaxed)) fail("abandoned: another block of the launch failed");
    Fiber* f = b->running;
    f->wait = W_NONE;
    if ((++b->polls & 63u) == 0) sched_yield();   // other blocks / emulated ranks share the cores
    yield_to_scheduler(b, f);
}

static void fiber_main() {
    Block* b = t_blk;
    (*b->fn)();
    Fiber* f = b->running;
    for (auto& g : f->groups) run_copies(g);                    // the hardware completes them; nobody may rely on it
    f->groups.clear();
    run_copies(f->open_group);
    f->done = true;
    WarpState& w = b->warps[f->warp];
    b->alive--;
    w.alive--;
    w.alive_mask &= ~(1u << f->lane);
    // threads that exit no longer take part in barriers
    if (b->alive > 0 && b->arrived == b->alive) { b->arrived = 0; b->gen++; }
    if (w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; w.mask[w.gen & 1] = w.alive_mask; w.gen++; }
    simt_switch(&f->sp, b->sched_sp);
    abort();
}

// ---- stacks ------------------------------------------------------------------------------------------------
static std::mutex g_stack_mu;
static std::vector<std::pair<size_t, char*>> g_stack_pool;

static char* stacks_acquire(size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_stack_mu);
        for (size_t i = 0; i < g_stack_pool.size(); ++i) {
            if (g_stack_pool[i].first == bytes) {
                char* p = g_stack_pool[i].second;
                g_stack_pool.erase(g_stack_pool.begin() + (long)i);
                return p;
            }
        }
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { fprintf(stderr, "[simt] cannot map %zu bytes of fiber stacks\n", bytes); abort(); }
    return static_cast<char*>(p);
}
static void stacks_release(size_t bytes, char* p) {
    std::lock_guard<std::mutex> lk(g_stack_mu);
    if (g_stack_pool.size() < 64) g_stack_pool.emplace_back(bytes, p);
    else munmap(p, bytes);
}

static inline bool runnable(const Block& b, const Fiber& f) {
    if (f.done) return false;
    if (f.wait == W_NONE) return true;
    if (f.wait == W_BLOCK) return b.gen != f.wait_gen;
    if (f.wait == W_NAMED) return b.named[f.wait_id].gen != f.wait_gen;
    return b.warps[f.warp].gen != f.wait_gen;
}

static void run_block(Block& b, char* stacks, unsigned linear_block, dim3 grid, dim3 block, size_t smem_bytes) {
    const int T = (int)(block.x * block.y * block.z);
    const int nwarps = (T + 31) / 32;
    b.gen = 0;
    b.arrived = 0;
    b.alive = T;
    b.failed = false;
    b.warps.assign((size_t)nwarps, WarpState());
    for (NamedBarrier& nb : b.named) nb = NamedBarrier();
    b.tmem_cols = 0;
    b.tma.clear();
    b.fault_injected = false;
    b.mbar_polls = 0;
    if (smem_bytes) memset(b.smem, 0xff, smem_bytes);   // poison: NaN patterns / huge integers
    uint3 bid;
    bid.x = linear_block % grid.x;
    bid.y = (linear_block / grid.x) % grid.y;
    bid.z = linear_block / (grid.x * grid.y);
    for (int t = 0; t < T; ++t) {
        Fiber& f = b.fibers[(size_t)t];
        f.ctx.tid.x = (unsigned)t % block.x;
        f.ctx.tid.y = ((unsigned)t / block.x) % block.y;
        f.ctx.tid.z = (unsigned)t / (block.x * block.y);
        f.ctx.bid = bid;
        f.ctx.bdim = uint3{block.x, block.y, block.z};
        f.ctx.gdim = uint3{grid.x, grid.y, grid.z};
        f.wait = W_NONE;
        f.done = false;
        f.groups.clear();
        f.open_group.clear();
        f.warp = t / 32;
        f.lane = t % 32;
        WarpState& w = b.warps[(size_t)f.warp];
        w.alive++;
        w.alive_mask |= 1u << f.lane;
        void** top = reinterpret_cast<void**>(stacks + (size_t)(t + 1) * kStackBytes - 64);
        for (int r = 0; r < 6; ++r) top[r] = nullptr;
        top[6] = reinterpret_cast<void*>(&fiber_main);
        top[7] = nullptr;
        f.sp = top;
    }
    const bool rnd = random_sched();
    unsigned rng = g_seed * 2654435761u + linear_block * 40503u + 12345u;
    int next = 0;
    std::vector<int> cand;
    while (b.alive > 0 && !b.failed) {
        int pick = -1;
        if (!rnd) {
            for (int k = 0; k < T; ++k) {
                const int i = (next + k) % T;
                if (runnable(b, b.fibers[(size_t)i])) { pick = i; break; }
            }
        } else {
            cand.clear();
            for (int i = 0; i < T; ++i)
                if (runnable(b, b.fibers[(size_t)i])) cand.push_back(i);
            if (!cand.empty()) {
                rng = rng * 1664525u + 1013904223u;
                pick = cand[(rng >> 8) % cand.size()];
            }
        }
        if (pick < 0) {
            int at_block = 0, at_warp = 0, at_named = 0;
            for (const Fiber& f : b.fibers) {
                if (f.done) continue;
                if (f.wait == W_BLOCK) ++at_block; else if (f.wait == W_WARP) ++at_warp; else if (f.wait == W_NAMED) ++at_named;
            }
            char msg[256];
            snprintf(msg, sizeof(msg),
                     "deadlock in block %u: %d threads alive, %d wait in __syncthreads, %d in a warp collective, "
                     "%d at a named barrier (divergent barrier?)", linear_block, b.alive, at_block, at_warp, at_named);
            record_error(msg);
            g_abort.store(true);
            b.failed = true;
            break;
        }
        next = pick + 1;
        Fiber& f = b.fibers[(size_t)pick];
        f.wait = W_NONE;
        b.running = &f;
        simt_switch(&b.sched_sp, f.sp);
        b.running = nullptr;
    }
    if (!b.failed && !b.tma.empty() && !b.fault_injected) {
        record_error("block exited with cp.async.bulk.tensor copies in flight (nobody waited for their mbarrier)");
        b.tma.clear();
    }
    if (!b.failed && b.tmem_cols) {
        record_error("block exited without tcgen05.dealloc: tensor memory would stay allocated on the SM");
        b.tmem_cols = 0;
    }
    if (!b.failed) {
        for (int id = 1; id < 16; ++id) {
            if (b.named[id].arrived != 0) {
                char msg[160];
                snprintf(msg, sizeof(msg), "block %u exited with %d of %d arrivals pending at named barrier %d",
                         linear_block, b.named[id].arrived, b.named[id].count, id);
                record_error(msg);
            }
        }
    }
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

void launch(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& fn) {
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    const int T = (int)(block.x * block.y * block.z);
    if (nblocks == 0 || T <= 0 || T > 1024) { record_error("invalid launch configuration"); return; }
    if (smem > 227 * 1024) { record_error("too much dynamic shared memory"); return; }
    if (t_blk && t_blk->running) { record_error("nested launch"); return; }
    g_abort.store(false);
    const size_t stack_bytes = (size_t)T * kStackBytes;
    std::atomic<size_t> next_block{0};
    auto worker = [&](bool one_block_only, size_t fixed_block) {
        Block b;
        b.fibers.resize((size_t)T);
        b.fn = &fn;
        char* stacks = stacks_acquire(stack_bytes);
        // the shared-memory window starts on a 1024-byte boundary; dynamic shared memory sits 128 bytes into it, as it
        // would behind some static shared memory: only 128-byte aligned, so kernels must align swizzled tiles themselves
        b.smem_window = smem ? static_cast<char*>(aligned_alloc(1024, (smem + 128 + 1023) / 1024 * 1024)) : nullptr;
        b.smem = smem ? b.smem_window + 128 : nullptr;
        t_blk = &b;
        if (one_block_only) {
            run_block(b, stacks, (unsigned)fixed_block, grid, block, smem);
        } else {
            for (;;) {
                const size_t i = next_block.fetch_add(1);
                if (i >= nblocks || g_abort.load()) break;
                run_block(b, stacks, (unsigned)i, grid, block, smem);
            }
        }
        t_blk = nullptr;
        free(b.smem_window);
        stacks_release(stack_bytes, stacks);
    };
    if (cooperative) {
        if (nblocks > 256) { record_error("cooperative launch too large for the emulator"); return; }
        std::vector<std::thread> th;
        for (size_t i = 0; i < nblocks; ++i) th.emplace_back(worker, true, i);
        for (auto& t : th) t.join();
        return;
    }
    int P = env_int("SIMT_WORKERS", 0);
    if (P <= 0) {
        P = (int)std::thread::hardware_concurrency();
        if (P > 8) P = 8;
        if (P < 1) P = 1;
    }
    if ((size_t)P > nblocks) P = (int)nblocks;
    if (P == 1) {
        worker(false, 0);
        return;
    }
    std::vector<std::thread> th;
    for (int i = 0; i < P; ++i) th.emplace_back(worker, false, (size_t)0);
    for (auto& t : th) t.join();
}

}  // namespace simt

// test observation: tcgen05.ld / tcgen05.st executed so far (per thread)
extern "C" long long simt_tmem_access_count(void) { return simt::g_tmem_accesses.load(); }

// test control: scheduling order of the emulated threads (0 round-robin, 1 random with `seed`)
extern "C" void simt_set_schedule(int random, unsigned seed) {
    simt::g_random_sched = random ? 1 : 0;
    simt::g_seed = seed;
}

// ---- fake CUDA runtime -----------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e) {
    if (e == simt::kSimtLaunchFailure) return simt::g_err_msg;
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorNotSupported: return "operation not supported (SIMT emulator)";
        default: return "unknown error";
    }
}
cudaError_t cudaGetLastError() { return simt::g_last_error.exchange(0); }
// libgfk caches the SM count per device id: the emulated "device id" is the emulated SM count, so a test that changes
// SIMT_SMS really gets the new launch geometry
cudaError_t cudaGetDevice(int* dev) { *dev = simt::env_int("SIMT_SMS", 4); return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    switch (a) {
        case cudaDevAttrMultiProcessorCount: *v = simt::env_int("SIMT_SMS", 4); return cudaSuccess;
        case cudaDevAttrCooperativeLaunch: *v = 1; return cudaSuccess;
        case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 227 * 1024; return cudaSuccess;
    }
    return cudaErrorInvalidValue;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "SIMT host emulator");
    p->major = 10;
    p->minor = 0;
    p->multiProcessorCount = simt::env_int("SIMT_SMS", 4);
    return cudaSuccess;
}
// Small allocations come from a per-process POSIX shared-memory arena so that cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle work between emulated ranks (one process per rank, like the real thing).
namespace simt {
constexpr size_t kArenaBytes = 8u << 20;
constexpr size_t kArenaMaxAlloc = 1u << 20;
static char* g_arena = nullptr;
static size_t g_arena_used = 0;
static char g_arena_name[64] = "";
static std::mutex g_arena_mu;
struct IpcHandle { unsigned magic; int pid; unsigned long long offset; };
static std::map<int, char*> g_peer_arenas;

static void arena_unlink() { if (g_arena_name[0]) shm_unlink(g_arena_name); }
static char* arena_base() {
    if (g_arena) return g_arena;
    snprintf(g_arena_name, sizeof(g_arena_name), "/simt_arena_%d", (int)getpid());
    const int fd = shm_open(g_arena_name, O_CREAT | O_RDWR | O_TRUNC, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)kArenaBytes) != 0) return nullptr;
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return nullptr;
    atexit(arena_unlink);
    g_arena = static_cast<char*>(p);
    return g_arena;
}
static bool in_arena(const void* p) {
    return g_arena && static_cast<const char*>(p) >= g_arena && static_cast<const char*>(p) < g_arena + kArenaBytes;
}
}  // namespace simt

cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (bytes <= simt::kArenaMaxAlloc) {
        std::lock_guard<std::mutex> lk(simt::g_arena_mu);
        char* base = simt::arena_base();
        const size_t need = (bytes + 255) / 256 * 256;
        if (base && simt::g_arena_used + need <= simt::kArenaBytes) {
            *p = base + simt::g_arena_used;
            simt::g_arena_used += need;
            return cudaSuccess;
        }
    }
    *p = aligned_alloc(256, (bytes + 255) / 256 * 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
    if (!simt::in_arena(p)) free(p);   // arena blocks are never reused (test processes are short-lived)
    return cudaSuccess;
}
// Asynchronous device-to-host copies keep their real contract: the host buffer holds the data only after the stream has
// been synchronised.  The value is snapshotted at enqueue time (everything enqueued earlier has run - the emulator
// executes launches synchronously), the destination is poisoned, and the next synchronising call delivers it; host
// code that reads the result of cudaMemcpyAsync without synchronising sees 0xA5 bytes instead of passing by accident.
namespace simt {
struct PendingD2H { void* dst; std::vector<unsigned char> data; };
static std::mutex g_d2h_mu;
static std::vector<PendingD2H> g_d2h;
static void flush_d2h() {
    std::lock_guard<std::mutex> lk(g_d2h_mu);
    for (PendingD2H& c : g_d2h) memcpy(c.dst, c.data.data(), c.data.size());
    g_d2h.clear();
}
}  // namespace simt
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
    simt::flush_d2h();
    memcpy(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) {
    if (kind == cudaMemcpyDeviceToHost && bytes > 0) {
        simt::PendingD2H c;
        c.dst = dst;
        c.data.assign(static_cast<const unsigned char*>(src), static_cast<const unsigned char*>(src) + bytes);
        memset(dst, 0xa5, bytes);
        std::lock_guard<std::mutex> lk(simt::g_d2h_mu);
        simt::g_d2h.push_back(std::move(c));
        return cudaSuccess;
    }
    memcpy(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { simt::flush_d2h(); return simt::g_last_error.load(); }
cudaError_t cudaDeviceSynchronize() { simt::flush_d2h(); return simt::g_last_error.load(); }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new SimtEvent{0.0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { simt::flush_d2h(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    if (!simt::in_arena(p)) return cudaErrorInvalidValue;
    simt::IpcHandle ih{0x51a7u, (int)getpid(), (unsigned long long)(static_cast<char*>(p) - simt::g_arena)};
    memset(h, 0, sizeof(*h));
    memcpy(h, &ih, sizeof(ih));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    simt::IpcHandle ih;
    memcpy(&ih, &h, sizeof(ih));
    if (ih.magic != 0x51a7u || ih.offset >= simt::kArenaBytes) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lk(simt::g_arena_mu);
    char*& base = simt::g_peer_arenas[ih.pid];
    if (!base) {
        char name[64];
        snprintf(name, sizeof(name), "/simt_arena_%d", ih.pid);
        const int fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) return cudaErrorInvalidValue;
        void* m = mmap(nullptr, simt::kArenaBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return cudaErrorInvalidValue;
        base = static_cast<char*>(m);
    }
    *p = base + ih.offset;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t simt_func_set_attribute(size_t bytes) { return bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t simt_occupancy(int* blocks, int threads, size_t smem) {
    *blocks = (threads <= 1024 && smem <= 227 * 1024) ? 1 : 0;
    return cudaSuccess;
}
