// comm.h - process-wide communicator state of libgfk (one process per GPU, NCCL over NVLink).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfk {
int comm_size();  // 1 when no communicator was initialised
int comm_rank();
// all-gather of one 16-byte (value, index) record per rank: recv[r] <- rank r's *send
int comm_allgather16(const void* send, void* recv, cudaStream_t st);

// ---- peer-memory mailboxes (CUDA IPC over NVLink) for the per-step (value, index) exchange ------------
// mail[job parity][step parity][rank][4] 64-bit words; each word = (tag << 32 | payload), so a reader that
// sees the expected tag in the three payload words of a slot has the whole record: no fences, no atomics,
// one NVLink store latency per step instead of an NCCL launch.
constexpr int kMailRanks = 16;
constexpr int kMailSlotWords = 2 * 2 * kMailRanks * 4;
// one extra 128-byte line after the slots: word kMailPoison holds the tag_base of a job that a peer aborted
// (written by that peer into every mailbox; polled by every wait of the job, see thin.cu "failure protocol")
constexpr int kMailPoison = kMailSlotWords;
constexpr int kMailWords = kMailSlotWords + 16;
struct MailArgs {
    unsigned long long* local;             // this rank's mailbox (peers write into it)
    unsigned long long* peer[kMailRanks];  // every rank's mailbox as mapped in this process (peer[rank] == local)
    unsigned tag_base;                     // tags of this job are tag_base + step + 1
    int job_parity;
    int rank;
    int enabled;
};
// fills `out` for the next collective job of `steps` exchanges (advances the tag epoch); enabled = 0 if
// the mailboxes were never opened
void p2p_next_job(MailArgs* out, int steps);
}  // namespace gfk
