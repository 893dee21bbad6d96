"""Host model of the flag-in-data exchange protocols of csrc/thin.cu, run under adversarial schedules.

The kernels exchange (value, index) records without fences or atomics: every 8-byte word carries a tag, a
reader accepts a record only when all its words carry the expected tag, buffers alternate by step parity
(intra-GPU slots, peer mailboxes) and by job parity (mailboxes), and mailbox tags come from a per-job epoch
so the mailboxes are never cleared.  The safety argument is "a writer can be at most one step (one job)
ahead of the slowest reader".  This test executes that argument: participants are generators that yield
before every single word access, a seeded scheduler interleaves them arbitrarily (including word-by-word
torn writes), and every reader must obtain exactly the record of the step it is waiting for.
"""
import random

import pytest

RANKS = 16          # kMailRanks


def _slot16_encode(tag16, value_bits, index):
    """csrc/thin.cu slot_store: two words, 16-bit tag over 48-bit payload."""
    t = (tag16 & 0xffff) << 48
    return t | (value_bits >> 16), t | ((value_bits & 0xffff) << 32) | index


def _slot16_decode(a, b):
    return ((a << 16) & 0xffffffffffffffff) | ((b >> 32) & 0xffff), b & 0xffffffff


def test_slot16_roundtrip_and_tag_isolation():
    rng = random.Random(1)
    for _ in range(2000):
        v, i, tag = rng.getrandbits(64), rng.getrandbits(32), rng.getrandbits(16)
        a, b = _slot16_encode(tag, v, i)
        assert (a >> 48) == tag and (b >> 48) == tag
        assert _slot16_decode(a, b) == (v, i)


# ---- intra-GPU: every CTA publishes, every CTA polls all slots (single-rank kernel) ---------------------
def _cta(cta, n_cta, steps, slots, log):
    for step in range(steps):
        tag = (step + 1) & 0xffff
        value = (cta * 1000003 + step * 7919) & 0xffffffffffffffff
        a, b = _slot16_encode(tag, value, cta)
        buf = slots[step & 1]
        yield
        buf[cta][0] = a                     # the two halves of the 16-byte store may land separately
        yield
        buf[cta][1] = b
        got = {}
        while len(got) < n_cta:
            for other in range(n_cta):
                if other in got:
                    continue
                yield
                w0 = buf[other][0]
                yield
                w1 = buf[other][1]
                if (w0 >> 48) == tag and (w1 >> 48) == tag:
                    got[other] = _slot16_decode(w0, w1)
        for other, (v, i) in got.items():
            assert v == (other * 1000003 + step * 7919) & 0xffffffffffffffff and i == other, (cta, step, other)
        log.append((cta, step))


def _run(generators, rng, max_ticks=4_000_000):
    live = list(generators)
    ticks = 0
    while live:
        g = rng.choice(live)
        # bursts make some participants run far ahead of others
        for _ in range(rng.choice((1, 1, 2, 5, 40))):
            try:
                next(g)
            except StopIteration:
                live.remove(g)
                break
        ticks += 1
        assert ticks < max_ticks, 'deadlock or livelock'


@pytest.mark.parametrize('seed', range(6))
def test_grid_slots_parity_double_buffer(seed):
    rng = random.Random(seed)
    n_cta, steps = 7, 40
    slots = [[[0, 0] for _ in range(n_cta)] for _ in range(2)]
    log = []
    _run([_cta(c, n_cta, steps, slots, log) for c in range(n_cta)], rng)
    assert len(log) == n_cta * steps


# ---- cross-GPU: mailboxes with job parity, step parity and epoch tags (never cleared) -------------------
class _Epoch:
    """csrc/comm.cu p2p_next_job: identical on every rank because every rank runs the same jobs."""

    def __init__(self):
        self.tag, self.jobs = 1, 0

    def next_job(self, steps):
        base, parity = self.tag, self.jobs & 1
        self.tag += steps + 2
        self.jobs += 1
        return base, parity


def _offset(job_parity, exchange, rank):
    return ((job_parity * 2 + (exchange & 1)) * RANKS + rank) * 4


def _rank(rank, n_ranks, jobs, mail, log):
    epoch = _Epoch()
    for job, steps in enumerate(jobs):
        base, jp = epoch.next_job(steps)
        for step in range(steps):
            tag = (base + step + 1) & 0xffffffff
            value = (job * 7 + rank * 13 + step * 31) & 0xffffffffffffffff
            words = [(tag << 32) | (value & 0xffffffff), (tag << 32) | (value >> 32), (tag << 32) | rank]
            for peer in range(n_ranks):                  # mail_post: three separate 8-byte stores per peer
                off = _offset(jp, step, rank)
                for k in range(3):
                    yield
                    mail[peer][off + k] = words[k]
            got = {}
            while len(got) < n_ranks:                    # mail_wait on the local mailbox
                for other in range(n_ranks):
                    if other in got:
                        continue
                    off = _offset(jp, step, other)
                    w = []
                    for k in range(3):
                        yield
                        w.append(mail[rank][off + k])
                    if all((x >> 32) == tag for x in w):
                        got[other] = (((w[1] & 0xffffffff) << 32) | (w[0] & 0xffffffff), w[2] & 0xffffffff)
            for other, (v, i) in got.items():
                assert v == (job * 7 + other * 13 + step * 31) and i == other, (rank, job, step, other)
            log.append((rank, job, step))


@pytest.mark.parametrize('seed', range(6))
def test_peer_mailboxes_epoch_tags_and_parities(seed):
    rng = random.Random(100 + seed)
    n_ranks = 3
    jobs = [5, 1, 4, 2, 6, 3]        # consecutive jobs of different lengths, odd and even: parities collide on purpose
    mail = [[0] * (2 * 2 * RANKS * 4) for _ in range(n_ranks)]
    log = []
    _run([_rank(r, n_ranks, jobs, mail, log) for r in range(n_ranks)], rng)
    assert len(log) == n_ranks * sum(jobs)


def test_model_detects_missing_parity_buffer():
    """Negative control: without the step-parity double buffer a fast CTA overwrites a record that a slow CTA
    has not read yet, and the model must notice (the slow reader never sees its tag again)."""
    caught = 0
    for seed in range(4):
        rng = random.Random(seed)
        n_cta, steps = 7, 40
        one = [[0, 0] for _ in range(n_cta)]
        log = []
        try:
            _run([_cta(c, n_cta, steps, [one, one], log) for c in range(n_cta)], rng, max_ticks=300_000)
        except AssertionError:
            caught += 1
    assert caught == 4
