"""Failure protocol of the multi-rank thinning exchange (csrc/thin.cu "failure protocol"), executed on the SIMT
emulator with one process per rank: a peer that never launches, a peer that leaves mid-job, and a stale record with
an out-of-range index must each end in GFK_ECUDA on every surviving rank within a bounded time - no hang, no
out-of-range read (the inputs sit in front of PROT_NONE pages).  The reference has no counterpart (single process);
the bar is SURVEY.md 8b "errors: non-zero return code, never abort"."""
import os
import subprocess
import sys

import pytest

import simt_harness as H

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(scenario, strategy, world, tmp_path):
    H.load()
    worker = os.path.join(HERE, 'simt_fault_worker.py')
    procs = [subprocess.Popen([sys.executable, worker, scenario, strategy, str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=400)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and 'SIMT_FAULT_OK %d' % r in out, (r, out[-3000:])
    return outs


@pytest.mark.parametrize('strategy', ['persistent', 'steps'])
def test_peer_never_launches(strategy, tmp_path):
    _run('absent', strategy, 2, tmp_path)


@pytest.mark.parametrize('strategy', ['persistent', 'steps'])
def test_peer_leaves_at_step_17_and_the_poison_word_releases_the_others(strategy, tmp_path):
    outs = _run('dies', strategy, 3, tmp_path)
    assert 'deadline' in outs[0] and 'peer rank aborted' in outs[1]


@pytest.mark.parametrize('strategy', ['persistent', 'steps'])
def test_record_with_an_out_of_range_index_is_refused(strategy, tmp_path):
    outs = _run('forged', strategy, 2, tmp_path)
    assert 'out of range' in outs[0]
