"""bench.py's JSON contract, checked without a GPU: the line assembly is a pure function (made-up measurements in,
keys and arithmetic out), and the whole `run_ours` / secondary code path is executed once over the SIMT-emulator build
of libgfk at toy sizes (torch.cuda is patched away here and only here), so a NameError or a wrong key cannot wait for
the first GPU run to show up.  No timing produced here means anything."""
import argparse
import contextlib
import ctypes
import json

import numpy as np
import pytest
import torch

import bench
import simt_harness as H


def _args(**kw):
    base = dict(gpus=1, steps=4, warmup=3, impl='ours', scaling='strong', precision='f64', cpu_baseline_m=3,
                no_cpu_baseline=False, no_secondary=True, secondary_only=None)
    base.update(kw)
    return argparse.Namespace(**base)


def test_assemble_line_contract_and_arithmetic():
    res = dict(dev_ms=4 * 16.0, e2e_ms=4 * 17.5, loop_ms=[15.0] * 4, launches=40, n_total=1_000_000,
               clocks={'sm_mhz': 1900.0, 'sm_max_mhz': 1965, 'reasons': []},
               other=dict(ms_per_step=12.0, loop_ms=11.0, same_indices=True))
    line = bench.assemble_line(_args(), 1, res, 36.2, traffic=74_865_920)
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert key in line, key
    assert line['value'] == pytest.approx(1e9 / 16e-3) and line['e2e']['value'] == pytest.approx(1e9 / 17.5e-3)
    assert line['e2e']['h2d_bytes_per_step'] == 80_000_000 and line['e2e']['d2h_bytes_per_step'] == 4000
    r = line['roofline']
    # the binding roof is the FP64 pipe: 51 algorithmic flops per pair over the launch time, against the measured peak
    assert r['bound'] == 'fp64' and r['unit'] == 'TFLOP/s' and r['flops_per_pair'] == 51
    assert r['achieved'] == pytest.approx(51 * 1e9 / 15e-3 / 1e12) and r['frac'] == pytest.approx(r['achieved'] / 36.2)
    assert r['frac'] < 0.2 and r['traffic'] == 74_865_920
    h = r['hbm_algorithmic']
    assert h['algorithmic_bytes_per_launch'] == 88e9 and h['algorithmic_to_measured_traffic'] == pytest.approx(88e9 / 74_865_920)
    assert h['achieved_gbs'] == pytest.approx(88e9 / 15e-3 / 1e9)
    assert 0.2 < r['ceiling']['max_frac_at_full_fp64_pipe'] < 0.3          # strict: ~101 FP64 instructions for 51 flops
    fast = line['f64_fast_mode']['roofline']
    assert fast['bound'] == 'fp64' and fast['ceiling']['max_frac_at_full_fp64_pipe'] > 0.5
    json.dumps(line)


def test_multi_gpu_headline_is_the_strong_scaled_baseline_config():
    res = dict(dev_ms=4 * 12.0, e2e_ms=4 * 14.0, loop_ms=[11.0] * 4, launches=8000, n_total=1_000_000,
               clocks={'sm_mhz': 1900.0, 'sm_max_mhz': 1965, 'reasons': []}, other=None)
    weak = dict(dev_ms=2 * 23.0, e2e_ms=2 * 30.0, loop_ms=[22.0] * 2, n_total=8_000_000)
    line = bench.assemble_line(_args(gpus=8), 8, res, 36.2, other_scaling=('weak', weak, 2))
    assert line['scaling'] == 'strong' and line['n_gpus'] == 8 and 'n=1000000' in line['config']['workload']
    assert line['value'] == pytest.approx(1e9 / 12e-3)
    assert line['roofline']['pairs_per_launch'] == 125_000 * 1000             # this rank's rows x m
    assert line['e2e']['h2d_bytes_per_step'] == 10_000_000                    # 1/8 of the replica per rank
    assert line['weak']['n'] == 8_000_000 and line['weak']['value'] == pytest.approx(8e9 / 23e-3)


class _FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)


@pytest.fixture
def emulated_cuda(monkeypatch):
    from gradient_free_mcmc_postprocessing_b200 import _engine, _lib, proxies
    for k in ('GFK_PERSISTENT', 'GFK_RESIDENT', 'SIMT_SMS', 'GFK_PRECISION'):
        monkeypatch.delenv(k, raising=False)
    H.set_schedule(False)
    monkeypatch.setattr(_lib, '_lib', H.load())
    cpu = torch.device('cpu')
    for mod in (_engine, proxies):
        monkeypatch.setattr(mod, '_require_cuda', lambda device=None: cpu)
        monkeypatch.setattr(mod, '_stream_ptr', lambda device=None: ctypes.c_void_p(0))
    monkeypatch.setattr(torch.cuda, 'device', lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a: None)
    monkeypatch.setattr(torch.cuda, 'Event', _FakeEvent)

    class _Stream:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _Stream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    real_device = torch.device
    monkeypatch.setattr(torch, 'device', lambda *a, **k: cpu if a and a[0] == 'cuda' else real_device(*a, **k))
    real_empty = torch.empty
    # the 512 MiB L2-flush buffer becomes 1 MiB; every other allocation keeps its size
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: real_empty(*(((1 << 20),) + a[1:] if a and isinstance(a[0], int)
                                                                    and a[0] >= (256 << 20) else a), **k))
    return cpu


def test_run_ours_executes_end_to_end_over_the_emulator(emulated_cuda, monkeypatch, capsys):
    monkeypatch.setattr(bench, 'N_C3', 3000)
    monkeypatch.setattr(bench, 'M', 12)
    bench.run_ours(_args(steps=1, warmup=1, cpu_baseline_m=3))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line['parity_prefix_ok'] is True and line['gpu_launches'] > 0
    assert line['roofline']['bound'] == 'fp64' and line['cpu_baseline']['kind'] == 'port'
    assert line['f64_fast_mode']['same_indices_as_f64'] in (True, False)


def test_secondary_records_execute_over_the_emulator(emulated_cuda, monkeypatch):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    monkeypatch.setattr(bench, 'N_C3', 2500)
    monkeypatch.setattr(bench, 'M', 8)
    c2, c4, c5 = synthetic.config_c2, synthetic.config_c4, synthetic.config_c5
    monkeypatch.setattr(synthetic, 'config_c2', lambda **k: c2(n=3000, m=10))
    monkeypatch.setattr(synthetic, 'config_c4', lambda **k: c4(n=700, d=10))
    monkeypatch.setattr(synthetic, 'config_c5', lambda **k: c5(n=1200, d=64, m=6))
    dev = emulated_cuda
    for rec in (bench.measure_thin_c2(dev, 36.2), bench.measure_thin_c5(dev), bench.measure_thin_c3_gaussian(dev, 36.2)):
        assert 'error' not in rec, rec
        assert rec['roofline']['frac'] > 0 and rec['pair_evals_per_s'] > 0
    assert bench.measure_thin_c5(dev)['roofline']['bound'] == 'hbm'
