#!/usr/bin/env python
"""bench.py - Stein-kernel pair evaluations / s for greedy gradient-free thinning on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]
                    [--precision f64|f64-fast|f32]

One "step" is one complete thinning job (n candidates -> m selections): standardisation, median
preconditioner, log-weights and the m-step greedy loop.  Workload: BASELINE.json config C3 (synthetic d=4
Gaussian mixture, n = 10^6, m = 1000, gradient-free kernel, 'med' preconditioner) - the configuration the
north-star target is quoted on.  The proxy q is the same mixture with its covariances widened by 1.3
(synthetic.config_c3): under SURVEY.md's moment-matched Gaussian proxy gradient-free thinning degenerates at this
n (the search returns the same one or two points), which would make the index-parity check vacuous; per-step
arithmetic is identical for either proxy, and the Gaussian-proxy job is timed next to it (`thin_c3_gaussian_proxy`).

At N > 1 (one process per GPU, torchrun) the HEADLINE is BASELINE's own config: the same n = 10^6 rows sharded
over the N ranks ("scaling": "strong"), one (value, index) pair per rank exchanged per greedy step; 10^6 rows PER
GPU is reported under the `weak` key (`--scaling weak` makes it the headline instead).

`value`    : pair-evals/s with the raw inputs already resident in HBM (public API, CUDA tensors in)
`e2e`      : the same call with pinned HOST tensors in and the numpy index vector out (H2D + D2H timed)
`roofline` : the thinning kernel against the roof that binds it.  The shard is read from HBM once per job and then
             lives in shared memory / L2 (ncu: DRAM traffic ~1000x below the algorithmic bytes), so the bound is the
             FP64 pipe: achieved = algorithmic flops (8d + 16 + 3 per pair) x pairs per launch /
             launch time (CUDA events around the launch), peak = the FP64 FMA rate measured on this box by
             gfk_fma_peak (MEASURED_PEAKS.json has no FP64 figure).  `ceiling` states what the executed-vs-algorithmic
             FP64 instruction ratio of the formulation allows at a fully busy pipe (static SASS counts,
             profiles/r02_static_sass_counts.json); `hbm_algorithmic` keeps the bytes-based figure of the streaming
             formulation next to the measured DRAM `traffic` and their ratio.
`f64_fast_mode`, `thin_c2`, `ksd_c4`, `thin_c5`, `thin_c3_gaussian_proxy` (N = 1 only): secondary records, each with
             its own roofline object - the FMA formulation of the headline job, BASELINE configs C2 / C4 (full
             n = 2 x 10^5 KSD) / C5 measured in a subprocess after the headline, and SURVEY's Gaussian-proxy C3.
`cpu_baseline`: the numpy oracle (restatement of the reference's stein_thinning path) timed on this
             host on a bounded sample of the same workload (same n, reduced m)
"""
import argparse
import json
import os
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, M, N_C3 = 4, 1000, 1_000_000
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def algo_bytes_per_pair(d, gf, f32=False):
    """read x_i, s_i (+ lw_i), read and write A_i (float64 running objective in every mode)."""
    e = 4 if f32 else 8
    return 2 * d * e + 16 + (e if gf else 0)


def algo_flops_per_pair(d, gf):
    """8d + 16 (+3 for the gradient-free weights), FMA = 2, rsqrt = 1."""
    return 8 * d + 16 + (3 if gf else 0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scaling', default='strong', choices=['weak', 'strong'])
    ap.add_argument('--precision', default=os.environ.get('GFK_PRECISION', 'f64'),
                    choices=['f64', 'f64-fast', 'f32'])
    ap.add_argument('--cpu-baseline-m', type=int, default=40)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary records (C2, C4, C5, weak scaling)')
    ap.add_argument('--secondary-only', type=float, default=None, metavar='FP64_PEAK_TFLOPS',
                    help='(internal) run only the secondary single-GPU measurements and print their JSON object')
    return ap.parse_args()


def make_inputs(n, seed=12345, proxy='wide-mixture'):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    c = synthetic.config_c3(n=n, d=D, seed=seed, m=M, proxy=proxy)
    return {k: np.ascontiguousarray(c[k]) for k in ('sample', 'log_p', 'log_q', 'gradient_q')}


def hbm_peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def static_ceiling(precision):
    """What the formulation's executed FP64 instructions per pair allow (tools/sass_counts.py, no GPU involved)."""
    path = os.path.join(ROOT, 'profiles', 'r02_static_sass_counts.json')
    try:
        c = json.load(open(path))['c3_update_loop'][precision]
    except Exception:   # noqa: BLE001 - absent or stale file: the line simply carries no ceiling
        return None
    return {'fp64_pipe_instructions_per_pair': c['fp64_pipe_instructions_per_pair'],
            'instructions_per_pair': c['instructions_per_pair'],
            'max_frac_at_full_fp64_pipe': c['max_frac_at_full_fp64_pipe'],
            'max_frac_at_full_issue': c['max_frac_at_full_issue'],
            'source': 'static SASS counts of the update loop, profiles/r02_static_sass_counts.json'}


def fp64_roofline(kernel, pairs_per_launch, flops_per_pair, launch_ms, fp64_peak_tflops, precision=None, extra=None):
    """The roofline object of a launch bound by the FP64 pipe (algorithmic flops, measured FMA peak)."""
    achieved = pairs_per_launch * flops_per_pair / (launch_ms * 1e-3) / 1e12
    r = {'bound': 'fp64', 'kernel': kernel, 'achieved': achieved, 'peak': fp64_peak_tflops, 'unit': 'TFLOP/s',
         'frac': achieved / fp64_peak_tflops if fp64_peak_tflops else None, 'traffic': None,
         'peak_source': 'measured on this box (gfk_fma_peak: register-resident DFMA chains on all SMs); '
                        'MEASURED_PEAKS.json has no FP64 figure',
         'flops_per_pair': flops_per_pair, 'pairs_per_launch': pairs_per_launch, 'launch_ms': launch_ms}
    if precision is not None:
        c = static_ceiling(precision)
        if c:
            r['ceiling'] = c
    if extra:
        r.update(extra)
    return r


def thinning_roofline(precision, rows_local, d, gf, loop_ms, m, fp64_peak_tflops, traffic=None):
    """Headline-style roofline of one persistent launch (= m greedy steps over this rank's rows)."""
    pairs = float(rows_local) * m
    f32 = precision == 'f32'
    r = fp64_roofline('thin_persistent_kernel', pairs, algo_flops_per_pair(d, gf), loop_ms, fp64_peak_tflops,
                      precision=precision if (d == 4 and gf and not f32) else None)
    r['traffic'] = traffic
    r['greedy_step_us'] = loop_ms / m * 1e3
    peak_gbs, peak_src = hbm_peak()
    bytes_launch = algo_bytes_per_pair(d, gf, f32) * pairs
    gbs = bytes_launch / (loop_ms * 1e-3) / 1e9
    r['hbm_algorithmic'] = {
        'achieved_gbs': gbs, 'peak_gbs': peak_gbs, 'frac': gbs / peak_gbs, 'peak_source': peak_src,
        'algorithmic_bytes_per_launch': bytes_launch,
        'algorithmic_to_measured_traffic': (bytes_launch / traffic) if traffic else None,
        'note': 'bytes of the STREAMING formulation (16d + 16 + 8 per pair); the persistent kernel reads the shard from '
                'HBM once per job and serves every later greedy step from shared memory / L2, so this fraction may '
                'exceed 1 and is not the binding roof - the FP64 fraction above is'}
    return r


# ---- reference arm: the numpy oracle on the host cores ----------------------------------------------
_SHARED_INPUTS = None      # set before fork so the workers inherit the arrays without pickling


def _oracle_job(args):
    inputs, m = args
    if inputs is None:
        inputs = _SHARED_INPUTS
    from oracle.stein_thinning_oracle.thinning import thin_gf
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t0 = time.perf_counter()
        thin_gf(inputs['sample'], inputs['log_p'], inputs['log_q'], inputs['gradient_q'], m, preconditioner='med')
        return time.perf_counter() - t0


def cpu_oracle_throughput(inputs, m, workers):
    """pair-evals/s of `workers` concurrent oracle jobs (the reference parallelises across chains only,
    code/src/utils/parallel.py:48-52; one job is single-threaded numpy)."""
    n = inputs['sample'].shape[0]
    if workers <= 1:
        dt = _oracle_job((inputs, m))
        return n * m / dt, dt
    import multiprocessing as mp
    global _SHARED_INPUTS
    _SHARED_INPUTS = inputs
    ctx = mp.get_context('fork')
    with ctx.Pool(workers) as pool:
        pool.map(_oracle_job, [(None, 1)] * workers)          # start the workers outside the timed region
        t0 = time.perf_counter()
        pool.map(_oracle_job, [(None, m)] * workers, chunksize=1)
        dt = time.perf_counter() - t0
    return workers * n * m / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    inputs = make_inputs(N_C3)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    m_sample = 20

    def one_step():
        # the path itself is single-threaded numpy; the only parallelism the reference has is across
        # independent chains, so time both and keep whichever uses the host better
        single = cpu_oracle_throughput(inputs, m_sample, 1)[0]
        multi = cpu_oracle_throughput(inputs, m_sample, workers)[0] if workers > 1 else 0.0
        return (single, 1) if single >= multi else (multi, workers)

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    results = [one_step() for _ in range(args.steps)]
    total = time.perf_counter() - t0
    value = float(np.mean([v for v, _ in results]))
    used = int(np.median([w for _, w in results]))
    sample = (f'best of 1 job and {workers} concurrent single-threaded numpy jobs (the reference parallelises '
              f'across chains only; best used {used} core(s)), each n={N_C3} d={D} GF med, '
              f'm={m_sample} of {M} greedy steps per step')
    workers = used
    line = {
        'impl': 'reference', 'metric': 'stein_kernel_pair_evals_per_s', 'value': value, 'unit': 'pair-evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'C3 gradient-free thinning n={N_C3} d={D} m={M} med (numpy oracle port of '
                               'stein_thinning; bounded sample, throughput is m-independent)', 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'pair-evals/s', 'cores': workers, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'pair-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ---- secondary single-GPU records (own subprocess) ------------------------------------------------------------
def _timed_thin(make_problem, m, reps=3):
    """best-of-`reps` whole-job time (problem construction + greedy loop) and loop-only time, CUDA events"""
    import torch
    best, best_loop, idx = 1e30, 1e30, None
    for rep in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timing = {}
        e0.record()
        idx = make_problem().thin(m, timing=timing)
        e1.record()
        torch.cuda.synchronize()
        if rep > 0:                                        # rep 0 warms up
            best = min(best, e0.elapsed_time(e1))
            best_loop = min(best_loop, timing['loop_ms'])
    return best, best_loop, idx


def measure_ksd_c4(dev, fp64_peak_tflops):
    """BASELINE config C4: the full O(n^2) gradient-free KSD of n = 2 x 10^5, d = 10 samples with the median-heuristic
    preconditioner (report/report.tex:309-313, code/src/utils/ksd.py:19-27); pair-evals = n (n + 1) / 2."""
    import torch
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    try:
        n, d = 200_000, 10
        c = synthetic.config_c4(n=n, d=d)
        inputs = {k: torch.from_numpy(np.ascontiguousarray(c[k])).to(dev) for k in ('sample', 'gradient_q', 'log_p', 'log_q')}
        prob = SteinProblem(inputs['sample'], inputs['gradient_q'], log_p=inputs['log_p'], log_q=inputs['log_q'],
                            preconditioner='med', precision='f64', device=dev)
        prob.ksd(None, 20_000)                                  # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ks = prob.ksd(None, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pairs = n * (n + 1) / 2
        return {'workload': f'C4 full gradient-free KSD n={n} d={d} med, f64', 'seconds': ms * 1e-3,
                'pair_evals_per_s': pairs / (ms * 1e-3), 'ksd': float(ks[-1]),
                'roofline': fp64_roofline('ksd_rowsum_kernel', pairs, algo_flops_per_pair(d, True), ms, fp64_peak_tflops)}
    except Exception as e:   # noqa: BLE001 - a secondary number must not take the headline line down with it
        return {'error': repr(e)[:300]}


def measure_thin_c2(dev, fp64_peak_tflops):
    """BASELINE config C2 (Lotka-Volterra-shaped chain, d = 4, n = 2 x 10^5, m = 1000, score-based kernel, 'med'; runs
    of duplicated rows) - the whole job from device-resident inputs, best of 3."""
    import torch
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    try:
        c = synthetic.config_c2()
        x = torch.from_numpy(np.ascontiguousarray(c['sample'])).to(dev)
        g = torch.from_numpy(np.ascontiguousarray(c['gradient'])).to(dev)
        n, m = x.shape[0], int(c['m'])
        ms, loop_ms, idx = _timed_thin(lambda: SteinProblem(x, g, preconditioner='med', precision='f64', device=dev), m)
        r = thinning_roofline('f64', n, 4, False, loop_ms, m, fp64_peak_tflops)
        r['note'] = 'the shard is fully shared-memory resident; the step is dominated by the per-step grid exchange latency'
        return {'workload': f'C2 score-based Stein thinning n={n} d=4 m={m} med, f64', 'seconds': ms * 1e-3,
                'pair_evals_per_s': n * m / (ms * 1e-3), 'unique_selected': int(len(np.unique(idx))), 'roofline': r}
    except Exception as e:   # noqa: BLE001
        return {'error': repr(e)[:300]}


def measure_thin_c5(dev):
    """BASELINE config C5 (d = 64, n = 5 x 10^5, m = 500, score-based, 'med'): one launch per greedy step, every step
    streams x and s from HBM (512 MB replica > L2): the HBM roofline binds, 16d + 16 = 1 040 bytes per pair."""
    import torch
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    try:
        c = synthetic.config_c5()
        x = torch.from_numpy(np.ascontiguousarray(c['sample'])).to(dev)
        g = torch.from_numpy(np.ascontiguousarray(c['gradient'])).to(dev)
        n, d, m = x.shape[0], x.shape[1], int(c['m'])
        ms, loop_ms, idx = _timed_thin(lambda: SteinProblem(x, g, preconditioner='med', precision='f64', device=dev), m, reps=2)
        peak_gbs, peak_src = hbm_peak()
        step_bytes = algo_bytes_per_pair(d, False) * n
        gbs = step_bytes / (loop_ms / m * 1e-3) / 1e9
        return {'workload': f'C5 score-based Stein thinning n={n} d={d} m={m} med, f64', 'seconds': ms * 1e-3,
                'pair_evals_per_s': n * m / (ms * 1e-3), 'unique_selected': int(len(np.unique(idx))),
                'roofline': {'bound': 'hbm', 'kernel': 'thin_step_chunked_kernel', 'achieved': gbs, 'peak': peak_gbs,
                             'unit': 'GB/s', 'frac': gbs / peak_gbs, 'traffic': None, 'peak_source': peak_src,
                             'algorithmic_bytes_per_launch': step_bytes, 'launch_ms': loop_ms / m,
                             'note': 'one launch = one greedy step over all n rows; launch time = loop time / m'}}
    except Exception as e:   # noqa: BLE001
        return {'error': repr(e)[:300]}


def measure_thin_c3_gaussian(dev, fp64_peak_tflops):
    """SURVEY.md section 8d's literal C3: the moment-matched Gaussian proxy (the headline uses the widened mixture, see
    the module docstring).  Same arithmetic per step; the selection degenerates to one or two points."""
    import torch
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    try:
        h = make_inputs(N_C3, proxy='gaussian')
        t = {k: torch.from_numpy(v).to(dev) for k, v in h.items()}
        ms, loop_ms, idx = _timed_thin(lambda: SteinProblem(t['sample'], t['gradient_q'], log_p=t['log_p'], log_q=t['log_q'],
                                                            preconditioner='med', precision='f64', device=dev), M, reps=2)
        return {'workload': f'C3 with the moment-matched Gaussian proxy n={N_C3} d={D} m={M} med, f64', 'seconds': ms * 1e-3,
                'pair_evals_per_s': N_C3 * M / (ms * 1e-3), 'unique_selected': int(len(np.unique(idx))),
                'roofline': thinning_roofline('f64', N_C3, D, True, loop_ms, M, fp64_peak_tflops)}
    except Exception as e:   # noqa: BLE001
        return {'error': repr(e)[:300]}


def secondary_in_subprocess(fp64_peak_tflops, timeout_s=420):
    """The secondary measurements run in their own process: whatever happens there (a CUDA error, a hang) cannot take
    the headline measurement - already complete at this point - down with it."""
    import subprocess
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--secondary-only', repr(float(fp64_peak_tflops))],
                             capture_output=True, text=True, timeout=timeout_s)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        return {'secondary_error': ('no output; rc=%d; ' % out.returncode) + out.stderr[-300:]}
    except Exception as e:   # noqa: BLE001
        return {'secondary_error': repr(e)[:300]}


# ---- the JSON line (pure: unit-tested on the CPU with made-up measurements) -------------------------------------
def assemble_line(args, world, res, fp64_peak_tflops, other_scaling=None, traffic=None):
    """`res`: dict(dev_ms, e2e_ms, loop_ms [per timed step], launches, clocks, n_total, other) of the headline case."""
    n_total = int(res['n_total'])
    pairs = float(n_total) * M
    ms_per_step = res['dev_ms'] / args.steps
    e2e_ms_per_step = res['e2e_ms'] / args.steps
    rows_local = -(-n_total // world)
    loop_ms = float(np.mean(res['loop_ms']))
    f32 = args.precision == 'f32'
    # host inputs: each rank uploads its 1/N block of rows over PCIe and all-gathers the rest over NVLink
    h2d = sum(int(np.prod(s)) * 8 for s in ((n_total, D), (n_total, D), (n_total,), (n_total,))) // world
    sharding = 'one GPU' if world == 1 else (
        f'{n_total} rows sharded over {world} GPUs ({rows_local} candidate rows per GPU, full replica of x / s / lw on '
        'every GPU), one (value, index) pair per rank exchanged per greedy step')
    line = {
        'metric': 'stein_kernel_pair_evals_per_s', 'value': pairs / (ms_per_step * 1e-3), 'unit': 'pair-evals/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32' if f32 else 'f64',
        'data': 'synthetic',
        'config': {'workload': f'C3 gradient-free Stein thinning: d={D} Gaussian mixture (proxy: same mixture, covariances '
                               f'x 1.3), n={n_total}, m={M}, preconditioner=med, precision={args.precision}; {sharding}',
                   'thinning_wall_time_ms': ms_per_step, 'l2': 'flushed between timed steps (512 MiB write)',
                   'step': 'one full thinning job n->m (standardise + median + log-weights + m greedy steps)',
                   'exchange': None if world == 1 else ('peer mailboxes (GFK_P2P=1)' if os.environ.get('GFK_P2P') == '1'
                                                        else 'ncclAllGather of 16 bytes per rank per greedy step (default)')},
        'e2e': {'value': pairs / (e2e_ms_per_step * 1e-3), 'unit': 'pair-evals/s', 'ms_per_step': e2e_ms_per_step,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4 * M, 'bytes_are': 'per rank',
                'copies': 'pinned host sample, log_p, log_q, gradient_q -> device inside the timed region; uint32 index '
                          'vector -> host inside the timed region'},
        'gpu_launches': int(res['launches']),
        'clocks': res['clocks'],
        'roofline': thinning_roofline(args.precision, rows_local, D, True, loop_ms, M, fp64_peak_tflops, traffic=traffic),
    }
    if world > 1:
        line['roofline']['note'] = ('strong scaling of a 1000-step sequential chain: the step is the per-step exchange '
                                    'latency plus 1/N of the update, so the FP64 fraction falls with N by construction')
    if res.get('other'):
        o = res['other']
        line['f64_fast_mode'] = {'value': pairs / (o['ms_per_step'] * 1e-3), 'unit': 'pair-evals/s',
                                 'ms_per_step': o['ms_per_step'], 'same_indices_as_f64': o['same_indices'],
                                 'roofline': thinning_roofline('f64-fast', rows_local, D, True, o['loop_ms'], M,
                                                               fp64_peak_tflops)}
    if other_scaling is not None:
        key, o, steps = other_scaling
        om = o['dev_ms'] / steps
        line[key] = {'n': int(o['n_total']), 'ms_per_step': om, 'value': float(o['n_total']) * M / (om * 1e-3),
                     'unit': 'pair-evals/s', 'e2e_ms_per_step': o['e2e_ms'] / steps,
                     'greedy_step_us': float(np.mean(o['loop_ms'])) / M * 1e3}
    return line


# ---- our arm ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gradient_free_mcmc_postprocessing_b200 import _lib, distributed as gdist
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    lib = _lib.load()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        gdist.init()                                   # NCCL exchange unless GFK_P2P=1 is set for every rank
    warnings.simplefilter('ignore')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_case(n_total):
        """identical replica on every rank: rank 0 generates, NCCL broadcasts; pinned host copies for e2e"""
        names = ('sample', 'log_p', 'log_q', 'gradient_q')
        shapes = {'sample': (n_total, D), 'log_p': (n_total,), 'log_q': (n_total,), 'gradient_q': (n_total, D)}
        devt = {}
        host_np = make_inputs(n_total) if rank == 0 else None
        for k in names:
            t = torch.from_numpy(host_np[k]).to(dev) if rank == 0 else torch.empty(shapes[k], dtype=torch.float64,
                                                                                     device=dev)
            if world > 1:
                dist.broadcast(t, src=0)
            devt[k] = t
        pinned = {k: devt[k].cpu().pin_memory() for k in names}
        return devt, pinned, host_np

    precision = {'value': args.precision}

    def job(inp, timing=None):
        prob = SteinProblem(inp['sample'], inp['gradient_q'], log_p=inp['log_p'], log_q=inp['log_q'],
                            preconditioner='med', precision=precision['value'], device=dev)
        return prob.thin(M, timing=timing)

    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        _lib.check(lib.gfk_flush_l2(flush_buf.data_ptr(), flush_buf.numel(), torch.cuda.current_stream().cuda_stream))

    def timed_steps(inp, steps, collect_loop=False):
        """sum of per-step CUDA-event times (L2 flushed between steps, outside the brackets); max over ranks"""
        total_ms, loop_ms, idx = 0.0, [], None
        for _ in range(steps):
            flush_l2()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            timing = {} if collect_loop else None
            e0.record()
            idx = job(inp, timing)
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            if collect_loop:
                loop_ms.append(timing['loop_ms'])
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), loop_ms, idx

    def measure(n_total, steps, warmup, with_other=True):
        devt, pinned, host_np = make_case(n_total)
        for _ in range(warmup):
            job(devt)
            job(pinned)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        lib.gfk_launch_count(1)
        dev_ms, loop_ms, idx_dev = timed_steps(devt, steps, collect_loop=True)
        launches = lib.gfk_launch_count(1)
        e2e_ms, _, idx_host = timed_steps(pinned, steps)
        clocks = sampler.result()
        assert np.array_equal(idx_dev, idx_host)
        other = None
        if with_other and world == 1 and args.precision == 'f64':
            # the FMA formulation of the same float64 job, reported next to the strict default
            precision['value'] = 'f64-fast'
            job(devt)
            o_ms, o_loop, o_idx = timed_steps(devt, 2, collect_loop=True)
            precision['value'] = args.precision
            other = dict(ms_per_step=o_ms / 2, loop_ms=float(np.mean(o_loop)),
                         same_indices=bool(np.array_equal(o_idx, idx_dev)))
        return dict(dev_ms=dev_ms, e2e_ms=e2e_ms, loop_ms=loop_ms, launches=launches, clocks=clocks, idx=idx_dev,
                    host_np=host_np, n_total=n_total, other=other)

    n_total = N_C3 * world if args.scaling == 'weak' else N_C3
    res = measure(n_total, args.steps, args.warmup)
    other_scaling = None
    if world > 1 and not args.no_secondary:
        # the scaling mode that is not the headline, with fewer steps
        steps2 = max(2, args.steps // 2)
        if args.scaling == 'strong':
            other_scaling = ('weak', measure(N_C3 * world, steps2, 2, with_other=False), steps2)
        else:
            other_scaling = ('strong', measure(N_C3, steps2, 2, with_other=False), steps2)

    # measured FP64 FMA peak (not in MEASURED_PEAKS.json)
    import ctypes
    fp64 = ctypes.c_double()
    _lib.check(lib.gfk_fma_peak(0, ctypes.byref(fp64), None))

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if world == 1 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f'thin_step:{args.precision}:bytes_per_launch')
        line = assemble_line(args, world, res, fp64.value, other_scaling=other_scaling, traffic=traffic)
        if world == 1 and not args.no_secondary:
            line.update(secondary_in_subprocess(fp64.value))   # thin_c2, ksd_c4, thin_c5, thin_c3_gaussian_proxy
        if world == 1 and not args.no_cpu_baseline:
            m_cpu = args.cpu_baseline_m
            v, dt = cpu_oracle_throughput(res['host_np'], m_cpu, 1)
            from oracle.stein_thinning_oracle.thinning import thin_gf as oracle_thin_gf  # checker
            line['cpu_baseline'] = {'value': v, 'unit': 'pair-evals/s', 'cores': 1, 'kind': 'port',
                                    'host_cores_available': os.cpu_count(), 'seconds': dt,
                                    'sample': f'numpy oracle thin_gf on the same n={n_total} d={D} inputs, first '
                                              f'{m_cpu} of {M} greedy steps (throughput is m-independent)'}
            h = res['host_np']
            expect = oracle_thin_gf(h['sample'], h['log_p'], h['log_q'], h['gradient_q'], 8, preconditioner='med')
            line['parity_prefix_ok'] = bool(np.array_equal(expect, res['idx'][:8]))
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        gdist.finalize()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.secondary_only is not None:
        import torch
        warnings.simplefilter('ignore')
        torch.cuda.set_device(0)
        dev0 = torch.device('cuda', 0)
        peak = args.secondary_only
        print(json.dumps({'thin_c2': measure_thin_c2(dev0, peak), 'ksd_c4': measure_ksd_c4(dev0, peak),
                          'thin_c5': measure_thin_c5(dev0), 'thin_c3_gaussian_proxy': measure_thin_c3_gaussian(dev0, peak)}),
              flush=True)
        return
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
