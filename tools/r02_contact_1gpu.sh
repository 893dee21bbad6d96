#!/usr/bin/env bash
# Round-2 re-qualification on ONE GPU, in strictly increasing blast radius.  Every stage is gated on the previous one:
# a sanitizer error, a failed test or a device that stops answering ends the script before anything larger runs.
#   gpurun --timeout 1500 -- 'bash tools/r02_contact_1gpu.sh'
set -u
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L > $O/smi.log 2>&1
alive() {
    timeout 180 python -c "import torch; x = torch.ones(8, device='cuda'); assert float(x.sum()) == 8.0" > $O/alive.log 2>&1
}
stop() { echo "STOP: $1"; nvidia-smi --query-gpu=index,name,clocks_throttle_reasons.active --format=csv >> $O/smi.log 2>&1; exit 1; }
clean() { grep -q "ERROR SUMMARY: 0 errors" "$1"; }
SMOKE='import __graft_entry__ as g; g.smoke()'
SUBSET='golden_g1 or golden_g4 or thin_indices_match_oracle or thin_gf_indices_match_oracle or launch_strategies or m_greater or pair_values or kmat_and_ksd or median_subsample or standardisation'

alive || stop "device does not answer before anything ran"
echo "== A1 memcheck smoke"
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python -c "$SMOKE" > $O/memcheck_smoke.log 2>&1; echo "rc=$?"; tail -3 $O/memcheck_smoke.log
clean $O/memcheck_smoke.log || stop "memcheck reported errors on smoke()"
grep -q "smoke ok" $O/memcheck_smoke.log || stop "smoke() failed under memcheck"
echo "== A2 synccheck smoke"
timeout 600 compute-sanitizer --tool synccheck --print-limit 30 python -c "$SMOKE" > $O/synccheck_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/synccheck_smoke.log
clean $O/synccheck_smoke.log || stop "synccheck reported errors on smoke()"
echo "== A3 racecheck smoke"
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python -c "$SMOKE" > $O/racecheck_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/racecheck_smoke.log
grep -q "RACECHECK SUMMARY: 0 hazards" $O/racecheck_smoke.log || echo "NOTE: racecheck did not print a clean summary (see log)"
alive || stop "device does not answer after the sanitizer runs"
echo "== B memcheck on a parity subset"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py tests/test_zz_next_rows.py -m gpu -x -q -k "$SUBSET or proxy or energy or chain" > $O/memcheck_subset.log 2>&1; echo "rc=$?"; tail -6 $O/memcheck_subset.log
clean $O/memcheck_subset.log || stop "memcheck reported errors on the parity subset"
alive || stop "device does not answer after memcheck subset"
echo "== C smoke + pytest -m gpu (no sanitizer)"
timeout 300 python -c "$SMOKE" > $O/smoke.log 2>&1; echo "rc=$?"; tail -1 $O/smoke.log
grep -q "smoke ok" $O/smoke.log || stop "smoke() failed"
timeout 1500 python -m pytest tests -m gpu -x -q -rxXs > $O/pytest_gpu.log 2>&1; rc=$?; echo "rc=$rc"; tail -12 $O/pytest_gpu.log
alive || stop "device does not answer after pytest"
[ $rc -eq 0 ] || stop "pytest -m gpu failed"
echo "== D bench"
timeout 900 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "rc=$?"; head -c 3000 $O/bench_n1.log; tail -5 $O/bench_n1.err
alive || stop "device does not answer after bench"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.log 2>&1; echo "rc=$?"; head -c 800 $O/bench_ref.log
echo "== E probes"
timeout 600 python tools/gpu_probe.py > $O/probe.log 2>&1; echo "rc=$?"; tail -30 $O/probe.log
alive || stop "device does not answer after probes"
timeout 600 python tools/next_rows_probe.py > $O/next_rows_probe.log 2>&1; echo "rc=$?"; tail -20 $O/next_rows_probe.log
echo "== F opt-in kernels that have never run on a GPU (TMA-fed step kernel, tcgen05 KSD, tensor-memory row store):"
echo "   each in its own process under its own timeout, memcheck first; every wait inside them is deadline-bounded"
OPT='tests/test_zz_next_rows.py::test_tensor_core_and_tma_variants_on_device'
GFK_TEST_OPTIN_KERNELS=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest $OPT -m gpu -x -q --runxfail > $O/memcheck_optin.log 2>&1; echo "rc=$?"; tail -6 $O/memcheck_optin.log
alive || stop "device does not answer after the opt-in kernels under memcheck"
if clean $O/memcheck_optin.log; then
    GFK_TEST_OPTIN_KERNELS=1 timeout 600 python -m pytest $OPT -m gpu -x -q --runxfail > $O/pytest_optin.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_optin.log
    alive || stop "device does not answer after the opt-in kernels"
    AB_TAG=default timeout 200 python tools/ab_probe.py > $O/ab_tmem.log 2>&1; echo "rc=$?"
    AB_TAG=tmem GFK_TMEM=1 timeout 200 python tools/ab_probe.py >> $O/ab_tmem.log 2>&1; echo "rc=$?"
    tail -6 $O/ab_tmem.log
fi
nvidia-smi --query-gpu=index,name,clocks_throttle_reasons.active --format=csv >> $O/smi.log 2>&1
echo "DONE"
