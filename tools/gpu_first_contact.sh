#!/usr/bin/env bash
# One gpurun call that (re)establishes every GPU fact this repository needs, in the order of what matters most if the
# call is cut short.  Usage (ONE GPU; never under ncu for bench values):
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_contact.sh'
# Everything lands in gpurun_out/; copy what is to be judged into profiles/ afterwards.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.log 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== compute-sanitizer memcheck on the small smoke job (misaligned / out-of-bounds accesses are reported with their source line)"
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_smoke.log
echo "== bench"; timeout 400 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; head -c 600 gpurun_out/bench_n1.log
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"
echo "== probes"; timeout 400 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "rc=$?"
timeout 400 python tools/next_rows_probe.py > gpurun_out/next_rows_probe.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/next_rows_probe.log
echo "== A/B: tensor-memory row store (opt-in, own process + timeout: first run of tcgen05.ld/st on a device)"
AB_TAG=default timeout 200 python tools/ab_probe.py > gpurun_out/ab_tmem.log 2>&1; echo "rc=$?"
AB_TAG=tmem GFK_TMEM=1 timeout 200 python tools/ab_probe.py >> gpurun_out/ab_tmem.log 2>&1; echo "rc=$?"
AB_TAG=device-exp GFK_EXP_DD=0 timeout 200 python tools/ab_probe.py >> gpurun_out/ab_tmem.log 2>&1; echo "rc=$?"
GFK_TMEM=1 timeout 300 python -m pytest tests/test_zy_full_size_golden.py -m gpu -x -q >> gpurun_out/ab_tmem.log 2>&1; echo "rc=$?"
cat gpurun_out/ab_tmem.log | tail -20
echo "== per-phase trace of the persistent kernel (needs: GFK_TRACE=1 bash gradient_free_mcmc_postprocessing_b200/csrc/build.sh)"
if [ -f libgfk_trace.so ]; then
    for cfg in "c3 f64" "c3 f64-fast" "c2 f64"; do
        GFK_LIB=$PWD/libgfk_trace.so timeout 200 python tools/trace_steps.py $cfg >> gpurun_out/trace.log 2>&1
    done
    tail -30 gpurun_out/trace.log
fi
echo "== ncu launch list of the bench"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_ncu.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu --set full: persistent thinning kernel, KSD row sums, d=64 step kernel, KDE, energy"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -c 1 -o gpurun_out/prof_thin_persistent -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/ncu_thin.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ksd_rowsum -c 2 -o gpurun_out/prof_ksd_rowsum -f \
    python tools/ncu_targets.py ksd > gpurun_out/ncu_ksd.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_step_chunked -s 5 -c 1 -o gpurun_out/prof_thin_chunked -f \
    python tools/ncu_targets.py c5 > gpurun_out/ncu_c5.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kde_partial|dist_rowsum" -c 2 -o gpurun_out/prof_next_rows -f \
    python tools/ncu_targets.py next > gpurun_out/ncu_next.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -30
