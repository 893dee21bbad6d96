#!/usr/bin/env bash
# Host builds for the CPU test suite (test infrastructure only; nothing here is loaded by the product package):
#   libpair_check*.so : the pair-kernel header stein_math.cuh compiled for the host
#   libgfk_sim.so     : the REAL kernel sources of libgfk (csrc/*.cu) compiled as C++ on top of the SIMT emulator
#                       in tests/host/simt, exporting the same C ABI with host pointers as "device" pointers
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
csrc="${here}/../../gradient_free_mcmc_postprocessing_b200/csrc"
g++ -O2 -ffp-contract=off -shared -fPIC -I"${csrc}" -o "${here}/libpair_check.so" "${here}/pair_check.cpp"
g++ -O2 -ffp-contract=off -shared -fPIC -DGFK_TEST_RSQRT_PERTURB -I"${csrc}" \
    -o "${here}/libpair_check_perturbed.so" "${here}/pair_check.cpp"
echo "built ${here}/libpair_check.so"

gen="${here}/_gen"
mkdir -p "${gen}"
objs=()
pids=()
for f in api comm prep thin ksd proxy energy ${GFK_SIM_EXTRA_UNITS:-}; do
    [ -f "${csrc}/${f}.cu" ] || continue
    python "${here}/simt/cu2cpp.py" "${csrc}/${f}.cu" "${gen}/${f}.cpp"
    g++ -std=c++17 -O2 -g -ffp-contract=off -fPIC -DGFK_SIMT_HOST=1 -Wno-unknown-pragmas -Wno-attributes -Wno-psabi ${GFK_SIM_CXXFLAGS:-} \
        -I"${here}/simt/include" -I"${csrc}" -c "${gen}/${f}.cpp" -o "${gen}/${f}.o" &
    pids+=($!)
done
g++ -std=c++17 -O2 -g -fPIC -c "${here}/simt/simt.cpp" -o "${gen}/simt.o" &
pids+=($!)
python "${here}/simt/cu2cpp.py" "${here}/simt/selftest.cu" "${gen}/selftest.cpp"
g++ -std=c++17 -O2 -g -fPIC -DGFK_SIMT_HOST=1 -Wno-unknown-pragmas -I"${here}/simt/include" \
    -c "${gen}/selftest.cpp" -o "${gen}/selftest.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
for f in api comm prep thin ksd proxy energy ${GFK_SIM_EXTRA_UNITS:-} simt selftest; do [ -f "${gen}/${f}.o" ] && objs+=("${gen}/${f}.o"); done
g++ -shared ${GFK_SIM_CXXFLAGS:-} -o "${here}/libgfk_sim.so" "${objs[@]}" -ldl -lpthread -lrt
g++ -std=c++17 -O2 -g -fPIC -shared -o "${here}/libfake_nccl.so" "${here}/simt/fake_nccl.cpp" -lrt -lpthread
echo "built ${here}/libgfk_sim.so"
