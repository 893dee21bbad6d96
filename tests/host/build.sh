#!/usr/bin/env bash
# Host build of the pair-kernel header for the CPU test suite (test infrastructure only).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
g++ -O2 -ffp-contract=off -shared -fPIC -I"${here}/../../gradient_free_mcmc_postprocessing_b200/csrc" \
    -o "${here}/libpair_check.so" "${here}/pair_check.cpp"
g++ -O2 -ffp-contract=off -shared -fPIC -DGFK_TEST_RSQRT_PERTURB -I"${here}/../../gradient_free_mcmc_postprocessing_b200/csrc" \
    -o "${here}/libpair_check_perturbed.so" "${here}/pair_check.cpp"
echo "built ${here}/libpair_check.so"
