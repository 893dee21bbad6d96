#!/usr/bin/env bash
# Builds libgfk.so for sm_100a in-tree (next to this script's parent package).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libgfk.so"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared ${GFK_NVCC_EXTRA:-} \
     -o "${out}" "${here}/api.cu" "${here}/comm.cu" "${here}/prep.cu" "${here}/thin.cu" "${here}/ksd.cu" \
     "${here}/proxy.cu" "${here}/energy.cu" -ldl
echo "built ${out}"
