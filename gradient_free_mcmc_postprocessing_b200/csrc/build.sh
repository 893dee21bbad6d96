#!/usr/bin/env bash
# Builds libgfk.so for sm_100a in-tree (next to this script's parent package).
#   GFK_TRACE=1 : development build with globaltimer stamps in the persistent kernel -> <repo>/libgfk_trace.so
#                 (GFK_LIB=$PWD/libgfk_trace.so python tools/trace_steps.py c3 f64)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libgfk.so"
extra="${GFK_NVCC_EXTRA:-}"
if [ "${GFK_TRACE:-0}" = "1" ]; then
    out="${here}/../../libgfk_trace.so"
    extra="${extra} -DGFK_TRACE"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared ${extra} \
     -o "${out}" "${here}/api.cu" "${here}/comm.cu" "${here}/prep.cu" "${here}/thin.cu" "${here}/ksd.cu" \
     "${here}/proxy.cu" "${here}/energy.cu" -ldl
echo "built ${out}"
