#!/usr/bin/env bash
# Builds libgfk.so for sm_100a in-tree (next to this script's parent package); translation units compile in parallel.
#   GFK_TRACE=1 : development build with globaltimer stamps in the persistent kernel -> <repo>/libgfk_trace.so
#                 (GFK_LIB=$PWD/libgfk_trace.so python tools/trace_steps.py c3 f64)
# The CUDA runtime is linked statically on purpose (the build that ran on a B200 in round 1): a shared-runtime build
# would resolve to the libcudart.so.12 torch has already loaded, a 12.8 / 12.9 combination that has not been tried.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libgfk.so"
extra="${GFK_NVCC_EXTRA:-}"
tag="rel"
if [ "${GFK_TRACE:-0}" = "1" ]; then
    out="${here}/../../libgfk_trace.so"
    extra="${extra} -DGFK_TRACE"
    tag="trace"
fi
obj="${here}/../../build/gfk_${tag}"
mkdir -p "${obj}"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${extra}"
pids=()
for f in api comm prep thin ksd proxy energy; do
    nvcc ${flags} -c "${here}/${f}.cu" -o "${obj}/${f}.o" &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${obj}"/{api,comm,prep,thin,ksd,proxy,energy}.o -ldl
echo "built ${out}"
