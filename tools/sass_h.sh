#!/bin/bash
# Development aid: compile only the workload-H kernels of native_cnvm.cu and dump the SASS of the RING one.
#   tools/sass_h.sh [extra nvcc flags]   ->  /tmp/h.sass, register count on stdout
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr \
  -DSP_FAST_BUILD "$@" -Xptxas -v -c sponet_b200/csrc/native_cnvm.cu -o /tmp/nc_fast.o 2>&1 | grep -A2 "Li2ELb1ELi2ELb0ELb1E" | grep -E "Used|spill"
FN=$(cuobjdump -sass /tmp/nc_fast.o | grep -o '_ZN[A-Za-z0-9_]*cnvm_native_kernelILi2ELb1ELi2ELb0ELb1EEEvNS_14NativeCnvmArgsE' | head -1)
cuobjdump -sass -fun "$FN" /tmp/nc_fast.o > /tmp/h.sass 2>/dev/null || true
grep -c "^\s*/\*[0-9a-f]*\*/" /tmp/h.sass
