#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the reference-derived checkers into oracle/_ref/
# from the sources WHERE THEY LIE under the read-only reference checkout ($1, default
# /root/reference).  No reference source is copied into the repository: the one patched
# header (pivot export) is generated into a mktemp dir and deleted.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

# 1. host checker: verify.hpp (pivot flavour), unmodified
g++ -O2 -std=c++17 -I"$REF/parallel_pivot" -shared -fPIC "$HERE/ref_verify_shim.cpp" -o "$OUT/libref_verify.so"

# 2. pivot-exporting patch of the two pivoting kernels: one extra `int* piv_out`
#    argument and one store of the shared-memory `pivots[]` before the write-back loop.
for v in serial_pivot parallel_pivot; do
  mkdir -p "$TMP/$v"
  python3 - "$REF/$v/luBatchedInplace.cuh" "$TMP/$v/luBatchedInplace.cuh" <<'PY'
import re, sys
src = open(sys.argv[1]).read()
sig = re.compile(r"batched_lu_subwarp\(T\*( __restrict__)? A\) \{")
assert len(sig.findall(src)) == 1, "kernel signature not found"
src = sig.sub(lambda m: "batched_lu_subwarp(T*%s A, int* piv_out) {" % (m.group(1) or ""), src)
tail = re.compile(r"(\n[ \t]*#pragma unroll\n[ \t]*for \(int k = threadIdInMatrix; k < numElements; k \+= threadsPerMatrix\) \{\n[ \t]*A\[k \+ mtrxOffset\] = sh_A\[k\];)")
assert len(tail.findall(src)) == 1, "write-back loop not found"
export = "\n        if (threadIdInMatrix < matrixSize) piv_out[globalMatrixId * matrixSize + threadIdInMatrix] = pivots[threadIdInMatrix];"
src = tail.sub(lambda m: export + m.group(1), src)
open(sys.argv[2], "w").write(src)
PY
done

# 3. the kernels, rebuilt for sm_100 with the sweep's arithmetic flags
#    (templated/run.py:37-87: -O3 --use_fast_math --std=c++17 --restrict)
FLAGS=(-O3 --use_fast_math --std=c++17 --restrict -arch=sm_100 -shared -Xcompiler -fPIC -w)
build() { # name incdir T smem_kind extra...
  local name="$1" inc="$2" t="$3" smem="$4"; shift 4
  "$NVCC" "${FLAGS[@]}" -I"$inc" -I"$REF/parallel_pivot" -DREF_T="$t" -DREF_SYM="$name" \
      -DREF_SMEM_KIND="$smem" "$@" "$HERE/ref_kernels_shim.cu" -o "$OUT/lib$name.so"
}
pids=()
for t in float double; do
  s=f32; [ "$t" = double ] && s=f64
  build "ref_none_$s"          "$REF/templated"      "$t" 0 & pids+=($!)
  build "ref_serial_$s"        "$REF/serial_pivot"   "$t" 1 & pids+=($!)
  build "ref_parallel_$s"      "$REF/parallel_pivot" "$t" 2 & pids+=($!)
  build "ref_serial_piv_$s"    "$TMP/serial_pivot"   "$t" 1 -DREF_PIVOUT & pids+=($!)
  build "ref_parallel_piv_$s"  "$TMP/parallel_pivot" "$t" 2 -DREF_PIVOUT & pids+=($!)
  # keep at most 5 nvcc jobs in flight
  for p in "${pids[@]}"; do wait "$p"; done; pids=()
done
ls -la "$OUT"
