#!/bin/bash
# developer shortcut: rebuild only the pivot_mode 3 translation units and relink (a full `make` is the reference build)
set -e
cd "$(dirname "$0")/../matrixinversion_b200/csrc"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp,-O2 $EXTRA"
nvcc $FL -DLUB_T=float -DLUB_TN=f32 -c lub_lapack_inst.cu -o build/lapack_f32.o &
nvcc $FL -DLUB_T=double -DLUB_TN=f64 -c lub_lapack_inst.cu -o build/lapack_f64.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../liblubatched.so build/api.o build/inst_*.o build/lapack_f32.o build/lapack_f64.o build/interleaved_f32.o build/interleaved_f64.o -Xcompiler -fopenmp -lgomp
ls -la ../liblubatched.so
