#!/bin/bash
# Tuning builds of the persistent PCG kernel: one library per resident-CTA target (register cap 32 / 40 / 48 / 64), selected with PF2_LIB.
#   bash tools/build_variants.sh 8 6 5 4   ->  pansfem2_b200/bin/libpf2_minb<N>.so   (git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/../pansfem2_b200/csrc"
mkdir -p ../bin build
python -m pansfem2_b200.build >/dev/null 2>&1 || (cd ../.. && python -m pansfem2_b200.build >/dev/null)
objs=$(ls build/*.o | grep -v "build/pcg.o" | grep -v "pcg_minb")
for m in "$@"; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DPF2_PCG_MINB=$m ${PF2_EXTRA_DEFS:-} -c pcg.cu -o build/pcg_minb$m.o &&
    nvcc -shared -o ../bin/libpf2_minb$m${PF2_TAG:-}.so $objs build/pcg_minb$m.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ldl && echo built minb$m ) &
done
wait
