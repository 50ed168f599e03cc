#!/bin/bash
# Builds A/B variants of the CartPole step kernel into gymnasium_b200/_ab/ (git-ignored; travels with gpurun):
#   soa4_inline  : round-1 state layout (four 8-byte streams), reset / sincos inlined   (what round 1 shipped)
#   soa4_cold    : round-1 layout, reset / sincos out of line
#   aos2_inline  : two 16-byte streams, inlined
#   (the shipped library is aos2_cold)
# Run on a GPU:  for v in ...; do B2E_LIB_PATH=$PWD/gymnasium_b200/_ab/libb200env_$v.so python scripts/block_sweep.py; done
set -e
cd "$(dirname "$0")/.."
mkdir -p gymnasium_b200/_ab build/ab
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=false -Iinclude -Igymnasium_b200/csrc"
O=build/csrc
build() {  # name, extra flags
  nvcc $FLAGS $2 -c gymnasium_b200/csrc/cartpole.cu -o build/ab/cartpole_$1.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 \
    -o gymnasium_b200/_ab/libb200env_$1.so $O/api.o build/ab/cartpole_$1.o $O/frozenlake.o $O/taxi.o $O/blackjack.o $O/classic.o \
    $O/lunarlander.o $O/humanoid.o $O/hopper.o $O/walker2d.o
}
build soa4_inline "-DB2E_CARTPOLE_AB_SOA4 -DB2E_CARTPOLE_AB_COLD=__forceinline__"
build soa4_cold "-DB2E_CARTPOLE_AB_SOA4"
build aos2_inline "-DB2E_CARTPOLE_AB_COLD=__forceinline__"
ls -la gymnasium_b200/_ab
