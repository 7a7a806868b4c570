#!/bin/bash
# A tuning build of ONE model's phase pipeline: compile mpcv_phase_inst.cu for model $2 with the extra flags $3.. into
# csrc/build_$1/ and link it with the objects of the main build into mpc_verde_b200/libmpcv_$1.so (load it with
# MPCV_LIB=...).    scripts/build_variant.sh rs8 0 -DMPCV_REPACK_SPLIT=8
set -e
tag=$1; model=$2; shift 2
cd "$(dirname "$0")/../mpc_verde_b200/csrc"
mkdir -p build_$tag
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -DMPCV_INST_MODEL=$model \
     -Xptxas -v -c mpcv_phase_inst.cu -o build_$tag/phase_$model.o 2> build_$tag/phase_$model.ptxas.log
objs=$(ls build/*.o | grep -v "build/phase_$model.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libmpcv_$tag.so $objs build_$tag/phase_$model.o
echo built ../libmpcv_$tag.so
