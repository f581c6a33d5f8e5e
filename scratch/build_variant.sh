#!/bin/bash
# scratch/build_variant.sh <name> "<extra nvcc flags>"  ->  scratch/variants/<name>.so (kernel A/B experiments; not product)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../dualip_b200/csrc"
mkdir -p ../../scratch/variants/obj_$name
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-ffp-contract=off -I../../include $flags"
$NV -c -o ../../scratch/variants/obj_$name/calc.o calc.cu &
for f in agd setup lp; do [ -f ../_lib/obj/$f.o ] || $NV -c -o ../_lib/obj/$f.o $f.cu; done
wait
$NV -shared -o ../../scratch/variants/$name.so ../../scratch/variants/obj_$name/calc.o ../_lib/obj/agd.o ../_lib/obj/setup.o ../_lib/obj/lp.o
echo built scratch/variants/$name.so
