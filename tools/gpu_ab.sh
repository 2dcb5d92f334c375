#!/bin/bash
# A/B of builds of libb2t.so: gpu_ab.sh <tag> <lib1> <lib2> ...  (kernel times of the default bench for each library)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for lib in "$@"; do
  name=$(basename $lib .so)
  B2T_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  python -c "
import json,sys
d=json.load(open('$out/${tag}_${name}.json')); print('$name', {k: round(v,3) for k,v in d['kernels_ms'].items()}, 'ms/step', round(d['ms_per_step'],2))"
done
