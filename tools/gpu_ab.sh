#!/bin/bash
# A/B of builds of libb2t.so through the whole bench line: gpu_ab.sh <tag> <lib1> <lib2> ...  (device-resident + e2e legs for each library)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for lib in "$@"; do
  name=$(basename $lib .so)
  B2T_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --no-configs --steps 4 --warmup 3 > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  python -c "
import json
d=json.load(open('$out/${tag}_${name}.json')); print('$name', 'dev', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'ids', round(d['e2e_ids_only']['value'],2), 'dense', round(d['e2e_dense_128']['value'],2))"
done
