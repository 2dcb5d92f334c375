#!/bin/bash
# parity tests + A/B of library builds in one call: gpu_ab2.sh <tag> <lib1> <lib2> ...   (gpt2 and llama3 corpora, 512 MB)
tag=$1; shift; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 400 python tools/ab_kernels.py --mb 512 --config gpt2 "$@" > $out/${tag}_ab_gpt2.txt 2>&1; cat $out/${tag}_ab_gpt2.txt | cut -c1-400
timeout 400 python tools/ab_kernels.py --mb 512 --config llama3 "$@" > $out/${tag}_ab_llama3.txt 2>&1; cat $out/${tag}_ab_llama3.txt | cut -c1-400
