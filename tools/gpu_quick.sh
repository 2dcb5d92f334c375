#!/bin/bash
# Quick GPU check: parity tests, one bench line for the streaming K1 and one for the round-1 tiled K1 (A/B), ncu of K1.
tag=${1:-q}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-600 $out/${tag}_bench.json
B2T_K1_TILED=1 timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $out/${tag}_bench_tiled.json 2>> $out/${tag}_bench.err; cut -c1-400 $out/${tag}_bench_tiled.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pretok_lean -c 1 -f -o $out/${tag}_k1 \
  python bench.py --mb 256 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_k1.log 2>&1
ls -la $out/${tag}_* | awk '{print $5, $9}'
