#!/bin/bash
# One gpurun call for a whole validation + measurement round (every call costs ~2-3 box-minutes before the command
# even starts, so do everything in one):   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r02 [quick]'
# Writes gpurun_out/<tag>_*: pytest log, bench JSON lines (N=1 headline + llama3 / wordpiece), the ncu launch list of
# the bench command and one `ncu --set full` capture each of the pre-tokenization scan and the page kernel.
# tools/refresh_profiles.py <tag> turns the pulled files into the summaries under profiles/.
tag=${1:-rXX}; quick=$2
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench.json
if [ -z "$quick" ]; then
  timeout 200 python bench.py --config llama3 --no-cpu > $out/${tag}_bench_llama3.json 2>> $out/${tag}_bench.err
  timeout 200 python bench.py --config wordpiece --no-cpu > $out/${tag}_bench_wordpiece.json 2>> $out/${tag}_bench.err
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:pretok_lean -c 1 -f -o $out/${tag}_k1 \
  python bench.py --mb 256 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_k1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:model_tile -c 1 -f -o $out/${tag}_k2 \
  python bench.py --mb 256 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_k2.log 2>&1
ls -la $out/${tag}_* | awk '{print $5, $9}'
