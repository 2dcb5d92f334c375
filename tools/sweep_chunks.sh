#!/bin/bash
# e2e throughput of b2t_encode_batch vs host-path chunk size (B2T_CHUNK_BYTES); developer tool
for cb in 16777216 33554432 67108864 134217728 268435456; do
  B2T_CHUNK_BYTES=$cb timeout 300 python bench.py --no-cpu --steps 4 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('chunk_bytes', $cb, 'e2e GB/s', round(d['e2e']['value'],2), 'ms', round(d['e2e']['ms_per_step'],2))"
done
