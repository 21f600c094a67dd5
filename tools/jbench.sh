python -m pytest tests/test_jpeg.py -x -q -m gpu 2>&1 | tail -3
for c in 64 128 256; do
  UF_JPEG_CHUNK=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_j$c.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_j$c.json').read().strip().splitlines()[-1])
j=d['e2e']['jpeg']
print($c, 'value',round(d['value']), 'e2e', round(d['e2e']['value']), 'jpeg', round(j['value']), 'inflight', round(j['calls_in_flight']), 'single', round(j['one_call_at_a_time']), 'host', round(j['host_huffman']['value']))
PY
done
