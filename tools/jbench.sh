# JPEG-path check on a GPU box: parity tests, the bench's JPEG leg, the per-family device times and an ncu launch list
python -m pytest tests/test_jpeg.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_j.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
j=d['e2e']['jpeg']
print('value',round(d['value']), 'e2e', round(d['e2e']['value']), 'jpeg', round(j['value']), 'inflight', round(j['calls_in_flight']), 'single', round(j['one_call_at_a_time']), 'host', round(j['host_huffman']['value']))
PY
python tools/jpeg_profile.py 512 2>&1 | tail -18
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"jhuff|jpeg" -s 100 -c 20 --csv --log-file gpurun_out/jhuff_launches.csv python tools/jpeg_profile.py 256 > gpurun_out/jp.log 2>&1
