# JPEG-path check on a GPU box: parity tests, the bench's JPEG legs, the per-family device times and an ncu launch list
python -m pytest tests/test_jpeg.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_j.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
e=d['e2e']; j=e['jpeg']
print('value',round(d['value']), 'e2e', round(e['value']), 'jpeg', round(j['value']), 'single', round(j['one_call_at_a_time']), 'host', round(j['host_huffman']['value']), 'ingest', round(e['ingest']['value']), 'worker', round(e['worker']['value']))
PY
python tools/jpeg_profile.py 512 2>&1 | grep -E "^batch|jpeg_|device total"
ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__inst_executed.sum --clock-control none -k regex:"jhuff|jpeg" -s 100 -c 20 --csv --log-file gpurun_out/jhuff_launches.csv python tools/jpeg_profile.py 256 > gpurun_out/jp.log 2>&1
