import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["one_call_at_a_time"]["value"]), d["latency_batch1_ms"])
for k in d["roofline"]["kernels"]: print(" ", k["name"], round(k["ms_per_step"],4), round(k["alg_GBps"]))
