mkdir -p gpurun_out
bash tools/ncu_capture.sh r02x_sasl16 --shaders sasl --aniso 16 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-variants --no-configs > gpurun_out/r02x_launches_bench.log 2>&1
(time python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err)
tail -2 gpurun_out/r02x_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02x_bench_ref.json 2> gpurun_out/r02x_bench_ref.err)
python - <<P
import json
d=json.loads(open("gpurun_out/r02x_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["bit_exact"]); r=d["roofline"]; print("frac", r["frac"], "traffic", r["traffic"], "frame", r["frame"]["frac"]); print(r["stage_ms_per_frame"])
for k,v in (d.get("variants") or {}).items(): print(k, v.get("frames_per_sec") if isinstance(v,dict) else v)
for k,v in (d.get("configs") or {}).items(): print(k, v.get("frames_per_sec"), v.get("error"))
P
