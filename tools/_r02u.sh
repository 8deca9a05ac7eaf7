mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 60 --no-parity --no-variants --no-configs --no-cpu-baseline > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; tail -3 gpurun_out/r02u_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/r02u_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"]); r=d["roofline"]; print("frac", r["frac"], "frame", r["frame"]["frac"], r["frame"]["terms"]); print(r["kernels"]["k_shade"]["what"])
P
