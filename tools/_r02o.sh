mkdir -p gpurun_out
for nb in 2 4; do
  SLV_BENCH_NBUF=$nb python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 --no-parity --no-variants --no-configs --no-cpu-baseline > gpurun_out/r02o_bench4_nbuf$nb.json 2> gpurun_out/r02o_bench4_nbuf$nb.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r02o_bench4_nbuf$nb.json").read().strip().splitlines()[-1])
print("nbuf $nb: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
P
done
