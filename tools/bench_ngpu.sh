# Developer tool (GPU box): the bench line and the reference arm on N GPUs of one box -> gpurun_out/r02r_bench<N>*.json.   bash tools/bench_ngpu.sh <N>
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02r_bench$N.json 2> gpurun_out/r02r_bench$N.err
python - <<P
import json
d=json.loads(open("gpurun_out/r02r_bench$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "host us", d.get("host_frame_loop_us_per_frame"))
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/r02r_bench${N}_ref.json 2>> gpurun_out/r02r_bench$N.err
tail -c 300 gpurun_out/r02r_bench${N}_ref.json
