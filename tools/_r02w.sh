mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu) > gpurun_out/r02w_pytest.log 2>&1
tail -6 gpurun_out/r02w_pytest.log
for bt in 0 1; do
  echo "== SLV_BIG_TILES=$bt"
  SLV_BIG_TILES=$bt FRAMES=300 REPS=2 timeout 300 python tools/frame_times.py 2>&1 | grep -v "^frame [1-6]"
  SLV_BIG_TILES=$bt SHARD=3,8 FRAMES=400 REPS=3 timeout 300 python tools/frame_times.py 2>&1 | grep -v "^frame [1-6]"
  SLV_BIG_TILES=$bt SHARD=0,8 FRAMES=400 REPS=2 NO_STAGES=1 timeout 300 python tools/frame_times.py 2>&1
  SLV_BIG_TILES=$bt SHARD=1,4 FRAMES=400 REPS=2 NO_STAGES=1 timeout 300 python tools/frame_times.py 2>&1
  SLV_BIG_TILES=$bt timeout 300 python tools/stress_10m.py 2>&1 | tail -2
done > gpurun_out/r02w_bigtiles.txt 2>&1
cat gpurun_out/r02w_bigtiles.txt
