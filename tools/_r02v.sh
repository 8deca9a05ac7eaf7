mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_geometry_split.py -x -q -m gpu) > gpurun_out/r02v_pytest.log 2>&1
tail -6 gpurun_out/r02v_pytest.log
for gs in 0 1; do
  echo "== SLV_GEOMETRY_SPLIT=$gs"
  SLV_GEOMETRY_SPLIT=$gs FRAMES=300 REPS=2 timeout 300 python tools/frame_times.py 2>&1 | grep -v "^frame [1-7]"
  SLV_GEOMETRY_SPLIT=$gs SHARD=3,8 FRAMES=400 REPS=3 timeout 300 python tools/frame_times.py 2>&1 | grep -v "^frame [1-7]"
  SLV_GEOMETRY_SPLIT=$gs SHARD=0,8 FRAMES=400 REPS=2 NO_STAGES=1 timeout 300 python tools/frame_times.py 2>&1
  SLV_GEOMETRY_SPLIT=$gs SHARD=1,4 FRAMES=400 REPS=2 NO_STAGES=1 timeout 300 python tools/frame_times.py 2>&1
  SLV_GEOMETRY_SPLIT=$gs timeout 300 python tools/stress_10m.py 2>&1 | tail -2
done > gpurun_out/r02v_split.txt 2>&1
cat gpurun_out/r02v_split.txt
