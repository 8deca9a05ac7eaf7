mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/r02l_pytest.log 2>&1
tail -8 gpurun_out/r02l_pytest.log
for vc in 0 1; do SLV_VERTEX_CACHE=$vc timeout 300 python tools/vtf_bench.py 2>&1 | tail -3; done > gpurun_out/r02l_vtf.txt 2>&1
cat gpurun_out/r02l_vtf.txt
