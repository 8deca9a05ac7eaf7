mkdir -p gpurun_out
timeout 900 python tools/variant_sweep.py > gpurun_out/r02n_sweep.txt 2>&1
cat gpurun_out/r02n_sweep.txt
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/r02n_pytest.log 2>&1
tail -6 gpurun_out/r02n_pytest.log
