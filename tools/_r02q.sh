mkdir -p gpurun_out
bash tools/ncu_capture.sh r02q_shard3of8 --shaders sasl --aniso 16 --shard 3,8 --frame 7 2>&1 | tail -2
sed -n 1,200p gpurun_out/r02q_shard3of8_summary.txt | grep -E "^## |duration|issue slots|achieved occupancy|warp instructions|grid  |threads per warp"
