bash tools/probe/final_n1.sh > gpurun_out/final_n1.log 2>&1
bash tools/probe/ncu_r2c.sh > gpurun_out/ncu_r2c.log 2>&1
tail -12 gpurun_out/final_n1.log
