# One gpurun call: plain run, launch list, one full capture of the Miller kernel (exports CSV pages, drops the .ncu-rep).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 $CMD > gpurun_out/plain_v8.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v8.csv $CMD > gpurun_out/ncu_l8.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_groth16_miller -s 3 -c 1 -f -o gpurun_out/v8m $CMD > gpurun_out/ncu_v8m.log 2>&1
ncu -i gpurun_out/v8m.ncu-rep --page raw --csv > gpurun_out/v8m_raw.csv 2>/dev/null
ncu -i gpurun_out/v8m.ncu-rep --page source --csv > gpurun_out/v8m_src.csv 2>/dev/null
rm -f gpurun_out/v8m.ncu-rep
timeout 800 ncu --set full --clock-control none -k regex:k_groth16_finish -s 3 -c 1 -f -o gpurun_out/v8f $CMD > gpurun_out/ncu_v8f.log 2>&1
ncu -i gpurun_out/v8f.ncu-rep --page raw --csv > gpurun_out/v8f_raw.csv 2>/dev/null
rm -f gpurun_out/v8f.ncu-rep
ls -la gpurun_out | tail -8
