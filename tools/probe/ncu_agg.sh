# Aggregate-check captures, one gpurun call: the plain run first (must exit 0), then its launch list and one
# `ncu --set full` capture of the per-proof Miller kernel.  Exports CSV pages, drops the .ncu-rep file.
set -x
A="python tools/probe/agg_only.py 16 2"
timeout 300 $A > gpurun_out/r2_agg_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_groth16_agg -c 80 --csv --log-file gpurun_out/r2_agg_launches.csv $A > gpurun_out/r2_agg_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_groth16_agg_miller -s 1 -c 1 -f -o gpurun_out/aggm $A > gpurun_out/r2_agg_ncu.log 2>&1
ncu -i gpurun_out/aggm.ncu-rep --page raw --csv > gpurun_out/r2_aggm_raw.csv 2>/dev/null
ncu -i gpurun_out/aggm.ncu-rep --page source --csv > gpurun_out/r2_aggm_src.csv 2>/dev/null
rm -f gpurun_out/aggm.ncu-rep
tail -2 gpurun_out/r2_agg_plain.log
