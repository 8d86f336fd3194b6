set -x
CMD="python tools/probe/plonk_only.py 14 2"
timeout 300 $CMD > gpurun_out/r2_e3_plain.log 2>&1 || exit 1
BN254V_TRIO_MAX=70000 timeout 300 python tools/probe/plonk_only.py 16 2 > gpurun_out/r2_e3_plain16.log 2>&1 || exit 1
BN254V_TRIO_MAX=70000 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_plonk -c 14 --csv --log-file gpurun_out/r2_e3_launches16.csv python tools/probe/plonk_only.py 16 2 > gpurun_out/r2_e3_ncu16.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_plonk_stage_e3 -s 1 -c 1 -f -o gpurun_out/e3 $CMD > gpurun_out/r2_e3_ncu.log 2>&1
ncu -i gpurun_out/e3.ncu-rep --page raw --csv > gpurun_out/r2_e3_raw.csv 2>/dev/null
ncu -i gpurun_out/e3.ncu-rep --page source --csv > gpurun_out/r2_e3_src.csv 2>/dev/null
rm -f gpurun_out/e3.ncu-rep
tail -3 gpurun_out/r2_e3_plain.log gpurun_out/r2_e3_plain16.log
