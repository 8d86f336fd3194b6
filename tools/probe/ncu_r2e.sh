# Last build of the round (dedicated Fq squaring in the G1 arithmetic): launch list of the default bench and the
# captures of the kernels that contain G1 arithmetic (Groth16 prepare, the two MSM term kernels at 2^16 and 2^14).
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
G="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
P16="python tools/probe/plonk_only.py 16 2"
P14="python tools/probe/plonk_only.py 14 2"
timeout 600 $B > gpurun_out/r2e_plain_bench.log 2>&1 || exit 1
timeout 300 $P16 > gpurun_out/r2e_plain_p16.log 2>&1 || exit 1
timeout 300 $P14 > gpurun_out/r2e_plain_p14.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2e_launches.csv $B > gpurun_out/r2e_ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/r2e_ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/r2e_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/r2e_${name}_src.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap prepare k_groth16_prepare 3 $G
cap terms0 k_plonk_terms 2 $P16
cap terms1 k_plonk_terms 3 $P16
cap terms0_14 k_plonk_terms 2 $P14
ls gpurun_out | grep r2e | head -30
