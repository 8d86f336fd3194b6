# Captures of the last round-2 build, one gpurun call: plain runs first (each must exit 0), then the launch list of the
# default bench and one `ncu --set full` capture per kernel family.  Exports CSV pages, drops the .ncu-rep files.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
G="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
P14="python tools/probe/plonk_only.py 14 2"
P16="python tools/probe/plonk_only.py 16 2"
Q="python tools/probe/pairing_only.py 20 2"
A="python tools/probe/agg_only.py 16 2"
timeout 600 $B > gpurun_out/r2c_plain_bench.log 2>&1 || exit 1
timeout 300 $P14 > gpurun_out/r2c_plain_p14.log 2>&1 || exit 1
timeout 300 $P16 > gpurun_out/r2c_plain_p16.log 2>&1 || exit 1
timeout 300 $Q > gpurun_out/r2c_plain_q.log 2>&1 || exit 1
timeout 300 $A > gpurun_out/r2c_plain_agg.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2c_launches.csv $B > gpurun_out/r2c_ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/r2c_ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/r2c_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/r2c_${name}_src.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap miller k_groth16_miller 3 $G
cap finish k_groth16_finish 3 $G
cap prepare k_groth16_prepare 3 $G
cap terms0 k_plonk_terms 2 $P16
cap terms1 k_plonk_terms 3 $P16
cap terms0_14 k_plonk_terms 2 $P14
cap stage_e k_plonk_stage_e 1 $P16
cap stage_e3 k_plonk_stage_e3 1 $P14
cap pairing_miller k_pairing_miller 1 $Q
cap pairing_finish k_pairing_finish 1 $Q
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_groth16_agg -c 80 --csv --log-file gpurun_out/r2c_agg_launches.csv $A > gpurun_out/r2c_ncu_agg_l.log 2>&1
cap agg_prepare k_groth16_agg_prepare 1 $A
ls -la gpurun_out | grep r2c
