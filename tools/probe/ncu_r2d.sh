# After the last change of the round (joint MSM form in groups of five): the new pairing test, the launch list of the
# default bench and the captures of the two MSM term kernels at 2^16 proofs again.
set -x
timeout 600 python -m pytest tests/test_gpu_pairing.py -x -q 2>&1 | tail -2 > gpurun_out/r2d_pairing_tests.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
P16="python tools/probe/plonk_only.py 16 2"
timeout 600 $B > gpurun_out/r2d_plain_bench.log 2>&1 || exit 1
timeout 300 $P16 > gpurun_out/r2d_plain_p16.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches.csv $B > gpurun_out/r2d_ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/r2d_ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/r2d_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/r2d_${name}_src.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap terms0 k_plonk_terms 2 $P16
cap terms1 k_plonk_terms 3 $P16
cat gpurun_out/r2d_pairing_tests.txt
