# A/B of experiment builds: every altlibs/*.so (built with BN254V_LIB=... BN254V_NVCC_EXTRA=... python build.py), then the
# in-tree library.  The extra flags are read back from the .flags stamp so that the staleness check accepts the build.
# Usage: bash tools/probe/ab.sh [extra bench.py flags]
for l in $(ls altlibs/*.so 2>/dev/null) snark-bn254-verifier_b200/libbn254v.so; do
  extra=$(sed 's/.*-fPIC//' $l.flags 2>/dev/null)
  echo "$l [$extra]"
  BN254V_LIB=$PWD/$l BN254V_NVCC_EXTRA="$extra" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('step ms', d['ms_per_step'], 'proofs/s', d['value'], 'miller', r['kernel_ms_per_launch'], 'prepare', r['step']['prepare_ms_per_launch'], 'finish', r['step']['finish_ms_per_launch'], 'frac', r['frac'])"; done
