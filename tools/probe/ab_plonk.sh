# A/B of experiment builds on the PlonK path: every altlibs/*.so, then the in-tree library.  Usage: bash tools/probe/ab_plonk.sh
for l in $(ls altlibs/*.so 2>/dev/null) snark-bn254-verifier_b200/libbn254v.so; do
  extra=$(sed 's/.*-fPIC//' $l.flags 2>/dev/null)
  echo "== $l [$extra]"
  for lg in 14 16 18; do
    BN254V_LIB=$PWD/$l BN254V_NVCC_EXTRA="$extra" timeout 300 python tools/probe/plonk_only.py $lg 4 2>&1 | tail -1
  done
done
