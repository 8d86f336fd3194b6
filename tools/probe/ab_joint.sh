for lg in 11 12 13 14 15; do
  for jm in 1 1000000000; do
    echo "== 2^$lg joint_min $jm"
    BN254V_PLONK_JOINT_MIN=$jm timeout 300 python tools/probe/plonk_only.py $lg 4 2>&1 | tail -1
  done
done
