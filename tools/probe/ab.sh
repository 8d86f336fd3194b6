for l in altlibs/lib_nolazy.so snark-bn254-verifier_b200/libbn254v.so; do echo $l; BN254V_LIB=$PWD/$l timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['step']['finish_ms_per_launch'], d['roofline']['frac'], d['plonk']['ms'], d['pairing_product_k4']['ms'], d['e2e']['value'])"; done
