# A/B of experiment builds: every altlibs/*.so, then the in-tree library (default bench headline fields)
for l in altlibs/*.so snark-bn254-verifier_b200/libbn254v.so; do echo $l; BN254V_LIB=$PWD/$l timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'proofs/s', d['value'], 'miller', d['roofline']['kernel_ms_per_launch'], 'finish', d['roofline']['step']['finish_ms_per_launch'], 'plonk ms', d['plonk']['ms'], 'plonk 2^17 proofs/s', d['plonk']['e2e_proofs_per_sec_2e17_batch'], 'pp4 ms', d['pairing_product_k4']['ms'], 'e2e', d['e2e']['value'])"; done
