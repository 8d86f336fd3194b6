# GPU tests + the default bench, printing the headline fields
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/quick_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('step ms', d['ms_per_step'], 'proofs/s', d['value'], 'e2e', d['e2e']['value'], 'miller', r['kernel_ms_per_launch'], 'prepare', r['step']['prepare_ms_per_launch'], 'finish', r['step']['finish_ms_per_launch'], 'frac', r['frac'])
print('plonk ms', d['plonk']['ms_per_step'], 'e2e', d['plonk']['e2e']['value'], 'pairing ms', d['pairing']['ms_per_step'], 'mixed items/s', d['mixed']['value'], 'e2e', d['mixed']['e2e']['value'])"
