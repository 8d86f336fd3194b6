# Final single-GPU records of the round: GPU tests, smoke, the default bench line and the reference arm.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_gpu_tests.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 >> gpurun_out/final_gpu_tests.txt
timeout 900 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/r2_bench_final_n1.json
timeout 900 python bench.py --impl reference 2>>gpurun_out/final_bench.err | tail -1 > gpurun_out/r2_bench_final_reference_arm.json
cat gpurun_out/final_gpu_tests.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_final_n1.json")); r = d["roofline"]
print("groth16", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", r["frac"], "stages", r["step"]["prepare_ms_per_launch"], r["kernel_ms_per_launch"], r["step"]["finish_ms_per_launch"], "cpu", d["cpu_baseline"])
for k in ("plonk", "pairing", "mixed", "all_valid"):
    x = d[k]; print(k, x["value"], x["ms_per_step"], "e2e", x["e2e"]["value"], "frac", x["roofline"].get("frac"))
print(json.load(open("gpurun_out/r2_bench_final_reference_arm.json")))
PY
