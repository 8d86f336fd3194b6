# One compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash tools/probe/sanitize.sh memcheck|racecheck
TOOL=$1
set -x
timeout 300 python tools/probe/sanitize_target.py big > gpurun_out/r2_san_plain.log 2>&1 || { tail -5 gpurun_out/r2_san_plain.log; exit 1; }
# three lanes per item (default for these batch sizes)
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/probe/sanitize_target.py > gpurun_out/r2_san_${TOOL}_trio.log 2>&1
tail -4 gpurun_out/r2_san_${TOOL}_trio.log
# one item per thread (small-batch shapes) and the 448-thread two-launch kernels
BN254V_TRIO_MAX=0 timeout 2400 compute-sanitizer --tool $TOOL --print-limit 20 python tools/probe/sanitize_target.py big > gpurun_out/r2_san_${TOOL}_thread.log 2>&1
tail -4 gpurun_out/r2_san_${TOOL}_thread.log
