#!/bin/bash
# Round-end evidence: peaks, full GPU test suite, smoke, benches, ncu launch list (+ DRAM bytes)
# of the bench command and ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
timeout 120 python - > gpurun_out/microbench.json 2> gpurun_out/microbench.err <<'PY'
import json
from bayeslim_b200 import _lib
out = dict(device=_lib.device_info(0))
for kind, it in (("fp32", 4096), ("fp32x2", 4096), ("rf3_fp32", 4096), ("rf3_fp32x2", 4096), ("mix_rot_mac", 4096), ("mix_mac", 4096), ("mix_rot", 4096), ("fp64", 1024), ("mufu", 2048)):
    g, ms = _lib.microbench(kind, it)
    out[kind] = dict(gops=g, ms=ms)
print(json.dumps(out))
PY
echo "microbench exit $?"
timeout 120 bayeslim_b200/csrc/tools/macbench > gpurun_out/macbench.jsonl 2>&1; echo "macbench exit $?"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-300 gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err
B200RIME_ANT=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default_blowned.json 2> gpurun_out/bench_default_blowned.err
echo "bench default (baseline-owned kernels) exit $?"; cut -c1-200 gpurun_out/bench_default_blowned.json
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cut -c1-200 gpurun_out/bench_c2.json
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "bench c1 exit $?"; cut -c1-200 gpurun_out/bench_c1.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference exit $?"; cut -c1-200 gpurun_out/bench_reference.json
CMD="python bench.py --workload c3 --nt 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
if [ "$1" = "full" ]; then
for which in fwd bwd; do
  timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ant_fringe_$which -s 1 -c 1 \
      -o gpurun_out/prof_bench_$which $CMD > gpurun_out/ncu_bench_$which.log 2>&1
  echo "ncu full $which exit $?"; tail -n 2 gpurun_out/ncu_bench_$which.log
done
fi
