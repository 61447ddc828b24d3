#!/bin/bash
# round-2 evidence: smoke, bench lines (C3 default, forward-only, C2, C1, C4, C5 at N=1),
# reference arms, then the ncu launch list of the default bench command
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/nvsmi.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
run() { # name, args...
  n=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  echo "bench $n rc=$? $(head -c 300 gpurun_out/bench_$n.json)"
}
run c3 --steps 3 --warmup 3
run c3_fwd --steps 3 --warmup 3 --pass fwd --no-cpu-baseline
run c2 --workload c2 --steps 3 --warmup 3
run c1 --workload c1 --steps 5 --warmup 3
B200RIME_TC=0 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_fp32.json 2> gpurun_out/bench_c3_fp32.err; echo "bench c3_fp32 rc=$?"
run ref_c3 --impl reference --steps 5 --warmup 1
