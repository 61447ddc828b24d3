#!/bin/bash
# Round-2 closing evidence on one B200: full GPU suite, smoke, bench lines of every workload
# (fwd+bwd and forward-only for the headline C3), the reference CPU arm, the ncu launch list (with
# DRAM bytes) of the default bench command, ncu --set full of the three tensor-core kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/nvsmi.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
run() { # name, args...
  n=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  echo "bench $n rc=$? $(head -c 300 gpurun_out/bench_$n.json)"
}
run c3 
run c3_fwd --pass fwd --no-cpu-baseline
run ref_c3 --impl reference
run c1 --workload c1 --steps 5 --warmup 3 --no-cpu-baseline
run c1_graph --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --graph
run refgpu_c3 --impl reference-gpu
run c2 --workload c2 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --workload c4 --steps 2 --warmup 3 --no-cpu-baseline
run c5 --workload c5 --steps 2 --warmup 3 --no-cpu-baseline
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python scripts/tc_probe.py prof > gpurun_out/tc_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fringe_fwd -c 1 \
    -o gpurun_out/r02_tc_fwd -f python scripts/tc_probe.py prof > gpurun_out/ncu_tc_fwd.log 2>&1
echo "ncu fwd rc=$?"; tail -1 gpurun_out/ncu_tc_fwd.log
timeout 300 python scripts/tc_probe.py profbwd > gpurun_out/tc_profbwd_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fringe_bwd -c 1 \
    -o gpurun_out/r02_tc_bwd -f python scripts/tc_probe.py profbwd > gpurun_out/ncu_tc_bwd.log 2>&1
echo "ncu bwd rc=$?"; tail -1 gpurun_out/ncu_tc_bwd.log
timeout 300 python scripts/alm_probe.py prof > gpurun_out/alm_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:alm_cgemm_kernel -s 2 -c 1 \
    -o gpurun_out/r02_alm_cgemm -f python scripts/alm_probe.py prof > gpurun_out/ncu_alm.log 2>&1
echo "ncu alm rc=$?"; tail -1 gpurun_out/ncu_alm.log
ls -la gpurun_out | tail -30
