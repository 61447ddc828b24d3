#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --workload c3 --nt 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu -k regex:build_interp --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/builder_metrics.csv $CMD > gpurun_out/ncu_builder.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_builder.log | cut -c1-200
