#!/bin/bash
# ncu evidence of one build (run under gpurun; one GPU): per-kernel tensor-pipe / DRAM counters of a whole forward,
# full captures of the stack kernels and the CNN, and the launch list of bench.py.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
set -x
python tools/ncu_forward.py --iters 3 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
M1=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M1 --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_forward_metrics.csv python tools/ncu_forward.py --iters 3 > gpurun_out/${tag}_ncu1.log 2>&1
M2=sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tmem.sum
ncu --metrics $M2 --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_forward_metrics_tensor.csv python tools/ncu_forward.py --iters 3 > gpurun_out/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xformer_stack_kernel -s 3 -c 3 -f -o gpurun_out/${tag}_stack python tools/ncu_forward.py --iters 3 > gpurun_out/${tag}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:visual_cnn_tc -s 1 -c 1 -f -o gpurun_out/${tag}_cnn python tools/ncu_forward.py --iters 3 > gpurun_out/${tag}_ncu4.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out/${tag}_*
