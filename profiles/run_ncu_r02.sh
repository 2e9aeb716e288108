#!/bin/bash
# profiles/run_ncu_r02.sh -- round-2 ncu passes, run under gpurun (numbers printed by profiled runs are NOT bench values).
#   gpurun_out/r02_early.csv     every launch of ONE early block (block 10: 90 sync segments): duration + DRAM bytes
#   gpurun_out/r02_steady.csv    every launch of ONE steady-state block (block 104: one 51 k-read segment + its sync): duration, DRAM bytes,
#                                L1/L2 sector counters (random-access sector efficiency), achieved occupancy
#   gpurun_out/r02_lookup.ncu-rep, r02_rough.ncu-rep   `--set full` of k_lookup (the kernel that carries the lookup half of the algorithmic
#                                bytes) and of k_rough (the longest kernel of an early segment), one launch each, with source lines
# Only the block between cudaProfilerStart/Stop is profiled (--profile-from-start off); the blocks before it run natively.
set -x
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --no-compress-e2e --no-phase-events --no-parity-check --steps 20 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_early.csv python bench.py $COMMON --max-blocks 12 --profile-block 10 > gpurun_out/r02_early.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
M=$M,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_steady.csv python bench.py $COMMON --max-blocks 106 --profile-block 104 > gpurun_out/r02_steady.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_lookup -c 1 -f -o gpurun_out/r02_lookup \
    python bench.py $COMMON --max-blocks 106 --profile-block 104 > gpurun_out/r02_lookup.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_rough -c 1 -f -o gpurun_out/r02_rough \
    python bench.py $COMMON --max-blocks 12 --profile-block 10 > gpurun_out/r02_rough.log 2>&1
ls -la gpurun_out | tail -8
