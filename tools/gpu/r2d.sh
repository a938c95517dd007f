#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python tools/run_trsv5.py time 128 2 > gpurun_out/r2d_plainA.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wave_kernel -s 1 -c 1 -f -o gpurun_out/r2d_wave128 python tools/run_trsv5.py time 128 2 > gpurun_out/r2d_ncuA.log 2>&1
echo "A rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum
python tools/run_spmv.py 512 2 kind=dot > gpurun_out/r2d_plainB.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:spmv_win -s 1 -c 1 --csv --log-file gpurun_out/r2d_spmv_dram_blocked.csv python tools/run_spmv.py 512 2 kind=dot > gpurun_out/r2d_ncuB.log 2>&1
echo "B rc=$?"
python tools/run_spmv.py 512 2 kind=dot spmv_l2_mb=100000 > gpurun_out/r2d_plainC.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:spmv_win -s 1 -c 1 --csv --log-file gpurun_out/r2d_spmv_dram_natural.csv python tools/run_spmv.py 512 2 kind=dot spmv_l2_mb=100000 > gpurun_out/r2d_ncuC.log 2>&1
echo "C rc=$?"
tail -n 3 gpurun_out/r2d_spmv_dram_blocked.csv gpurun_out/r2d_spmv_dram_natural.csv
