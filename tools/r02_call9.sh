#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/k1_variants.py 10000 1000000 0,512,1024,2048,3072,0 > gpurun_out/r02_s9_variants.log 2>&1
for f in 0 1024 2048; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:table_gram_kernel3 -c 1 --csv --log-file gpurun_out/r02_s9_ncu_f$f.csv \
    python tools/k1_variants.py --one $f 10000 1000000 > /dev/null 2>&1
done
cat gpurun_out/r02_s9_variants.log; grep -h "table_gram" gpurun_out/r02_s9_ncu_f*.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
