#!/bin/bash
mkdir -p gpurun_out
./build/peaks_r02 > gpurun_out/r02_s5_peaks.json 2> gpurun_out/r02_s5_peaks.err
timeout 600 python tools/k1_variants.py > gpurun_out/r02_s5_variants.log 2>&1
for f in 0 16 32 48 128 144; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:table_gram_kernel3 -c 1 --csv --log-file gpurun_out/r02_s5_ncu_f$f.csv \
    python tools/k1_variants.py --one $f 10000 250000 > /dev/null 2>&1
done
cat gpurun_out/r02_s5_peaks.json; cat gpurun_out/r02_s5_variants.log; grep -h "table_gram" gpurun_out/r02_s5_ncu_f*.csv | cut -d, -f1,13- | head -40
