#!/bin/sh
# ncu passes of round 2 (run under gpurun on ONE GPU); outputs land in gpurun_out/ and are summarised in profiles/r02_summary.md
#   launch lists (device time per launch; cold-cache and serialised: compare shares)   -> r02_launches_{detect,train}.csv
#   --set full captures of the dominant kernels                                         -> r02_{hog,predict,syrk}.ncu-rep
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_detect.csv \
    python bench.py --no-cpu --no-train --steps 2 --warmup 1 --batch 2048 > gpurun_out/r02_prof_detect.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hog_patch_kernel -s 4 -c 1 -o gpurun_out/r02_hog \
    python bench.py --no-cpu --no-train --steps 1 --warmup 1 --batch 2048 > gpurun_out/r02_prof_hog.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:predict_rows_kernel -s 4 -c 1 -o gpurun_out/r02_predict \
    python bench.py --no-cpu --no-train --steps 1 --warmup 1 --batch 2048 > gpurun_out/r02_prof_predict.log 2>&1
LEVELS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_train.csv \
    python tools/train_once.py > gpurun_out/r02_prof_train.log 2>&1
LEVELS=1 ncu --set full --clock-control none --import-source on -k regex:syrk_tc2_kernel -s 0 -c 1 -o gpurun_out/r02_syrk \
    python tools/train_once.py > gpurun_out/r02_prof_syrk.log 2>&1
ls -la gpurun_out | tail -12
