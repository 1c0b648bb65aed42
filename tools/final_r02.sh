# final single-GPU validation of round 2 (run under gpurun); outputs -> gpurun_out/fin_*
export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/fin_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fin_smoke.log 2>&1
python bench.py > gpurun_out/fin_bench.log 2>&1
python bench.py --workload train --no-cpu > gpurun_out/fin_train.log 2>&1
python bench.py --workload train --no-cpu --solve cg > gpurun_out/fin_train_cg.log 2>&1
LEVELS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/fin_launches_train.csv python tools/train_once.py > gpurun_out/fin_prof_train.log 2>&1
LEVELS=1 ncu --set full --clock-control none --import-source on -k regex:syrk_tc2_kernel -s 0 -c 1 -o gpurun_out/fin_syrk python tools/train_once.py > gpurun_out/fin_prof_syrk.log 2>&1
python bench.py --workload train5 --no-cpu > gpurun_out/fin_train5.log 2>&1
tail -2 gpurun_out/fin_gpu_tests.log; cat gpurun_out/fin_smoke.log | tail -2
