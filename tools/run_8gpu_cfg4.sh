export PYTHONUNBUFFERED=1
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29655 bench.py --gpus 8 --workload train --no-cpu --solve cg > gpurun_out/r22_train_8_cg.log 2>&1
echo rc=$?
