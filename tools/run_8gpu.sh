export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29600
TAG=${TAG:-r20}
run() { # name nproc args...
  name=$1; np=$2; shift 2; P=$((P+1))
  timeout 400 $TR --nproc-per-node $np --master-port $P bench.py --gpus $np "$@" > gpurun_out/${TAG}_$name.log 2>&1
  echo "$name rc=$?"
}
run train_8_cg 8 --workload train --no-cpu --solve cg
run train5_8_distributed 8 --workload train5 --no-cpu --solve distributed
run train_4_cg 4 --workload train --no-cpu --solve cg
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
