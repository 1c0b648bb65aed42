export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29600
run() { # name nproc args...
  name=$1; np=$2; shift 2; P=$((P+1))
  timeout 300 $TR --nproc-per-node $np --master-port $P bench.py --gpus $np "$@" > gpurun_out/r10_$name.log 2>&1
  echo "$name rc=$?"
}
run train_8_cg 8 --workload train --no-cpu --solve cg
run train_8_replicated 8 --workload train --no-cpu --solve replicated
run train5_8_distributed 8 --workload train5 --no-cpu --solve distributed
run train5_8_cg 8 --workload train5 --no-cpu --solve cg
run train5_8_replicated 8 --workload train5 --no-cpu --solve replicated
run train_4_cg 4 --workload train --no-cpu --solve cg
run train5_4_distributed 4 --workload train5 --no-cpu --solve distributed
run detect_8 8 --no-cpu
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv > gpurun_out/r10_smi.log 2>&1
