#!/bin/bash
# 8 GPUs: is the plane exchange of the sharded sliding window limited by NCCL's channels per P2P peer?
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
    bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 2 > gpurun_out/r02_sw8_$name.json 2> gpurun_out/r02_sw8_$name.err
  echo "$name rc=$?"
}
run default X=1
run p2p16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
