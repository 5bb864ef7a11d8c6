N=2
FGB_SHARD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_t$N.json 2> gpurun_out/r2_scale_t$N.err
grep "shard" gpurun_out/r2_scale_t$N.err | tail -4
