mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "exit $?" >> gpurun_out/bench_g2.err
cat gpurun_out/bench_g2.json | cut -c1-400; tail -5 gpurun_out/bench_g2.err
