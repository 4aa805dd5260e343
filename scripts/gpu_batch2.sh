python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python bench.py --steps 20 --image-size 1024 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python scripts/stage_table.py gpurun_out/bench_2gpu.json gpurun_out/bench_1024.json; tail -3 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_ref.json | cut -c1-600
