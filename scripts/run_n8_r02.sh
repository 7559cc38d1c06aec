# Round-2 8-GPU measurements (one box, one process per GPU): headline scaling point, C5, C3, C4.
#   bash scripts/run_n8_r02.sh          everything
#   bash scripts/run_n8_r02.sh flat     only the flat runs (batch-1 headline + C5)
mkdir -p gpurun_out/n8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8/bench_n8.json 2> gpurun_out/n8/bench_n8.err
timeout 600 $T bench.py --gpus 8 --rows 200000000 --batch 1024 --mixed --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n8/c5_flat_200Mx768_batch1024.json 2> gpurun_out/n8/c5.err
if [ "$1" != "flat" ]; then
timeout 600 $T scripts/bench_configs.py --config c3 > gpurun_out/n8/c3_ivf_50Mx512.jsonl 2> gpurun_out/n8/c3.err
timeout 400 $T scripts/bench_configs.py --config c4 --niter 10 > gpurun_out/n8/c4_kmeans_10Mx1024.jsonl 2> gpurun_out/n8/c4.err
fi
head -c 400 gpurun_out/n8/bench_n8.json; echo; head -c 400 gpurun_out/n8/c5_flat_200Mx768_batch1024.json; echo
