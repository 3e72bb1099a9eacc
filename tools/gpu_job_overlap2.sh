set -x
timeout 400 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -5
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/bench_train.py --batch 8 --steps 10 --warmup 4"
timeout 300 $T --overlap > gpurun_out/train_2gpu_overlap.json 2> gpurun_out/train_2gpu_overlap.err; tail -c 1500 gpurun_out/train_2gpu_overlap.json; tail -3 gpurun_out/train_2gpu_overlap.err
timeout 300 $T > gpurun_out/train_2gpu_nooverlap.json 2> gpurun_out/train_2gpu_nooverlap.err; cut -c1-330 gpurun_out/train_2gpu_nooverlap.json
timeout 200 python tools/bench_train.py --batch 8 --steps 10 --warmup 4 > gpurun_out/train_1gpu_overlapcode.json 2>&1; cut -c1-330 gpurun_out/train_1gpu_overlapcode.json
