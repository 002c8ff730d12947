set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_feature_tower.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/run_rowband.py --H 2000 --W 3000 --D 400 --preset mb:fast --check --iters 2 --cheap-inputs 2>/dev/null | tail -1
