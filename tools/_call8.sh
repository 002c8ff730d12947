set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/run_rowband.py --H 2000 --W 3000 --D 400 --preset mb:fast --check --iters 3 --cheap-inputs > gpurun_out/r2_final_rowband_mb_n8.json 2> gpurun_out/r2_final_rowband_mb_n8.err
echo rc=$?
cat gpurun_out/r2_final_rowband_mb_n8.json
grep -v -i "warn" gpurun_out/r2_final_rowband_mb_n8.err | tail -3
