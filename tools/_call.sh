set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 > gpurun_out/r2_s5_tests_full.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_s5_tests_full.log | head -20
timeout 600 python bench.py > gpurun_out/r2_s5_bench_n1.json 2> gpurun_out/r2_s5_bench_n1.err
cut -c1-200 gpurun_out/r2_s5_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_s5_bench_reference_n1.json 2> /dev/null
cut -c1-200 gpurun_out/r2_s5_bench_reference_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_s5_launches_fast.csv python tools/time_pipeline.py --iters 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_s5_launches_exact.csv python tools/time_pipeline.py --iters 1 --exact > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_s5_launches_fast.csv | tail -26
