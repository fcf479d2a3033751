#!/bin/bash
# Round-2 evidence run on one B200 (run under gpurun): tests, smoke, headline bench, reference arm,
# the other BASELINE configs, per-method lines, ncu DRAM-traffic capture of the dominant kernel.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/unpinned_rows_measured.log
nvidia-smi --query-gpu=name,memory.total,driver_version --format=csv > $O/r2_box.txt; nproc >> $O/r2_box.txt
(time python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
python __graft_entry__.py --smoke > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2_smoke.log
python bench.py --steps 5 --warmup 3 > $O/bench_r2_default_1gpu.json 2> $O/bench_r2_default_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
python bench.py --method gptq --model llama3-8b --bits 3 --steps 2 --warmup 2 > $O/bench_r2_llama3_8b_gptq_w3.json 2>/dev/null
python bench.py --method gptq --model opt-125m --steps 3 --warmup 3 > $O/bench_r2_opt125m_gptq.json 2>/dev/null
: > $O/bench_r2_matrix_sweep.jsonl
for M in 4096x4096 11008x4096 4096x11008 14336x4096 8192x8192 28672x8192 8192x28672; do
  python bench.py --model matrix-$M --steps 3 --warmup 3 --no-cpu-baseline >> $O/bench_r2_matrix_sweep.jsonl 2>/dev/null
done
: > $O/bench_r2_methods_1gpu.jsonl
for m in awq_fixed gptq_fast smoothquant pot apot smoothquant_search; do
  python bench.py --method $m --steps 2 --warmup 3 >> $O/bench_r2_methods_1gpu.jsonl 2>/dev/null
done
python tools/prof_hessian_traffic.py > $O/r2_traffic_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:hessian_gemm --csv --log-file $O/r2_hessian_traffic.csv python tools/prof_hessian_traffic.py > /dev/null 2>&1
tail -4 $O/r2_pytest_gpu.log; tail -2 $O/r2_smoke.log
