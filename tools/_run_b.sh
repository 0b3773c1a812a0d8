set -x
(timeout 1200 python -m pytest tests/test_gpu_quantsim.py tests/test_gpu_bench_shape.py tests/test_gpu_stats_multi.py tests/test_gpu_distributed.py tests/test_adaround.py tests/test_auto_quant.py tests/test_quant_analyzer.py tests/test_gpu_train_graph.py tests/test_gpu_ddp_qat.py tests/test_gpu_percentile.py tests/test_gpu_parity.py tests/test_broadcast_qdq.py -x -q -m gpu 2>&1 | tail -6) > gpurun_out/pytest_b.log 2>&1
cat gpurun_out/pytest_b.log
python tools/phase_profile.py 8 2>&1 | head -3
S="--no-kernels --no-cpu-baseline --no-reference-python --no-strong"
for st in 8 32; do
  python bench.py --steps $st --warmup 3 $S > gpurun_out/bench_g_s$st.json 2> gpurun_out/bench_g_s$st.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_g_s$st.json').read().strip().splitlines()[-1]); print('steps$st', d['value'], d['ms_per_step'], d['e2e']['value'], d['encodings_sha256'][:12], d['parity'].get('equals_oracle_checked_golden'), d.get('forward_tf32'))"
done
python tools/job_timeline.py 8
python tools/kernel_sweep.py --sizes-mb 4096 --filter per_channel > gpurun_out/kernels_4g.txt 2>&1; grep bf16 gpurun_out/kernels_4g.txt
