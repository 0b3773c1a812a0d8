set -x
python tools/profile_kernels.py qdq_bf16 pc > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"per_channel_fast_kernel|per_tensor_kernel" -s 6 -c 3 -o gpurun_out/r2_qdq_bf16 -f python tools/profile_kernels.py qdq_bf16 pc > gpurun_out/ncu_qdq_bf16.log 2>&1
echo rc=$?
ls -la gpurun_out | grep r2_qdq
