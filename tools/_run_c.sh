python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_forward_profile.py 6 > gpurun_out/sharded_profile.txt 2>&1
grep -v "^\*\*\*\|NCCL\|OMP_NUM" gpurun_out/sharded_profile.txt | head -70
