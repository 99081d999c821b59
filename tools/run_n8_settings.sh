set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline"
$R --no-secondary > gpurun_out/n8_default.json 2> gpurun_out/n8_default.err
NCCL_MAX_CTAS=8 $R --no-secondary > gpurun_out/n8_ctas8.json 2> gpurun_out/n8_ctas8.err
NCCL_MAX_CTAS=16 $R --no-secondary > gpurun_out/n8_ctas16.json 2> gpurun_out/n8_ctas16.err
$R --no-secondary --grad-comm bf16 > gpurun_out/n8_bf16.json 2> gpurun_out/n8_bf16.err
$R > gpurun_out/n8_full.json 2> gpurun_out/n8_full.err
