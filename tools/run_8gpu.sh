#!/bin/bash
# One 8-GPU session (gpurun --gpus 8): the pure-copy ceiling of the box at 1/2/4/8 GPUs, then bench.py at 8 GPUs for all four
# shipped profiles (BASELINE.json configs[4]).  Results land in gpurun_out/<tag>_*; copy what is to be kept into profiles/.
tag=${1:-r02}
mkdir -p gpurun_out
(nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv; nvidia-smi topo -m; nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)") > gpurun_out/${tag}_8gpu_box.txt 2>&1
tools/build/d2h_probe --gpus 1,2,4,8 --quick --seconds 1.2 > gpurun_out/${tag}_d2h_ceiling.json 2> gpurun_out/${tag}_d2h_ceiling.err
port=29500
for p in XTen GAIIx HiSeq2000 HiSeq2500; do
  port=$((port + 1))
  extra="--no-cpu-baseline --no-file"
  [ "$p" = "XTen" ] || extra="$extra --no-gzip"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --steps 10 --warmup 3 --profile $p $extra > gpurun_out/${tag}_bench8_$p.json 2> gpurun_out/${tag}_bench8_$p.err
done
tail -c 300 gpurun_out/${tag}_bench8_*.json
