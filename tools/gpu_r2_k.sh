#!/bin/bash
mkdir -p gpurun_out
run2() { ( while sleep 2; do ps -eo rss,comm | awk '/python/{s+=$1} END{print "rss_mb", s/1024}'; done > gpurun_out/n2_$4.rss & echo $! > /tmp/sampler.pid ); 
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 $2 $3 > gpurun_out/n2_$4.log 2> gpurun_out/n2_$4.err; echo "$4 rc=$?"; kill $(cat /tmp/sampler.pid) 2>/dev/null; sort -k2 -n gpurun_out/n2_$4.rss | tail -1; tail -c 1500 gpurun_out/n2_$4.log; echo; grep -i "error\|Traceback\|killed\|signal\|bench\]" gpurun_out/n2_$4.err | head -8; }
run2 29521 --only cfg2 cfg2
run2 29522 --only cfg3 cfg3
run2 29523 --only cfg4 cfg4
run2 29524 --only cfg5 cfg5
