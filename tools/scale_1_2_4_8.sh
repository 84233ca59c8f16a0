for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale3_n1.json 2> gpurun_out/scale3_n1.err;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n > gpurun_out/scale3_n$n.json 2> gpurun_out/scale3_n$n.err; fi
  tail -c 300 gpurun_out/scale3_n$n.json | head -c 300; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/mg8c.log 2>&1; tail -3 gpurun_out/mg8c.log
# BASELINE config 5: BoseFS{100,100}, 2^24 walkers sharded over the GPUs, 1000 steps per block, g = 1.0 and 0.33
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) tools/run_configs.py --cfg5 > gpurun_out/cfg5_n$n.json 2> gpurun_out/cfg5_n$n.err
  tail -c 400 gpurun_out/cfg5_n$n.json; echo
done
