for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale3_n1.json 2> gpurun_out/scale3_n1.err;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n > gpurun_out/scale3_n$n.json 2> gpurun_out/scale3_n$n.err; fi
  tail -c 300 gpurun_out/scale3_n$n.json | head -c 300; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/mg8c.log 2>&1; tail -3 gpurun_out/mg8c.log
