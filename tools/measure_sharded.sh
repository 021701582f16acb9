# strong scaling of one dense SDP over 1, 2, 4 GPUs of one box: bash tools/measure_sharded.sh
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --workload sharded-dense --steps 2 --warmup 3 > gpurun_out/r1b_sharded_dense_n1.json 2> gpurun_out/r1b_sh_n1.err
for N in ${NLIST:-2 4}; do
  DEV=$(seq -s, 0 $((N-1)))
  CUDA_VISIBLE_DEVICES=$DEV $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --workload sharded-dense --steps 2 --warmup 3 > gpurun_out/r1b_sharded_dense_n$N.json 2> gpurun_out/r1b_sh_n$N.err
done
for N in 1 ${NLIST:-2 4}; do cut -c1-170 gpurun_out/r1b_sharded_dense_n$N.json; done
python -m pytest tests/test_gpu_solver.py -m gpu -q -k "sharded or shares" 2>&1 | tail -1
