# multi-GPU measurement pass (one box, N GPUs): bash tools/measure_multi.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port 29601 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r1b_bench_n$N.json 2> gpurun_out/r1b_bench_n$N.err
head -c 120 gpurun_out/r1b_bench_n$N.json; echo
$TR --nproc-per-node $N --master-port 29602 bench.py --gpus $N --workload sharded-dense --steps 2 --warmup 3 > gpurun_out/r1b_sharded_dense_n$N.json 2> gpurun_out/r1b_sh_n$N.err
cut -c1-200 gpurun_out/r1b_sharded_dense_n$N.json
if [ "$N" = "8" ]; then
  CUDA_VISIBLE_DEVICES=0,1,2,3 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --workload sharded-dense --steps 2 --warmup 3 > gpurun_out/r1b_sharded_dense_n4.json 2> gpurun_out/r1b_sh_n4.err
  cut -c1-200 gpurun_out/r1b_sharded_dense_n4.json
fi
for w in frontier-mkp120 frontier-tt500; do
  $TR --nproc-per-node $N --master-port 29604 bench.py --gpus $N --workload $w --no-cpu-baseline > gpurun_out/r1b_${w}_n$N.json 2> gpurun_out/r1b_fr_n$N.err
  cut -c1-200 gpurun_out/r1b_${w}_n$N.json
done
