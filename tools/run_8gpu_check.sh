set -x
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; nproc >> gpurun_out/r02_topo_8gpu.txt
(timeout 600 python -m pytest tests/test_gpu_multi.py -q) 2>&1 | tail -3
for n in 1 2 4 8; do python tools/pcie_floor_multi.py $n; done > gpurun_out/r02_pcie_floor_multi.txt 2>&1
cat gpurun_out/r02_pcie_floor_multi.txt
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_D_${n}gpu.json 2> gpurun_out/r02_bench_D_${n}gpu.err
tail -2 gpurun_out/r02_bench_D_${n}gpu.err
done
python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_8.json 2>&1
