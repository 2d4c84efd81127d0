run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err; echo "$name rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2r_$name.json | head -1)"; }
run simple NCCL_PROTO=Simple
run nvls NCCL_ALGO=NVLS
