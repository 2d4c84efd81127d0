run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err; echo "$name rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_$name.json | head -1)"; }
run base KLAB_X=1
run ctas8 NCCL_MAX_CTAS=8
run ctas32 NCCL_MAX_CTAS=32
run bucket32 KLAB_BUCKET_MB=32
run bucket128 KLAB_BUCKET_MB=128
run static KLAB_DYNAMIC_SCHED=0
run static_res0 KLAB_DYNAMIC_SCHED=0 KLAB_SM_RESERVE=0
run ctas8_b128 NCCL_MAX_CTAS=8 KLAB_BUCKET_MB=128
