run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err; grep -h "^{" gpurun_out/r2p_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    b=json.loads(l); print('$tag', round(b['ms_per_step'],3), round(b['value'],1), min(b.get('rank_ms_per_step',[0])))
"; }
run base A=1
run bucket1000 SHM_DP_BUCKET_MB=1000
run nooverlap1000 SHM_DP_BUCKET_MB=1000 SHM_DP_OVERLAP=0
run nooverlap25 SHM_DP_OVERLAP=0
run chan8 NCCL_MAX_NCHANNELS=8
run noreduce SHM_DP_NOREDUCE=1
