N=${N:-2}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-parity-check "$@" 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1:], d['t_eff_per_gpu'], d['roofline']['us_per_launch'])" "$@"; }
run --opt p2p_halo=1
run --opt p2p_halo=1 --opt graphs=0
run --opt p2p_halo=0
