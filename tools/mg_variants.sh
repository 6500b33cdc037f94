N=${N:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-parity-check "$@" 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1:], round(d['t_eff_per_gpu'],1), round(d['ms_per_step'],2), d['pt_iters_per_step'])" "$@"; }
run --opt tb2=1
run --opt tb2=0
run --workload E --fixed-iters 510 --steps 1 --opt tb2=1
run --workload E --fixed-iters 510 --steps 1 --opt tb2=0
