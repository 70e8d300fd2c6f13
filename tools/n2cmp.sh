run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 3 --no-other-configs --no-dp-parity "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['name'], 'graph_collective', d['config']['graph_collective'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"; }
run 29501 --workload imagenet_resnet18_w4a4
run 29502 --workload imagenet_resnet18_w4a4 --no-graph-collective
run 29503 --workload cifar100_resnet20_w4a4
run 29504 --workload cifar100_resnet20_w4a4 --no-graph-collective
python bench.py --steps 30 --warmup 3 --no-other-configs --no-baselines --no-cpu-baseline --workload cifar100_resnet20_w4a4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['config']['name'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3))"
