# e2e chunk schedule (model.chunk_schedule tail rule) with the round-2 cascade: TRL_SCHED_A / _B / _MIN_DIV / chunk
run() { python bench.py --no-cpu-baseline --no-facenet-sweep --steps 6 --chunk $4 2>/dev/null > gpurun_out/sched.json
  python -c "import json; d=json.load(open('gpurun_out/sched.json')); print('A=$1 B=$2 div=$3 chunk=$4', 'e2e ms', round(d['e2e']['ms_per_step'],3), 'h2d-only', round(d['e2e']['h2d_only_ms_per_step'],3))"; }
TRL_SCHED_A=0.6 TRL_SCHED_B=8 TRL_SCHED_MIN_DIV=4 run 0.6 8 4 90
TRL_SCHED_A=0.4 TRL_SCHED_B=8 TRL_SCHED_MIN_DIV=8 run 0.4 8 8 90
TRL_SCHED_A=0.35 TRL_SCHED_B=6 TRL_SCHED_MIN_DIV=10 run 0.35 6 10 90
TRL_SCHED_A=0.4 TRL_SCHED_B=8 TRL_SCHED_MIN_DIV=8 run 0.4 8 8 120
TRL_SCHED_A=0.4 TRL_SCHED_B=8 TRL_SCHED_MIN_DIV=8 run 0.4 8 8 64
TRL_SCHED_A=0.6 TRL_SCHED_B=8 TRL_SCHED_MIN_DIV=4 run 0.6 8 4 90
