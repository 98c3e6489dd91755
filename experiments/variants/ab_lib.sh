# A/B of two builds of the library on the bench workloads: bash experiments/variants/ab_lib.sh <lib A> <lib B>
for lib in "$@"; do for w in bundled_360p 720p30_single 1080p60_multi clips1080p; do
  TRL_LIB_PATH=$PWD/truely-real-time-ai-generated-video-detection-framework-for-social-platforms_b200/$lib python bench.py --workload $w --no-facenet-sweep --no-cpu-baseline --steps 4 2>/dev/null > gpurun_out/ab.json
  python -c "import json; d=json.load(open('gpurun_out/ab.json')); print('$lib', '$w', round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['stages'].items() if k in ('rnet','onet')})"
done; done
