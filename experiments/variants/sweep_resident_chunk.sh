for rc in 450 225 150 113 90; do
  python bench.py --no-cpu-baseline --no-facenet-sweep --steps 6 --resident-chunk $rc 2>/dev/null > gpurun_out/rc_$rc.json
  python -c "import json; d=json.load(open('gpurun_out/rc_$rc.json')); print($rc, round(d['value']), round(d['ms_per_step'],3))"
done
