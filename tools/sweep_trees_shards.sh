for cfg in "16384 2" "16384 4" "16384 8" "32768 4" "32768 8" "32768 16" "65536 8" "131072 8" "131072 16"; do
  set -- $cfg
  python bench.py --games-per-gpu $1 --shards $2 --no-cnn --no-selfplay --no-split --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('G=$1 shards=$2', round(d['value']/1e9,3), 'G sims/s', round(d['ms_per_step'],3), 'ms/step')"
done
