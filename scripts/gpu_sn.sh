#!/bin/bash
# stream-mode pipeline variants: dense streams and priorities with the FPS stream
cd "$(dirname "$0")/.."
for v in "--dense-streams 1" "--dense-streams 2" "--dense-streams 3" "--dense-streams 2 --prio -1,0" "--dense-streams 2 --prio 0,-1" "--dense-streams 3 --prio -1,0"; do
  echo "$v"
  timeout 200 python bench.py --steps 20 --warmup 5 --no-extras $v 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
