for r in 1 2 5 25; do
  python tools/stream_cfg5.py --scale 0.08 --repeat 5 --regions-per-job $r --check 0 2>/dev/null | tail -1 > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('regions/job', $r, 'jobs', d['jobs'], 'wall %.3f s' % d['wall_s'], 'e2e %.0f GCUPS' % d['e2e_gcups'])"
done
