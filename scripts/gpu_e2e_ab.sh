for v in 0 1 0 1; do
  if [ "$v" = "1" ]; then export PBRGPU_PLAIN_READBACK=1; else unset PBRGPU_PLAIN_READBACK; fi
  python bench.py --spp 128 --steps 8 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
v=d['value']; e=d['e2e']['value']; n=1920*1080*128/1e6
print('plain=%s 128 spp: value %.1f (%.2f ms)  e2e %.1f (%.2f ms)  gap %.2f ms' % (os.environ.get('PBRGPU_PLAIN_READBACK','0'), v, n/v*1e3, e, n/e*1e3, n/e*1e3-n/v*1e3))
"
done
