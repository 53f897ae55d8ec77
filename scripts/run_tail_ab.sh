# A/B of the tail render modes on one box.  Output: gpurun_out/tail_ab.txt
mkdir -p gpurun_out
out=gpurun_out/tail_ab.txt
: > $out
run() {
  env $1 python bench.py --no-clocks --no-cpu $2 2>>gpurun_out/tail_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$1 $2', '| value %.0f e2e %.0f step_ms %.3f render_ms %s launches %d' % (d['value'], d['e2e']['value'], r['kernel_ms'], (r['render_kernel'] or {}).get('kernel_ms'), d['gpu_launches']))" >> $out
}
run MOOG_TAIL_RENDER=0 "--e2e-frames chunked"
run MOOG_TAIL_RENDER=2 "--e2e-frames mapped"
run "MOOG_TAIL_RENDER=2 MOOG_TAIL_BUSY_THR=2" "--e2e-frames mapped"
run "MOOG_TAIL_RENDER=2 MOOG_TAIL_BUSY_THR=3" "--e2e-frames mapped"
run MOOG_TAIL_RENDER=2 "--e2e-frames mapped"
run "MOOG_TAIL_RENDER=2 MOOG_TAIL_BUSY_THR=2" "--e2e-frames mapped"
cat $out
