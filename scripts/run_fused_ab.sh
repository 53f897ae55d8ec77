# A/B of the frame paths on one box: frames drawn inside the step kernel vs the render kernel,
# and the three ways step_to_host brings them to the host.  Output: gpurun_out/fused_ab.txt
mkdir -p gpurun_out
out=gpurun_out/fused_ab.txt
: > $out
for variant in "--render separate --e2e-frames chunked" "--render auto --e2e-frames mapped" "--render auto --e2e-frames device" "--render separate --e2e-frames chunked" "--render auto --e2e-frames mapped"; do
  python bench.py --no-clocks --no-cpu $variant 2>>gpurun_out/fused_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$variant', '| value %.0f e2e %.0f step_ms %.3f render_ms %s launches %d | %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], (r['render_kernel'] or {}).get('kernel_ms'), d['gpu_launches'], d['config']['frames']))" >> $out
done
cat $out
