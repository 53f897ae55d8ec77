run() { python bench.py --no-clocks --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'], d['config']['envs_per_gpu'], 'value %.0f e2e %.0f step_ms %.3f render_ms %.3f call_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['render_kernel']['kernel_ms'], d['roofline']['timed_step_call_ms']))"; }
run
run --scene pacman64 --envs 8192 --episode 200 --burn-in 60 --pool 256
run --scene colliding_predators84 --envs 16384 --episode 200 --burn-in 60 --pool 512
run --scene cleanup64 --envs 8192 --episode 200 --burn-in 60 --pool 256
