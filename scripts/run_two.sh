run() { timeout 300 python bench.py --no-clocks --no-cpu "$@" 2>gpurun_out/other.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'], d['config']['envs_per_gpu'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'step_ms', round(d['roofline']['kernel_ms'],3), 'render_ms', round(d['roofline']['render_kernel']['kernel_ms'],3))" || tail -3 gpurun_out/other.err; }
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run --scene falling_balls20 --envs 4096
run --scene pacman64 --envs 8192 --episode 200 --burn-in 60 --pool 256
run --scene synthetic32 --envs 65536 --episode 200 --burn-in 60 --pool 512
