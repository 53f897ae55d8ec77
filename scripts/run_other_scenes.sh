# the other BASELINE configs through bench.py (not the headline line): device-resident + e2e
run() { python bench.py --no-clocks --no-cpu "$@" 2>gpurun_out/other.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'scene': d['config']['workload'], 'envs': d['config']['envs_per_gpu'], 'sprites': d['config']['sprites'], 'substeps': d['config']['substeps'], 'image': d['config']['image'], 'value': round(d['value']), 'e2e': round(d['e2e']['value']), 'step_ms': round(d['roofline']['kernel_ms'],3), 'render_ms': round(d['roofline']['render_kernel']['kernel_ms'],3), 'launch': d['config']['step_launch']}))" || tail -3 gpurun_out/other.err; }
run --scene falling_balls20 --envs 4096
run --scene falling_balls20 --envs 32768
run --scene colliding_predators84 --envs 16384 --episode 200 --burn-in 60 --pool 512
run --scene cleanup64 --envs 8192 --episode 200 --burn-in 60 --pool 256
run --scene pacman64 --envs 8192 --episode 200 --burn-in 60 --pool 256
run --scene synthetic32 --envs 4096 --episode 200 --burn-in 60
run --scene synthetic32 --envs 65536 --episode 200 --burn-in 60 --pool 512
