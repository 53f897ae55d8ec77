# every BASELINE.json config through bench.py's own arms (same code path as the headline line):
# one JSON line per scene -> profiles/r02_other_scenes.jsonl
OUT=${OUT:-gpurun_out/r02_other_scenes.jsonl}
: > $OUT
run() { python bench.py --no-clocks "$@" 2>gpurun_out/other.err >> $OUT || { echo "FAILED: $@"; tail -5 gpurun_out/other.err; }; }
run --scene falling_balls20
run --scene falling_balls20 --envs 32768 --no-cpu
run --scene colliding_predators84
run --scene cleanup64
run --scene pacman64
run --scene synthetic32
run --scene synthetic32 --state-only
run --scene synthetic32 --envs 4096 --no-cpu
python - <<'PY'
import json, os
for line in open(os.environ.get('OUT', 'gpurun_out/r02_other_scenes.jsonl')):
    d = json.loads(line)
    c = d['config']
    print('%-22s %6d envs %-10s value %9.0f e2e %9.0f (one step per call %9.0f) step %.3f ms render %s ms cpu %s launch %s' % (
        c['workload'], c['envs_per_gpu'], c['image'], d['value'], d['e2e']['value'], d['e2e_one_step_per_call']['value'],
        d['roofline']['kernel_ms'], d['roofline']['render_kernel']['kernel_ms'],
        round(d['cpu_baseline']['value']) if 'cpu_baseline' in d else '-', c['step_launch']))
PY
