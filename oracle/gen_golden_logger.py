"""Records what the UNMODIFIED reference's LoggingEnvironment (moog/env_wrappers/logger.py) writes
for three episodes of the test_simulation.py environment (tests/kat.py SIM_ACTIONS), as the wire-format
fixture of tests/golden/logger_sim_timing.json.  Build container only (needs /root/reference).
`time` stamps are zeroed; everything else is the reference's JSON verbatim."""
import json
import os
import shutil
import sys
import tempfile

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.join(_ROOT, 'oracle')]
sys.path.insert(0, _ROOT)
from oracle import refenv  # noqa: E402

refenv.activate()
import numpy as np  # noqa: E402
from moog import environment  # noqa: E402
from moog.env_wrappers import logger  # noqa: E402
from oracle.gen_golden_episodes import _sim_timing_config, _SIM_ACTIONS  # noqa: E402


def main():
    np.random.seed(0)
    tmp = tempfile.mkdtemp()
    env = logger.LoggingEnvironment(environment.Environment(**_sim_timing_config()), log_dir=tmp,
                                    log_vertices='WHEN_NECESSARY')
    env.reset()
    for episode in range(3):
        for a in _SIM_ACTIONS:
            env.step(a)
        env.step(4)          # the step after a termination is the reset (FIRST)
    (stamp,) = os.listdir(tmp)
    d = os.path.join(tmp, stamp)
    out = {'attributes': json.load(open(os.path.join(d, 'attributes.txt'))),
           'description': open(os.path.join(d, 'description.txt')).read(), 'episodes': []}
    for fn in sorted(f for f in os.listdir(d) if f.isdigit()):
        ep = json.load(open(os.path.join(d, fn)))
        for step in ep:
            step[0][1] = 0.0
        out['episodes'].append(ep)
    shutil.rmtree(tmp)
    path = os.path.join(_ROOT, 'tests', 'golden', 'logger_sim_timing.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('{} episodes, {} steps -> {}'.format(len(out['episodes']), sum(len(e) for e in out['episodes']), path))


if __name__ == '__main__':
    main()
