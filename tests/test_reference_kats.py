"""The reference's own known-answer vectors (tests/kat.py) against the CPU
oracle (here, no GPU) and against the CUDA path (-m gpu)."""
import numpy as np
import pytest

from tests import kat

_CASES = None


def _cases():
    global _CASES
    if _CASES is None:
        _CASES = kat.all_cases()
    return _CASES


def _names():
    return [n for n, _ in _cases()]


@pytest.mark.parametrize('name', _names())
def test_oracle_reproduces_reference_kat(name):
    from oracle.oracle import Oracle
    case = dict(_cases())[name]
    prog, arrays = kat.compile_case(case)
    orc = Oracle(prog, arrays)
    orc.post_reset()
    for step in range(1, case['steps'] + 1):
        orc.step(None)
        if step in case['checks']:
            kat.check_state(case, prog, orc.dyn[0], step, name)


def test_oracle_reproduces_rare_collision_branches():
    """_make_disjoint / _position_correction (crossed bars) and the un-negated
    perpendicular (App. A, C7): values recorded from the reference."""
    from oracle.oracle import Oracle
    for name, case in kat.rare_branch_cases():
        prog, arrays = kat.compile_case(case)
        orc = Oracle(prog, arrays)
        orc.post_reset()
        orc.step(None)
        kat.check_rare(case, prog, orc.dyn[0], name)


@pytest.mark.gpu
def test_cuda_reproduces_rare_collision_branches():
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    for name, case in kat.rare_branch_cases():
        prog, arrays = kat.compile_case(case)
        eng = Engine(prog, 1, 'cuda:0')
        eng.state.upload(arrays)
        eng.post_reset()
        eng.env_step(None, auto_reset=False, want_counters=True)
        dev = eng.state.download()
        kat.check_rare(case, prog, dev['dyn'][0], name)
        orc = Oracle(prog, arrays)
        orc.post_reset()
        orc.step(None)
        assert np.array_equal(eng.counters.cpu().numpy()[0, :4], orc.counters[0]), name


@pytest.mark.gpu
def test_cuda_reproduces_reference_kats():
    """All KAT cases at once, one engine per case; the CUDA state must satisfy the
    reference's known answers AND agree with the oracle (bit-exact when nothing
    rotates, 1e-5 relative otherwise)."""
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    from tests import util
    for name, case in _cases():
        prog, arrays = kat.compile_case(case)
        eng = Engine(prog, 1, 'cuda:0')
        eng.state.upload(arrays)
        eng.post_reset()
        orc = Oracle(prog, arrays)
        orc.post_reset()
        rotates = bool(np.any(arrays['dyn'][0, 5] != 0)) or name.startswith(('triangles', 'tether'))
        for step in range(1, case['steps'] + 1):
            if rotates:   # one step at a time from the oracle's state
                st = {k: v.copy() for k, v in orc.arrays().items()}
                eng.state.upload(st)
            eng.env_step(None, auto_reset=False, want_counters=True)
            orc.step(None)
            dev = eng.state.download()
            live = util.live_mask(prog, orc.cnt[0])
            err = max(util.rel_err(dev['dyn'][0][:, live], orc.dyn[0][:, live]),
                      util.rel_err(dev['vtx'][0], orc.vtx[0]))
            assert err <= (util.RTOL if rotates else 0.0), (name, step, err)
            c = eng.counters.cpu().numpy()[0, :4]
            assert np.array_equal(c, orc.counters[0]), (name, step, 'overlap pair set / contacts')
            if step in case['checks']:
                kat.check_state(case, prog, dev['dyn'][0], step, name)


@pytest.mark.reference
def test_reference_own_tests_pass_through_the_shims():
    """Build container only: the UNMODIFIED reference's physics / environment
    KATs run green on top of oracle/shims (the stand-ins for the matplotlib and
    dm_env surface the reference imports), i.e. the shims that recorded
    tests/golden/ compute what the reference's own tests expect."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = '/root/reference'
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1',
               PYTHONPATH=os.pathsep.join([os.path.join(root, 'oracle', 'shims'), ref]))
    files = [os.path.join(ref, 'tests', 'moog', 'physics', 'test_collisions.py'),
             os.path.join(ref, 'tests', 'moog', 'physics', 'test_tether_physics.py'),
             os.path.join(ref, 'tests', 'moog', 'env_wrappers', 'test_simulation.py')]
    out = subprocess.run([sys.executable, '-m', 'pytest', '-p', 'no:cacheprovider', '-q'] + files,
                         cwd='/tmp', env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '24 passed' in out.stdout, out.stdout[-500:]
