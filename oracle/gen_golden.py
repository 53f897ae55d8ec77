"""Generate golden vectors by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs
/root/reference, imported through oracle/shims).  Usage:

    python oracle/gen_golden.py            # writes tests/golden/*.npz

For each scene the reference `Environment` (moog/environment.py) is reset and
stepped with seeded random actions.  Recorded per scene:

    blob                 program compiled from the reference's own config objects
    init_*               state record packed from state_initializer()'s output
                         (before the reference steps its rules at reset)
    reset_*              state record after Environment.reset()
    actions[T, A]        action fed at step t
    noise[T, K, ND]      unit uniforms behind RandomForce's draws at step t
    dyn/stat/meta/vtx/cnt[T, ...]  state record after step t
    reward[T], last[T]   TimeStep.reward / TimeStep.last()
    n_calls[T], n_true[T], true_hash[T]
                         Sprite.overlaps_sprite calls of step t: count, number
                         that returned True, order-sensitive hash of the True
                         (slot_a, slot_b) events
    frames[F, H, W, 3]   PILRenderer output at steps frame_steps[F]
"""

import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.join(_ROOT, 'oracle')]
sys.path.insert(0, _ROOT)

from oracle import refenv  # noqa: E402

refenv.activate()

import moog_b200  # noqa: E402,F401  (adds the configs package path)
from moog_b200 import compiler  # noqa: E402
from moog import environment  # noqa: E402

MASK64 = (1 << 64) - 1


def true_event_hash(h, a, b):
    """Same mixing as oracle/moog_oracle.c `overlaps` (True events only)."""
    x = ((a * 1315423911) & 0xFFFFFFFF) ^ ((b * 2654435761) & 0xFFFFFFFF) ^ 1
    return ((h ^ x) * 1099511628211) & MASK64


SCENES = {
    # name: (module, level, seed, T, frame_every)
    'pong': ('moog_demos.example_configs.pong', None, 0, 60, 6),
    'falling_balls': ('moog_demos.example_configs.falling_balls', None, 1, 40, 8),
    'falling_balls20': ('moog_b200.configs.falling_balls20', None, 2, 45, 9),
    'colliding_predators': ('moog_demos.example_configs.colliding_predators', None, 3, 50, 10),
    'predators_arena': ('moog_demos.example_configs.predators_arena', 3, 4, 40, 10),
    'synthetic32': ('moog_b200.configs.synthetic32', None, 5, 30, 10),
    # a ball-ball contact whose back-projected vertex has no valid crossing: the
    # reference's argmax picks the NaN entry (collisions.py:207-209) and the sprite's
    # position becomes NaN for the rest of the episode
    'falling_balls20_nan': ('moog_b200.configs.falling_balls20', None, 297, 8, 2),
    # contact-triggered rules (ConditionalRule x ModifySprites(sample_one), ModifyOnContact),
    # ContactReward with a pair condition, Composite of 3 Joysticks; the agents are
    # steered towards fruits / fountains so that the rules actually fire
    'cleanup': ('moog_demos.example_configs.cleanup', None, 6, 150, 25),
    # TorusGeometry renderer (9 copies per sprite), ModifySprites assigning `position`
    # (np.remainder wrap), ConstantSpeed, RandomForce, Joystick(control_velocity),
    # Reset on `len(state['prey']) == 0`; the agent chases the prey so that it vanishes
    'chase_avoid_torus': ('moog_demos.example_configs.chase_avoid_torus', 0, 7, 120, 5),
    # BASELINE config 4: pacman.get_config(0) with the PILRenderer at 64x64 (BASELINE.json
    # quotes the maze tasks at 64x64; the shipped 256x256 canvas exceeds one CTA's shared
    # memory): RandomMazeWalk (np.random.rand(2, 2) recorded as noise), MazePhysics
    # (constant_speed, turning sprites), Grid(control_velocity), VanishOnContact,
    # ConditionalRule on `state['agent'][0].velocity`
    'pacman': ('moog_demos.example_configs.pacman', 0, 12, 120, 10),
    # TimedRule / TemporaryRule / DelayedRule, KeepNearCenter, FirstPersonAgent renderer
    'timed_center': ('moog_b200.configs.timed_center', None, 13, 40, 3),
    # a shipped first-person config: Tether corrective, KeepNearCenter, ModifyOnContact,
    # FirstPersonAgent renderer, grid-line sprites and the 102-vertex annulus occluder
    # (a draw-only outline, MOOG_MAX_OUTLINE)
    'parallelogram_catch': ('moog_demos.example_configs.parallelogram_catch', 0, 14, 90, 6),
    # pairwise Gravity (incl. its dist == 0 branch), KineticFriction (rest / infinite mass / overshoot),
    # DistanceForce(spring_force_fn), TetherZippedLayers (both modes), SetPosition with inertia
    'forces_zoo': ('moog_b200.configs.forces_zoo', None, 15, 40, 5),
    # modifiers assigning scale / aspect_ratio / angle: the outline is rebuilt from the shape and the
    # rotational inertia compounds (sprite.py:411-424, 516-558)
    'reshape_zoo': ('moog_b200.configs.reshape_zoo', None, 16, 40, 5),
    # Portal (square and circular portal pairs) and ChangeLayer into a layer that starts empty
    'portal_zoo': ('moog_b200.configs.portal_zoo', None, 17, 45, 5),
    # a shipped config whose ContactReward reward_fn branches on `sprite.metadata[...]` (red_green.py:171-176)
    # and whose state initializer rolls the physics forward on the host; the agent waits, then walks into
    # the red response box
    'red_green': ('moog_demos.example_configs.red_green', 1, 18, 200, 10),
    # this repo's small version of the same pattern: an if / elif / else reward over two metadata keys
    'predict_zoo': ('moog_b200.configs.predict_zoo', None, 19, 69, 6),
    # a shipped config whose Reset task's condition and reward_fn pick single sprites out of the state, test
    # overlaps between them and branch on metadata (bounce_box_contact_prediction.py:94-110): a decision tree;
    # translucent occluder (opacity 128), a TimedRule that removes the screen
    'bounce_box': ('moog_demos.example_configs.bounce_box_contact_prediction', True, 20, 120, 8),
    # a shipped config with a rule class of its own (functional_maze.py:18-67 `Booster`: a countdown attribute,
    # an overlap test over a layer, mass / colour changes applied and reverted 60 steps later), Portal,
    # RandomForce, DistanceForce; the agent is steered into a booster, then after the prey
    'functional_maze': ('moog_demos.example_configs.functional_maze', None, 127, 130, 10),
    # a shipped config built on PhaseSequence / Phase (end conditions on meta_state, a duration drawn with
    # np.random.randint at reset), two Fixation rules, a dict meta_state whose 'phase' entry the Reset task
    # reads, a rule class of its own, SetPosition, TetherZippedLayers; the agent fixates the cross, then target 0
    'multi_tracking': ('moog_demos.example_configs.multi_tracking_with_feature', 3, 22, 220, 10),
    # a shipped config whose own rule draws random numbers and does vector algebra on the sprites it loops over
    # (match_to_sample.py:45-71 BeginMotion: np.random.uniform / randint, np.matmul, np.linalg.norm, zip over two
    # layers), PhaseSequence with one-time and continual rules, a Reset condition on a sprite AND meta_state,
    # metadata rewards, TetherZippedLayers with an anchor, infinite and zero masses, transparent sprites
    'match_to_sample': ('moog_demos.example_configs.match_to_sample', 4, 23, 220, 10),
}


def _sample_one_rules(rules):
    """ModifySprites(sample_one=True) rules in the compiler's traversal order
    (= the order of their rule-noise columns)."""
    out = []
    for r in rules:
        if type(r).__name__ == 'ConditionalRule':
            out += _sample_one_rules(r._rules)  # pylint: disable=protected-access
        elif type(r).__name__ == 'ModifySprites' and r._sample_one:  # pylint: disable=protected-access
            out.append(r)
    return out


def _seek_action(env, t):
    """cleanup: every agent heads for the nearest fruit (even steps of 40) or
    fountain, full stick."""
    act = {}
    for k, name in enumerate(('agent_0', 'agent_1', 'agent_2')):
        a = env.state[name][0]
        targets = env.state['fruits'] if ((t // 40) + k) % 2 == 0 else env.state['fountains']
        d = [np.array(s.position) - np.array(a.position) for s in targets]
        d = min(d, key=lambda v: float(np.dot(v, v)))
        n = float(np.linalg.norm(d))
        act[name] = d / n if n > 0 else np.zeros(2)   # float64, like random_action()
    return act


def _chase_action(env, t):
    """chase_avoid_torus: full stick towards the first prey (away from it every
    fourth step, so that the path crosses the arena edge)."""
    a = env.state['agent'][0]
    if not env.state['prey']:
        return np.zeros(2)
    d = np.array(env.state['prey'][0].position) - np.array(a.position)
    n = float(np.linalg.norm(d))
    d = d / n if n > 0 else np.zeros(2)
    return -d if t % 4 == 3 else d


def _booster_action(env, t):
    """functional_maze: full stick towards the nearest booster for 45 steps, then towards the nearest prey."""
    a = env.state['agent'][0]
    targets = env.state['boosters'] if t < 45 else (env.state['prey'] or env.state['boosters'])
    d = [np.array(s.position) - np.array(a.position) for s in targets]
    d = min(d, key=lambda v: float(np.dot(v, v)))
    n = float(np.linalg.norm(d))
    return d / n if n > 0 else np.zeros(2)


def _fixate_action(env, t):
    """multi_tracking_with_feature: SetPosition on the fixation cross while it exists, then on target 0."""
    del t
    target = env.state['fixation'][0] if env.state['fixation'] else env.state['targets'][0]
    return np.array(target.position, dtype=np.float64)


def _match_action(env, t):
    """match_to_sample: full stick towards the cover that hides the prey (the agent is glued until the
    response phase)."""
    del t
    d = np.array(env.state['covers'][0].position) - np.array(env.state['agent'][0].position)
    n = float(np.linalg.norm(d))
    return d / n if n > 0 else np.zeros(2)


def _pacman_action(env, t):
    """pacman: hold a Grid direction for 6 steps, cycling left, up, right, down,
    so that the agent runs into walls, turns at intersections and eats prey."""
    del env
    return [0, 3, 1, 2][(t // 6) % 4] if t % 23 != 22 else 4


def _slot_map(prog, state):
    m = {}
    for l, name in enumerate(prog.layer_names):
        for k, sp in enumerate(state[name]):
            m[id(sp)] = int(prog.layer_off[l]) + k
    return m


def _flat_action(prog, action):
    out = np.zeros(max(prog.action_dim, 1))
    for key, kind, off, width in prog.action_layout:
        a = action if key is None else action[key]
        out[off:off + width] = np.asarray(a, dtype=np.float64).reshape(-1)[:width]
    return out


AA_SCENES = ('pong', 'falling_balls20', 'colliding_predators', 'cleanup')
AA_FACTORS = (2, 3)


def _aa_renderers(renderer):
    """The scene's PILRenderer again, with anti_aliasing 2 and 3
    (pil_renderer.py:37-86: supersampled canvas + Image.resize(LANCZOS))."""
    from moog.observers import pil_renderer
    out = {}
    for aa in AA_FACTORS:
        out[aa] = pil_renderer.PILRenderer(
            image_size=tuple(renderer._image_size), anti_aliasing=aa,  # pylint: disable=protected-access
            bg_color=renderer._canvas_bg.getpixel((0, 0)),  # pylint: disable=protected-access
            color_to_rgb=renderer.color_to_rgb,
            polygon_modifier=renderer._polygon_modifier)  # pylint: disable=protected-access
    return out


BIG_SCENES = ('pong', 'colliding_predators')
BIG_SIZES = ((256, 256), (512, 512), (136, 200), (1024, 1024))      # (width, height); 64 x 64 x ... does not fit one CTA beyond ~128^2


def _big_renderers(renderer):
    """The scene's PILRenderer again at canvas sizes that exceed one CTA's shared memory (the shipped
    pacman draws 256 x 256, tests/runtime_benchmark.py times up to 1024 x 1024)."""
    from moog.observers import pil_renderer
    return {size: pil_renderer.PILRenderer(
        image_size=size, anti_aliasing=1, bg_color=renderer._canvas_bg.getpixel((0, 0)),  # pylint: disable=protected-access
        color_to_rgb=renderer.color_to_rgb, polygon_modifier=renderer._polygon_modifier)  # pylint: disable=protected-access
        for size in BIG_SIZES}


def generate(name, out_dir):
    module, level, seed, T, frame_every = SCENES[name]
    np.random.seed(seed)
    mod = importlib.import_module(module)
    config = mod.get_config(level)
    if name == 'pacman':
        from moog.observers import pil_renderer
        config['observers'] = {'image': pil_renderer.PILRenderer(
            image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')}
    env = environment.Environment(**config)

    # capture the state initializer's output before the reference steps rules
    init_state = env.state_initializer()
    samples = [init_state]
    prog = compiler.compile_config(config, samples)
    init = compiler.pack_states(prog, [init_state])
    env.state_initializer = lambda: init_state
    # Phase durations drawn as np.random.randint(lo, hi) in Phase.reset (task_phases.py:71-77): the outcome d
    # is replayed through the phase's rule-noise column as a uniform u with lo + int(u * (hi - lo)) == d
    reset_draws = []
    real_randint = np.random.randint

    def _reset_randint(low, high=None, size=None, **kwargs):
        out = real_randint(low, high, size, **kwargs)
        if size is None and not kwargs:
            reset_draws.append(int(out))
        return out
    np.random.randint = _reset_randint
    try:
        ts = env.reset()
    finally:
        np.random.randint = real_randint
    reset_rule_noise = np.zeros(max(prog.rule_noise_dim, 1))
    assert len(reset_draws) == len(prog.duration_draws), (reset_draws, prog.duration_draws)
    for d, (col, lo, hi) in zip(reset_draws, prog.duration_draws):
        assert lo <= d < hi
        reset_rule_noise[col] = (d - lo + 0.5) / (hi - lo)
    table = init['shape_table']
    after_reset = compiler.pack_states(prog, [env.state], table)
    renderer = config.get('observers', {}).get('image')

    # uniforms behind RandomForce (random_force.py:22-26)
    draws = []
    orig_uniform = np.random.uniform

    # draws made inside a rule class of the config's own (lambdas.trace_rule: np.random.uniform / randint become
    # rule-noise columns, in call order): the uniform behind each one goes to its column
    plan_now = [None, 0]
    orig_randint = np.random.randint

    def _next_planned(kind):
        plan, k = plan_now
        assert k < len(plan) and plan[k][0] == kind, (kind, plan, k)
        plan_now[1] = k + 1
        return plan[k]

    def _uniform(low=0.0, high=1.0, size=None):
        if size is not None:
            return orig_uniform(low, high, size)
        u = np.random.random_sample()
        if plan_now[0] is not None:
            _, col, _ = _next_planned('uniform')
            rule_draws[col] = u
        else:
            draws.append(u)
        return low + (high - low) * u

    def _randint(low, high=None, size=None, **kwargs):
        out = orig_randint(low, high, size, **kwargs)
        if plan_now[0] is not None and size is None and not kwargs:
            _, col, n = _next_planned('randint')
            lo = 0 if high is None else low
            rule_draws[col] = (int(out) - lo + 0.5) / n
        return out

    for rule_obj, plan in prog.rule_draws:
        def _planned_step(state, meta_state, _orig=rule_obj.step, _plan=plan):
            plan_now[0], plan_now[1] = _plan, 0
            try:
                return _orig(state, meta_state)
            finally:
                plan_now[0] = None
        rule_obj.step = _planned_step

    # the uniform behind ModifySprites(sample_one)'s np.random.choice
    # (modify_sprites.py:48-49): element int(u * len) of the filtered list
    so_rules = _sample_one_rules(config.get('game_rules', ()))
    rule_draws = {}
    current_rule = [None]
    orig_choice = np.random.choice

    def _choice(seq, *args, **kwargs):
        if current_rule[0] is None or args or kwargs:
            return orig_choice(seq, *args, **kwargs)
        u = np.random.random_sample()
        col = so_rules.index(current_rule[0])
        assert col not in rule_draws, 'a sample_one rule fired twice in one step'
        rule_draws[col] = u
        return seq[min(int(u * len(seq)), len(seq) - 1)]

    for r in so_rules:
        def _wrapped(state, meta_state, _r=r, _orig=r.step):
            current_rule[0] = _r
            try:
                return _orig(state, meta_state)
            finally:
                current_rule[0] = None
        r.step = _wrapped

    # RandomMazeWalk (maze_walk.py:185-186): the four np.random.rand(2, 2) uniforms of a
    # sprite that picks a new direction go to the noise columns [col + 4 k, col + 4 k + 4)
    # of sprite k of the walking layer
    walk_draws = {}
    walkers = []
    f_maze_walk = compiler.F_MAZE_WALK
    walk_ops = [o for o in prog.ops if o['kind'] == f_maze_walk]
    walk_forces = [entry[0] for entry in config['physics']._forces  # pylint: disable=protected-access
                   if type(entry[0]).__name__ == 'RandomMazeWalk']
    assert len(walk_ops) == len(walk_forces), 'one layer per RandomMazeWalk entry expected'
    current_walk = [None]
    orig_rand = np.random.rand

    def _rand(*shape):
        if current_walk[0] is None or shape != (2, 2):
            return orig_rand(*shape)
        u = np.random.random_sample(4)
        col, k = current_walk[0]
        walk_draws[col + 4 * k] = u
        return u.reshape(2, 2)

    for wf, wo in zip(walk_forces, walk_ops):
        def _wrapped_walk(sprite_, updates_per_env_step=1, _orig=wf._step_sprite, _op=wo):  # pylint: disable=protected-access
            layer = prog.layer_names[_op['i'][0]]
            k = [id(x) for x in env.state[layer]].index(id(sprite_))
            current_walk[0] = (_op['i'][2], k)
            try:
                return _orig(sprite_, updates_per_env_step=updates_per_env_step)
            finally:
                current_walk[0] = None
        wf._step_sprite = _wrapped_walk  # pylint: disable=protected-access
        walkers.append(wf)

    rec = {k: [] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'reward', 'last',
                           'actions', 'noise', 'rule_noise', 'n_calls', 'n_true', 'true_hash')}
    frames, frame_steps = [], []
    aa_r = _aa_renderers(renderer) if (renderer is not None and name in AA_SCENES) else {}
    aa_frames = {aa: [] for aa in aa_r}
    big_r = _big_renderers(renderer) if (renderer is not None and name in BIG_SCENES and
                                          os.environ.get('MOOG_GOLDEN_BIG_ONLY')) else {}
    big_frames = {size: [] for size in big_r}
    if renderer is not None:
        frames.append(np.asarray(ts.observation['image']))
        frame_steps.append(-1)
        for aa, r in aa_r.items():
            aa_frames[aa].append(np.asarray(r(env.state)))
        for size, r in big_r.items():
            big_frames[size].append(np.asarray(r(env.state)))
    K, nd = prog.K, prog.noise_dim
    for t in range(T):
        if name == 'cleanup':
            action = _seek_action(env, t)
        elif name == 'chase_avoid_torus':
            action = _chase_action(env, t)
        elif name == 'pacman':
            action = _pacman_action(env, t)
        elif name == 'functional_maze':
            action = _booster_action(env, t)
        elif name == 'multi_tracking':
            action = _fixate_action(env, t)
        elif name == 'match_to_sample':
            action = _match_action(env, t)
        elif name == 'bounce_box':
            action = 4 if t < 25 else 0             # wait, then walk into the left response box
        elif name == 'predict_zoo':
            action = 4 if t < 30 else (0 if t < 50 else 1)      # wait, walk into the left box, then back to the right one
        elif name == 'red_green':
            action = (4 if t % 9 else 3 - (t // 9) % 2) if t < 100 else 1      # wait (a little up / down), then go right
        elif name == 'timed_center':
            action = np.array([1.0, 0.6]) if t < 25 else np.array([-1.0, -1.0])   # leaves the centre cell repeatedly
        else:
            action = env.action_space.random_action()
        flat = _flat_action(prog, action)
        slots = _slot_map(prog, env.state)
        del draws[:]
        rule_draws.clear()
        walk_draws.clear()
        np.random.uniform = _uniform
        np.random.choice = _choice
        np.random.rand = _rand
        np.random.randint = _randint
        try:
            with refenv.OverlapLog() as log:
                ts = env.step(action)
        finally:
            np.random.uniform = orig_uniform
            np.random.choice = orig_choice
            np.random.rand = orig_rand
            np.random.randint = orig_randint
        h, n_true = 0, 0
        for a, b, r in log.calls:
            if r:
                n_true += 1
                h = true_event_hash(h, slots[id(a)], slots[id(b)])
        noise = np.zeros((K, max(nd, 1)))
        if walk_ops:
            assert K == 1 and not draws
            for col, u in walk_draws.items():
                noise[0, col:col + 4] = u
        elif nd:
            # draws arrive substep by substep, in force order, 2 per sprite
            per = len(draws) // K
            assert per * K == len(draws) and per <= nd, (per, nd, len(draws))
            noise[:, :per] = np.array(draws).reshape(K, per)
        st = compiler.pack_states(prog, [env.state], table)
        for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt'):
            rec[k].append(st[k][0])
        rec['reward'].append(0.0 if ts.reward is None else float(ts.reward))
        rec['last'].append(bool(ts.last()))
        rec['actions'].append(flat)
        rec['noise'].append(noise)
        rn = np.zeros(max(prog.rule_noise_dim, 1))
        for col, u in rule_draws.items():
            rn[col] = u
        rec['rule_noise'].append(rn)
        rec['n_calls'].append(len(log.calls))
        rec['n_true'].append(n_true)
        rec['true_hash'].append(np.uint64(h))
        if renderer is not None and (t % frame_every == 0 or ts.last()):
            frames.append(np.asarray(ts.observation['image']))
            frame_steps.append(t)
            for aa, r in aa_r.items():
                aa_frames[aa].append(np.asarray(r(env.state)))
            if t % (3 * frame_every) == 0:
                for size, r in big_r.items():
                    big_frames[size].append(np.asarray(r(env.state)))
        if ts.last():
            break

    shape_verts, shape_nv = table.arrays()
    out = dict(
        blob=np.frombuffer(prog.blob, dtype=np.uint8),
        layer_names=np.array(prog.layer_names),
        shape_verts=shape_verts, shape_nv=shape_nv,
        frames=np.array(frames, dtype=np.uint8).reshape(
            (len(frames),) + ((renderer._image_size[1], renderer._image_size[0], 3)  # pylint: disable=protected-access
                              if renderer is not None else (0, 0, 3))),
        frame_steps=np.array(frame_steps, dtype=np.int32),
    )
    for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'envi', 'envf'):
        out['init_' + k] = init[k][0]
        out['reset_' + k] = after_reset[k][0]
    for k, v in rec.items():
        out[k] = np.array(v)
    if prog.duration_draws:
        out['reset_rule_noise'] = reset_rule_noise
    path = os.path.join(out_dir, name + '.npz')
    if os.environ.get('MOOG_GOLDEN_BIG_ONLY'):
        # big-canvas frames only, for states the main fixture already holds
        assert big_r, name
        prev = np.load(path)
        assert np.array_equal(prev['frames'], out['frames']) and np.array_equal(prev['dyn'], out['dyn']), \
            'the trajectory changed; regenerate the main fixture first'
        steps = [-1] + [t for t in out['frame_steps'][1:] if t % (3 * frame_every) == 0]
        path = os.path.join(out_dir, name + '_big.npz')
        np.savez_compressed(path, frame_steps=np.array(steps, dtype=np.int32),
                            **{'frames_%dx%d' % size: np.array(v, dtype=np.uint8) for size, v in big_frames.items()})
        print('{:22s} {} big-canvas frames x {} -> {} ({} KB)'.format(name, len(steps), list(big_frames), path,
                                                                    os.path.getsize(path) // 1024))
        return
    if os.environ.get('MOOG_GOLDEN_AA_ONLY'):
        # anti-aliased frames only, for the states the main fixture already holds
        assert aa_r, name
        prev = np.load(path)
        assert np.array_equal(prev['frames'], out['frames']) and np.array_equal(prev['dyn'], out['dyn']), \
            'the trajectory changed; regenerate the main fixture first'
        path = os.path.join(out_dir, name + '_aa.npz')
        np.savez_compressed(path, frame_steps=out['frame_steps'],
                            **{'frames_aa%d' % aa: np.array(v, dtype=np.uint8) for aa, v in aa_frames.items()})
        print('{:22s} {} anti-aliased frames x {} -> {}'.format(name, len(frame_steps), list(aa_frames), path))
        return
    np.savez_compressed(path, **out)
    print('{:22s} T={:3d} slots={:3d} vtx={:4d} true/step={:.1f} -> {} ({} KB)'.format(
        name, len(rec['reward']), prog.n_slots, prog.n_vtx,
        float(np.mean(rec['n_true'])), path, os.path.getsize(path) // 1024))


def main():
    out_dir = os.path.join(_ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    names = sys.argv[1:] or list(SCENES)
    for n in names:
        generate(n, out_dir)


if __name__ == '__main__':
    main()
