"""predict_zoo: a state initializer that rolls the physics forward ON THE HOST before it accepts an
initial state, and a reward that reads what it found from `sprite.metadata` -- the two things
red_green.py:87-101,144-176 and bounce_box_contact_prediction.py:26-36,94-110 need.

A puck bounces between four walls and two goals.  The initializer draws the puck, then calls
`physics.step(state)` until the puck touches a goal; trials that end too early or too late are drawn
again; the agent learns nothing of this except through `agent.metadata = {'goal': 0 | 1, 'when': steps}`.
The agent answers by walking into the left (goal 0) or right (goal 1) answer box: the reward function
branches on the metadata.  A rule stops the puck on the goal it reaches.

`Physics.step` on the host runs the CUDA physics kernel on a batch of one env
(moog_b200/host_physics.py): the initializer needs the GPU.
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks

STEP_RANGE = (8, 45)      # a trial is kept when the puck reaches a goal within this many steps


def get_config(level=None):
    del level
    physics = physics_lib.Physics(
        (physics_lib.Collision(elasticity=1., symmetric=False, update_angle_vel=False), 'puck', 'walls'),
        updates_per_env_step=5)

    def _wall(x0, y0, x1, y1):
        return sprite.Sprite(x=0., y=0., shape=np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]]),
                             c0=0.6, c1=0.2, c2=0.5)

    def _first_goal(state):
        """(goal index, steps until the puck touches it) or None."""
        puck = state['puck'][0]
        for step in range(STEP_RANGE[1]):
            touching = [puck.overlaps_sprite(g) for g in state['goals']]
            if any(touching):
                return (touching.index(True), step) if step >= STEP_RANGE[0] else None
            physics.step(state)
        return None

    def state_initializer():
        while True:
            walls = [_wall(-1., -1., 2., 0.15), _wall(-1., 0.95, 2., 2.), _wall(-1., -1., 0.05, 2.), _wall(0.95, -1., 2., 2.)]
            goals = [sprite.Sprite(x=0.2, y=0.8, shape='square', scale=0.14, c0=0.0, c1=1., c2=1.),
                     sprite.Sprite(x=0.8, y=0.8, shape='square', scale=0.14, c0=0.33, c1=1., c2=1.)]
            theta = np.random.uniform(0., 2 * np.pi)
            puck = sprite.Sprite(x=np.random.uniform(0.3, 0.7), y=np.random.uniform(0.3, 0.6), shape='circle', scale=0.06,
                                 x_vel=0.04 * np.cos(theta), y_vel=0.04 * np.sin(theta), c0=0.6, c1=1., c2=1.)
            boxes = [sprite.Sprite(x=0.35, y=0.07, shape='square', scale=0.06, c0=0.0, c1=1., c2=0.7),
                     sprite.Sprite(x=0.65, y=0.07, shape='square', scale=0.06, c0=0.33, c1=1., c2=0.7)]
            agent = sprite.Sprite(x=0.5, y=0.07, shape='triangle', scale=0.04, c0=0.15, c1=0.3, c2=1.)
            state = collections.OrderedDict([('walls', walls), ('goals', goals), ('puck', [puck]), ('boxes', boxes),
                                             ('agent', [agent])])
            start = (np.copy(puck.position), np.copy(puck.velocity))
            found = _first_goal(state)
            if found is None:
                continue
            puck.position, puck.velocity = start
            agent.metadata = {'goal': found[0], 'when': found[1]}
            return state

    def _answer(agent, box):
        said_right = box.x > 0.5
        if agent.metadata['goal'] == said_right:
            return 1.
        elif agent.metadata['when'] > 30:
            return -0.5           # a late trial costs less
        else:
            return -1.

    def _stop(s):
        s.velocity = np.zeros(2)

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': tasks.CompositeTask(
            tasks.ContactReward(reward_fn=_answer, layers_0='agent', layers_1='boxes', reset_steps_after_contact=3),
            timeout_steps=70),
        'action_space': action_spaces.Grid(scaling_factor=0.02, action_layers='agent', control_velocity=True),
        'observers': {'image': observers.PILRenderer(image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': (game_rules.ModifyOnContact(layers_0='puck', layers_1='goals', modifier_0=_stop),),
    }
