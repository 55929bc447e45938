"""Loaders for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import ast
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ROLLOUTS = ['roll_cfg1', 'roll_full', 'roll_fs4', 'roll_big', 'roll_single', 'roll_cap',
            'roll_rect', 'roll_crowd', 'roll_coop', 'roll_human', 'roll_cfg4', 'roll_map_cross', 'roll_map_12', 'roll_map_ml2']


def unpack_obs(packed):
    """uint8 [..., fs] -> uint8 0/1 [..., 8*fs] (inverse of make_golden.pack_obs)."""
    p = np.asarray(packed, dtype=np.uint8)
    bits = np.unpackbits(p[..., None], axis=-1, bitorder='little')
    return bits.reshape(*p.shape[:-1], p.shape[-1] * 8)


class Rollout:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.name = name
        self.kwargs = ast.literal_eval(str(z['meta_kwargs']))
        self.num_envs = int(z['meta_num_envs'])
        self.steps = int(z['meta_steps'])
        self.seed = int(z['meta_seed'])
        self.env_id = str(z['meta_env_id']) if 'meta_env_id' in z.files else 'Snake-v1'
        self.done_mode = 1 if self.env_id == 'SnakeCoop-v1' else 0
        self.env = [{k[len(f'e{e}_'):]: z[k] for k in z.files if k.startswith(f'e{e}_')}
                    for e in range(self.num_envs)]

    def stacked(self, key):
        return np.stack([e[key] for e in self.env])


class Scenario:
    def __init__(self, z, name):
        self.name = name
        g = lambda k: z[f'{name}__{k}']  # noqa: E731
        self.H, self.W = int(g('H')), int(g('W'))
        self.kwargs = ast.literal_eval(str(g('kw')))
        for k in ('grid0', 'alive0', 'dir0', 'len0', 'cells0', 'actions', 'grid', 'rewards', 'dones',
                  'obs', 'counter', 'draws', 'draws_end'):
            setattr(self, k, g(k))
        self.counter0 = int(g('counter0'))
        self.num_snakes = len(self.alive0)


def load_scenarios():
    z = np.load(os.path.join(GOLDEN_DIR, 'scenarios.npz'))
    return [Scenario(z, str(n)) for n in z['names']]


def load_spawn_tables():
    z = np.load(os.path.join(GOLDEN_DIR, 'spawn_tables.npz'))
    tables = {tuple(int(x) for x in k.split('_')[1:]): z[k] for k in z.files if k.startswith('c_')}
    return tables, z['turn']


def load_map_spawn_tables():
    """Spawn tables the reference enumerates on its asset maps: name -> (walls uint8 [H, W], snake_length, cands)."""
    z = np.load(os.path.join(GOLDEN_DIR, 'spawn_tables.npz'))
    names = sorted({k[len('map_'):].split('__')[0] for k in z.files if k.startswith('map_')})
    return {n: (z[f'map_{n}__walls'], int(z[f'map_{n}__k']), z[f'map_{n}__cands']) for n in names}
