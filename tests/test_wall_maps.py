"""custom wall layouts (the reference's assets/*.txt maps, core/grid_util.py:23-33).  CPU part: the C ABI's
spawn enumeration on a map against the reference's dfs_sweep_empty (golden), the text loader.  GPU part: the CUDA path
on a map against the oracle (Philox mode) and the host build of the rule source in both tile modes; the golden replay
of the map rollouts runs with the other fixtures (test_gpu_rollout_replay / test_hostsim_rollout_replay)."""
import ctypes as C
import os

import numpy as np
import pytest

from golden_util import load_map_spawn_tables

CROSS = """\
####################
#..................#
#..................#
#........##........#
#........##........#
#..................#
#..#####....#####..#
#..................#
#........##........#
#........##........#
#..................#
####################
"""


def test_c_abi_spawn_table_on_maps_matches_reference_enumeration():
    import marl_snake_b200 as m
    for name, (walls, k, ref) in load_map_spawn_tables().items():
        H, W = walls.shape
        w = np.ascontiguousarray(walls, dtype=np.uint8)
        n = m.lib.snk_spawn_count_map(H, W, k, w.ctypes.data_as(C.c_void_p))
        assert n == len(ref), name
        out = np.zeros((n, k), dtype=np.int32)
        assert m.lib.snk_spawn_cells_map(H, W, k, w.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n) == 0
        assert np.array_equal(out, ref[..., 0].astype(np.int32) * W + ref[..., 1]), name
    # the walled box given as a map is the default table
    box = np.ones((9, 11), dtype=np.uint8)
    box[1:-1, 1:-1] = 0
    assert m.lib.snk_spawn_count_map(9, 11, 3, box.ctypes.data_as(C.c_void_p)) == m.lib.snk_spawn_count(9, 11, 3)


def test_text_map_loader(tmp_path):
    from marl_snake_b200.grid_util import as_wall_plane, make_grid_from_txt
    p = tmp_path / 'cross.txt'
    p.write_text(CROSS)                                   # with a final newline, as most of the reference's assets
    g = make_grid_from_txt(str(p), {'#': 1, '.': 0})
    assert g.shape == (12, 20) and g[0].all() and g[3, 9] == 1 and g[1, 1] == 0
    assert np.array_equal(as_wall_plane(str(p)), (g != 0).astype(np.uint8))
    (tmp_path / 'ragged.txt').write_text('####\n#..#\n###\n')
    with pytest.raises(ValueError):
        make_grid_from_txt(str(tmp_path / 'ragged.txt'))
    with pytest.raises(ValueError):
        as_wall_plane(g, height=20, width=20)


def cross_map():
    return np.array([[1 if ch == '#' else 0 for ch in line] for line in CROSS.strip().split('\n')])


@pytest.mark.gpu
@pytest.mark.parametrize('compact', ['0', '1'])
@pytest.mark.parametrize('coop', ['0', '1'])
@pytest.mark.parametrize('kw,N,steps', [
    (dict(num_snakes=4, snake_length=3, vision_range=4), 40, 120),
    (dict(num_snakes=3, snake_length=4, frame_stack=2, num_fruits=6), 17, 90),
    (dict(num_snakes=8, snake_length=2, vision_range=3, max_episode_steps=40), 9, 90),     # direction-plane record
])
def test_gpu_wall_map_philox_matches_oracle(monkeypatch, coop, kw, N, steps, compact):
    from gpu_backend import GpuBackend
    from parity_util import check_against_oracle_philox
    monkeypatch.setenv('SNK_COOP', coop)
    monkeypatch.setenv('SNK_COMPACT', compact)
    walls = cross_map()
    kw = dict(height=walls.shape[0], width=walls.shape[1], wall_map=walls, **kw)
    be = check_against_oracle_philox(GpuBackend, kw, num_envs=N, steps=steps, seed=77, env_id_offset=11)
    assert be.errors() == 0
    be.close()


@pytest.mark.gpu
def test_gpu_wall_map_large_batch_against_rule_source():
    """The reference's 40x40 logo map on 3000 envs: walls survive every step and reset, everything else equals the
    host build of the rule source."""
    from gpu_backend import GpuBackend
    from hostsim_util import HostSim
    walls = load_map_spawn_tables()['40x40_ml2'][0]
    kw = dict(height=40, width=40, num_snakes=6, snake_length=3, vision_range=5, max_episode_steps=30, wall_map=walls)
    N, ns = 3000, 6
    hs = HostSim(N, kw, rng_mode=0, auto_reset=1, seed=5)
    be = GpuBackend(N, kw, rng_mode=0, auto_reset=1, seed=5)
    assert np.array_equal(hs.reset(), be.reset())
    rng = np.random.RandomState(2)
    for t in range(70):
        a = rng.randint(0, 3, size=(N, ns)).astype(np.uint8)
        o1, r1, d1, i1 = hs.step(a)
        o2, r2, d2, i2 = be.step(a)
        assert np.array_equal(r1, r2) and np.array_equal(d1, d2.astype(np.uint8)) and np.array_equal(o1, o2), t
    g = be.grid()[0].reshape(N, 40, 40)
    assert np.array_equal(hs.grid()[0].reshape(N, 40, 40), g)
    assert ((g % 10 == 1) == (walls != 0)[None]).all()
    assert be.errors() == 0
    be.close()


@pytest.mark.gpu
def test_make_snake_with_a_map_file(tmp_path):
    """Drop-in spelling: make_snake(..., wall_map=<path or array>) for one env and for a vector of envs."""
    from marl_snake_b200 import SnkError, SnakeBatch, make_snake
    p = tmp_path / 'cross.txt'
    p.write_text(CROSS)
    walls = cross_map()
    env, _, _, props = make_snake(num_envs=1, num_snakes=3, snake_length=3, vision_range=3, wall_map=str(p))
    obs = env.reset()
    assert obs.shape == (3, 7, 7, 8) and env.grid_shape == (12, 20)
    assert np.array_equal(env.grid % 10 == 1, walls != 0)
    for _ in range(30):
        obs, rew, done, info = env.step([env.action_space.sample() for _ in range(3)])
        if all(done):
            env.reset()
    assert np.array_equal(env.grid % 10 == 1, walls != 0)
    env.close()
    venv, _, _, _ = make_snake(num_envs=5, num_snakes=2, snake_length=3, wall_map=walls)
    assert venv.reset().shape == (5, 2, 12, 20, 8)
    venv.close()
    bad = walls.copy()
    bad[0, 4] = 0                                           # a hole in the outer ring: refused
    with pytest.raises(SnkError):
        SnakeBatch(2, wall_map=bad)
