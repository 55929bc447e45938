"""Shared comparison drivers: run any backend that looks like (reset, step, grid, set_state) against the
golden fixtures or the oracle.  Used with the host simulation on CPU and with libsnk.so on the GPU."""
import numpy as np

from golden_util import unpack_obs


def check_rollout_replay(backend, g):
    """backend: object with set_replay/reset/step/grid over g.num_envs envs in replay mode with auto-reset."""
    backend.set_replay([e['draws'] for e in g.env])
    obs = backend.reset()
    assert np.array_equal(obs, np.stack([unpack_obs(e['obs0']) for e in g.env])), 'reset obs'
    grid, counter, cursor = backend.grid()
    H, W = g.env[0]['grid0'].shape
    assert np.array_equal(grid.reshape(-1, H, W), g.stacked('grid0')), 'reset grid'
    assert np.array_equal(cursor, g.stacked('draws_end')[:, 0]), 'reset draw count'
    nterm = np.zeros(g.num_envs, dtype=int)
    for t in range(g.steps):
        acts = np.stack([e['actions'][t] for e in g.env])
        obs, rew, done, info = backend.step(acts)
        tag = (g.name, t)
        assert np.array_equal(rew, np.stack([e['rewards'][t] for e in g.env])), ('rewards', tag)
        assert np.array_equal(done.astype(np.uint8), np.stack([e['dones'][t] for e in g.env])), ('dones', tag)
        grid, counter, cursor = backend.grid()
        assert np.array_equal(grid.reshape(-1, H, W), np.stack([e['grid_live'][t] for e in g.env])), ('grid', tag)
        assert np.array_equal(obs, np.stack([unpack_obs(e['obs'][t]) for e in g.env])), ('obs', tag)
        assert np.array_equal(cursor, g.stacked('draws_end')[:, t + 1]), ('draw count', tag)
        for e, ge in enumerate(g.env):
            ended = bool(ge['dones'][t].all())
            assert bool(info['finished'][e]) == ended, ('finished', tag, e)
            if ended:
                k = nterm[e]
                assert ge['info_step'][k] == t
                assert np.array_equal(info['rank'][e], ge['info_rank'][k]), ('rank', tag, e)
                assert np.array_equal(info['episode_scores'][e], ge['info_scores'][k]), ('scores', tag, e)
                assert np.array_equal(info['episode_steps'][e], ge['info_steps'][k]), ('steps', tag, e)
                assert np.array_equal(info['episode_fruits'][e], ge['info_fruits'][k]), ('fruits', tag, e)
                assert np.array_equal(info['episode_kills'][e], ge['info_kills'][k]), ('kills', tag, e)
                nterm[e] += 1
            else:
                assert counter[e] == ge['counter'][t], ('alive counter', tag, e)
    assert all(nterm[e] == len(ge['info_step']) for e, ge in enumerate(g.env))


def check_scenario(backend, sc):
    """backend: 1 env, replay mode, no auto-reset."""
    backend.set_replay([sc.draws])
    L = sc.cells0.shape[-1]
    backend.set_state(sc.grid0[None], sc.alive0[None], sc.dir0[None], sc.len0[None],
                      np.maximum(sc.cells0, 0).reshape(1, sc.num_snakes, L), np.array([sc.counter0]))
    for t, a in enumerate(sc.actions):
        obs, rew, done, info = backend.step(a[None])
        grid, counter, cursor = backend.grid()
        tag = (sc.name, t)
        assert np.array_equal(grid.reshape(sc.H, sc.W), sc.grid[t]), ('grid', tag)
        assert np.array_equal(rew[0], sc.rewards[t]), ('rewards', tag, rew[0], sc.rewards[t])
        assert np.array_equal(done[0].astype(np.uint8), sc.dones[t]), ('dones', tag)
        assert counter[0] == sc.counter[t], ('counter', tag)
        assert np.array_equal(obs[0], unpack_obs(sc.obs[t])), ('obs', tag)
        assert cursor[0] == sc.draws_end[t], ('draw count', tag)


def check_against_oracle_philox(backend_factory, kw, num_envs, steps, seed, env_id_offset=0, action_seed=0):
    """Philox mode: the backend and the Python oracle (PhiloxDraws) must agree step for step."""
    from oracle.snake_oracle import OracleSnakeEnv, PhiloxDraws
    ns = kw.get('num_snakes', 4)
    backend = backend_factory(num_envs, kw, rng_mode=0, seed=seed, env_id_offset=env_id_offset)
    draws = [PhiloxDraws(seed, env_id_offset + e) for e in range(num_envs)]
    envs = [OracleSnakeEnv(draws=draws[e], **kw) for e in range(num_envs)]
    for dr in draws:
        dr.tick()
    ref_obs = np.stack([e.reset() for e in envs])
    obs = backend.reset()
    assert np.array_equal(obs, ref_obs), 'reset obs'
    rng = np.random.RandomState(action_seed)
    n_actions = 5 if kw.get('observer') == 'human' else 3
    for t in range(steps):
        acts = rng.randint(0, n_actions, size=(num_envs, ns)).astype(np.uint8)
        obs, rew, done, info = backend.step(acts)
        for e, env in enumerate(envs):
            draws[e].tick()
            o, r, d, inf = env.step([int(a) for a in acts[e]])
            if all(d):
                o = env.reset()
            assert np.array_equal(rew[e], np.asarray(r)), ('rew', t, e)
            assert np.array_equal(done[e].astype(bool), np.asarray(d)), ('done', t, e)
            assert np.array_equal(obs[e], o), ('obs', t, e)
            assert bool(info['finished'][e]) == bool(inf), ('finished', t, e)
            if inf:
                assert np.array_equal(info['rank'][e], inf['rank'])
                assert np.array_equal(info['episode_scores'][e], inf['episode_scores'])
        grid, counter, _ = backend.grid()
        H, W = envs[0].grid_shape
        assert np.array_equal(grid.reshape(-1, H, W), np.stack([e.grid for e in envs])), ('grid', t)
    return backend
