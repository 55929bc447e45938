"""GPU parity at BASELINE.json's real sizes ("What's weak" 1).

(a) cfg2 at 4 096, cfg3 at 65 536 and cfg4 at 16 384 environments -- the batch sizes whose tile mode,
    tile size and grid shape the create-time heuristics pick for the real run -- bit-exact against the host
    build of the rule source (Philox streams are keyed by global env id, so slices of the batch can be
    stepped alone), plus size-independent invariants over the whole batch.
(b) replay mode -- the only mode that is bit-comparable with the reference -- on 1 024 environments x 512
    steps, draws recorded at test time from the oracle under np.random.seed(base + env), once
    with warp-private tiles (SNK_COOP=0, the bench's own mode at cfg5) and once with cooperative tiles
    (SNK_COOP=1), comparing grid, rewards, dones, observations, info and draw cursors at every step.
"""
import numpy as np
import pytest
import torch

from golden_util import unpack_obs
from gpu_backend import GpuBackend

pytestmark = pytest.mark.gpu

CFG4_REW = {'fruit': 10.0, 'kill': 1.0, 'lose': -1.0, 'win': 0.1, 'time': -0.001}
CFG = {
    'cfg2': (dict(height=20, width=20, num_snakes=4, snake_length=3), 4096),
    'cfg3': (dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5, frame_stack=4), 65536),
    'cfg4': (dict(height=64, width=64, num_snakes=16, snake_length=5, vision_range=7, reward_dict=CFG4_REW), 16384),
}
# (first env, count) slices stepped by the host build; None = the whole batch
SLICES = {'cfg2': None, 'cfg3': [(0, 1500), (32768 - 250, 500), (65536 - 1000, 1000)],
          'cfg4': [(0, 100), (8192 - 20, 40), (16384 - 60, 60)]}
STEPS = {'cfg2': 150, 'cfg3': 60, 'cfg4': 50}


def _invariants(b, obs, done, info, kw):
    """Size-independent properties of the whole batch (one HEAD and TAIL per live snake, body cell count ==
    length, walls intact, own head at the window centre / on channel 5, 0/1 observation)."""
    N, ns = b.num_envs, b.num_snakes
    H, W = b.grid_shape
    st = b.get_state()
    grid = st['grid'].reshape(N, -1).long()
    kind, owner = grid % 10, grid // 10
    alive = st['alive'].bool()
    for i in range(ns):
        own = owner == i
        assert torch.equal(((kind == 3) & own).sum(1), alive[:, i].long())
        assert torch.equal(((kind == 5) & own).sum(1), alive[:, i].long())
        assert torch.equal(((kind >= 3) & own).sum(1), st['length'][:, i].long())
    border = torch.zeros(H, W, dtype=torch.bool, device=grid.device)
    border[0] = border[-1] = True
    border[:, 0] = border[:, -1] = True
    assert bool((grid.view(N, H, W)[:, border] == 1).all())
    assert int(obs.max()) == 1
    V = kw.get('vision_range')
    cur = obs[..., -8:]                                    # newest frame
    if V:
        assert torch.equal(cur[:, :, V, V, 5].bool(), alive)
    else:
        assert torch.equal(cur[..., 5].sum((2, 3)).bool(), alive)
    live_env = ~info['finished']
    assert torch.equal(done[live_env], ~alive[live_env])


@pytest.mark.parametrize('name', ['cfg2', 'cfg3', 'cfg4'])
def test_gpu_baseline_config_full_size(name):
    from hostsim_util import HostSim
    from marl_snake_b200 import SnakeBatch
    kw, N = CFG[name]
    ns = kw['num_snakes']
    b = SnakeBatch(N, seed=23, **kw)
    slices = SLICES[name] or [(0, N)]
    sims = [HostSim(n, kw, rng_mode=0, auto_reset=1, seed=23, env_id_offset=lo) for lo, n in slices]
    obs = b.reset()
    for sim, (lo, n) in zip(sims, slices):
        assert np.array_equal(obs[lo:lo + n].cpu().numpy(), sim.reset()), (name, 'reset', lo)
    g = torch.Generator(device='cuda').manual_seed(5)
    for t in range(STEPS[name]):
        a = torch.randint(0, 3, (N, ns), dtype=torch.uint8, device='cuda', generator=g)
        obs, rew, done, info = b.step(a)
        for sim, (lo, n) in zip(sims, slices):
            sl = slice(lo, lo + n)
            o, r, d, i = sim.step(a[sl].cpu().numpy())
            assert np.array_equal(rew[sl].cpu().numpy(), r) and np.array_equal(done[sl].cpu().numpy(), d.astype(bool)), (name, t, lo)
            assert np.array_equal(obs[sl].cpu().numpy(), o), (name, t, lo)
            fin = i['finished'].astype(bool)
            assert np.array_equal(info['finished'][sl].cpu().numpy(), fin), (name, t, lo)
            for k in ('rank', 'episode_scores', 'episode_steps', 'episode_fruits', 'episode_kills'):
                assert np.array_equal(info[k][sl].cpu().numpy()[fin], i[k][fin]), (name, k, t, lo)
    grid = b.get_state()['grid'].reshape(N, -1)
    for sim, (lo, n) in zip(sims, slices):
        assert np.array_equal(grid[lo:lo + n].cpu().numpy(), sim.grid()[0]), (name, 'grid', lo)
    _invariants(b, obs, done, info, kw)
    assert b.device_errors() == 0
    assert b.stats()['episodes'] > 0


REPLAY_KW = dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5)
REPLAY_N, REPLAY_T, REPLAY_BASE = 1024, 512, 4000


@pytest.mark.parametrize('compact', ['0', '1'])
@pytest.mark.parametrize('coop', ['0', '1'])
def test_gpu_replay_at_scale(monkeypatch, coop, compact):
    """1 024 envs x 512 steps of the cfg1/cfg5 shape in replay mode against the oracle's recorded
    trajectories, in both tile modes."""
    from replay_scale_util import check_replay_at_scale, oracle_trajectories
    envs = oracle_trajectories(REPLAY_KW, REPLAY_N, REPLAY_T, REPLAY_BASE)
    monkeypatch.setenv('SNK_COOP', coop)
    monkeypatch.setenv('SNK_COMPACT', compact)             # both record layouts (HBM record with / without the grid)
    be = GpuBackend(REPLAY_N, REPLAY_KW, rng_mode=1, auto_reset=1)
    episodes = check_replay_at_scale(be, envs, REPLAY_T, unpack_obs)
    assert episodes > 4 * REPLAY_N                       # mean episode ~70 steps: every env restarts several times
    assert be.errors() == 0
    be.close()


@pytest.mark.parametrize('kw,N,T', [
    (dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5, frame_stack=4), 256, 200),   # cfg3 shape
    (dict(height=20, width=20, num_snakes=4, snake_length=3), 256, 200),                                  # cfg2 shape
    (dict(height=64, width=64, num_snakes=16, snake_length=5, vision_range=7, reward_dict=CFG4_REW), 24, 120),   # cfg4
])
def test_gpu_replay_other_shapes(kw, N, T):
    """The other BASELINE shapes in replay mode against freshly recorded oracle trajectories (more
    environments and steps than the committed golden rollouts hold)."""
    from replay_scale_util import check_replay_at_scale, oracle_trajectories
    envs = oracle_trajectories(kw, N, T, 9000)
    be = GpuBackend(N, kw, rng_mode=1, auto_reset=1)
    check_replay_at_scale(be, envs, T, unpack_obs)
    assert be.errors() == 0
    be.close()
