"""CPU check of the DEVICE rule source: marl-snake_b200/csrc/snk_core.cuh compiled for the host
(tests/hostsim) must reproduce the reference's golden trajectories bit for bit, and its Philox stream
must match the oracle's Python restatement of the same spec."""
import numpy as np
import pytest

from golden_util import ROLLOUTS, Rollout, load_scenarios
from hostsim_util import HostSim
from parity_util import check_against_oracle_philox, check_rollout_replay, check_scenario


@pytest.mark.parametrize('name', ROLLOUTS)
def test_hostsim_rollout_replay(name):
    g = Rollout(name)
    hs = HostSim(g.num_envs, g.kwargs, rng_mode=1, auto_reset=1, done_mode=g.done_mode)
    check_rollout_replay(hs, g)
    assert hs.errors() == 0
    hs.close()


@pytest.mark.parametrize('sc', load_scenarios(), ids=lambda s: s.name)
def test_hostsim_scenarios(sc):
    kw = dict(height=sc.H, width=sc.W, num_snakes=sc.num_snakes, snake_length=2, **sc.kwargs)
    hs = HostSim(1, kw, rng_mode=1, auto_reset=0)
    check_scenario(hs, sc)
    assert hs.errors() == 0
    hs.close()


@pytest.mark.parametrize('kw', [
    dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5),
    dict(height=9, width=12, num_snakes=3, snake_length=4, vision_range=2, frame_stack=3, max_episode_steps=30),
])
def test_hostsim_philox_matches_oracle(kw):
    hs = check_against_oracle_philox(HostSim, kw, num_envs=6, steps=120, seed=0xC0FFEE1234, env_id_offset=5)
    assert hs.errors() == 0
    hs.close()


@pytest.mark.parametrize('case', range(24))
def test_hostsim_fuzz_matches_oracle(case):
    """Seeded random small shapes, Philox mode: the device rule source (host build) against the oracle.  The
    GPU suite fuzzes libsnk.so against the same host build on larger shapes, closing the chain
    reference -> oracle -> rule source -> kernels."""
    from test_oracle_vs_golden import small_fuzz_configs
    kw, _ = small_fuzz_configs(24, 4242)[case]
    hs = check_against_oracle_philox(HostSim, kw, num_envs=5, steps=60, seed=31337 + case, env_id_offset=case,
                                     action_seed=case)
    assert hs.errors() == 0
    hs.close()


def test_bad_action_flag():
    hs = HostSim(1, dict(num_snakes=2), rng_mode=0, auto_reset=1)
    hs.reset()
    hs.step(np.array([[0, 7]], dtype=np.uint8))
    assert hs.errors() & 1
    hs.close()


@pytest.mark.parametrize('kw,N,T', [
    (dict(height=20, width=20, num_snakes=4, snake_length=3, vision_range=5), 96, 160),
    (dict(height=12, width=14, num_snakes=3, snake_length=3, vision_range=2, frame_stack=3, max_episode_steps=40), 40, 100),
])
def test_hostsim_replay_recorded_at_test_time(kw, N, T):
    """The replay-at-scale driver of the GPU suite (tests/replay_scale_util.py: oracle trajectories recorded
    under np.random.seed(base + env), draws fed to the replay mode) exercised on the host build."""
    from golden_util import unpack_obs
    from replay_scale_util import check_replay_at_scale, oracle_trajectories
    envs = oracle_trajectories(kw, N, T, 1234, procs=4)
    hs = HostSim(N, kw, rng_mode=1, auto_reset=1)
    episodes = check_replay_at_scale(hs, envs, T, unpack_obs)
    assert episodes > N and hs.errors() == 0
    hs.close()
