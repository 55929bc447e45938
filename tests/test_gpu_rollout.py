"""the device rollout (`rollout.collect`) against a host loop written the way the reference trainer's inner
loop is (train_dqn.py:268-308): per-snake forward + argmax, action 0 for dead snakes, a transition only for snakes that
were alive before the step, the early-death penalty, the MAX_STEPS_PER_EPISODE cut, reset when all snakes are done."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class ExactNet(torch.nn.Module):
    """Small integer-weight conv net: every intermediate is an exactly representable float, so a batched forward
    and a per-snake forward agree bit for bit and argmax ties break the same way."""

    def __init__(self, channels, h, w, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.conv = torch.nn.Conv2d(channels, 4, 3, padding=1, bias=False)
        self.head = torch.nn.Linear(4 * h * w, 3, bias=False)
        with torch.no_grad():
            self.conv.weight.copy_(torch.randint(-2, 3, self.conv.weight.shape, generator=g).float())
            self.head.weight.copy_(torch.randint(-2, 3, self.head.weight.shape, generator=g).float())

    def forward(self, x):
        return self.head(torch.relu(self.conv(x)).flatten(1))


def reference_style_loop(batch, net, steps, threshold, penalty, max_steps):
    """train_dqn.py:268-308 per environment, on a SnakeBatch stepped from the host."""
    N, ns = batch.num_envs, batch.num_snakes
    obs = batch.reset().cpu().numpy()
    dones = np.zeros((N, ns), dtype=bool)
    age = np.zeros(N, dtype=np.int64)
    out = []
    for _ in range(steps):
        actions = np.zeros((N, ns), dtype=np.uint8)
        for e in range(N):
            for i in range(ns):
                if dones[e, i]:
                    actions[e, i] = 0
                else:                                                     # SharedAgent.select_action, epsilon 0
                    x = torch.tensor(obs[e, i], dtype=torch.float32).cuda().unsqueeze(0).permute(0, 3, 1, 2)
                    with torch.no_grad():
                        actions[e, i] = net(x).argmax().item()
        next_obs, rew, next_dones, info = batch.step(torch.as_tensor(actions).cuda())
        next_obs, rew, next_dones = next_obs.cpu().numpy(), rew.cpu().numpy(), next_dones.cpu().numpy()
        fin = info['finished'].cpu().numpy().astype(bool)
        for e in range(N):
            for i in range(ns):
                if not dones[e, i]:
                    r = rew[e, i]
                    if next_dones[e, i] and age[e] < threshold:
                        r += penalty
                    out.append((obs[e, i], int(actions[e, i]), np.float32(r), next_obs[e, i], bool(next_dones[e, i])))
        obs, dones = next_obs, next_dones.copy()
        age += 1
        cut = np.zeros(N, dtype=bool)
        for e in range(N):
            if fin[e]:                        # `while not all(dones)` ended; the batch already reset this env
                assert next_dones[e].all()
                dones[e], age[e] = False, 0
            elif age[e] >= max_steps:         # `step < MAX_STEPS_PER_EPISODE` ended the episode: env.reset()
                cut[e] = True
                dones[e], age[e] = False, 0
        if cut.any():
            obs = batch.reset(mask=torch.as_tensor(cut.astype(np.uint8)).cuda()).cpu().numpy()
    return out


@pytest.mark.parametrize('kw,N', [(dict(num_snakes=4, height=12, width=12, snake_length=3), 6),
                                  (dict(num_snakes=3, height=10, width=14, snake_length=4, vision_range=3, frame_stack=2), 5)])
def test_collect_matches_the_reference_trainer_loop(kw, N):
    from marl_snake_b200 import DeviceReplayBuffer, SnakeBatch, collect
    steps, threshold, penalty, max_steps = 90, 10, -1.0, 40
    rew = {'fruit': 1.0, 'kill': 0.5, 'lose': -0.25, 'win': 2.0, 'time': -0.125}
    a = SnakeBatch(N, seed=11, reward_dict=rew, **kw)
    b = SnakeBatch(N, seed=11, reward_dict=rew, **kw)
    ns, oh, ow, ch = a.obs_shape
    net = ExactNet(ch, oh, ow).cuda()
    want = reference_style_loop(a, net, steps, threshold, penalty, max_steps)
    buf = DeviceReplayBuffer(len(want) + 100, b.obs_shape[1:], b.device)
    # two calls: the second continues the rollout (dead snakes stay dead across the call boundary)
    n = int(collect(b, net, 37, buf, epsilon=0.0, early_death=(threshold, penalty), max_steps=max_steps))
    n += int(collect(b, net, steps - 37, buf, epsilon=0.0, early_death=(threshold, penalty), max_steps=max_steps))
    assert n == len(want) == buf.size
    assert any(t[4] for t in want) and not all(t[4] for t in want)
    assert n < steps * N * ns                                   # some snakes were dead for a while: no transitions for them
    o = np.stack([t[0] for t in want]); o2 = np.stack([t[3] for t in want])
    assert np.array_equal(buf.obs[:n].cpu().numpy(), o)
    assert np.array_equal(buf.next_obs[:n].cpu().numpy(), o2)
    assert np.array_equal(buf.action[:n].cpu().numpy(), np.array([t[1] for t in want], dtype=np.uint8))
    assert np.array_equal(buf.reward[:n].cpu().numpy(), np.array([t[2] for t in want], dtype=np.float32))
    assert np.array_equal(buf.done[:n].cpu().numpy(), np.array([t[4] for t in want]))
    assert a.device_errors() == 0 and b.device_errors() == 0


def test_collect_resumes_from_device_state_after_foreign_steps():
    """a collect() that follows plain step() calls must not invent transitions for dead snakes."""
    from marl_snake_b200 import DeviceReplayBuffer, SnakeBatch, collect
    N, ns = 64, 4
    b = SnakeBatch(N, num_snakes=ns, height=10, width=10, vision_range=3, seed=5)
    b.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    for _ in range(25):
        b.step(torch.randint(0, 3, (N, ns), dtype=torch.uint8, device='cuda', generator=g))
    alive = b.get_state()['alive'].bool()
    assert not bool(alive.all())
    net = ExactNet(8, 7, 7).cuda()
    buf = DeviceReplayBuffer(1000, b.obs_shape[1:], b.device)
    n = int(collect(b, net, 1, buf, epsilon=0.0))
    assert n == int(alive.sum())
    assert bool((buf.obs[:n, 3, 3, 5] == 1).all())              # every stored state has its own head at the window centre
