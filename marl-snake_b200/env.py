"""Host-side mirror of the reference's SnakeEnv interface over libsnk.so.

`SnakeBatch`  N environments resident on one GPU; step()/reset() take and return CUDA tensors
              (the fast path; one kernel launch per step on the current torch stream).
`SnakeEnv`    drop-in for `gym.make('Snake-v1', **kw)` (marlenv/envs/snake_env.py:31): one
              environment, NumPy observation without batch dimension, list rewards / dones, dict info,
              same constructor kwargs, same exceptions.

PyTorch is used for device memory and streams only; every rule of the game runs in the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SnkConfig, SnkStateView, SnkStepExtra, check, lib
from .spaces import Box, Discrete

DEFAULT_REWARD_DICT = {'fruit': 10.0, 'kill': 0.0, 'lose': -0.5, 'win': 0.0, 'time': -0.001}   # snake_env.py:46-52
ACTION_ANGLE_DICT = {0: 0.0, 1: np.pi / 2.0, 2: -np.pi / 2.0}                                   # snake_env.py:40-44
DEFAULT_ACTION_DICT = {'noop': 0, 'left': 1, 'right': 2, 'down': 3, 'up': 4}                    # snake_env.py:32-38
DEFAULT_MAX_EPISODE_STEPS = 1e4                                                                 # snake_env.py:56


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class SnakeBatch:
    """A shard of `num_envs` Snake-v1 environments on one GPU."""

    def __init__(self, num_envs, height=20, width=20, num_snakes=4, snake_length=3, vision_range=None,
                 frame_stack=1, observer='snake', reward_dict=None, num_fruits=None,
                 max_episode_steps=DEFAULT_MAX_EPISODE_STEPS, device=0, seed=0, rng='philox',
                 auto_reset=True, env_id_offset=0, done_mode='all', wall_map=None, **ignored):
        reward_dict = DEFAULT_REWARD_DICT if reward_dict is None else reward_dict
        self.wall_map = None
        if wall_map is not None:                      # custom wall layout: the map decides the grid shape
            from .grid_util import as_wall_plane
            self.wall_map = as_wall_plane(wall_map)
            height, width = self.wall_map.shape
        if reward_dict.keys() != DEFAULT_REWARD_DICT.keys():                      # snake_env.py:77-80
            raise KeyError(f'reward dict keys must correspond to {DEFAULT_REWARD_DICT.keys()}')
        if observer not in ('snake', 'human'):
            raise ValueError("observer must be 'snake' (three relative actions) or 'human' (five absolute)")
        self.observer = observer
        self.action_dict = ACTION_ANGLE_DICT if observer == 'snake' else DEFAULT_ACTION_DICT   # snake_env.py:101-104
        if rng not in ('philox', 'replay'):
            raise ValueError("rng must be 'philox' or 'replay'")
        if not torch.cuda.is_available():
            raise RuntimeError('SnakeBatch needs a CUDA device: the step path has no CPU fallback')
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self.num_envs, self.num_snakes = int(num_envs), int(num_snakes)
        self.grid_shape = (int(height), int(width))
        self.snake_length = int(snake_length)
        self.vision_range = vision_range
        self.frame_stack = int(frame_stack)
        self.reward_dict = dict(reward_dict)
        self.num_fruits = int(round(num_snakes * 0.8)) if num_fruits is None else int(num_fruits)
        self.max_episode_steps = max_episode_steps
        self.auto_reset = bool(auto_reset)
        self.rng = rng

        cfg = SnkConfig(abi_version=_lib.SNK_ABI_VERSION, device=self.device.index, num_envs=self.num_envs,
                        height=height, width=width, num_snakes=num_snakes, snake_length=snake_length,
                        vision_range=int(vision_range or 0), frame_stack=frame_stack,
                        num_fruits=self.num_fruits, auto_reset=int(self.auto_reset),
                        done_mode={'all': 0, 'any': 1}[done_mode],
                        rng_mode=_lib.SNK_RNG_REPLAY if rng == 'replay' else _lib.SNK_RNG_PHILOX,
                        observer=1 if observer == 'human' else 0,
                        seed=int(seed), env_id_offset=int(env_id_offset),
                        max_episode_steps=float(max_episode_steps),
                        reward_fruit=float(reward_dict['fruit']), reward_kill=float(reward_dict['kill']),
                        reward_lose=float(reward_dict['lose']), reward_win=float(reward_dict['win']),
                        reward_time=float(reward_dict['time']))
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            if self.wall_map is None:
                check(lib.snk_create(C.byref(cfg), C.byref(self._h)))
            else:
                check(lib.snk_create_map(C.byref(cfg), self.wall_map.ctypes.data_as(C.c_void_p), C.byref(self._h)))
        shape = (C.c_int32 * 4)()
        check(lib.snk_obs_shape(self._h, C.byref(shape)))
        self.obs_shape = tuple(shape)                                  # (ns, oh, ow, 8*fs)
        self.obs_ch = self.obs_shape[-1]
        self.bits_shape = self.obs_shape[:-1] + (self.obs_shape[-1] // 8,)   # channel-bit form: one byte per cell and frame
        self._bits = None
        N, ns, dev = self.num_envs, self.num_snakes, self.device
        self._obs = torch.empty((N,) + self.obs_shape, dtype=torch.uint8, device=dev)
        self._rew = torch.empty((N, ns), dtype=torch.float64, device=dev)
        self._done = torch.empty((N, ns), dtype=torch.uint8, device=dev)
        self._fin = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self._rank = torch.zeros((N, ns), dtype=torch.int32, device=dev)
        self._ep_scores = torch.zeros((N, ns), dtype=torch.float64, device=dev)
        self._ep_steps = torch.zeros((N, ns), dtype=torch.int32, device=dev)
        self._ep_fruits = torch.zeros((N, ns), dtype=torch.int32, device=dev)
        self._ep_kills = torch.zeros((N, ns), dtype=torch.int32, device=dev)
        self._extra = SnkStepExtra(self._fin.data_ptr(), self._rank.data_ptr(), self._ep_scores.data_ptr(),
                                   self._ep_steps.data_ptr(), self._ep_fruits.data_ptr(),
                                   self._ep_kills.data_ptr())
        self.observation_space = Box(0, 1, (N,) + self.obs_shape, np.uint8)
        self.action_space = Discrete(len(self.action_dict))
        self._has_obs = False            # self._obs holds the current observation of every env
        self._rollout_alive = None       # rollout.collect's (alive, episode age) cache; dropped by any other mutation
        self._rollout_age = None

    # ---- lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            lib.snk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- hot path ----------------------------------------------------------------------------------
    def _bits_buffer(self):
        if self._bits is None:
            self._bits = torch.zeros((self.num_envs,) + self.bits_shape, dtype=torch.uint8, device=self.device)
        return self._bits

    def reset(self, mask=None, copy=False, bits=False):
        """SnakeEnv.reset for all envs (or those with mask != 0). Returns uint8 [N, ns, oh, ow, 8*fs], or with
        bits=True the channel-bit form uint8 [N, ns, oh, ow, fs] (see step_bits)."""
        if mask is not None:
            if mask.numel() != self.num_envs:
                raise ValueError('mask must have one entry per environment')
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            if bits:
                out = self._bits_buffer()
                check(lib.snk_reset_bits(self._h, _ptr(mask), _ptr(out), self._stream()))
            else:
                out = self._obs
                check(lib.snk_reset(self._h, _ptr(mask), _ptr(out), self._stream()))
                self._has_obs = True
        self._rollout_alive = None
        return out.clone() if copy else out

    def step_bits(self, actions, copy=False, want_info=True):
        """As step(), but the observation comes as channel bits: uint8 [N, ns, oh, ow, fs], one byte per window cell
        and frame, bit c = channel c (np.packbits(obs.reshape(..., fs, 8), -1, bitorder='little')).  The encode
        writes them directly -- no NHWC block, an eighth of the observation traffic; `unpack_obs` widens them."""
        if actions.dtype != torch.uint8 or actions.device != self.device or not actions.is_contiguous() \
                or actions.numel() != self.num_envs * self.num_snakes:
            raise ValueError('actions must be a contiguous uint8 CUDA tensor of shape [num_envs, num_snakes]')
        out = self._bits_buffer()
        check(lib.snk_step_bits(self._h, _ptr(actions), _ptr(out), _ptr(self._rew), _ptr(self._done),
                                C.byref(self._extra) if want_info else None, self._stream()))
        info = {}
        if want_info:
            info = dict(finished=self._fin.view(torch.bool), rank=self._rank, episode_scores=self._ep_scores,
                        episode_steps=self._ep_steps, episode_fruits=self._ep_fruits,
                        episode_kills=self._ep_kills)
        res = (out, self._rew, self._done.view(torch.bool), info)
        if copy:
            res = (out.clone(), res[1].clone(), res[2].clone(), {k: v.clone() for k, v in info.items()})
        return res

    def step(self, actions, copy=False, want_obs=True, want_info=True):
        """actions: uint8 CUDA tensor [N, ns] in {0,1,2} (observer 'snake') or {0..4} ('human').  Returns (obs, rewards f64, dones bool, info)
        where info holds per-env `finished` and, for finished envs, the terminal `rank` / `episode_*`
        arrays the reference puts in its info dict (snake_env.py:396-410).  The returned tensors are
        the batch's own buffers, overwritten by the next call unless copy=True."""
        if actions.dtype != torch.uint8 or actions.device != self.device or not actions.is_contiguous() \
                or actions.numel() != self.num_envs * self.num_snakes:
            raise ValueError('actions must be a contiguous uint8 CUDA tensor of shape [num_envs, num_snakes]')
        check(lib.snk_step(self._h, _ptr(actions), _ptr(self._obs if want_obs else None), _ptr(self._rew),
                           _ptr(self._done), C.byref(self._extra) if want_info else None, self._stream()))
        self._has_obs, self._rollout_alive = self._has_obs and want_obs, None
        info = {}
        if want_info:
            info = dict(finished=self._fin.view(torch.bool), rank=self._rank, episode_scores=self._ep_scores,
                        episode_steps=self._ep_steps, episode_fruits=self._ep_fruits,
                        episode_kills=self._ep_kills)
        out = (self._obs if want_obs else None, self._rew, self._done.view(torch.bool), info)
        if copy:
            out = (None if out[0] is None else out[0].clone(), out[1].clone(), out[2].clone(),
                   {k: v.clone() for k, v in info.items()})
        return out

    def step_many(self, actions, obs=None, rewards=None, dones=None, bits=None, finished=None):
        """T steps in one call (snk_step_many) for open-loop action streams.  actions: uint8 CUDA [T, N, ns].
        rewards f64 / dones uint8 [T, N, ns] are allocated when not given.  obs (uint8) / bits: a [T, N, ...] tensor
        receives every step's block, an [N, ...] tensor only the last step's; None skips that output.  finished:
        optional uint8 [T, N].  Returns (obs, rewards, dones, bits).  Bit-identical to T calls of step()."""
        N, ns = self.num_envs, self.num_snakes
        if actions.dtype != torch.uint8 or actions.device != self.device or not actions.is_contiguous() \
                or actions.dim() != 3 or tuple(actions.shape[1:]) != (N, ns):
            raise ValueError('actions must be a contiguous uint8 CUDA tensor of shape [T, num_envs, num_snakes]')
        T = actions.shape[0]
        if rewards is None:
            rewards = torch.empty((T, N, ns), dtype=torch.float64, device=self.device)
        if dones is None:
            dones = torch.empty((T, N, ns), dtype=torch.uint8, device=self.device)
        every = None
        for name, buf, shape in (('obs', obs, self.obs_shape), ('bits', bits, self.bits_shape)):
            if buf is None:
                continue
            if buf.dtype != torch.uint8 or buf.device != self.device or not buf.is_contiguous():
                raise ValueError(f'{name} must be a contiguous uint8 CUDA tensor')
            if tuple(buf.shape) == (T, N) + shape and (T > 1 or buf.dim() == len(shape) + 2):
                e = True
            elif tuple(buf.shape) == (N,) + shape:
                e = False
            else:
                raise ValueError(f'{name} must have shape [T, N, ...] (every step) or [N, ...] (last step only)')
            if every is not None and every != e:
                raise ValueError('obs and bits must both hold every step or both hold the last step only')
            every = e
        extra = None
        if finished is not None:
            if tuple(finished.shape) != (T, N) or finished.dtype != torch.uint8 or not finished.is_contiguous():
                raise ValueError('finished must be a contiguous uint8 CUDA tensor of shape [T, N]')
            extra = SnkStepExtra(finished.data_ptr(), 0, 0, 0, 0, 0)
        check(lib.snk_step_many(self._h, T, _ptr(actions), _ptr(obs), _ptr(bits), int(bool(every)), _ptr(rewards),
                                _ptr(dones), C.byref(extra) if extra is not None else None, self._stream()))
        return obs, rewards, dones, bits

    def step_host(self, actions, obs, rewards, dones):
        """Reference-shaped call on HOST buffers (pinned NumPy/torch CPU arrays): copies actions in,
        steps, copies obs/rewards/dones out, synchronises.  `obs` may be None."""
        check(lib.snk_step_host(self._h, C.c_void_p(actions.data_ptr()),
                                C.c_void_p(obs.data_ptr()) if obs is not None else C.c_void_p(0),
                                C.c_void_p(rewards.data_ptr()), C.c_void_p(dones.data_ptr())))

    def step_host_bits(self, actions, bits, rewards, dones):
        """As step_host, but `bits` (uint8, obs shape with the last dimension divided by 8) receives channel
        bits -- one byte per (cell, frame) -- and nothing is widened on the host."""
        check(lib.snk_step_host_bits(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(bits.data_ptr()),
                                     C.c_void_p(rewards.data_ptr()), C.c_void_p(dones.data_ptr())))

    def reset_host(self, obs):
        check(lib.snk_reset_host(self._h, C.c_void_p(obs.data_ptr()) if obs is not None else C.c_void_p(0)))
        self._host_reset_done = True

    def set_host_transport(self, mode, threads=0):
        """How step_host / reset_host move the observation block over PCIe: 'raw' (one D2H of the NHWC
        bytes) or 'packed' (channel bits on the link, widened by `threads` host threads; 0 = all cores)."""
        check(lib.snk_set_host_transport(self._h, {'raw': _lib.SNK_XFER_RAW, 'packed': _lib.SNK_XFER_PACKED}[mode],
                                         int(threads)))

    def host_transport(self):
        mode, threads = C.c_int(0), C.c_int(0)
        check(lib.snk_get_host_transport(self._h, C.byref(mode), C.byref(threads)))
        return ('raw', 'packed')[mode.value], threads.value

    def pack_obs(self, obs=None, out=None):
        """Channel-bit form of an observation tensor (uint8 0/1, last dim a multiple of 8): one byte per
        (cell, frame), bit c = channel c.  8x smaller for device replay buffers; `unpack_obs` inverts it."""
        obs = self._obs if obs is None else obs
        if obs.dtype != torch.uint8 or not obs.is_cuda or not obs.is_contiguous() or obs.shape[-1] % 8:
            raise ValueError('obs must be a contiguous uint8 CUDA tensor whose last dimension is a multiple of 8')
        shape = (*obs.shape[:-1], obs.shape[-1] // 8)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=obs.device)
        check(lib.snk_pack_obs(self._h, _ptr(obs), _ptr(out), obs.numel(), self._stream()))
        return out

    # ---- parity / checkpoint interface ---------------------------------------------------------------
    def get_state(self, max_cells=None):
        N, ns, dev = self.num_envs, self.num_snakes, self.device
        H, W = self.grid_shape
        i32 = dict(dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        st = dict(grid=torch.empty((N, H * W), **u8), head=torch.empty((N, ns), **i32),
                  tail=torch.empty((N, ns), **i32), length=torch.empty((N, ns), **i32),
                  dir=torch.empty((N, ns), **u8), alive=torch.empty((N, ns), **u8),
                  alive_counter=torch.empty((N,), **i32), episode_length=torch.empty((N,), **i32))
        if max_cells:
            st['cells'] = torch.empty((N, ns, max_cells), **i32)
        view = SnkStateView(**{k: v.data_ptr() for k, v in st.items()}, max_cells=int(max_cells or 0))
        check(lib.snk_get_state(self._h, C.byref(view), self._stream()))
        st['grid'] = st['grid'].view(N, H, W)
        return st

    def set_state(self, grid, alive, dir, length, cells, alive_counter, episode_length=None):
        """Force a state (scenario tests / restore). cells: int32 [N, ns, L] head-first flat cell indices."""
        N, ns, dev = self.num_envs, self.num_snakes, self.device

        def t(x, dtype, shape):
            return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).reshape(shape).to(dev).contiguous()
        cells = t(cells, torch.int32, (N, ns, -1))
        H, W = self.grid_shape
        length_h = np.asarray(length).reshape(N, ns)
        if cells.shape[-1] < 2 or int(length_h.max(initial=0)) > cells.shape[-1] or int(length_h.min(initial=0)) < 0:
            raise ValueError('length must be within 0..cells.shape[-1] (and cells hold at least two columns)')
        if int(cells.min()) < 0 or int(cells.max()) >= H * W:
            raise ValueError('cells must be flat indices into the H x W grid (pad unused entries with 0)')
        if episode_length is None:
            episode_length = np.zeros(N, dtype=np.int32)
        keep = dict(grid=t(grid, torch.uint8, (N, -1)), alive=t(alive, torch.uint8, (N, ns)),
                    dir=t(dir, torch.uint8, (N, ns)), length=t(length, torch.int32, (N, ns)), cells=cells,
                    alive_counter=t(alive_counter, torch.int32, (N,)),
                    episode_length=t(episode_length, torch.int32, (N,)))
        view = SnkStateView(**{k: v.data_ptr() for k, v in keep.items()}, max_cells=cells.shape[-1])
        check(lib.snk_set_state(self._h, C.byref(view), _ptr(self._obs), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()      # `keep` must outlive the kernels
        self._has_obs, self._rollout_alive = True, None
        return self._obs

    def save_checkpoint(self):
        """Exact state of the shard as a NumPy byte array (records, frame histories, statistics); a batch
        created with the same arguments continues bit for bit after load_checkpoint()."""
        n = int(lib.snk_checkpoint_bytes(self._h))
        blob = np.empty(n, dtype=np.uint8)
        check(lib.snk_checkpoint_save(self._h, blob.ctypes.data_as(C.c_void_p), n))
        return blob

    def load_checkpoint(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        check(lib.snk_checkpoint_load(self._h, blob.ctypes.data_as(C.c_void_p), blob.size))
        self._has_obs, self._rollout_alive = False, None

    def set_replay(self, draws_per_env):
        """draws_per_env: sequence of N int arrays -- recorded draw outputs in consumption order."""
        assert len(draws_per_env) == self.num_envs
        off = np.zeros(self.num_envs + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(d) for d in draws_per_env])
        flat = np.concatenate([np.asarray(d, dtype=np.int32).reshape(-1) for d in draws_per_env]) \
            if off[-1] else np.zeros(1, dtype=np.int32)
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        check(lib.snk_set_replay(self._h, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))

    def replay_cursors(self):
        out = np.zeros(self.num_envs, dtype=np.int32)
        check(lib.snk_replay_cursors(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def device_errors(self, clear=False):
        bits = C.c_uint32(0)
        check(lib.snk_device_errors(self._h, C.byref(bits), int(clear)))
        return bits.value

    def raise_on_device_errors(self):
        bits = self.device_errors(clear=True)
        if bits & 1:
            raise KeyError('action outside {0, 1, 2}')                 # the reference's KeyError (:606)
        if bits:
            raise _lib.SnkError('; '.join(m for b, m in _lib.DEV_ERRORS.items() if bits & b))

    # ---- rollout statistics --------------------------------------------------------------------------
    def stats(self, clear=False):
        out = np.zeros(8, dtype=np.float64)
        check(lib.snk_stats(self._h, out.ctypes.data_as(C.c_void_p), int(clear)))
        return dict(zip(_lib.STAT_NAMES[:7], out[:7]))

    def stats_tensor(self):
        """The device-resident statistics vector (8 doubles) as a torch tensor sharing memory, ready
        for an in-place torch.distributed.all_reduce at the end of a rollout."""
        p = C.c_void_p()
        check(lib.snk_stats_dev(self._h, C.byref(p)))
        return _tensor_from_ptr(p.value, 8, torch.float64, self.device)

    def algorithmic_bytes_per_env_step(self):
        return int(lib.snk_algorithmic_bytes_per_env_step(self._h))


def _tensor_from_ptr(ptr, n, dtype, device):
    class _Holder:
        pass
    h = _Holder()
    itemsize = torch.empty((), dtype=dtype).element_size()
    h.__cuda_array_interface__ = dict(shape=(n,), typestr={8: '<f8', 4: '<f4'}[itemsize], data=(ptr, False),
                                      version=2)
    return torch.as_tensor(h, device=device)


class SnakeEnv:
    """Drop-in for the reference's SnakeEnv (one environment, host-side return types).

    reset() -> np.uint8 [ns, oh, ow, 8*fs]; step(actions) -> (obs, list[float], list[bool], dict).
    Unknown kwargs are ignored as the reference does (snake_env.py:68-69); `reward_func` is accepted as
    an alias of `reward_dict` (README spelling)."""

    default_reward_dict = DEFAULT_REWARD_DICT
    default_action_dict = DEFAULT_ACTION_DICT
    action_angle_dict = ACTION_ANGLE_DICT
    reward_keys = DEFAULT_REWARD_DICT.keys()
    metadata = {}
    _done_mode = 'all'

    def __init__(self, height=20, width=20, num_snakes=4, snake_length=3, vision_range=None, frame_stack=1,
                 observer='snake', *args, **kwargs):
        if 'reward_dict' not in kwargs and isinstance(kwargs.get('reward_func'), dict):
            kwargs['reward_dict'] = kwargs.pop('reward_func')
        reward_dict = kwargs.pop('reward_dict', DEFAULT_REWARD_DICT)
        self._batch = SnakeBatch(1, height=height, width=width, num_snakes=num_snakes,
                                 snake_length=snake_length, vision_range=vision_range, frame_stack=frame_stack,
                                 observer=observer, reward_dict=reward_dict,
                                 num_fruits=kwargs.pop('num_fruits', None),
                                 max_episode_steps=kwargs.pop('max_episode_steps', DEFAULT_MAX_EPISODE_STEPS),
                                 device=kwargs.pop('device', 0), seed=kwargs.pop('seed', 0),
                                 rng=kwargs.pop('rng', 'philox'), auto_reset=False, done_mode=self._done_mode,
                                 wall_map=kwargs.pop('wall_map', None))
        b = self._batch
        self.reward_dict = b.reward_dict
        self.max_episode_steps = b.max_episode_steps
        self.num_snakes, self.num_fruits = b.num_snakes, b.num_fruits
        self.grid_shape, self.snake_length = b.grid_shape, b.snake_length
        self.vision_range, self.frame_stack, self.observer = vision_range, b.frame_stack, observer
        self.low, self.high, self.image_obs = 0, 1, False
        self.action_dict = b.action_dict
        self.obs_ch = b.obs_ch
        self.action_space = Discrete(len(self.action_dict) * self.num_snakes)             # snake_env.py:107-109
        self.observation_space = Box(self.low, self.high, b.obs_shape, np.uint8)          # snake_env.py:115-129
        # host buffers of the single-environment call: one C-ABI call per step (snk_step_host_info: actions
        # in, step, observation / rewards / dones / terminal info out, one synchronisation)
        ns = self.num_snakes
        self._h_act = np.zeros((1, ns), dtype=np.uint8)
        self._h_obs = np.empty((1,) + b.obs_shape, dtype=np.uint8)
        self._h_rew = np.empty((1, ns), dtype=np.float64)
        self._h_done = np.empty((1, ns), dtype=np.uint8)
        self._h_fin = np.zeros(1, dtype=np.uint8)
        self._h_rank = np.zeros((1, ns), dtype=np.int32)
        self._h_scores = np.zeros((1, ns), dtype=np.float64)
        self._h_counts = np.zeros((3, 1, ns), dtype=np.int32)
        self._xh = SnkStepExtra(self._h_fin.ctypes.data, self._h_rank.ctypes.data, self._h_scores.ctypes.data,
                                self._h_counts[0].ctypes.data, self._h_counts[1].ctypes.data,
                                self._h_counts[2].ctypes.data)
        # the buffers never move: their addresses (and the handle) are converted for ctypes once, not on every step
        self._step_args = (self._batch._h,) + tuple(x.ctypes.data_as(C.c_void_p) for x in
                                                    (self._h_act, self._h_obs, self._h_rew, self._h_done)) + (C.byref(self._xh),)

    @property
    def unwrapped(self):
        return self

    def seed(self, seed=42):
        """As in the reference (snake_env.py:161-163) this has no effect on the game: there it only creates
        `self.np_random`, which nothing reads -- spawns and fruits draw from the module-global NumPy generator.  The
        stream of this environment is fixed by the constructor's `seed=` (Philox key) or by `set_replay`."""
        return [seed]

    def close(self):
        self._batch.close()

    def reset(self):
        check(lib.snk_reset_host(self._batch._h, self._h_obs.ctypes.data_as(C.c_void_p)))
        self.frame_buffer = []                                                            # snake_env.py:151
        return self._h_obs[0].copy()

    def _alive_now(self):
        return self._batch.get_state()['alive'][0].cpu().numpy().astype(bool)

    def step(self, actions):
        if isinstance(actions, int):
            actions = [actions]
        assert len(actions) == self.num_snakes                                            # snake_env.py:313
        alive = None
        for i, ac in enumerate(actions):
            if isinstance(ac, np.ndarray):
                ac = ac.item()
            if self.observer == 'human':              # unknown actions keep the direction (:610-632)
                self._h_act[0, i] = int(ac) if ac in (0, 1, 2, 3, 4) else 0
                continue
            if ac not in self.action_dict:            # the reference looks the action up only for live snakes
                alive = self._alive_now() if alive is None else alive
                if alive[i]:
                    raise KeyError(ac)                                                    # snake_env.py:606
                ac = 0
            self._h_act[0, i] = int(ac)
        check(lib.snk_step_host_info(*self._step_args))
        info = {}
        if self._h_fin[0]:
            info['rank'] = [int(r) for r in self._h_rank[0]]
            info['episode_scores'] = self._h_scores[0].copy()
            for k, name in enumerate(('episode_steps', 'episode_fruits', 'episode_kills')):
                info[name] = self._h_counts[k, 0].astype(np.float64)
        return (self._h_obs[0].copy(), self._h_rew[0].tolist(), self._h_done[0].astype(bool).tolist(), info)

    @property
    def grid(self):
        """The H x W cell-code grid as the reference keeps it (np.int64)."""
        return self._batch.get_state()['grid'][0].cpu().numpy().astype(np.int64)

    @property
    def alive_snakes(self):
        return int(self._batch.get_state()['alive_counter'][0])

    def render(self, mode='ascii'):
        if mode == 'ascii':
            sym = {0: '.', 1: '#', 2: 'o', 3: 'H', 4: 'b', 5: 't'}
            print('\n'.join(''.join(sym[int(v) % 10] for v in row) for row in self.grid))
        elif mode == 'rgb_array':
            from .render import rgb_from_grid
            return rgb_from_grid(self.grid)
        elif mode == 'gif':                       # append a frame; save_gif() writes them out (snake_env.py:283-288)
            from .render import image_from_grid
            self.frame_buffer.append(image_from_grid(self.grid))

    def save_gif(self, fp=None):
        """Write the frames collected by render('gif') as an animated GIF (snake_env.py:419-436): to `fp`
        (path or file object), or to ./tmp/<timestamp>.gif.  Returns fp."""
        import datetime
        import os
        import warnings
        if fp is None:
            save_dir = os.path.join(os.getcwd(), 'tmp')
            os.makedirs(save_dir, exist_ok=True)
            fp = os.path.join(save_dir, '{}.gif'.format(datetime.datetime.now().strftime('%Y%m%d%H%M%S')))
        if not getattr(self, 'frame_buffer', None):
            warnings.warn("You must call render('gif') first. No images to save.")
        else:
            self.frame_buffer[0].save(fp, save_all=True, append_images=self.frame_buffer[1:], format='GIF', loop=0)
        return fp

    def render_fancy(self, cell_size=40, save_path=None):
        """RGB frame of the current state, pixel for pixel the reference's drawing (snake_env.py:165-263; see
        render.py).  The state -- grid, per-snake body cells, headings -- is exported from the device."""
        from .render import render_fancy
        longest = max(int(self._batch.get_state()['length'].max()), 2)
        st = self._batch.get_state(max_cells=longest)
        img = render_fancy(st['grid'][0].cpu().numpy(), st['cells'][0].cpu().numpy(), st['alive'][0].cpu().numpy(),
                           st['dir'][0].cpu().numpy(), cell_size)
        if save_path:
            from PIL import Image
            Image.fromarray(img).save(save_path)
            print(f"Saved fancy render to {save_path}")
        return img


class CoopSnakeEnv(SnakeEnv):
    """SnakeCoop-v1: the episode ends as soon as any snake is done (coop_snake_env.py:14-22)."""
    _done_mode = 'any'
