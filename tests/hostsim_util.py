"""Builds and binds tests/hostsim/libsnk_hostsim.so: the device rule source (snk_core.cuh) compiled for
the host.  TEST INFRASTRUCTURE -- lets `-m "not gpu"` check the kernel's rule code against golden."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'hostsim', 'snk_hostsim.cpp')
OUT = os.path.join(ROOT, 'tests', 'hostsim', 'libsnk_hostsim.so')
DEPS = [SRC, os.path.join(ROOT, 'marl-snake_b200', 'csrc', 'snk_core.cuh'),
        os.path.join(ROOT, 'marl-snake_b200', 'csrc', 'snk_spawn.cpp')]


def build():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I/usr/local/cuda/include',
                               '-o', OUT, SRC, DEPS[2]])
    return OUT


def make_config(num_envs, kw, rng_mode=1, auto_reset=1, seed=0, env_id_offset=0, done_mode=0):
    from marl_snake_b200._lib import SnkConfig
    rd = kw.get('reward_dict', {'fruit': 10.0, 'kill': 0.0, 'lose': -0.5, 'win': 0.0, 'time': -0.001})
    ns = kw.get('num_snakes', 4)
    H, W = kw.get('height', 20), kw.get('width', 20)
    if kw.get('wall_map') is not None:
        H, W = np.asarray(kw['wall_map']).shape
    return SnkConfig(abi_version=1, device=0, num_envs=num_envs, height=H,
                     width=W, num_snakes=ns, snake_length=kw.get('snake_length', 3),
                     vision_range=int(kw.get('vision_range') or 0), frame_stack=kw.get('frame_stack', 1),
                     num_fruits=kw.get('num_fruits', int(round(ns * 0.8))), auto_reset=auto_reset,
                     done_mode=done_mode, rng_mode=rng_mode, observer=int(kw.get('observer', 'snake') == 'human'),
                     seed=seed, env_id_offset=env_id_offset,
                     max_episode_steps=float(kw.get('max_episode_steps', 1e4)),
                     reward_fruit=rd['fruit'], reward_kill=rd['kill'], reward_lose=rd['lose'],
                     reward_win=rd['win'], reward_time=rd['time'])


class HostSim:
    def __init__(self, num_envs, kw, **cfg_kw):
        self.lib = C.CDLL(build())
        self.lib.hs_create.restype = C.c_void_p
        self.lib.hs_create_map.restype = C.c_void_p
        self.cfg = make_config(num_envs, kw, **cfg_kw)
        if kw.get('wall_map') is not None:
            walls = np.ascontiguousarray(np.asarray(kw['wall_map']) != 0, dtype=np.uint8)
            self.h = C.c_void_p(self.lib.hs_create_map(C.byref(self.cfg), walls.ctypes.data_as(C.c_void_p)))
        else:
            self.h = C.c_void_p(self.lib.hs_create(C.byref(self.cfg)))
        self.N, self.ns = num_envs, self.cfg.num_snakes
        H, W = self.cfg.height, self.cfg.width
        self.HW = H * W
        vr, fs = self.cfg.vision_range, self.cfg.frame_stack
        oh, ow = (2 * vr + 1, 2 * vr + 1) if vr else (H, W)
        self.obs_shape = (num_envs, self.ns, oh, ow, 8 * fs)

    def set_replay(self, draws):
        off = np.zeros(self.N + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(d) for d in draws])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.int32) for d in draws] +
                                                   [np.zeros(1, np.int32)]), dtype=np.int32)
        self.lib.hs_set_replay(self.h, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p))

    def reset(self):
        obs = np.zeros(self.obs_shape, dtype=np.uint8)
        self.lib.hs_reset(self.h, obs.ctypes.data_as(C.c_void_p))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8).reshape(self.N, self.ns)
        obs = np.zeros(self.obs_shape, dtype=np.uint8)
        rew = np.zeros((self.N, self.ns), dtype=np.float64)
        done = np.zeros((self.N, self.ns), dtype=np.uint8)
        fin = np.zeros(self.N, dtype=np.uint8)
        rank = np.zeros((self.N, self.ns), dtype=np.int32)
        sc = np.zeros((self.N, self.ns), dtype=np.float64)
        st, fr, kl = (np.zeros((self.N, self.ns), dtype=np.int32) for _ in range(3))
        p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.lib.hs_step(self.h, p(a), p(obs), p(rew), p(done), p(fin), p(rank), p(sc), p(st), p(fr), p(kl))
        return obs, rew, done, dict(finished=fin, rank=rank, episode_scores=sc, episode_steps=st,
                                    episode_fruits=fr, episode_kills=kl)

    def grid(self):
        g = np.zeros((self.N, self.HW), dtype=np.uint8)
        c = np.zeros(self.N, dtype=np.int32)
        cur = np.zeros(self.N, dtype=np.int32)
        p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.lib.hs_get_grid(self.h, p(g), p(c), p(cur))
        return g, c, cur

    def set_state(self, grid, alive, dir, length, cells, counter, ep_len=None):
        N, ns = self.N, self.ns
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(N, ns, -1)
        ep_len = np.zeros(N, np.int32) if ep_len is None else np.ascontiguousarray(ep_len, np.int32)
        args = [np.ascontiguousarray(grid, np.uint8).reshape(N, -1), np.ascontiguousarray(alive, np.uint8),
                np.ascontiguousarray(dir, np.uint8), np.ascontiguousarray(length, np.int32), cells]
        obs = np.zeros(self.obs_shape, dtype=np.uint8)
        p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.lib.hs_set_state(self.h, p(args[0]), p(args[1]), p(args[2]), p(args[3]), p(args[4]),
                              C.c_int(cells.shape[-1]), p(np.ascontiguousarray(counter, np.int32).reshape(N)),
                              p(ep_len), p(obs))
        return obs

    def errors(self):
        self.lib.hs_errors.restype = C.c_uint32
        return self.lib.hs_errors(self.h)

    def close(self):
        self.lib.hs_destroy(self.h)
