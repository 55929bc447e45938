"""the single-env drawings are pixel-identical to the reference's (tests/golden/render.npz holds frames of
the unmodified reference: render_fancy, render('rgb_array'), one render('gif') frame, with the states they show)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'render.npz')


def frames():
    z = np.load(GOLDEN)
    return [(str(n), {k.split('__')[1]: z[k] for k in z.files if k.startswith(str(n) + '__')}) for n in z['names']]


FRAMES = frames()


@pytest.mark.parametrize('name,f', FRAMES, ids=[n for n, _ in FRAMES])
def test_drawings_match_the_reference_pixels(name, f):
    from marl_snake_b200 import render
    cs = int(f['cell_size'])
    assert np.array_equal(render.rgb_from_grid(f['grid']), f['rgb'])
    assert np.array_equal(np.asarray(render.image_from_grid(f['grid'])), f['gif'])
    got = render.render_fancy(f['grid'], f['cells'], f['alive'], f['dir'], cs)
    assert got.shape == f['fancy'].shape and got.dtype == np.uint8
    assert np.array_equal(got, f['fancy']), int((got != f['fancy']).any(-1).sum())


def test_golden_frames_cover_dead_snakes_and_the_colour_wheel():
    assert any((f['alive'] == 0).any() for _, f in FRAMES)
    assert any(len(f['alive']) > 4 for _, f in FRAMES)
    assert any((f['grid'] == 2).any() for _, f in FRAMES)


@pytest.mark.gpu
@pytest.mark.parametrize('name,f', FRAMES[:5], ids=[n for n, _ in FRAMES[:5]])
def test_env_render_calls_from_device_state(name, f):
    """SnakeEnv.render_fancy / render('rgb_array') through the C ABI's state export: force the golden state onto the
    device, draw, compare with the reference's pixels."""
    from marl_snake_b200 import SnakeEnv
    H, W = f['grid'].shape
    ns = len(f['alive'])
    env = SnakeEnv(height=H, width=W, num_snakes=ns, snake_length=3)
    env.reset()
    L = f['cells'].shape[-1]
    env._batch.set_state(f['grid'][None], f['alive'][None], f['dir'][None], f['len'][None],
                         np.maximum(f['cells'], 0).reshape(1, ns, L), np.array([int(f['alive'].sum())]))
    assert np.array_equal(env.render_fancy(cell_size=int(f['cell_size'])), f['fancy'])
    assert np.array_equal(env.render('rgb_array'), f['rgb'])
    env.render('gif')
    assert np.array_equal(np.asarray(env.frame_buffer[-1]), f['gif'])
    env.close()
