"""correlate1d / correlate2d (timelapse kernel, config 5) and the norms against
the numpy restatement of signal/correlate.py and signal/normalize.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcorr_b200 import signal
    from oracle import cc_oracle as orc
    return torch, signal, orc


def spectrogram_like(rng, shape):
    from scipy.ndimage import uniform_filter
    return uniform_filter(rng.standard_normal(shape) ** 2, size=3) + 0.01


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("shape", [(5, 7), (20, 12), (64, 33), (3, 40, 16)])
def test_correlate2d_matches_reference_restatement(env, dtype, tol, shape):
    torch, signal, orc = env
    rng = np.random.default_rng(sum(shape))
    f = spectrogram_like(rng, shape)
    g = np.roll(f, (2, -3), axis=(-2, -1)) + 0.05 * spectrogram_like(rng, shape)
    got = signal.correlate2d(f.astype(dtype), g.astype(dtype))
    ref = orc.correlate2d(f, g)
    assert got.shape == ref.shape and got.dtype == dtype
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    a = np.argmax(got.reshape(*got.shape[:-2], -1), axis=-1)
    b = np.argmax(ref.reshape(*ref.shape[:-2], -1), axis=-1)
    assert np.array_equal(a, b)
    raw = signal.correlate2d(f.astype(dtype), g.astype(dtype), normalize=False)
    assert np.max(np.abs(raw - orc.correlate2d(f, g, normalize=False))) <= tol * np.max(np.abs(raw))


def test_correlate1d_and_norms(env):
    torch, signal, orc = env
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 1000))
    y = np.roll(x, 11, axis=-1) + 0.1 * rng.standard_normal((4, 1000))
    got = signal.correlate1d(x, y)
    ref = orc.correlate1d(x, y)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.all(np.argmax(got, axis=-1) == 999 - 11)
    xt = torch.from_numpy(x).cuda()
    assert np.allclose(signal.norm1d(xt).cpu().numpy(), orc.norm1d(x), rtol=1e-14)
    img = rng.standard_normal((3, 8, 9))
    assert np.allclose(signal.norm2d(torch.from_numpy(img).cuda()).cpu().numpy(), orc.norm2d(img), rtol=1e-14)


def test_timelapse_shape_day_pairs(env):
    """Config 5 shape: 1024 x 128 spectrograms, every image transformed once and
    reused by its pairs; known (dF, dT) shifts between days are recovered."""
    torch, signal, orc = env
    rng = np.random.default_rng(5)
    F, T, D = 1024, 128, 6
    base = spectrogram_like(rng, (F + 40, T + 40))
    shifts = [(0, 0), (1, 2), (3, -1), (-2, 4), (5, 5), (-4, -3)]
    days = np.stack([base[20 + a:20 + a + F, 20 + b:20 + b + T] + 0.02 * spectrogram_like(rng, (F, T))
                     for a, b in shifts]).astype(np.float32)
    pairs = [(i, j) for i in range(D) for j in range(i + 1, D)]
    plan = signal.Correlate2dPlan(F, T, D, pairs, dtype=torch.float32)
    assert plan.n == 262_144 and int(np.prod(plan.plan.dims)) == 262_144
    cc2 = plan(torch.from_numpy(days).cuda())
    val, dF, dT = plan.argmax(cc2)
    for q, (i, j) in enumerate(pairs):
        # in1 = day i, in2 = day j: day_j[u] = base[u + s_j]  ->  peak at s_j - s_i ... measured from in2 to in1
        assert (int(dF[q]), int(dT[q])) == (shifts[j][0] - shifts[i][0], shifts[j][1] - shifts[i][1]), (i, j)
    ref = orc.correlate2d(days[1].astype(np.float64), days[4].astype(np.float64))
    q = pairs.index((1, 4))
    got = cc2[q].cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert np.unravel_index(np.argmax(ref), ref.shape) == (int(dF[q]) + F - 1, int(dT[q]) + T - 1)
    assert abs(float(val[q]) - ref.max()) <= 1e-5
