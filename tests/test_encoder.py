"""Observation-encoder hook: the CUDA kernel behind jb_encode_obs against the numpy restatement
of Jitterbug.encode_obs (jitterbug_dmc/jitterbug.py:927-993) in oracle/oracle.py.  fp32 network vs fp64 oracle: 2e-6."""
import ctypes as C

import numpy as np
import pytest

from jitterbug_b200 import _native
from oracle import oracle as O

TOL = 2e-6


@pytest.fixture(scope="module")
def lib():
    from jitterbug_b200 import build as jb_build
    jb_build.build()
    return _native.load()


def _desc(obs_dim, k, dims, vae=0, n_models=1):
    d = _native.JbEncoderDesc()
    d.obs_dim, d.k_encode, d.n_layers, d.vae, d.n_models = obs_dim, k, len(dims) - 1, vae, n_models
    for i, v in enumerate(dims):
        d.dims[i] = v
    return d


def test_descriptor_arithmetic(lib):
    d = _desc(16, 16, [16, 12])
    assert lib.jb_encoder_num_params(C.byref(d)) == 17 * 12
    assert lib.jb_encoder_out_dim(C.byref(d)) == 12
    d = _desc(19, 15, [15, 12])
    assert lib.jb_encoder_out_dim(C.byref(d)) == 12 + 4
    d = _desc(16, 16, [16, 4, 1, 2], vae=1)
    assert lib.jb_encoder_num_params(C.byref(d)) == 17 * 4 + 5 * 1 + 2 * 4
    d.n_layers = 9
    assert lib.jb_encoder_num_params(C.byref(d)) == -1


def test_oracle_encode_obs_modes():
    rng = np.random.RandomState(0)
    obs = rng.uniform(-1, 1, 16)
    ex = np.stack([-1 - rng.rand(16), 1 + rng.rand(16)], 1)
    w, b = rng.randn(15, 12), rng.randn(12)
    out = O.encode_obs(obs, lambda x: O.dae_encode(x, w, b), k_encode=15, extrema=ex)
    n = O.normalize_obs(obs, ex)
    assert out.shape == (13,) and out[12] == n[15]
    np.testing.assert_allclose(out[:12], np.tanh(n[:15] @ w + b))
    assert np.all(np.abs(n) <= 1) and np.all((O.normalize_obs(obs, ex, True) >= 0))
    two = O.encode_obs(obs, [lambda x: O.dae_encode(x, w, b), lambda x: O.dae_encode(x, -w, b)], k_encode=15)
    np.testing.assert_allclose(two[:12], 0.5 * (np.tanh(obs[:15] @ w + b) + np.tanh(-(obs[:15] @ w) + b)))
    assert two[12] == obs[15]


@pytest.mark.gpu
@pytest.mark.parametrize("k,normalize", [(16, None), (15, "pm1"), (13, "01")])
def test_dae_encoder_matches_oracle(k, normalize):
    import torch
    from jitterbug_b200.encoder import ObservationEncoder
    rng = np.random.RandomState(1)
    n, D = 5000, 16
    obs = rng.uniform(-1, 1, (n, 19)).astype(np.float32)
    ex = np.stack([-1 - rng.rand(D), 1 + rng.rand(D)], 1).astype(np.float32)
    w, b = rng.randn(k, 12).astype(np.float32), rng.randn(12).astype(np.float32)
    enc = ObservationEncoder(D, [(w, b)], ["tanh"], k_encode=k, extrema=ex if normalize else None, normalize=normalize)
    out = enc(torch.from_numpy(obs).cuda()).cpu().numpy()
    assert out.shape == (n, 12 + D - k)
    ref = np.stack([O.encode_obs(o[:D], lambda x: O.dae_encode(x, w.astype(np.float64), b.astype(np.float64)), k_encode=k,
                                 extrema=ex if normalize else None, zero_one=normalize == "01") for o in obs])
    assert np.abs(out - ref).max() < TOL * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
def test_vae_encoder_mean_and_noise():
    import torch
    from jitterbug_b200.encoder import ObservationEncoder
    rng = np.random.RandomState(2)
    n, D, L = 200000, 16, 3
    mk = lambda i, o: (rng.randn(i, o).astype(np.float32) * 0.5, rng.randn(o).astype(np.float32) * 0.1)
    fc1, fc2, om, os_ = mk(D, 4), mk(4, 2), mk(2, L), mk(2, L)
    obs = np.tile(rng.uniform(-1, 1, (1, D)).astype(np.float32), (n, 1))
    obs_t = torch.from_numpy(obs).cuda()
    zero = (np.zeros_like(os_[0]), np.zeros_like(os_[1]))
    enc0 = ObservationEncoder(D, [fc1, fc2, om + zero], ["relu", "relu", "linear"], vae=True)
    f64 = lambda l: (l[0].astype(np.float64), l[1].astype(np.float64))
    mean = O.vae_encode(obs[0], f64(fc1), f64(fc2), f64(om), f64(os_), 0.0)
    out0 = enc0(obs_t).cpu().numpy()
    assert np.abs(out0 - mean).max() < TOL * max(1.0, np.abs(mean).max())
    enc = ObservationEncoder(D, [fc1, fc2, om + os_], ["relu", "relu", "linear"], vae=True, seed=7)
    z = enc(obs_t).cpu().numpy().astype(np.float64)
    std = O.vae_encode(obs[0], f64(fc1), f64(fc2), f64(om), f64(os_), 1.0) - mean
    eps = (z - mean) / std                                   # identical obs rows: the spread is the noise alone
    assert np.abs(eps.mean(0)).max() < 0.01 and np.abs(eps.std(0) - 1).max() < 0.01
    assert np.abs(np.corrcoef(eps.T) - np.eye(L)).max() < 0.01
    z2 = enc(obs_t).cpu().numpy()                            # new call, new counter -> fresh noise
    assert np.abs(z2 - z).max() > 0.1
    again = ObservationEncoder(D, [fc1, fc2, om + os_], ["relu", "relu", "linear"], vae=True, seed=7)
    assert np.array_equal(again(obs_t).cpu().numpy().astype(np.float64), z)


@pytest.mark.gpu
def test_ensemble_and_vec_env_hook():
    import torch
    from jitterbug_b200 import BatchedJitterbugEnv
    from jitterbug_b200.encoder import ObservationEncoder
    from jitterbug_b200.rollout import VecJitterbugEnv
    rng = np.random.RandomState(3)
    members = [[(rng.randn(15, 12).astype(np.float32), rng.randn(12).astype(np.float32))] for _ in range(3)]
    enc = ObservationEncoder.ensemble(16, members, ["tanh"], k_encode=15)
    env = BatchedJitterbugEnv(task="face_direction", num_envs=256, seed=5)
    vec = VecJitterbugEnv(env, encoder=enc)
    obs = vec.reset()
    assert obs.shape == (256, 13) and vec.obs_dim == 13
    raw = env.reset().cpu().numpy()[:, :16]
    enc_fns = [(lambda x, w=w, b=b: O.dae_encode(x, w.astype(np.float64), b.astype(np.float64))) for ((w, b),) in members]
    ref = np.stack([O.encode_obs(o, enc_fns, k_encode=15) for o in raw])
    got = enc(torch.from_numpy(raw).cuda().contiguous()).cpu().numpy()
    assert np.abs(got - ref).max() < TOL * max(1.0, np.abs(ref).max())
    o2, r, done, _ = vec.step(torch.zeros(256, device="cuda"))
    assert o2.shape == (256, 13) and torch.isfinite(o2).all()


@pytest.mark.gpu
def test_encoder_rejects_bad_arguments():
    import torch
    from jitterbug_b200.encoder import ObservationEncoder
    w, b = np.zeros((16, 12), np.float32), np.zeros(12, np.float32)
    with pytest.raises(ValueError):
        ObservationEncoder(16, [(w, b)], ["tanh"], k_encode=15)           # W rows != k_encode
    with pytest.raises(ValueError):
        ObservationEncoder(16, [(w, b)], ["tanh"], normalize="pm1")       # no extrema
    enc = ObservationEncoder(16, [(w, b)], ["tanh"])
    with pytest.raises(ValueError):
        enc(torch.zeros(4, 16))                                             # host tensor
    with pytest.raises(ValueError):
        enc(torch.zeros(4, 8, device="cuda"))                               # obs narrower than obs_dim
