"""Observation-encoder hook on the device.

Drop-in for ``Jitterbug.encode_obs`` (jitterbug_dmc/jitterbug.py:927-993): the flat observation is (optionally)
scaled by recorded per-feature extrema, its first ``k_encode`` features are pushed through a small encoder network and
the remaining features are appended unchanged -- ``k_encode`` = 15 / 13 are the reference's ``use_autoencoder15`` /
``use_autoencoder13`` modes, ``k_encode = obs_dim`` its plain ``use_autoencoder`` mode.  Two encoder families, as in the
reference: the hidden layer of the tanh (denoising) auto-encoder (benchmarks/denoising_autoencoder.py:66-98, ``encode``
:195-205) and the VAE encoder (benchmarks/VAE.py:13-46, ``encode`` :134-146).  ``ObservationEncoder.ensemble`` averages
several encoders (``use_several_autoencoders``, jitterbug.py:983-991).

One fused CUDA launch per batch (csrc/jb_encoder.cu, through the C-ABI ``jb_encode_obs``); there is no torch / CPU
fallback.  Weights come from anywhere the user trained them (a ``torch.nn.Linear`` stack, a TF checkpoint dumped to
numpy): only the arrays cross the boundary.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native
from ._native import JbEncoderDesc

ACT = {"linear": 0, "tanh": 1, "relu": 2}
NORM = {None: 0, "none": 0, "pm1": 1, "01": 2}
MAX_LAYERS, MAX_WIDTH = 4, 64


class ObservationEncoder:
    """``layers``: list of ``(W[in, out], b[out])`` (the TF ``X @ W + b`` orientation; pass ``linear.weight.T`` for a
    torch ``nn.Linear``).  ``activations``: one name per layer.  ``vae=True``: the last entry of ``layers`` is the pair
    of heads ``(W_mean, b_mean, W_std, b_std)`` and the output is ``mean + std * eps``.  ``extrema``: ``[obs_dim, 2]``
    (min, max) with ``normalize="pm1"`` (the reference's ``normalize_obs``: ``-1 + 2 (x - min) / (max - min)``) or
    ``"01"`` (``normalize_obs01``)."""

    def __init__(self, obs_dim: int, layers: Sequence, activations: Sequence[str], k_encode: Optional[int] = None,
                 vae: bool = False, extrema=None, normalize: Optional[str] = None, device="cuda", seed: int = 0,
                 env_id_offset: int = 0, _members: Optional[Sequence[Sequence]] = None):
        members = [layers] if _members is None else list(_members)
        k_encode = obs_dim if k_encode is None else int(k_encode)
        if len(activations) != len(layers):
            raise ValueError("one activation per layer")
        if not 1 <= len(layers) <= MAX_LAYERS:
            raise ValueError(f"1..{MAX_LAYERS} layers supported")
        dims, packed = None, []
        for mem in members:
            d, flat = [k_encode], []
            for li, lay in enumerate(mem):
                arrs = [np.asarray(a, dtype=np.float32) for a in lay]
                if vae and li == len(mem) - 1:
                    wm, bm, ws, bs = arrs
                    w, b = np.concatenate([wm, ws], axis=1), np.concatenate([bm, bs])
                    out = wm.shape[1]
                else:
                    w, b = arrs
                    out = w.shape[1]
                if w.shape[0] != d[-1] or b.shape != (w.shape[1],):
                    raise ValueError(f"layer {li}: expected W[{d[-1]}, out], b[out]; got {w.shape}, {b.shape}")
                if w.shape[1] > MAX_WIDTH:
                    raise ValueError(f"layer {li}: width {w.shape[1]} > {MAX_WIDTH}")
                flat += [w.reshape(-1), b]
                d.append(out)
            if dims is not None and d != dims:
                raise ValueError("ensemble members must have identical shapes")
            dims = d
            packed.append(np.concatenate(flat))
        self.desc = JbEncoderDesc()
        self.desc.obs_dim, self.desc.k_encode, self.desc.n_layers = int(obs_dim), k_encode, len(layers)
        for i, v in enumerate(dims):
            self.desc.dims[i] = v
        for i, a in enumerate(activations):
            self.desc.act[i] = ACT[a]
        self.desc.vae, self.desc.normalize, self.desc.n_models = int(vae), NORM[normalize], len(members)
        self.desc.env_id_offset = int(env_id_offset)
        self._lib = _native.load()
        n_params = self._lib.jb_encoder_num_params(C.byref(self.desc))
        self.out_dim = self._lib.jb_encoder_out_dim(C.byref(self.desc))
        params = np.concatenate(packed)
        if n_params < 0 or params.size != n_params * len(members):
            raise ValueError("encoder description rejected by libjbsim")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ObservationEncoder needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.params = torch.from_numpy(params).to(self.device)
        self.extrema = None
        if self.desc.normalize:
            if extrema is None:
                raise ValueError("normalize needs extrema[obs_dim, 2]")
            ex = np.asarray(extrema, dtype=np.float32).reshape(obs_dim, 2)
            self.extrema = torch.from_numpy(np.ascontiguousarray(ex)).to(self.device)
        self.seed, self.counter = int(seed), 0

    @classmethod
    def ensemble(cls, obs_dim, members, activations, **kw):
        """Mean of several identically shaped encoders (``use_several_autoencoders``)."""
        return cls(obs_dim, members[0], activations, _members=members, **kw)

    @classmethod
    def from_torch_vae(cls, enc, obs_dim, **kw):
        """From a module shaped like the reference's ``VAE.Encoder`` (fc1, fc2, out_mean, out_std; VAE.py:13-46)."""
        t = lambda lin: (lin.weight.detach().cpu().numpy().T, lin.bias.detach().cpu().numpy())
        layers = [t(enc.fc1), t(enc.fc2), t(enc.out_mean) + t(enc.out_std)]
        return cls(obs_dim, layers, ["relu", "relu", "linear"], vae=True, **kw)

    def __call__(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """obs ``[N, >= obs_dim]`` float32 CUDA -> ``[N, out_dim]`` (latent, then the passed-through features)."""
        if obs.device != self.device or obs.dtype != torch.float32 or obs.dim() != 2 or obs.stride(1) != 1:
            raise ValueError("obs must be a float32 [N, D] tensor with unit inner stride on the encoder's device")
        if obs.shape[1] < self.desc.obs_dim:
            raise ValueError(f"obs has {obs.shape[1]} features per env, the encoder reads {self.desc.obs_dim}")
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, self.out_dim), dtype=torch.float32, device=self.device)
        elif out.device != self.device or out.dtype != torch.float32 or out.shape != (n, self.out_dim) or out.stride(1) != 1:
            raise ValueError("out must be a float32 [N, out_dim] tensor on the encoder's device")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self._lib.jb_encode_obs(
            C.byref(self.desc), self.params.data_ptr(), self.extrema.data_ptr() if self.extrema is not None else None,
            obs.data_ptr(), n, obs.stride(0), out.data_ptr(), out.stride(0), self.seed, self.counter, stream))
        self.counter += 1
        return out
