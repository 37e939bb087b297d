"""Model loading, folding and packing for the B200 engine (host side).

* `PoseModel`   -- the pose-VAE decoder folded to three dense layers
  24 -> 40 -> 60 -> 92 (LeakyReLU 0.2 between), the quaternion / displacement
  statistics, and the folded encoder used once per clip for the initial latent.
  Reference behaviour: `python/src/autoencoder.py:146-256` (decoder),
  `:56-143` (encoder), `python/src/skeleton.py:117-130,244-245` (masked conv with
  kernel 1 == masked Linear; unpool == constant 0/1 matmul),
  `python/src/drag_pose.py:27-34` (first 4 of every 8 dual-quat statistics).
* `TemporalModel` -- weights of the temporal predictor (`nn.Transformer`, d=48,
  4 heads, 3+3 layers, FF 2048; `python/src/temporal_transformer.py:6-78`,
  `python/src/train_temporal.py:17-37`).

Checkpoint formats are the reference's (`python/src/train.py:285-319`):
`generator.pt` = {"model_state_dict"}, `data.pt` = {"means","stds"} each with
"dqs" (176,) and "displacement" (3,), `temporal.pt` = {"model_state_dict",
"means_latent","stds_latent"}.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import topology

LATENT = 24
JOINTS = 22
N_OUT = 92  # 22 quats + 3 displacement + 1 pad
DEC_DIMS = (24, 40, 60, 92)
ENC_DIMS = (176, 112, 72, 48)
PAST_ROWS = 60  # train_temporal.param["future_frames"][0]
SAMPLE_STEP = 4
HEIGHT_JOINTS = (0, 4, 8, 13, 17, 21)

# skeleton of python/data/example/eval/example.bvh:2-122 (train.py:338 sets parents[0] = 0)
DEFAULT_PARENTS = (0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 11, 18, 19, 20)


@dataclass
class PoseModel:
    A: list  # [A0 (40,24), A1 (60,40), A2 (92,60)] float32
    b: list  # [b0 (40,), b1 (60,), b2 (92,)]
    mean_q: np.ndarray  # (88,)
    std_q: np.ndarray  # (88,)
    mean_d: np.ndarray  # (3,)
    std_d: np.ndarray  # (3,)
    enc_A: list = field(default_factory=list)  # folded encoder [(112,176),(72,112),(48,72)]
    enc_b: list = field(default_factory=list)
    enc_mu: tuple = None  # (W (24,48), b (24,))
    enc_logvar: tuple = None
    parents: tuple = DEFAULT_PARENTS
    mean_dqs: np.ndarray = None  # full (176,) statistics (kept for the encoder side)
    std_dqs: np.ndarray = None

    # ---- host-side numpy evaluation (data generation and init only) -------
    def decode_np(self, z):
        """Folded decoder forward: (B,24) -> standardised y (B,92)."""
        a = np.asarray(z, dtype=np.float32)
        for l in range(3):
            a = a @ self.A[l].T + self.b[l]
            if l < 2:
                a = np.where(a > 0, a, np.float32(0.2) * a)
        return a

    def encode_np(self, dqs_std):
        """Folded encoder forward: standardised dual quats (B,176) -> mu, logvar (B,24)."""
        a = np.asarray(dqs_std, dtype=np.float32)
        for l in range(3):
            a = a @ self.enc_A[l].T + self.enc_b[l]
            a = np.where(a > 0, a, np.float32(0.2) * a)
        mu = a @ self.enc_mu[0].T + self.enc_mu[1]
        logvar = a @ self.enc_logvar[0].T + self.enc_logvar[1]
        return mu, logvar


def _np(t):
    return t.detach().cpu().double().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float64)


def fold_generator_state(sd, means, stds, parents=DEFAULT_PARENTS) -> PoseModel:
    """Fold a reference `Generator_Model.state_dict()` (keys in SURVEY appendix A)."""
    pre = "autoencoder.decoder."
    Wf, bf = _np(sd[pre + "f_latent.weight"]), _np(sd[pre + "f_latent.bias"])
    A, b = [], []
    for l in range(3):
        U = _np(sd[f"{pre}layers.{l}.0.weight"])
        W = _np(sd[f"{pre}layers.{l}.1.weight"])[..., 0] * _np(sd[f"{pre}layers.{l}.1.mask"])[..., 0]
        A.append(W @ U)
        b.append(_np(sd[f"{pre}layers.{l}.1.bias"]))
    b[0] = A[0] @ bf + b[0]
    A[0] = A[0] @ Wf
    pre = "autoencoder.encoder."
    eA, eb = [], []
    for l in range(3):
        W = _np(sd[f"{pre}layers.{l}.0.weight"])[..., 0] * _np(sd[f"{pre}layers.{l}.0.mask"])[..., 0]
        P = _np(sd[f"{pre}layers.{l}.1.weight"])
        eA.append(P @ W)
        eb.append(P @ _np(sd[f"{pre}layers.{l}.0.bias"]))
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    mean_dqs, std_dqs = _np(means["dqs"]), _np(stds["dqs"])
    return PoseModel(
        A=[f32(x) for x in A],
        b=[f32(x) for x in b],
        mean_q=f32(mean_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        std_q=f32(std_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        mean_d=f32(_np(means["displacement"])),
        std_d=f32(_np(stds["displacement"])),
        enc_A=[f32(x) for x in eA],
        enc_b=[f32(x) for x in eb],
        enc_mu=(f32(_np(sd[pre + "f_mu.weight"])), f32(_np(sd[pre + "f_mu.bias"]))),
        enc_logvar=(f32(_np(sd[pre + "f_logvar.weight"])), f32(_np(sd[pre + "f_logvar.bias"]))),
        parents=tuple(int(p) for p in parents),
        mean_dqs=f32(mean_dqs),
        std_dqs=f32(std_dqs),
    )


def load_pose_model(model_dir, parents=DEFAULT_PARENTS) -> PoseModel:
    """Read `generator.pt` + `data.pt` (reference layout) or a folded `.npz`."""
    import torch

    if model_dir.endswith(".npz"):
        return load_folded_npz(model_dir)
    gen = torch.load(os.path.join(model_dir, "generator.pt"), map_location="cpu", weights_only=True)
    data = torch.load(os.path.join(model_dir, "data.pt"), map_location="cpu", weights_only=True)
    return fold_generator_state(gen["model_state_dict"], data["means"], data["stds"], parents)


def save_folded_npz(model: PoseModel, path, offsets=None):
    arrs = {f"A{l}": model.A[l] for l in range(3)}
    arrs.update({f"b{l}": model.b[l] for l in range(3)})
    arrs.update({f"encA{l}": model.enc_A[l] for l in range(3)})
    arrs.update({f"encb{l}": model.enc_b[l] for l in range(3)})
    arrs.update(
        enc_mu_w=model.enc_mu[0], enc_mu_b=model.enc_mu[1],
        enc_lv_w=model.enc_logvar[0], enc_lv_b=model.enc_logvar[1],
        mean_dqs=model.mean_dqs, std_dqs=model.std_dqs,
        mean_d=model.mean_d, std_d=model.std_d,
        parents=np.asarray(model.parents, dtype=np.int32),
    )
    if offsets is not None:
        arrs["offsets"] = np.asarray(offsets, dtype=np.float32)
    np.savez_compressed(path, **arrs)


def load_folded_npz(path) -> PoseModel:
    z = np.load(path)
    mean_dqs, std_dqs = z["mean_dqs"], z["std_dqs"]
    return PoseModel(
        A=[z[f"A{l}"] for l in range(3)],
        b=[z[f"b{l}"] for l in range(3)],
        mean_q=np.ascontiguousarray(mean_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        std_q=np.ascontiguousarray(std_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        mean_d=z["mean_d"], std_d=z["std_d"],
        enc_A=[z[f"encA{l}"] for l in range(3)],
        enc_b=[z[f"encb{l}"] for l in range(3)],
        enc_mu=(z["enc_mu_w"], z["enc_mu_b"]),
        enc_logvar=(z["enc_lv_w"], z["enc_lv_b"]),
        parents=tuple(int(p) for p in z["parents"]),
        mean_dqs=mean_dqs, std_dqs=std_dqs,
    )


def random_pose_model(seed=2222, parents=DEFAULT_PARENTS) -> PoseModel:
    """Random-init weights of the same architecture (kaiming-uniform per joint
    block like `skeleton.py:67-111`, default nn.Linear init for the dense heads);
    statistics are made up but plausible (unit quaternion mean, small stds)."""
    rng = np.random.default_rng(seed)

    def masked(mask):
        fan_in = np.maximum(mask.sum(axis=1, keepdims=True), 1.0)
        bound = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-1, 1, mask.shape) * bound * mask, rng.uniform(-1, 1, mask.shape[0]) * bound[:, 0]

    def linear(n_out, n_in):
        bound = 1.0 / math.sqrt(n_in)
        return rng.uniform(-bound, bound, (n_out, n_in)), rng.uniform(-bound, bound, n_out)

    dec_layers, primal = topology.decoder_plan(parents)
    Wf, bf = linear(primal, LATENT)
    A, b = [], []
    for U, mask in dec_layers:
        W, bias = masked(mask)
        A.append(W @ U)
        b.append(bias)
    b[0] = A[0] @ bf + b[0]
    A[0] = A[0] @ Wf
    enc_layers, primal_e = topology.encoder_plan(parents)
    eA, eb = [], []
    for mask, P in enc_layers:
        W, bias = masked(mask)
        eA.append(P @ W)
        eb.append(P @ bias)
    mu_w, mu_b = linear(LATENT, primal_e)
    lv_w, lv_b = linear(LATENT, primal_e)
    lv_w = np.zeros_like(lv_w)  # autoencoder.py:131-134
    mean_dqs = np.zeros((JOINTS, 8))
    mean_dqs[:, 0] = 0.9
    mean_dqs[:, 1:4] = rng.normal(0, 0.1, (JOINTS, 3))
    std_dqs = np.full((JOINTS, 8), 0.1)
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    mean_dqs, std_dqs = mean_dqs.reshape(-1), std_dqs.reshape(-1)
    return PoseModel(
        A=[f32(x) for x in A], b=[f32(x) for x in b],
        mean_q=f32(mean_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        std_q=f32(std_dqs.reshape(-1, 8)[:, :4].reshape(-1)),
        mean_d=f32(np.zeros(3)), std_d=f32(np.full(3, 6e-3)),
        enc_A=[f32(x) for x in eA], enc_b=[f32(x) for x in eb],
        enc_mu=(f32(mu_w), f32(mu_b)), enc_logvar=(f32(lv_w), f32(lv_b)),
        parents=tuple(int(p) for p in parents), mean_dqs=f32(mean_dqs), std_dqs=f32(std_dqs),
    )


# --------------------------------------------------------------------------
# Temporal predictor
# --------------------------------------------------------------------------
D_MODEL = 48
N_HEADS = 4
D_FF = 2048
N_ENC = 3
N_DEC = 3
ENC_IN = 33  # 24 latent + 3 accumulated displacement + 6 heights
ENC_TOKENS = 14
PE_LEN = 30


@dataclass
class TemporalModel:
    sd: dict  # name -> float32 ndarray, nn.Transformer key names (Temporal.state_dict())
    means_latent: np.ndarray
    stds_latent: np.ndarray
    trained: bool = False  # read from a temporal.pt (as opposed to the seeded random-init stand-in)


def positional_table(max_len=PE_LEN, d=D_MODEL):
    """sin/cos table of `python/src/positional_encoding.py:15-25` (float32 math)."""
    import torch

    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0)) / d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.numpy()


def random_temporal_state(seed=2222):
    """Seeded random-init predictor with the reference's module order
    (`temporal_transformer.py:15-34`): in_proj_encoder, in_proj_decoder,
    nn.Transformer, out_proj.  Used because `temporal.pt` is a missing blob."""
    import torch
    import torch.nn as nn
    import warnings

    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc = nn.Linear(ENC_IN, D_MODEL)
        dec = nn.Linear(LATENT, D_MODEL)
        tr = nn.Transformer(d_model=D_MODEL, nhead=N_HEADS, num_encoder_layers=N_ENC,
                            num_decoder_layers=N_DEC, dim_feedforward=D_FF, dropout=0.1)
        out = nn.Linear(D_MODEL, LATENT)
    torch.random.set_rng_state(g)
    sd = {"positional_encoding.pos_encoding": torch.from_numpy(positional_table())}
    sd.update({"in_proj_encoder." + k: v for k, v in enc.state_dict().items()})
    sd.update({"in_proj_decoder." + k: v for k, v in dec.state_dict().items()})
    sd.update({"temporal." + k: v for k, v in tr.state_dict().items()})
    sd.update({"out_proj." + k: v for k, v in out.state_dict().items()})
    return sd


def temporal_from_state(sd, means_latent=None, stds_latent=None) -> TemporalModel:
    arr = {k: np.ascontiguousarray(_np(v), dtype=np.float32) for k, v in sd.items()}
    ml = np.zeros(LATENT, np.float32) if means_latent is None else np.asarray(_np(means_latent), np.float32).reshape(-1)
    sl = np.ones(LATENT, np.float32) if stds_latent is None else np.asarray(_np(stds_latent), np.float32).reshape(-1)
    return TemporalModel(arr, ml, sl)


def load_temporal_model(model_dir, allow_random=False) -> TemporalModel:
    """Read `temporal.pt` (`train_temporal.py:474-482`, written by `train.py:311-319`: {"model_state_dict", "means_latent",
    "stds_latent"}).  A missing file raises like the reference's `torch.load` does; only callers that pass
    `allow_random=True` (bench, tests, smoke, `--random-temporal` on the CLIs -- the blob is absent from the reference
    checkout and BASELINE.json allows random-init weights) get the seed-2222 random-init predictor with means 0 / stds 1."""
    import torch

    path = os.path.join(model_dir, "temporal.pt")
    if os.path.isfile(path):
        ck = torch.load(path, map_location="cpu", weights_only=True)
        tm = temporal_from_state(ck["model_state_dict"], ck["means_latent"], ck["stds_latent"])
        tm.trained = True
        return tm
    if not allow_random:
        raise FileNotFoundError(f"{path} (pass allow_random=True / --random-temporal to run with a random-init predictor)")
    return temporal_from_state(random_temporal_state())


def save_temporal_model(tm: TemporalModel, path):
    """Writes a predictor in the reference's `temporal.pt` format (train.py:311-319)."""
    import torch

    torch.save({"model_state_dict": {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tm.sd.items()},
                "means_latent": torch.from_numpy(np.asarray(tm.means_latent, np.float32).copy()),
                "stds_latent": torch.from_numpy(np.asarray(tm.stds_latent, np.float32).copy())}, path)
