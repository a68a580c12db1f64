"""Synthetic BiCodec checkpoint + token streams (no real checkpoint exists offline).

The state dict uses exactly the key names / shapes of the reference's ``BiCodec/model.safetensors``
for the detokenize path (SURVEY.md §8a "Weights/keys"; weight-normed layers carry
``weight_g``/``weight_v`` as torch.nn.utils.weight_norm stores them, dim=0), so that the same dict
loads into the reference modules (oracle validation, golden generation) and into this package.

Every 1-D parameter (biases, LayerNorm affine, ConvNeXt gamma, Snake alpha, weight_g) is perturbed
away from its trivial 0/1 init so that a kernel which drops one of them cannot pass parity.
Generation is a pure function of ``seed`` (torch CPU generator), so the GPU box regenerates
bit-identical weights without /root/reference.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .config import BiCodecConfig


def _tn(g, shape, std=0.02):
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, std=std, a=-2 * std, b=2 * std, generator=g)
    return w


def _n(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _u(g, shape, lo, hi):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _wn(sd, g, prefix, shape, std=0.02):
    """weight_norm(dim=0) parametrisation: weight = g * v / ||v|| (norm over all dims but 0)."""
    v = _tn(g, shape, std)
    norm = v.flatten(1).norm(dim=1).reshape([shape[0]] + [1] * (len(shape) - 1))
    sd[prefix + ".weight_v"] = v
    sd[prefix + ".weight_g"] = norm * (1.0 + _n(g, norm.shape, 0.1))


def _backbone(sd, g, cfg, prefix: str, layers: int, ada: bool, c_in: int = 0):
    """One VocosBackbone's tensors (vocos.py:293-322); c_in = input channels of the embed conv (default vocos_dim)."""
    D, C, H = cfg.d_model, cfg.vocos_dim, cfg.vocos_intermediate_dim
    sd[prefix + ".embed.weight"] = _tn(g, (C, c_in or C, 7), 0.02)
    sd[prefix + ".embed.bias"] = _n(g, (C,), 0.05)

    def norm(p):
        if ada:
            # reference init is ones_/zeros_ then overwritten by trunc_normal(0.02) (vocos.py:319-322);
            # use a scale around 1 so conditioning matters but activations stay O(1)
            sd[p + ".scale.weight"] = _tn(g, (C, D), 0.02)
            sd[p + ".scale.bias"] = 1.0 + _n(g, (C,), 0.05)
            sd[p + ".shift.weight"] = _tn(g, (C, D), 0.02)
            sd[p + ".shift.bias"] = _n(g, (C,), 0.05)
        else:
            sd[p + ".weight"] = 1.0 + _n(g, (C,), 0.05)
            sd[p + ".bias"] = _n(g, (C,), 0.05)

    norm(prefix + ".norm")
    for i in range(layers):
        p = f"{prefix}.convnext.{i}"
        sd[p + ".gamma"] = (1.0 / layers) * (1.0 + _n(g, (C,), 0.2))
        sd[p + ".dwconv.weight"] = _tn(g, (C, 1, 7), 0.2)
        sd[p + ".dwconv.bias"] = _n(g, (C,), 0.05)
        norm(p + ".norm")
        sd[p + ".pwconv1.weight"] = _tn(g, (H, C), 0.05)
        sd[p + ".pwconv1.bias"] = _n(g, (H,), 0.05)
        sd[p + ".pwconv2.weight"] = _tn(g, (C, H), 0.05)
        sd[p + ".pwconv2.bias"] = _n(g, (C,), 0.05)
    sd[prefix + ".final_layer_norm.weight"] = 1.0 + _n(g, (C,), 0.05)
    sd[prefix + ".final_layer_norm.bias"] = _n(g, (C,), 0.05)


def synthetic_state_dict(cfg: BiCodecConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D, C, H = cfg.d_model, cfg.vocos_dim, cfg.vocos_intermediate_dim

    # quantizer (factorized_vector_quantize.py:59-68)
    sd["quantizer.codebook.weight"] = _n(g, (cfg.codebook_size, cfg.codebook_dim), 1.0)
    _wn(sd, g, "quantizer.out_project", (D, cfg.codebook_dim, 1), std=0.3)
    sd["quantizer.out_project.bias"] = _n(g, (D,), 0.05)

    # speaker encoder detokenize side (speaker_encoder.py:61-69, residual_fsq.py:66-71)
    nl = len(cfg.fsq_levels)
    sd["speaker_encoder.quantizer.project_out.weight"] = _n(g, (cfg.latent_dim, nl), 0.3)
    sd["speaker_encoder.quantizer.project_out.bias"] = _n(g, (cfg.latent_dim,), 0.05)
    sd["speaker_encoder.project.weight"] = _tn(g, (D, cfg.latent_dim * cfg.token_num), 0.02)
    sd["speaker_encoder.project.bias"] = _n(g, (D,), 0.05)

    # prenet (feat_decoder.py:47-76)
    sd["prenet.linear_pre.weight"] = _tn(g, (C, D), 0.02)
    sd["prenet.linear_pre.bias"] = _n(g, (C,), 0.05)

    def backbone(prefix: str, layers: int, ada: bool):
        _backbone(sd, g, cfg, prefix, layers, ada)

    for i in range(len(cfg.sample_ratios)):
        backbone(f"prenet.downsample.{i}.1", cfg.downsample_layers, ada=False)
    backbone("prenet.vocos_backbone", cfg.vocos_num_layers, ada=True)
    sd["prenet.linear.weight"] = _tn(g, (D, C), 0.05)
    sd["prenet.linear.bias"] = _n(g, (D,), 0.05)

    # WaveGenerator (wave_generator.py:56-83, layers.py:24-67)
    ch = cfg.dec_channels
    _wn(sd, g, "decoder.model.0", (ch, D, 7), 1.0 / (7 * D) ** 0.5)
    sd["decoder.model.0.bias"] = _n(g, (ch,), 0.05)
    cin = ch
    for i, (k, s) in enumerate(zip(cfg.kernel_sizes, cfg.rates)):
        cout = ch // 2 ** (i + 1)
        p = f"decoder.model.{i + 1}.block"
        sd[p + ".0.alpha"] = _u(g, (1, cin, 1), 0.5, 2.0)
        # ConvTranspose1d weight is (C_in, C_out, k): weight_norm dim=0 normalises per INPUT channel
        _wn(sd, g, p + ".1", (cin, cout, k), 1.0 / (cin * k / s) ** 0.5)
        sd[p + ".1.bias"] = _n(g, (cout,), 0.05)
        for j in range(3):
            q = f"{p}.{j + 2}.block"
            sd[q + ".0.alpha"] = _u(g, (1, cout, 1), 0.5, 2.0)
            _wn(sd, g, q + ".1", (cout, cout, 7), 1.0 / (7 * cout) ** 0.5)
            sd[q + ".1.bias"] = _n(g, (cout,), 0.05)
            sd[q + ".2.alpha"] = _u(g, (1, cout, 1), 0.5, 2.0)
            _wn(sd, g, q + ".3", (cout, cout, 1), 0.5 / cout ** 0.5)
            sd[q + ".3.bias"] = _n(g, (cout,), 0.05)
        cin = cout
    n = len(cfg.rates)
    sd[f"decoder.model.{n + 1}.alpha"] = _u(g, (1, cin, 1), 0.5, 2.0)
    _wn(sd, g, f"decoder.model.{n + 2}", (1, cin, 7), 0.15 / (7 * cin) ** 0.5)
    sd[f"decoder.model.{n + 2}.bias"] = _n(g, (1,), 0.05)
    return sd


def synthetic_tokens(cfg: BiCodecConfig, batch: int, frames: int, seed: int = 1234
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(semantic (B,T) int64 in [0,codebook_size), global (B,1,token_num) int32 in [0,prod(levels)))
    -- the dtypes of the ONNX ``bicodec_vocoder`` contract (export_sparktts_onnx.py:819-840)."""
    g = torch.Generator().manual_seed(seed)
    n_glob = 1
    for l in cfg.fsq_levels:
        n_glob *= l
    sem = torch.randint(0, cfg.codebook_size, (batch, frames), generator=g, dtype=torch.int64)
    glob = torch.randint(0, n_glob, (batch, 1, cfg.token_num), generator=g, dtype=torch.int64).to(torch.int32)
    return sem, glob


def synthetic_encoder_state_dict(cfg: BiCodecConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Encode-side tensors of the semantic half of ``BiCodec.tokenize`` (bicodec.py:151-169): the feature encoder
    (feat_encoder.py:27-77, same widths as the prenet) and the quantizer's ``in_project``
    (factorized_vector_quantize.py:59-61).  Drawn from its own generator so that ``synthetic_state_dict`` -- and
    every golden made from it -- stays bit-identical; merge the two dicts to get a checkpoint that can do both."""
    g = torch.Generator().manual_seed(seed + 7919)
    sd: Dict[str, torch.Tensor] = {}
    D, C = cfg.d_model, cfg.vocos_dim
    _backbone(sd, g, cfg, "encoder.encoder", cfg.vocos_num_layers, ada=False, c_in=D)
    for i in range(len(cfg.sample_ratios)):
        _backbone(sd, g, cfg, f"encoder.downsample.{i}.1", cfg.downsample_layers, ada=False)
    sd["encoder.project.weight"] = _tn(g, (D, C), 0.05)
    sd["encoder.project.bias"] = _n(g, (D,), 0.05)
    _wn(sd, g, "quantizer.in_project", (cfg.codebook_dim, D, 1), std=0.05)
    sd["quantizer.in_project.bias"] = _n(g, (cfg.codebook_dim,), 0.05)
    return sd


def synthetic_features(cfg: BiCodecConfig, batch: int, frames: int, seed: int = 4321) -> torch.Tensor:
    """(B, T, d_model) fp32 stand-in for the wav2vec2 feature mix ``BiCodec.tokenize`` receives, with the
    frame-to-frame correlation a real feature sequence has (so the 7-tap convolutions see structure)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, frames, cfg.d_model), generator=g)
    if frames > 1:
        x[:, 1:] = 0.6 * x[:, 1:] + 0.4 * x[:, :-1]
    return x.contiguous()


def synthetic_speaker_state_dict(cfg: BiCodecConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Encode-side tensors of the speaker half of ``BiCodec.tokenize`` (bicodec.py:162-167): the ECAPA-TDNN trunk up to
    its ``latent`` output (ecapa_tdnn.py:152-214; the pooling / x-vector head is computed and thrown away by
    ``SpeakerEncoder.tokenize``, so it is not part of this dict), the perceiver resampler (perceiver_encoder.py:297-350)
    and the FSQ ``project_in`` (residual_fsq.py:66-68).  Own generator, like ``synthetic_encoder_state_dict``.
    BatchNorm layers carry non-trivial running statistics; every bias / affine is perturbed."""
    g = torch.Generator().manual_seed(seed + 104729)
    sd: Dict[str, torch.Tensor] = {}
    ch, width, F_in = 512, 64, cfg.num_mels

    def bn(p, c):
        sd[p + ".weight"] = 1.0 + _n(g, (c,), 0.1)
        sd[p + ".bias"] = _n(g, (c,), 0.1)
        sd[p + ".running_mean"] = _n(g, (c,), 0.1) + 0.3
        sd[p + ".running_var"] = _u(g, (c,), 0.5, 1.5)
        sd[p + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.int64)

    def conv(p, cout, cin, k, gain=1.4):
        sd[p + ".weight"] = _n(g, (cout, cin, k), gain / (cin * k) ** 0.5)
        sd[p + ".bias"] = _n(g, (cout,), 0.1)

    e = "speaker_encoder.speaker_encoder"
    conv(e + ".layer1.conv", ch, F_in, 5, gain=8.0)          # mel magnitudes of a 0.1-rms signal are ~0.1
    bn(e + ".layer1.bn", ch)
    for name in ("layer2", "layer3", "layer4"):
        p = f"{e}.{name}.se_res2block"
        conv(p + ".0.conv", ch, ch, 1)
        bn(p + ".0.bn", ch)
        for i in range(7):
            conv(f"{p}.1.convs.{i}", width, width, 3)
            bn(f"{p}.1.bns.{i}", width)
        conv(p + ".2.conv", ch, ch, 1)
        bn(p + ".2.bn", ch)
        sd[p + ".3.linear1.weight"] = _n(g, (128, ch), 1.0 / ch ** 0.5)
        sd[p + ".3.linear1.bias"] = _n(g, (128,), 0.1)
        sd[p + ".3.linear2.weight"] = _n(g, (ch, 128), 1.0 / 128 ** 0.5)
        sd[p + ".3.linear2.bias"] = _n(g, (ch,), 0.1)
    conv(e + ".conv", 3 * ch, 3 * ch, 1)

    s = "speaker_encoder.perceiver_sampler"
    L, dim, inner = cfg.token_num, cfg.latent_dim, 512
    sd[s + ".latents"] = _n(g, (L, dim), 0.5)
    sd[s + ".proj_context.weight"] = _n(g, (dim, 3 * ch), 1.0 / (3 * ch) ** 0.5)
    sd[s + ".proj_context.bias"] = _n(g, (dim,), 0.1)
    ff_inner = int(dim * 4 * 2 / 3)
    for i in range(2):
        a = f"{s}.layers.{i}.0"
        sd[a + ".to_q.weight"] = _n(g, (inner, dim), 2.0 / dim ** 0.5)
        sd[a + ".to_kv.weight"] = _n(g, (2 * inner, dim), 2.0 / dim ** 0.5)
        sd[a + ".to_out.weight"] = _n(g, (dim, inner), 1.0 / inner ** 0.5)
        f = f"{s}.layers.{i}.1"
        sd[f + ".0.weight"] = _n(g, (2 * ff_inner, dim), 1.0 / dim ** 0.5)
        sd[f + ".0.bias"] = _n(g, (2 * ff_inner,), 0.1)
        sd[f + ".2.weight"] = _n(g, (dim, ff_inner), 1.0 / ff_inner ** 0.5)
        sd[f + ".2.bias"] = _n(g, (dim,), 0.1)
    sd[s + ".norm.gamma"] = 1.0 + _n(g, (dim,), 0.1)
    nl = len(cfg.fsq_levels)
    sd["speaker_encoder.quantizer.project_in.weight"] = _n(g, (nl, dim), 1.2 / dim ** 0.5)
    sd["speaker_encoder.quantizer.project_in.bias"] = _n(g, (nl,), 0.1)
    return sd


def synthetic_ref_wav(cfg: BiCodecConfig, batch: int, seconds: float = 6.0, seed: int = 777) -> torch.Tensor:
    """(B, n) fp32 stand-in for the reference clip ``BiCodec.tokenize`` receives: a few drifting harmonics plus noise
    per utterance (so that the mel frames differ along time and between utterances), rms ~ 0.1, n = the
    ``get_ref_clip`` length (a multiple of latent_hop_length)."""
    import math
    g = torch.Generator().manual_seed(seed)
    n = int(cfg.sample_rate * seconds) // cfg.latent_hop_length * cfg.latent_hop_length
    t = torch.arange(n, dtype=torch.float32) / cfg.sample_rate
    out = torch.zeros((batch, n))
    for b in range(batch):
        f0 = 90.0 + 160.0 * float(torch.rand((), generator=g))
        vib = 1.0 + 0.03 * torch.sin(2 * math.pi * (3.0 + 2.0 * float(torch.rand((), generator=g))) * t)
        phase = 2 * math.pi * torch.cumsum(f0 * vib, 0) / cfg.sample_rate
        for k in range(1, 9):
            out[b] += (0.6 ** k) * float(torch.rand((), generator=g) + 0.3) * torch.sin(k * phase)
        env = 0.55 + 0.45 * torch.sin(2 * math.pi * (0.7 + float(torch.rand((), generator=g))) * t)
        out[b] = out[b] * env + 0.05 * torch.randn(n, generator=g)
        out[b] *= 0.1 / out[b].pow(2).mean().sqrt()
    return out.contiguous()
