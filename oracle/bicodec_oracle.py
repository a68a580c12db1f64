"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference's
BiCodec detokenize path, written as plain functional torch over a checkpoint state dict.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product path (``spark-tts_b200/``) never does.

Parity pinning: the reference holds no golden vectors for this path.  This
restatement is pinned instead against the reference's own modules imported from /root/reference
(``oracle/validate_against_reference.py``: outputs agree to the fp32 round-off of re-associated
sums) and against fixtures generated from those modules (``tests/golden/make_golden.py`` ->
``tests/golden/*.npz``).  All arithmetic bottoms out in ATen (oneDNN on CPU), as the reference's does.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _wn_weight(sd, prefix: str) -> torch.Tensor:
    """Fold weight_norm(dim=0): w = v * (g / ||v||), norm over all dims but 0.
    sparktts/models/bicodec.py:213-221 (remove_weight_norm), sparktts/modules/blocks/layers.py:24-29."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return torch._weight_norm(v, g, 0)


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """sparktts/modules/blocks/layers.py:32-39: x + (alpha+1e-9)^-1 * sin(alpha*x)^2, alpha (1,C,1)."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def fsq_codes(indices: torch.Tensor, levels) -> torch.Tensor:
    """sparktts/modules/fsq/finite_scalar_quantization.py:143-162:
    level_j = (idx // basis_j) % levels_j ; code_j = (level_j - levels_j//2) / (levels_j//2)."""
    lv = torch.tensor(levels, dtype=torch.int32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32)
    level_idx = (indices.unsqueeze(-1) // basis) % lv
    half = lv // 2
    return (level_idx - half) / half  # int / int -> float32 true division


def vq_detokenize(sd, semantic: torch.Tensor) -> torch.Tensor:
    """sparktts/modules/vq/factorized_vector_quantize.py:154-167: codebook gather, transpose,
    out_project (weight-normed 1x1 conv).  semantic (B,T) int -> (B, d_model, T)."""
    e = F.embedding(semantic.long(), sd["quantizer.codebook.weight"]).transpose(1, 2)
    return F.conv1d(e, _wn_weight(sd, "quantizer.out_project"), sd["quantizer.out_project.bias"])


def speaker_detokenize(sd, global_tokens: torch.Tensor, levels) -> torch.Tensor:
    """sparktts/modules/speaker/speaker_encoder.py:107-112 + residual_fsq.py:112-199 (num_quantizers=1,
    scales == 1): global (B,1,N) int -> d_vector (B, d_model)."""
    idx = global_tokens.long().transpose(1, 2).squeeze(-1)          # (B, N)
    codes = fsq_codes(idx, levels)                                   # (B, N, 6)
    z = F.linear(codes.to(sd["speaker_encoder.quantizer.project_out.weight"].dtype),
                 sd["speaker_encoder.quantizer.project_out.weight"],
                 sd["speaker_encoder.quantizer.project_out.bias"])   # (B, N, latent)
    z = z.transpose(1, 2)                                            # (B, latent, N)
    x = z.reshape(z.shape[0], -1)                                    # flat index = c*N + n
    return F.linear(x, sd["speaker_encoder.project.weight"], sd["speaker_encoder.project.bias"])


def _norm(sd, prefix, x, cond):
    """nn.LayerNorm(eps 1e-6) or AdaLayerNorm (sparktts/modules/blocks/vocos.py:87-110)."""
    C = x.shape[-1]
    if prefix + ".scale.weight" in sd:
        scale = F.linear(cond, sd[prefix + ".scale.weight"], sd[prefix + ".scale.bias"])
        shift = F.linear(cond, sd[prefix + ".shift.weight"], sd[prefix + ".shift.bias"])
        x = F.layer_norm(x, (C,), eps=1e-6)
        return x * scale.unsqueeze(1) + shift.unsqueeze(1)
    return F.layer_norm(x, (C,), sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-6)


def vocos_backbone(sd, prefix, x, cond, taps: Optional[dict] = None):
    """sparktts/modules/blocks/vocos.py:324-335 (VocosBackbone.forward) and :65-84 (ConvNeXtBlock.forward).
    x (B,C,T) -> (B,T,C)."""
    x = F.conv1d(x, sd[prefix + ".embed.weight"], sd[prefix + ".embed.bias"], padding=3)
    x = _norm(sd, prefix + ".norm", x.transpose(1, 2), cond).transpose(1, 2)
    if taps is not None:
        taps[prefix + ".norm"] = x.transpose(1, 2)
    i = 0
    while f"{prefix}.convnext.{i}.gamma" in sd:
        p = f"{prefix}.convnext.{i}"
        C = x.shape[1]
        r = x
        y = F.conv1d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=3, groups=C)
        y = _norm(sd, p + ".norm", y.transpose(1, 2), cond)
        y = F.linear(y, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
        y = F.gelu(y)
        y = F.linear(y, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
        y = sd[p + ".gamma"] * y
        x = r + y.transpose(1, 2)
        if taps is not None:
            taps[p] = x.transpose(1, 2)
        i += 1
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[prefix + ".final_layer_norm.weight"],
                        sd[prefix + ".final_layer_norm.bias"], eps=1e-6)


def prenet(sd, z_q, d_vector, n_downsample: int = 2, taps: Optional[dict] = None):
    """sparktts/modules/encoder_decoder/feat_decoder.py:78-94 with sample_ratios [1,1]; the
    SamplingBlock with both ratios 1 returns conv_res + skip1_res + skip2_res = 3*x
    (sparktts/modules/blocks/samper.py:79-100).  z_q (B,D,T), d (B,D) -> (B,D,T)."""
    x = F.linear(z_q.transpose(1, 2), sd["prenet.linear_pre.weight"], sd["prenet.linear_pre.bias"])
    for i in range(n_downsample):
        xt = x.transpose(1, 2)
        xt = xt + xt + xt                     # SamplingBlock(ratio 1): three aliases of x summed
        x = vocos_backbone(sd, f"prenet.downsample.{i}.1", xt, None, taps)
        if taps is not None:
            taps[f"prenet.downsample.{i}"] = x
    x = vocos_backbone(sd, "prenet.vocos_backbone", x.transpose(1, 2), d_vector, taps)
    if taps is not None:
        taps["prenet.vocos_backbone"] = x
    x = F.linear(x, sd["prenet.linear.weight"], sd["prenet.linear.bias"]).transpose(1, 2)
    return x


def wave_generator(sd, x, rates, kernel_sizes, taps: Optional[dict] = None):
    """sparktts/modules/encoder_decoder/wave_generator.py:29-88; ResidualUnit layers.py:51-67.
    x (B,D,T) -> (B,1,prod(rates)*T)."""
    x = F.conv1d(x, _wn_weight(sd, "decoder.model.0"), sd["decoder.model.0.bias"], padding=3)
    if taps is not None:
        taps["decoder.model.0"] = x.transpose(1, 2)
    for i, (k, s) in enumerate(zip(kernel_sizes, rates)):
        p = f"decoder.model.{i + 1}.block"
        x = snake(x, sd[p + ".0.alpha"])
        x = F.conv_transpose1d(x, _wn_weight(sd, p + ".1"), sd[p + ".1.bias"], stride=s,
                               padding=(k - s) // 2)
        if taps is not None:
            taps[p + ".1"] = x.transpose(1, 2)
        for j, dil in enumerate((1, 3, 9)):
            q = f"{p}.{j + 2}.block"
            y = snake(x, sd[q + ".0.alpha"])
            y = F.conv1d(y, _wn_weight(sd, q + ".1"), sd[q + ".1.bias"], dilation=dil, padding=3 * dil)
            y = snake(y, sd[q + ".2.alpha"])
            y = F.conv1d(y, _wn_weight(sd, q + ".3"), sd[q + ".3.bias"])
            x = x + y
            if taps is not None:
                taps[f"{p}.{j + 2}"] = x.transpose(1, 2)
    n = len(rates)
    x = snake(x, sd[f"decoder.model.{n + 1}.alpha"])
    x = F.conv1d(x, _wn_weight(sd, f"decoder.model.{n + 2}"), sd[f"decoder.model.{n + 2}.bias"], padding=3)
    return torch.tanh(x)


@torch.no_grad()
def detokenize(sd: Dict[str, torch.Tensor], cfg, semantic_tokens: torch.Tensor,
               global_tokens: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """sparktts/models/bicodec.py:171-189 (BiCodec.detokenize).
    semantic (B,T), global (B,1,N) -> waveform (B,1,hop*T) float32."""
    z_q = vq_detokenize(sd, semantic_tokens)
    d = speaker_detokenize(sd, global_tokens, cfg.fsq_levels)
    if taps is not None:
        taps["z_q"] = z_q.transpose(1, 2)
        taps["d_vector"] = d
    x = prenet(sd, z_q, d, len(cfg.sample_ratios), taps)
    x = x + d.unsqueeze(-1)
    if taps is not None:
        taps["prenet_plus_d"] = x.transpose(1, 2)
    return wave_generator(sd, x, cfg.rates, cfg.kernel_sizes, taps)


@torch.no_grad()
def tokenizer_detokenize(sd, cfg, global_tokens: torch.Tensor, semantic_tokens: torch.Tensor):
    """sparktts/models/audio_tokenizer.py:132-146 (BiCodecTokenizer.detokenize): global (B,N),
    semantic (B,T) -> numpy float32 (B, hop*T), squeezed to (hop*T,) when B == 1."""
    wav = detokenize(sd, cfg, semantic_tokens, global_tokens.unsqueeze(1))
    return wav.detach().squeeze().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# Encode side, semantic half: BiCodec.tokenize's `quantizer.tokenize(encoder(feat^T))`
# ---------------------------------------------------------------------------------------------
def feat_encoder(sd, feat: torch.Tensor, n_downsample: int = 2) -> torch.Tensor:
    """sparktts/modules/encoder_decoder/feat_encoder.py:79-90 (Encoder.forward) with sample_ratios [1,1]:
    VocosBackbone(D -> C, 12 layers) -> n x [SamplingBlock(ratio 1) = 3*x, VocosBackbone(2 layers)] -> Linear C -> D.
    feat (B,T,D) -> z (B,D,T)."""
    x = vocos_backbone(sd, "encoder.encoder", feat.transpose(1, 2), None)
    for i in range(n_downsample):
        xt = x.transpose(1, 2)
        xt = xt + xt + xt                     # samper.py:79-100, both ratios 1
        x = vocos_backbone(sd, f"encoder.downsample.{i}.1", xt, None)
    return F.linear(x, sd["encoder.project.weight"], sd["encoder.project.bias"]).transpose(1, 2)


def vq_tokenize(sd, z: torch.Tensor):
    """sparktts/modules/vq/factorized_vector_quantize.py:147-152 (tokenize) and :169-187 (decode_latents):
    in_project (weight-normed 1x1 conv D -> 8), L2-normalise encodings and codebook, squared distance,
    indices = (-dist).max(1)[1].  z (B,D,T) -> (indices (B,T) int64, margin (B,T) fp32), margin = second-best
    minus best distance (how far an index is from flipping under rounding noise)."""
    z_e = F.conv1d(z, _wn_weight(sd, "quantizer.in_project"), sd["quantizer.in_project.bias"])
    B = z_e.shape[0]
    enc = F.normalize(z_e.transpose(1, 2).reshape(-1, z_e.shape[1]))
    cb = F.normalize(sd["quantizer.codebook.weight"])
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
    idx = (-dist).max(1)[1]
    two = dist.topk(2, dim=1, largest=False).values
    return idx.reshape(B, -1), (two[:, 1] - two[:, 0]).reshape(B, -1)


@torch.no_grad()
def tokenize_semantic(sd, cfg, feat: torch.Tensor):
    """The semantic half of sparktts/models/bicodec.py:151-169 (BiCodec.tokenize): feat (B,T,D) -> (indices, margin)."""
    return vq_tokenize(sd, feat_encoder(sd, feat, len(cfg.sample_ratios)))


# ---------------------------------------------------------------------------------------------
# Encode side, speaker half: BiCodec.tokenize's
# `speaker_encoder.tokenize(mel_transformer(ref_wav).squeeze(1).transpose(1, 2))`
# ---------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    """torchaudio.functional._hz_to_mel(mel_scale="slaney")."""
    import math
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney"): (n_freqs, n_mels) triangular filters,
    area-normalised.  (The reference builds it with torchaudio.transforms.MelSpectrogram, bicodec.py:191-211.)"""
    import math
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    f_pts = f_sp * m_pts
    log_t = m_pts >= min_log_mel
    f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def mel_spectrogram(wav: torch.Tensor, cfg) -> torch.Tensor:
    """TT.MelSpectrogram(sample_rate, n_fft, win_length, hop_length, f_min, f_max, n_mels, power=1, norm="slaney",
    mel_scale="slaney") as BiCodec.init_mel_transformer builds it (bicodec.py:191-211): centred reflect-padded STFT
    with a periodic Hann window of win_length (zero-padded to n_fft by torch.stft), magnitude, mel filterbank.
    wav (B, n) -> (B, n_mels, 1 + n // hop)."""
    win = torch.hann_window(cfg.mel_win_length, periodic=True, dtype=torch.float32)
    spec = torch.stft(wav, cfg.mel_n_fft, hop_length=cfg.mel_hop_length, win_length=cfg.mel_win_length, window=win,
                      center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True).abs()
    f_max = cfg.mel_fmax if cfg.mel_fmax is not None else float(cfg.sample_rate // 2)
    fb = mel_filterbank(cfg.mel_n_fft // 2 + 1, cfg.mel_fmin, f_max, cfg.num_mels, cfg.sample_rate)
    return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)


def _bn(sd, p, x):
    """nn.BatchNorm1d in eval mode on (B, C, T)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _conv_relu_bn(sd, p, x, padding=0, dilation=1):
    """Conv1dReluBn (ecapa_tdnn.py:89-113): bn(relu(conv(x)))."""
    return _bn(sd, p + ".bn", F.relu(F.conv1d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=padding,
                                             dilation=dilation)))


def _se_res2block(sd, p, x, dilation: int, scale: int = 8):
    """SE_Res2Block (ecapa_tdnn.py:135-149) = x + SE(Conv1dReluBn(Res2Conv1dReluBn(Conv1dReluBn(x))))."""
    y = _conv_relu_bn(sd, p + ".se_res2block.0", x)
    width = y.shape[1] // scale
    spx = torch.split(y, width, 1)
    out, sp = [], spx[0]
    for i in range(scale - 1):                                    # Res2Conv1dReluBn (ecapa_tdnn.py:28-83)
        if i >= 1:
            sp = sp + spx[i]
        q = f"{p}.se_res2block.1"
        sp = F.conv1d(sp, sd[f"{q}.convs.{i}.weight"], sd[f"{q}.convs.{i}.bias"], padding=dilation, dilation=dilation)
        sp = _bn(sd, f"{q}.bns.{i}", F.relu(sp))
        out.append(sp)
    out.append(spx[scale - 1])
    y = torch.cat(out, dim=1)
    y = _conv_relu_bn(sd, p + ".se_res2block.2", y)
    q = p + ".se_res2block.3"                                     # SE_Connect (ecapa_tdnn.py:119-132)
    s = y.mean(dim=2)
    s = F.relu(F.linear(s, sd[q + ".linear1.weight"], sd[q + ".linear1.bias"]))
    s = torch.sigmoid(F.linear(s, sd[q + ".linear2.weight"], sd[q + ".linear2.bias"]))
    return x + y * s.unsqueeze(2)


def ecapa_latent(sd, mels: torch.Tensor) -> torch.Tensor:
    """The `latent` output of ECAPA_TDNN.forward(x, return_latent=True) (ecapa_tdnn.py:196-214), which is all that
    SpeakerEncoder.tokenize keeps (speaker_encoder.py:100-105).  mels (B, T, F) -> (B, 1536, T)."""
    p = "speaker_encoder.speaker_encoder"
    x = mels.permute(0, 2, 1)
    out1 = _conv_relu_bn(sd, p + ".layer1", x, padding=2)
    out2 = _se_res2block(sd, p + ".layer2", out1, 2)
    out3 = _se_res2block(sd, p + ".layer3", out2, 3)
    out4 = _se_res2block(sd, p + ".layer4", out3, 4)
    out = torch.cat([out2, out3, out4], dim=1)
    return F.relu(F.conv1d(out, sd[p + ".conv.weight"], sd[p + ".conv.bias"]))


def perceiver_resampler(sd, x: torch.Tensor, heads: int = 8) -> torch.Tensor:
    """PerceiverResampler.forward (perceiver_encoder.py:297-350): proj_context, then depth x [cross attention of the
    latents over cat(latents, context) (no pre-norm, cross_attn_include_queries) + GEGLU feed-forward], RMSNorm.
    x (B, T, dim_context) -> (B, num_latents, dim)."""
    p = "speaker_encoder.perceiver_sampler"
    x = F.linear(x, sd[p + ".proj_context.weight"], sd[p + ".proj_context.bias"])
    B = x.shape[0]
    lat = sd[p + ".latents"].unsqueeze(0).expand(B, -1, -1)
    i = 0
    while f"{p}.layers.{i}.0.to_q.weight" in sd:
        a = f"{p}.layers.{i}.0"
        ctx = torch.cat((lat, x), dim=-2)
        q = F.linear(lat, sd[a + ".to_q.weight"])
        k, v = F.linear(ctx, sd[a + ".to_kv.weight"]).chunk(2, dim=-1)
        sp = lambda t: t.reshape(B, t.shape[1], heads, -1).transpose(1, 2)          # b n (h d) -> b h n d
        q, k, v = sp(q), sp(k), sp(v)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * q.shape[-1] ** -0.5
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        o = o.transpose(1, 2).reshape(B, lat.shape[1], -1)
        lat = F.linear(o, sd[a + ".to_out.weight"]) + lat
        f = f"{p}.layers.{i}.1"
        h, gate = F.linear(lat, sd[f + ".0.weight"], sd[f + ".0.bias"]).chunk(2, dim=-1)
        lat = F.linear(F.gelu(gate) * h, sd[f + ".2.weight"], sd[f + ".2.bias"]) + lat
        i += 1
    return F.normalize(lat, dim=-1) * (lat.shape[-1] ** 0.5) * sd[p + ".norm.gamma"]


def fsq_quantize_indices(sd, x: torch.Tensor, levels):
    """ResidualFSQ.forward with one quantizer (residual_fsq.py:213-283) -> FSQ.forward
    (finite_scalar_quantization.py:101-118, 126-141, 190-250): project_in, bound (tanh with the half-level shift of
    even level counts), round, codes_to_indices.  x (B, N, dim) -> (indices (B, N) int32, margin (B, N)): margin =
    distance of the closest pre-rounding coordinate to a rounding boundary (how far an index is from flipping)."""
    p = "speaker_encoder.quantizer"
    z = F.linear(x, sd[p + ".project_in.weight"], sd[p + ".project_in.bias"])
    lv = torch.tensor(levels, dtype=torch.int32)
    half_l = (lv - 1) * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    bounded = (z + shift).tanh() * half_l - offset
    q = bounded.round()
    half_width = lv // 2
    codes = q / half_width
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32)
    idx = ((codes * half_width + half_width) * basis).sum(dim=-1).to(torch.int32)
    frac = bounded - torch.floor(bounded)
    margin = (frac - 0.5).abs().min(dim=-1).values
    return idx, margin


@torch.no_grad()
def tokenize_speaker(sd, cfg, ref_wav: torch.Tensor):
    """The speaker half of BiCodec.tokenize (bicodec.py:162-167): ref_wav (B, n) -> (global_tokens (B, 1, N) int32,
    margin (B, N))."""
    mel = mel_spectrogram(ref_wav, cfg)                                        # (B, n_mels, T)
    feats = ecapa_latent(sd, mel.transpose(1, 2))                              # (B, 1536, T)
    lat = perceiver_resampler(sd, feats.transpose(1, 2))                       # (B, N, latent)
    idx, margin = fsq_quantize_indices(sd, lat, cfg.fsq_levels)
    return idx.unsqueeze(1), margin


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    ref = ref.double().flatten()
    err = test.double().flatten() - ref
    return float(10.0 * torch.log10(ref.pow(2).sum() / err.pow(2).sum().clamp_min(1e-300)))
