"""The decoder width is configuration, not a constant (BiCodec/config.yaml `decoder.channels`,
/root/reference sparktts/modules/encoder_decoder/wave_generator.py:56-83): other widths take other kernel
instantiations -- the ResidualUnits outside C in {96, 192, 384} run as k7 conv + 1x1 conv instead of the fused kernel,
the conv kernel gets other BLOCK_N / BK choices, the waveform head other channel counts (C = 32 / 64 / 128) -- and
must meet the same bars against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

FP32_MAX_ABS, FP32_SNR = 1e-3, 60.0
BF16_MAX_ABS, BF16_SNR = 5e-2, 30.0


@pytest.mark.parametrize("dec_channels", [1024, 2048])
def test_other_decoder_widths_match_the_oracle(dec_channels):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import bicodec_oracle as O
    from spark_tts_b200 import BiCodec, BiCodecConfig
    from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

    dev = torch.device("cuda:0")
    cfg = BiCodecConfig(dec_channels=dec_channels)
    sd = synthetic_state_dict(cfg, seed=3)
    m = BiCodec.from_state_dict(cfg, sd, device=dev)
    for batch, frames in ((2, 20), (1, 37)):          # 37 frames: ragged last row tile at every stage
        sem, glob = synthetic_tokens(cfg, batch, frames, seed=7)
        ref = O.detokenize(sd, cfg, sem, glob)
        for prec, max_abs, snr_min in (("fp32", FP32_MAX_ABS, FP32_SNR), ("bf16", BF16_MAX_ABS, BF16_SNR)):
            wav = m.detokenize(sem.to(dev), glob.to(dev), precision=prec).cpu()
            assert wav.shape == ref.shape
            err = (ref - wav).abs().max().item()
            snr = O.snr_db(ref, wav)
            assert err <= max_abs and snr >= snr_min, f"C={dec_channels} {batch}x{frames} {prec}: max_abs={err:.3e} snr={snr:.1f} dB"


def test_unsupported_decoder_width_fails_loudly():
    """A width whose last stage has 32 channels has no tcgen05 tile (N must be a multiple of 64 or 96): the library says
    so when the weights are finalised -- there is no CPU or library fallback to fall into."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spark_tts_b200 import BiCodec, BiCodecConfig
    from spark_tts_b200.synthetic import synthetic_state_dict

    cfg = BiCodecConfig(dec_channels=512)
    with pytest.raises(ValueError, match="no tile width divides"):
        BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, seed=3), device=torch.device("cuda:0"))
