"""First-contact GPU diagnostic: runs each check in its own subprocess with a timeout (a trapped kernel
poisons its CUDA context), prints one line per check.  Usage on the GPU box:
    python tools/gpu_diag.py [group ...]        groups: ops_simt ops_tc taps_simt taps_tc taps_tc_unfused full
    (SPARKCODEC_FP32_TERMS=3 in the environment selects the three-term bf16 split of the fp32 mode)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name: str):
    import torch
    import torch.nn.functional as F
    from oracle import bicodec_oracle as O
    from spark_tts_b200 import BiCodec, BiCodecConfig, ops
    from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

    torch.manual_seed(0)
    dev = torch.device("cuda:0")

    def conv_case(tag, c_in, c_out, k, param, transposed, B, L, impl, prec, act="none", residual=False):
        g = torch.Generator().manual_seed(hash(tag) % 1000)
        if transposed:
            w = torch.randn(c_in, c_out, k, generator=g) / (c_in * k / param) ** 0.5
        else:
            w = torch.randn(c_out, c_in, k, generator=g) / (c_in * k) ** 0.5
        b = torch.randn(c_out, generator=g) * 0.1
        x = torch.randn(B, L, c_in, generator=g)
        alpha = torch.rand(c_out, generator=g) + 0.5
        xt = x.transpose(1, 2).double()
        if transposed:
            ref = F.conv_transpose1d(xt, w.double(), b.double(), stride=param, padding=(k - param) // 2)
        else:
            ref = F.conv1d(xt, w.double(), b.double(), dilation=param, padding=(k - 1) // 2 * param)
        ref = ref.transpose(1, 2)
        res = None
        if residual:
            res = torch.randn(ref.shape, generator=g)
            ref = ref + res.double()
        if act == "snake":
            ref = ref + torch.sin(alpha.double() * ref) ** 2 / (alpha.double() + 1e-9)
        elif act == "gelu":
            ref = F.gelu(ref)
        t0 = time.time()
        y = ops.conv(x.to(dev), w, b, transposed=transposed, param=param, act=act, alpha=alpha,
                     residual=res.to(dev) if res is not None else None, precision=prec, impl=impl)
        torch.cuda.synchronize()
        y = y.cpu()
        snr = O.snr_db(ref, y)
        print(f"  {tag:34s} {impl:4s} {prec} act={act:5s} shape={tuple(y.shape)} snr={snr:6.1f} dB "
              f"maxerr={(y.double() - ref).abs().max().item():.2e} {time.time() - t0:.2f}s", flush=True)

    if name.startswith("ops_"):
        impl = name.split("_")[1]
        for prec in ("fp32", "bf16"):
            conv_case("1x1 96->96 (N96,BK32)", 96, 96, 1, 1, False, 2, 300, impl, prec)
            conv_case("k7 d1 96->96", 96, 96, 7, 1, False, 2, 300, impl, prec, act="snake")
            conv_case("k7 d9 192->192 (N192)", 192, 192, 7, 9, False, 2, 200, impl, prec, residual=True)
            conv_case("k7 d3 768->768 (N256)", 768, 768, 7, 3, False, 1, 260, impl, prec, act="snake")
            conv_case("1x1 384->2048 gelu", 384, 2048, 1, 1, False, 2, 70, impl, prec, act="gelu")
            conv_case("1x1 2048->384 res", 2048, 384, 1, 1, False, 2, 70, impl, prec, residual=True)
            conv_case("convT k16 s8 1536->768", 1536, 768, 16, 8, True, 2, 50, impl, prec, act="snake")
            conv_case("convT k11 s5 768->384", 768, 384, 11, 5, True, 1, 150, impl, prec)
            conv_case("convT k8 s4 384->192", 384, 192, 8, 4, True, 1, 130, impl, prec)
            conv_case("convT k4 s2 192->96", 192, 96, 4, 2, True, 1, 129, impl, prec, act="snake")
            conv_case("k7 d1 1024->1536", 1024, 1536, 7, 1, False, 1, 50, impl, prec)
            conv_case("1x1 128->64 (N64)", 128, 64, 1, 1, False, 1, 5, impl, prec)
        return

    cfg = BiCodecConfig()
    sd = synthetic_state_dict(cfg, 0)
    model = BiCodec.from_state_dict(cfg, sd, device=dev)
    if name.startswith("taps_"):
        impl = name.split("_", 1)[1]          # simt | tc | tc_unfused
        model.set_impl(impl)
        B, T = 2, 40
        sem, glob = synthetic_tokens(cfg, B, T, 77)
        taps = {}
        ref = O.detokenize(sd, cfg, sem, glob, taps)
        names = ["d_vector", "z_q", "prenet.downsample.0.1.norm", "prenet.downsample.0.1.convnext.0",
                 "prenet.downsample.0.1.convnext.1", "prenet.downsample.0", "prenet.downsample.1",
                 "prenet.vocos_backbone.norm", "prenet.vocos_backbone.convnext.0", "prenet.vocos_backbone.convnext.11",
                 "prenet.vocos_backbone", "prenet_plus_d", "decoder.model.0", "decoder.model.1.block.1",
                 "decoder.model.1.block.2", "decoder.model.1.block.4", "decoder.model.2.block.1",
                 "decoder.model.2.block.4", "decoder.model.3.block.1", "decoder.model.3.block.4",
                 "decoder.model.4.block.1", "decoder.model.4.block.3", "decoder.model.4.block.4"]
        for prec in ("fp32", "bf16"):
            for tn in names:
                wav, t = model.detokenize_tap(sem.to(dev), glob.to(dev), tn, precision=prec)
                r = taps[tn]
                if tn == "d_vector":
                    r = r.unsqueeze(1)
                if tn == "prenet.downsample.0":
                    r = r * 3.0   # the following SamplingBlock's x3 is folded into this LayerNorm
                print(f"  {impl} {prec} tap {tn:38s} shape={tuple(t.shape)} snr={O.snr_db(r, t.cpu()):6.1f} dB", flush=True)
            print(f"  {impl} {prec} WAV snr={O.snr_db(ref, wav.cpu()):6.1f} dB maxerr={(ref - wav.cpu()).abs().max().item():.2e}",
                  flush=True)
        return
    if name == "full":
        for (B, T) in [(1, 1), (3, 16), (2, 130), (1, 500)]:
            sem, glob = synthetic_tokens(cfg, B, T, 5 + T)
            ref = O.detokenize(sd, cfg, sem, glob)
            for prec in ("fp32", "bf16"):
                t0 = time.time()
                wav = model.detokenize(sem.to(dev), glob.to(dev), precision=prec)
                torch.cuda.synchronize()
                dt = time.time() - t0
                w = wav.cpu()
                print(f"  full B={B} T={T} {prec}: snr={O.snr_db(ref, w):6.1f} dB maxerr={(ref - w).abs().max().item():.2e} "
                      f"{dt * 1e3:.1f} ms launches={model.launch_count()}", flush=True)
        return
    raise SystemExit(f"unknown group {name}")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    groups = sys.argv[1:] or ["ops_simt", "ops_tc", "taps_simt", "taps_tc", "full"]
    for g in groups:
        print(f"== {g}", flush=True)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], timeout=420,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out, rc = p.stdout, p.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
        print(out[-12000:], flush=True)
        print(f"== {g} rc={rc} {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
