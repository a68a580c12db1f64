"""GPU (needs two devices, skipped otherwise): one process driving two B200s.  Function attributes and cluster
occupancy are per device, so the launchers keep their one-time state per device; both handles must give the same
bits for the same tokens, and BiCodec.to(other_device) must rebuild the handle there."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_handles_on_two_devices_agree(cfg, state_dict):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from spark_tts_b200 import BiCodec
    from spark_tts_b200.synthetic import synthetic_tokens
    sem, glob = synthetic_tokens(cfg, 3, 140, 91)
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    m0 = BiCodec.from_state_dict(cfg, state_dict, device=d0)
    a = m0.detokenize(sem.to(d0), glob.to(d0)).cpu()
    m1 = BiCodec.from_state_dict(cfg, state_dict, device=d1)          # first launches on device 1 come AFTER device 0's
    b = m1.detokenize(sem.to(d1), glob.to(d1)).cpu()
    assert torch.equal(a, b)
    for prec in ("fp32", "bf16"):                                      # interleave the devices
        x0 = m0.detokenize(sem.to(d0), glob.to(d0), precision=prec)
        x1 = m1.detokenize(sem.to(d1), glob.to(d1), precision=prec)
        assert torch.equal(x0.cpu(), x1.cpu())
    m0.to(d1)                                                          # nn.Module-style move: handle rebuilt on cuda:1
    assert m0.device == d1
    assert torch.equal(m0.detokenize(sem.to(d1), glob.to(d1)).cpu(), a)


def test_calls_restore_the_callers_current_device(cfg, state_dict):
    """every C-ABI entry point switches to the handle's device for the call only."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from spark_tts_b200 import BiCodec
    from spark_tts_b200.synthetic import synthetic_tokens
    d1 = torch.device("cuda:1")
    torch.cuda.set_device(0)
    m1 = BiCodec.from_state_dict(cfg, state_dict, device=d1)          # create + finalize on cuda:1
    assert torch.cuda.current_device() == 0
    sem, glob = synthetic_tokens(cfg, 2, 30, 92)
    m1.detokenize(sem.to(d1), glob.to(d1))                             # detokenize + check_tokens
    assert torch.cuda.current_device() == 0
    m1.prenet(sem.to(d1), glob.to(d1))
    m1.profile(True); m1.detokenize(sem.to(d1), glob.to(d1)); m1.profile_read(); m1.profile(False)
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device="cuda").device.index == 0
    del m1                                                             # destroy
    import gc
    gc.collect()
    assert torch.cuda.current_device() == 0
