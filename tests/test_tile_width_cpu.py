"""Host logic of the conv kernel's N tile choice (tc_gemm.cu launch_block_n, through the no-GPU test hook
sparkcodec_tile_width): the packed width is the largest of 256/192/128/96/64 that divides a polyphase branch; launches
that would keep at most half of the SMs busy narrow it so the layer's weight stream is spread over more SMs; the
batched configs never narrow."""
import pytest

from spark_tts_b200 import _lib

SMS = 148


def width(n_total, cols, m_tiles, sms=SMS):
    return _lib.load().sparkcodec_tile_width(n_total, cols, m_tiles, sms)


@pytest.mark.parametrize("n_total,cols,packed", [
    (2048, 2048, 256),    # pw1
    (384, 384, 192),      # pw2 / embed convs
    (768, 768, 256),      # C = 768 stage
    (6144, 768, 256),     # 1536 -> 768 up-sampler, 8 polyphase branches
    (1920, 384, 192),     # 768 -> 384, 5 branches
    (768, 192, 192),      # 384 -> 192, 4 branches
    (192, 96, 96),        # 192 -> 96, 2 branches: a tile never straddles two branches
    (1024, 1024, 256),
])
def test_batched_launches_keep_the_packed_width(n_total, cols, packed):
    # config 2: 64 utterances x 500 frames = 250 row tiles at the frame rate, more further down
    assert width(n_total, cols, 250) == packed
    assert width(n_total, cols, 64 * 1250) == packed


def test_few_tile_launches_narrow_the_tile():
    # one 10 s utterance: 4 row tiles at the frame rate
    assert width(2048, 2048, 4) == 64        # pw1: 32 -> 128 tiles
    assert width(384, 384, 4) == 64          # pw2: 8 -> 24 tiles
    assert width(1536, 1536, 4) == 64        # conv-in: 24 -> 96 tiles
    assert width(1024, 1024, 4) == 64
    # narrowing stops as soon as more than half of the SMs are busy
    assert width(768, 768, 32) == 256        # 96 tiles already
    assert width(768, 768, 16) == 128        # 48 tiles -> 96
    assert width(768, 768, 8) == 64          # 24 -> 48 -> 64 (96 wide) -> 96 tiles
    assert width(6144, 768, 4) == 256        # 96 tiles already
    # a 96-column branch has nothing narrower that divides it
    assert width(192, 96, 1) == 96
    assert width(96, 96, 1) == 96


def test_every_choice_divides_the_branch_and_is_a_kernel_instantiation():
    for cols in (64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048):
        for phases in (1, 2, 4, 5, 8):
            for m in (1, 2, 3, 4, 7, 16, 40, 157, 2000):
                w = width(cols * phases, cols, m)
                assert w in (64, 96, 128, 192, 256) and cols % w == 0
                # never fewer tiles than the packed width gives
                assert w <= width(cols * phases, cols, 10 ** 6)


def test_bad_shapes_are_refused():
    assert width(32, 32, 4) < 0              # no tile width divides 32 columns
    assert b"no tile width divides" in _lib.load().sparkcodec_last_error()
    assert width(100, 64, 4) < 0             # n_total is not a whole number of branches
    assert width(0, 64, 4) < 0
