

def test_colour_closed_forms_equal_the_lifting_chain():
    """The level-0 forward kernel evaluates RGBtoYCoCg<shift> (ric.cpp:76-91) through closed forms that are dot
    products of the pixel bytes plus one mask (ric_fwd.cuh convert_raw).  Exhaustive over all 2^24 pixels."""
    import numpy as np
    r = np.arange(256, dtype=np.int32).reshape(256, 1, 1)
    g = np.arange(256, dtype=np.int32).reshape(1, 256, 1)
    b = np.arange(256, dtype=np.int32).reshape(1, 1, 256)
    co = r - b + 0 * g
    t = b + (co >> 1)
    cg = g - t
    y = t + (cg >> 1) - 128
    assert np.array_equal(co << 3, 8 * r - 8 * b + 0 * g)
    assert np.array_equal(cg << 3, (8 * g - 4 * r - 4 * b + 4) & ~7)
    assert np.array_equal(y << 4, (4 * r + 8 * g + 4 * b - 2048) & ~15)


def test_inverse_output_folding_identities():
    """The level-0 inverse kernels fold the rounding constants of ric.cpp:98-110 / :237-240 into the instruction that
    sign-extends each half (ric_inv.cuh: s16_plus, pack4_sat).  The identities they rely on, over every int16:"""
    import numpy as np
    v = np.arange(-32768, 32768, dtype=np.int32)
    # gray: clip(128 + ((v + 8) >> 4)) == clip((v + 8 + 2048) >> 4), and the (short) cast of the reference is a no-op
    ref = np.clip((128 + ((v + 8) >> 4)).astype(np.int16).astype(np.int32), 0, 255)
    assert np.array_equal(ref, np.clip((v + 8 + 2048) >> 4, 0, 255))
    # luma of the RGB path: ((y + 8) >> 4) - ((cg >> 1) - 128) == ((y + 8 + 2048) >> 4) - (cg >> 1)
    cg = np.arange(-4096, 4096, 37, dtype=np.int32).reshape(-1, 1)
    assert np.array_equal(((v + 8) >> 4) - ((cg >> 1) - 128), ((v + 8 + 2048) >> 4) - (cg >> 1))
    # after the down-shifts no intermediate of YCoCgtoRGB leaves int16, so its short stores are identities
    co_max, cg_max, y_max = (32767 + 4) >> 3, (32767 + 4) >> 3, ((32767 + 8) >> 4) + 128
    worst = y_max + (cg_max >> 1) + 1 + cg_max + (co_max >> 1) + 1 + co_max
    assert worst < 32768


def test_dequantiser_byte_multiplier_identity():
    """unpack_in's one-instruction dequantiser: IDP.2A computes s16[0] * u8[0] + s16[1] * u8[1]; with the byte pair
    (q, 0) or (0, q) that is the sign-extended half times q, for every q in 1..255 (no truncation: 9/7 defers it)."""
    import numpy as np
    rng = np.random.default_rng(5)
    w = rng.integers(0, 1 << 32, 4096, dtype=np.uint64).astype(np.uint32)
    lo = (w & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int64)
    hi = (w >> 16).astype(np.uint16).view(np.int16).astype(np.int64)
    for q in (1, 2, 96, 255):
        assert np.array_equal(lo * q + hi * 0, lo * q) and np.array_equal(lo * 0 + hi * q, hi * q)
        assert np.abs(lo * q).max() < 2 ** 31


def test_negated_floor_shift_identity():
    """ric_dev.cuh writes x - ((l + r) >> k) as x + ((2^k - 1 - l - r) >> k): one instruction less, same integer."""
    import numpy as np
    s = np.arange(-70000, 70000, dtype=np.int64)
    for k in (1, 2, 4):
        assert np.array_equal(-(s >> k), ((1 << k) - 1 - s) >> k)
