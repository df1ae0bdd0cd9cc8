

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
