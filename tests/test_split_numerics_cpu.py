"""CPU model of the tensor-core operand arithmetic: the numerical claims the CUDA engine rests on.

  * x = hi + lo with hi = fp16(x), lo = fp16(x - hi) keeps >= 21 mantissa bits while lo stays a normal fp16 number;
  * the three products hi*hi + hi*lo + lo*hi with fp32 accumulation reproduce an fp32 dot product to ~1e-7 relative;
  * weights are pre-scaled by a power of two so that max|w| lands in [2^13, 2^14): exact, and keeps `lo` out of the subnormals;
  * the in-kernel operand transform re-splits  y = silu(scale * (hi + lo) + shift): its operand error stays at the 2^-21 level.
No GPU involved; numpy fp16 rounding is IEEE round-to-nearest-even like `__float2half_rn`."""
import numpy as np


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def test_split_keeps_21_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 16) * 3).astype(np.float32)
    hi, lo = split(x)
    r = hi.astype(np.float32) + lo.astype(np.float32)
    big = np.abs(x) > 0.25                      # lo is a normal fp16 number there (|lo| >= 2^-13 * |x| ... > 6.1e-5)
    rel = np.abs(r - x)[big] / np.abs(x)[big]
    assert rel.max() <= 2.0 ** -21
    assert np.abs(r - x).max() <= 2.0 ** -21 * np.abs(x).max()      # small values: absolute error of a subnormal lo (<= 3e-8)


def test_weight_prescale_is_exact_and_normalises_lo():
    rng = np.random.default_rng(1)
    w = (rng.standard_normal(4096) * 0.02).astype(np.float32)
    e = int(np.frexp(np.abs(w).max())[1])
    scale = np.float32(2.0 ** (14 - e))
    ws = w * scale
    assert 2.0 ** 13 <= np.abs(ws).max() < 2.0 ** 14 and np.array_equal(ws / scale, w)      # power of two: exact both ways
    hi, lo = split(ws)
    assert np.isfinite(hi.astype(np.float32)).all()
    r = (hi.astype(np.float32) + lo.astype(np.float32)) / scale
    # relative to the largest weight the pair carries > 21 bits even for the small weights
    assert np.abs(r - w).max() <= 2.0 ** -22 * np.abs(w).max()


def test_three_pass_dot_product_matches_fp32():
    rng = np.random.default_rng(2)
    K = 1152                                     # 3x3 x 128 channels
    a = rng.standard_normal((64, K)).astype(np.float32)
    b = (rng.standard_normal((K, 32)) / np.sqrt(K)).astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    ah, al = split(a)
    bh, bl = split(b * np.float32(2.0 ** 10))
    f = lambda t: t.astype(np.float32)           # noqa: E731  (fp32 accumulation of exact fp16 x fp16 products)
    got = (f(ah) @ f(bh) + (f(ah) @ f(bl) + f(al) @ f(bh))) * np.float32(2.0 ** -10)
    single = (f(ah) @ f(bh)) * np.float32(2.0 ** -10)
    scale = np.abs(exact).max()
    assert np.abs(got - exact).max() / scale < 2e-6       # fp32-accumulation level
    assert np.abs(single - exact).max() / scale > 1e-4    # one fp16 pass alone misses the 1e-3 end-to-end bar by far


def test_in_kernel_transform_operand_error():
    rng = np.random.default_rng(3)
    h = (rng.standard_normal(1 << 16) * 2).astype(np.float32)            # raw conv output
    sc, sh = np.float32(1.7), np.float32(-0.3)                             # GroupNorm folded into scale / shift
    silu = lambda v: v / (1 + np.exp(-v))                                  # noqa: E731
    exact = silu(sc.astype(np.float64) * h.astype(np.float64) + sh)
    hi, lo = split(h)                                                      # raw planes written by the producing epilogue
    y = silu((hi.astype(np.float32) + lo.astype(np.float32)) * sc + sh).astype(np.float32)
    yh, yl = split(y)                                                      # what the MMAs read
    got = yh.astype(np.float64) + yl.astype(np.float64)
    assert np.abs(got - exact).max() <= 4e-6 * max(1.0, np.abs(exact).max())
