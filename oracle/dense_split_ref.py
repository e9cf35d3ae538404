"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the operand encodings of the dense stepper
(scrutiny_b200/csrc/dense_kernel.cuh): the three-product split-precision form of X W^T with fp32 accumulation.

  fp16 split   hi = fp16(s v), lo = fp16(s v - hi), power-of-two scales (state 64; W: max|W| into [2^14, 2^15))
  3xTF32       hi = fp32 word with the low 13 mantissa bits cleared, lo = v - hi (what kind::tf32 reads)

Products of 11-bit operands are exact in fp32; the accumulation order of the tensor core is not specified, so this
restatement sums in float64 and reports the REPRESENTATION error of the scheme (what bounds the GPU result from below);
tests/test_dense_encoding_cpu.py compares it with the float64 product."""
import numpy as np

STATE_SCALE = 64.0


def w_scale(W):
    wmax = float(np.max(np.abs(W))) if W.size else 0.0
    if not np.isfinite(wmax) or wmax <= 0:
        return 1.0
    e = int(np.floor(np.log2(np.float32(wmax))))
    return float(2.0 ** max(-100, min(100, 14 - e)))


def split_f16(v, scale):
    vs = (np.asarray(v, dtype=np.float32) * np.float32(scale)).astype(np.float32)
    hi = vs.astype(np.float16)
    lo = (vs - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def split_tf32(v):
    v = np.ascontiguousarray(v, dtype=np.float32)
    hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = v - hi
    lo_t = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)   # the tensor core truncates lo as well
    return hi, lo_t


def product(X, W, kind="f16"):
    """Z[c, g] = sum_k X[c, k] W[g, k] through the three products lo(X)hi(W) + hi(X)lo(W) + hi(X)hi(W)."""
    X = np.asarray(X, dtype=np.float32)
    W = np.asarray(W, dtype=np.float32)
    if kind == "f16":
        sw = w_scale(W)
        xh, xl = split_f16(X, STATE_SCALE)
        wh, wl = split_f16(W, sw)
        f = np.float64
        z = xl.astype(f) @ wh.astype(f).T + xh.astype(f) @ wl.astype(f).T + xh.astype(f) @ wh.astype(f).T
        return z / (STATE_SCALE * sw)
    xh, xl = split_tf32(X)
    wh, wl = split_tf32(W)
    f = np.float64
    return xl.astype(f) @ wh.astype(f).T + xh.astype(f) @ wl.astype(f).T + xh.astype(f) @ wh.astype(f).T
