"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the counter-based generator the
CUDA kernels use in free-running mode: Philox4x32-10 (Salmon et al., SC'11; Random123) and the
Box-Muller map from two 23-bit uniforms.  The reference itself uses MT19937 streams that a
parallel run cannot reproduce; this file pins OUR generator: known-answer
vectors from Random123's kat_vectors plus exact agreement with the device words.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 counters; returns four uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & MASK for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in c]


def unit_float(bits):
    """f in [1,2) from the top 23 bits of a word (device: 0x3f800000 | (w >> 9))."""
    return ((np.asarray(bits, dtype=np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)


def box_muller(wa, wb, sigma):
    """Two normals (cos branch, sin branch) of std ``sigma`` from two words, float64 maths on the
    same 23-bit uniforms the device uses: u1 = 2 - f(wa) in (0,1], theta = 2*pi*f(wb) - 3*pi."""
    u1 = 2.0 - unit_float(wa).astype(np.float64)
    theta = 2.0 * np.pi * unit_float(wb).astype(np.float64) - 3.0 * np.pi
    r = sigma * np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(theta), r * np.sin(theta)
