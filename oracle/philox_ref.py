"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the counter-based generator the
CUDA kernels use in free-running mode: Philox4x32-10 (Salmon et al., SC'11; Random123), the
Box-Muller map from two 23-bit uniforms (dense stepper) and from the two 16-bit halves of one word
(sparse stepper, noise stream v2: ``stepper_noise``).  The reference itself uses MT19937 streams that a
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


# ---- noise stream v2 of the sparse stepper (csrc/step_kernel.cuh "Noise stream v2") ---------------------------------
ANGLE_SCALE = np.float32(2.0 * np.pi / 65536.0)
ANGLE_BIAS = np.float32(-(2.0 ** 23 * float(ANGLE_SCALE)) - np.pi + np.pi / 65536.0)


def box_muller16(w, sigma):
    """Two normals (cos branch, sin branch) of std ``sigma`` from ONE word: the top 16 bits k give the radius,
    u1 = 1 - k / 65536 in (0, 1]; the low 16 bits m the angle, theta = fma(2^23 + m, ANGLE_SCALE, ANGLE_BIAS) rounded
    to float32 like the device's FMA (= 2 pi (m + 1/2) / 65536 - pi up to a constant rotation < 1e-5 rad)."""
    w = np.asarray(w, dtype=np.uint32)
    k = (w >> np.uint32(16)).astype(np.float64)
    m = (w & np.uint32(0xFFFF)).astype(np.float64)
    u1 = 1.0 - k / 65536.0
    theta = ((2.0 ** 23 + m) * float(ANGLE_SCALE) + float(ANGLE_BIAS)).astype(np.float32).astype(np.float64)
    r = sigma * np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(theta), r * np.sin(theta)


def stepper_noise(global_cell, step, n_pad, perm, sigma, seed, stream=0):
    """sigma * xi the free-running sparse stepper adds for (global cell, absolute step), in ORIGINAL gene order (length
    n = number of non-negative entries of ``perm``).  ``perm[i]`` = original gene at internal position i (or -1):
    internal gene i = 128 b + 32 j + lane draws from the call with counter (16 b + lane // 2, step, cell lo, cell hi)
    under key (seed lo, seed hi ^ stream); the lane whose parity equals the cell id's parity takes words 0, 1, its
    partner words 2, 3; the first word of the half -> genes j = 0 (cos), 1 (sin), the second -> j = 2, 3."""
    perm = np.asarray(perm)
    nb = n_pad // 128
    lane = np.arange(32)
    out = np.zeros(int((perm >= 0).sum()))
    for b in range(nb):
        w = philox4x32_10(16 * b + lane // 2, step, global_cell & 0xFFFFFFFF, global_cell >> 32, seed & 0xFFFFFFFF,
                          (seed >> 32) ^ stream)
        partner = ((lane ^ (global_cell & 1)) & 1).astype(bool)
        wa = np.where(partner, w[2], w[0])
        wb = np.where(partner, w[3], w[1])
        a0, a1 = box_muller16(wa, sigma)
        a2, a3 = box_muller16(wb, sigma)
        for j, a in enumerate((a0, a1, a2, a3)):
            idx = 128 * b + 32 * j + lane
            o = perm[idx]
            out[o[o >= 0]] = a[o >= 0]
    return out
