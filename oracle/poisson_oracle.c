/* TEST INFRASTRUCTURE ONLY (oracle) -- C restatement of the count read-out sampler.
 *
 * The reference draws every count with numpy's legacy np.random.poisson (MatriSeq.py:884).  That
 * generator's algorithm (numpy/random/src/legacy + distributions.c, not vendored in the reference;
 * numpy is unpinned there) is: 0 for lam == 0, the multiplication method for lam < 10
 * and Hoermann's PTRS transformed rejection (Insurance: Mathematics and Economics 12, 1993) for
 * lam >= 10.  Its MT19937 stream is sequential, so a parallel run can only agree with it in
 * distribution; what CAN be exact is "same uniforms in, same integers out".  This file restates
 * the sampler over an explicit array of uniforms consumed in order; tests feed the same uniforms
 * to the CUDA kernel (scr_counts) and require bit-identical counts.
 *
 * Pinning: tests/test_oracle_poisson.py checks this file against np.random.poisson itself by
 * replaying the uniforms numpy's legacy generator consumes (random_sample stream of the same seed).
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* returns the count; *used = uniforms consumed; -1 if the array ran out */
static long poisson_one(double lam, const double* u, int avail, int* used) {
  int p = 0;
  long k = 0;
  if (!(lam > 0.0)) { *used = 0; return 0; }
  if (lam < 10.0) {
    const double enlam = exp(-lam);
    double prod = 1.0;
    for (;;) {
      if (p >= avail) { *used = p; return -1; }
      prod *= u[p++];
      if (prod > enlam) k++; else { *used = p; return k; }
    }
  }
  {
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
      double U, V, us, kf;
      if (p + 2 > avail) { *used = p; return -1; }
      U = u[p++] - 0.5;
      V = u[p++];
      us = 0.5 - fabs(U);
      kf = floor((2.0 * a / us + b) * U + lam + 0.43);
      if (us >= 0.07 && V <= vr) { *used = p; return (long)kf; }
      if (kf < 0.0 || (us < 0.013 && V > us)) continue;
      if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + kf * loglam - lgamma(kf + 1.0))) {
        *used = p;
        return (long)kf;
      }
    }
  }
}

/* counts[i] for lam[i] with uniforms u[i*draws .. i*draws+draws); returns number of entries that
 * ran out of uniforms (their count is set to -1) */
long oracle_poisson_batch(long entries, const double* lam, const double* u, int draws, int32_t* counts,
                          int32_t* used_out) {
  long bad = 0;
  for (long i = 0; i < entries; ++i) {
    int used = 0;
    const long k = poisson_one(lam[i], u + (size_t)i * draws, draws, &used);
    counts[i] = (int32_t)k;
    if (used_out) used_out[i] = used;
    if (k < 0) bad++;
  }
  return bad;
}

/* sequential-stream variant: consumes one shared stream like numpy does (for pinning against
 * np.random.poisson); returns uniforms consumed in total or -1 */
long oracle_poisson_stream(long entries, const double* lam, const double* u, long avail, int64_t* counts) {
  long pos = 0;
  for (long i = 0; i < entries; ++i) {
    int used = 0;
    const long room = avail - pos;
    const long k = poisson_one(lam[i], u + pos, room > 2000000000L ? 2000000000 : (int)room, &used);
    if (k < 0) return -1;
    counts[i] = k;
    pos += used;
  }
  return pos;
}

/* ---- extensions of the read-out (north_star (4): negative-binomial draws and dropout; DEFAULT OFF, not in the
 * reference, which only has np.random.poisson at MatriSeq.py:884 and a commented-out dropout at MatriSeq.py:887-889).
 * Own oracle, same contract: explicit uniforms consumed in a fixed order per entry --
 *   1. dropout_p > 0:  one uniform d; the entry is dropped (count 0) when d < dropout_p, AFTER all its draws were made
 *      (so the stream position does not depend on the outcome);
 *   2. nb_phi > 0:     G ~ Gamma(shape 1/phi, scale phi) (mean 1, variance phi) by Marsaglia & Tsang (ACM TOMS 26, 2000):
 *      shape < 1 -> one uniform for the boost U^(1/shape), shape += 1; then per attempt two uniforms for a Box-Muller normal
 *      (cos branch) and one for the acceptance test; the rate becomes lam * G (count ~ NB with mean lam, var lam + phi lam^2);
 *   3. the Poisson sampler above on what is left of the uniforms.
 * CUDA counterpart: counts_kernel's extension path (scr_counts_ext). */
static double gamma_mt(double shape, double scale, const double* u, int avail, int* p_io, int* ok) {
  int p = *p_io;
  double boost = 1.0;
  *ok = 1;
  if (shape < 1.0) {
    if (p >= avail) { *ok = 0; return 0.0; }
    boost = pow(u[p++], 1.0 / shape);
    shape += 1.0;
  }
  {
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
      double z, v, U;
      if (p + 3 > avail) { *ok = 0; *p_io = p; return 0.0; }
      z = sqrt(-2.0 * log(u[p])) * cos(6.283185307179586 * u[p + 1]);
      U = u[p + 2];
      p += 3;
      v = 1.0 + c * z;
      if (v <= 0.0) continue;
      v = v * v * v;
      if (U < 1.0 - 0.0331 * (z * z) * (z * z) || log(U) < 0.5 * z * z + d * (1.0 - v + log(v))) {
        *p_io = p;
        return boost * d * v * scale;
      }
    }
  }
}

long oracle_counts_ext_batch(long entries, const double* lam, const double* u, int draws, double nb_phi, double dropout_p,
                             int32_t* counts, int32_t* used_out) {
  long bad = 0;
  for (long i = 0; i < entries; ++i) {
    const double* ui = u + (size_t)i * draws;
    int p = 0, used = 0, ok = 1, drop = 0;
    double rate = lam[i];
    long k;
    if (dropout_p > 0.0) {
      if (p >= draws) ok = 0; else drop = ui[p++] < dropout_p;
    }
    if (ok && nb_phi > 0.0 && rate > 0.0) rate *= gamma_mt(1.0 / nb_phi, nb_phi, ui, draws, &p, &ok);
    k = ok ? poisson_one(rate, ui + p, draws - p, &used) : -1;
    if (k >= 0 && drop) k = 0;
    counts[i] = (int32_t)k;
    if (used_out) used_out[i] = p + used;
    if (k < 0) bad++;
  }
  return bad;
}
