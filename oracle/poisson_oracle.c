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
