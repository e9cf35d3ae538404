// Layout, reduction, read-out and debug kernels (sm_100a).
//   * pack/unpack between the reference's (genes, cells) float64 host layout and the cell-major,
//     gene-permuted device state (MatriSeq.py:340-365 builds the former);
//   * per-gene sums for transformData's gene means (MatriSeq.py:851);
//   * the count read-out exp -> size factor -> clamp -> Poisson (MatriSeq.py:873-885).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "step_kernel.cuh"

namespace scr {

// staging (n x cnt, row-major float64, original gene order) -> state[(cell0+c)][internal gene]
template <typename T>
__global__ void pack_states_kernel(const double* __restrict__ stage, int64_t cnt, T* __restrict__ state,
                                   int64_t ld, int64_t cell0, const int* __restrict__ perm, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // internal gene
  if (i >= n_pad) return;
  const int o = perm[i];
  for (int64_t c = blockIdx.y; c < cnt; c += gridDim.y)
    state[(cell0 + c) * ld + i] = o >= 0 ? static_cast<T>(stage[static_cast<int64_t>(o) * cnt + c]) : T(0);
}

template <typename T>
__global__ void unpack_states_kernel(double* __restrict__ stage, int64_t cnt, const T* __restrict__ state,
                                     int64_t ld, int64_t cell0, const int* __restrict__ perm, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int o = perm[i];
  if (o < 0) return;
  for (int64_t c = blockIdx.y; c < cnt; c += gridDim.y)
    stage[static_cast<int64_t>(o) * cnt + c] = static_cast<double>(state[(cell0 + c) * ld + i]);
}

template <typename T>
__global__ void broadcast_state_kernel(const double* __restrict__ x0, T* __restrict__ state, int64_t ld,
                                       int64_t cells, const int* __restrict__ perm, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int o = perm[i];
  const T v = o >= 0 ? static_cast<T>(x0[o]) : T(0);
  for (int64_t c = blockIdx.y; c < cells; c += gridDim.y) state[c * ld + i] = v;
}

// Per-gene sums over the context's cells in a FIXED order (run-to-run deterministic): block row y accumulates cells
// y, y + gridDim.y, ... in float64 into partial[y][gene]; gene_sums_finish_kernel adds the partials in index order.
template <typename T>
__global__ void gene_sums_kernel(const T* __restrict__ state, int64_t ld, int64_t cells, int n_pad,
                                 double* __restrict__ partial) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // internal gene
  if (i >= n_pad) return;
  double acc = 0.0;
  for (int64_t c = blockIdx.y; c < cells; c += gridDim.y) acc += static_cast<double>(state[c * ld + i]);
  partial[static_cast<int64_t>(blockIdx.y) * n_pad + i] = acc;
}
__global__ void gene_sums_finish_kernel(const double* __restrict__ partial, int rows, int n_pad, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  double acc = 0.0;
  for (int y = 0; y < rows; ++y) acc += partial[static_cast<int64_t>(y) * n_pad + i];
  sums[i] = acc;
}

// ------------------------------------------------------------------------------------------
// Poisson read-out.  Same algorithm family numpy's legacy generator uses behind
// np.random.poisson (MatriSeq.py:884): multiplication method below lambda = 10, Hoermann's PTRS
// transformed rejection above (W. Hoermann, Insurance: Mathematics and Economics 12 (1993)).
// Restated in oracle/poisson_oracle.c; uniforms are consumed strictly in order so both sides
// agree draw for draw.
// ------------------------------------------------------------------------------------------
struct UniformSource {
  const double* supplied;   // nullptr -> Philox
  int64_t sup_base;
  int sup_count, used;
  uint32_t c0, c2, c3, call;
  const RoundKeys* rk;      // Philox round keys of (seed, SCR_STREAM_COUNTS): kernel parameter, constant bank
  uint32_t w0, w1, w2, w3;  // words of the current call, consumed front to back (scalars: a dynamically indexed
  int have;                 //  array would live in local memory)
  int* err;
  __device__ double next() {
    if (supplied) {
      if (used >= sup_count) { *err = 1; return 0.5; }
      return supplied[sup_base + used++];
    }
    if (have == 0) {
      uint32_t w[4];
      philox4x32_10_rk(c0, call++, c2, c3, *rk, w);
      w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
      have = 4;
    }
    const uint32_t v = w0;
    w0 = w1; w1 = w2; w2 = w3;
    --have;
    ++used;
    return (static_cast<double>(v) + 0.5) * 2.3283064365386963e-10;   // (v + 1/2) / 2^32 in (0,1)
  }
};

__device__ inline int poisson_draw(double lam, UniformSource& u) {
  if (!(lam > 0.0)) return 0;
  if (lam > 2.0e9) lam = 2.0e9;   // counts are int32 across the ABI: rates beyond 2e9 (or +inf from an overflowing exp) saturate there
  if (lam < 10.0) {
    const double enlam = exp(-lam);
    double prod = 1.0;
    int k = 0;
    for (;;) {
      prod *= u.next();
      if (prod > enlam) ++k; else return k;
      if (k > 100000) return k;   // unreachable for lam < 10 unless the uniforms are degenerate
    }
  }
  const double slam = sqrt(lam);   // log(lam) is only needed by the full acceptance test (~1 draw in 8): formed there
  const double b = 0.931 + 2.53 * slam;
  const double a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  const double vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int guard = 0; guard < 100000; ++guard) {
    const double U = u.next() - 0.5;
    const double V = u.next();
    const double us = 0.5 - fabs(U);
    const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return static_cast<int>(kf);
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + kf * log(lam) - lgamma(kf + 1.0)))
      return static_cast<int>(kf);
  }
  return static_cast<int>(lam);
}

// Extension of the read-out (north_star (4), default off; restated in oracle/poisson_oracle.c gamma_mt): G ~ Gamma(shape,
// scale) by Marsaglia & Tsang (ACM TOMS 26, 2000) over the entry's uniform stream -- shape < 1: one uniform for the boost
// U^(1/shape); per attempt two uniforms for a Box-Muller normal (cos branch) and one for the acceptance test.
__device__ inline double gamma_mt(double shape, double scale, UniformSource& u) {
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(u.next(), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int guard = 0; guard < 1000; ++guard) {
    const double u1 = u.next(), u2 = u.next(), U = u.next();
    const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    double v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (U < 1.0 - 0.0331 * (z * z) * (z * z) || log(U) < 0.5 * z * z + d * (1.0 - v + log(v))) return boost * d * v * scale;
  }
  return shape * scale;   // unreachable unless the uniforms are degenerate
}

// Float32 screening of the multiplication method (lambda < 10), EXACT by construction: the same Philox words, the same
// draw order, but product and threshold carried in float32 with a guard band.  With eps = 2^-24:
//   lambda_f = ex2(t2) * size  (t2 = log2(e) * expScale * (x + shift) rounded to float32, |t2| < 64, size in [0.05, 8]):
//              |lambda_f - lambda| <= lambda (2^-22 + 2 eps) + lambda |t2| eps ln 2 <= 6.8e-6 for lambda < 10;
//   enlam_f  = ex2(-log2(e) lambda_f): relative error <= 6.8e-6 + 14.5 * 2 eps * ln 2 + 2^-22 <= 8e-6;
//   prod_f   : each uniform (v + 1/2) / 2^32 is one I2F and one FMA (2 eps), each multiply eps: <= 64 * 3 eps = 1.2e-5
// so prod_f and enlam_f together are within 2e-5 of the float64 quantities the exact sampler compares; a comparison
// is taken only when prod_f lies outside enlam_f (1 +- 1e-4), otherwise (about 2e-4 of the entries, or any input outside
// the ranges above) -1 is returned and the caller runs the float64 sampler from the first uniform again.
// tests/test_counts_guard.py checks the bound on the host; the GPU suite compares both paths entry for entry.
__device__ __forceinline__ int poisson_small_f32(float lam_f, uint32_t gene, uint64_t gid, const RoundKeys& rk) {
  const float enlam = ptx_ex2(lam_f * -1.44269504088896340736f);
  const float hi = enlam * (1.0f + 1e-4f), lo = enlam * (1.0f - 1e-4f);
  float prod = 1.0f;
  int k = 0;
  for (uint32_t call = 0; call < 16; ++call) {   // 64 uniforms: P(Poisson(10) >= 64) < 1e-28
    uint32_t w[4];
    philox4x32_10_rk(gene, call, static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), rk, w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      prod *= fmaf(__uint2float_rn(w[j]), 2.3283064365386963e-10f, 1.16415321826934814e-10f);   // (v + 1/2) / 2^32
      if (prod > hi) ++k;
      else return prod < lo ? k : -1;
    }
  }
  return -1;
}

// Float32 screening of Hoermann's PTRS sampler (10 < lambda < kPtrsLamMax), EXACT by construction like poisson_small_f32:
// the same Philox words in the same order -- uniforms 2i, 2i+1 are (U, V) of draw i -- every quantity the float64 sampler
// compares is formed in float32 and a decision is only taken when it clears a guard band wider than the float32 error;
// anything closer, and any draw that needs more than 16 pairs, returns -1 and the caller runs the float64 sampler from the
// first uniform again.  With eps = 2^-24 and lam_f within 1e-6 lambda of the float64 rate (see poisson_small_f32):
//   b_f, a_f, vr_f, invalpha_f: relative error <= 1e-6 (sqrt.approx 1 ulp, FMAs), a_f absolute <= 2e-6;
//   u_f, V_f (one I2F + one FMA): absolute <= 6e-8, so us_f = 0.5 - |u_f - 0.5| absolute <= 1.2e-7;
//   q = 2 a / us (rcp.approx 1 ulp): relative <= 1e-5 + 1.2e-7 / us;
//   x = (q + b) U + lambda + 0.43: absolute <= 1e-6 lambda + 1.2e-4 + q (5e-6 + 6e-8 / us)
//       -> floor(x) is trusted when frac(x) is further than g = 2e-6 lambda + 3e-4 + q (1e-5 + 2e-7 / us) from 0 and 1;
//   "us >= 0.07", "us < 0.013": decided outside +-1e-6;  "V <= vr": outside +-2e-6;  "V > us": outside +-1e-6;
//   full test ln V + ln invalpha - ln(a / us^2 + b) <= -lambda + k ln lambda - lgamma(k + 1) (lg2.approx absolute 2.4e-7,
//   lgamma from a float32 table of the exact values, cancellation at magnitude <= 8200 -> ulp 4.9e-4): both sides together
//   within 1.5e-3 + 2e-6 k + 1e-7 / V + 5e-7 / us; decided outside G = 3e-3 + 4e-6 k + 2e-7 / V + 1e-6 / us.
// tests/test_counts_guard.py replays this arithmetic on the host against the C oracle; the GPU suite compares both device
// paths entry for entry (3e8 entries in test_count_readout_moments_at_scale).
constexpr float kPtrsLamMax = 1000.0f;
constexpr int kLgamTable = 2048;                     // lgamma(k + 1), k < 2048 (lambda + 10 sqrt(lambda) < 1320)
__device__ __forceinline__ int poisson_ptrs_f32(float lam, uint32_t gene, uint64_t gid, const RoundKeys& rk,
                                                const float* __restrict__ lgam) {
  const float slam = ptx_sqrt(lam);
  const float b = fmaf(2.53f, slam, 0.931f);
  const float a = fmaf(0.02483f, b, -0.059f);
  const float vr = fmaf(-3.6224f, ptx_rcp(b - 2.0f), 0.9277f);
  const float g0 = fmaf(2e-6f, lam, 3e-4f);
  const float a2 = a + a;
  for (uint32_t call = 0; call < 8; ++call) {
    uint32_t w[4];
    philox4x32_10_rk(gene, call, static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), rk, w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u = fmaf(__uint2float_rn(w[2 * h]), 2.3283064365386963e-10f, 1.16415321826934814e-10f);
      const float V = fmaf(__uint2float_rn(w[2 * h + 1]), 2.3283064365386963e-10f, 1.16415321826934814e-10f);
      const float U = u - 0.5f;
      const float us = 0.5f - fabsf(U);
      if (fabsf(us - 0.07f) < 1e-6f || fabsf(us - 0.013f) < 1e-6f) return -1;
      const float rus = ptx_rcp(us);
      const float q = a2 * rus;
      const float x = fmaf(q + b, U, lam + 0.43f);
      const float g = fmaf(q, fmaf(2e-7f, rus, 1e-5f), g0);
      const float kf = floorf(x);
      const float fr = x - kf;
      const bool k_exact = fr > g && fr < 1.0f - g;
      if (us > 0.07f) {
        if (V < vr - 2e-6f) return k_exact ? static_cast<int>(kf) : -1;
        if (V < vr + 2e-6f) return -1;
      }
      if (x < g) {                       // kf < 0 -> next draw (only the sign matters here)
        if (x < -g) continue;
        return -1;
      }
      if (us < 0.013f) {                 // squeeze: V > us -> next draw
        if (V > us + 1e-6f) continue;
        if (V > us - 1e-6f) return -1;
      }
      if (!k_exact || kf >= static_cast<float>(kLgamTable)) return -1;
      const float lhs = (ptx_lg2(V) + ptx_lg2(fmaf(1.1328f, ptx_rcp(b - 3.4f), 1.1239f)) - ptx_lg2(fmaf(a, rus * rus, b))) * 0.69314718055994531f;
      const float rhs = fmaf(kf, ptx_lg2(lam) * 0.69314718055994531f, -lam) - lgam[static_cast<int>(kf)];
      const float G = fmaf(4e-6f, kf, 3e-3f) + 2e-7f * ptx_rcp(V) + 1e-6f * rus;
      if (lhs < rhs - G) return static_cast<int>(kf);
      if (lhs > rhs + G) continue;
      return -1;
    }
  }
  return -1;
}

// one gene of the read-out, in the order the kernel walks them (sorted by shift): one 16-byte load per entry instead of
// the chain order -> inverse permutation -> shift
struct __align__(16) CountGene { int gene; int state_idx; double shift; };

// one CTA per cell of the chunk.  Thread t handles gene `order[t]`, where `order` sorts the genes by
// their shift (the dominant term of log lambda): the 32 lanes of a warp then see similar lambda and leave
// the sampler's loops together instead of waiting for the largest one.  Counts are staged in shared
// memory and written as one coalesced int32 row in the caller's gene order.
// FAST (free-running Philox stream, no lambda dump): entries with lambda < 10 go through poisson_small_f32 first and
// fall back to the float64 sampler only when it declines; results are identical to FAST = false.
#ifndef SCR_COUNTS_PTRS_F32
#define SCR_COUNTS_PTRS_F32 1   // 0: lambda >= 10 always through the float64 sampler (A/B builds)
#endif
#ifndef SCR_COUNTS_THREADS
#define SCR_COUNTS_THREADS 128     // threads of a CTA = of one cell's read-out (64 / 128 / 256: 76.4 / 86.4 / 83.2 G entries/s: fewer warps wait at a cell's barrier)
#endif
#ifndef SCR_COUNTS_CTAS_PER_SM
#define SCR_COUNTS_CTAS_PER_SM (1024 / SCR_COUNTS_THREADS)   // 64 registers: the sampler is latency-bound, more warps pay (2 / 3 / 4 CTAs of 256: 28.8 / 22.5 / 21.5 ms at 400k cells)
#endif
// OUT = int32_t, or a compact type (uint16_t / uint8_t): a count that does not fit below the type's largest value is
// stored as that value (the escape code) and appended as {local cell, gene, count} to the overflow list (capacity
// ovf_cap triplets; *ovf_count keeps counting beyond it so the host can tell).  Mean lambda of the read-out is ~2
// (MatriSeq.py:855-873 defaults), so a byte per entry holds all but a handful of the 3e9 entries of config c5.
template <typename OUT> struct CountLimit { static constexpr int kEscape = 0x7fffffff; };
template <> struct CountLimit<uint16_t> { static constexpr int kEscape = 0xffff; };
template <> struct CountLimit<uint8_t> { static constexpr int kEscape = 0xff; };

template <typename T, bool FAST, typename OUT>
__global__ void __launch_bounds__(SCR_COUNTS_THREADS, SCR_COUNTS_CTAS_PER_SM)
counts_kernel(const T* __restrict__ state, int64_t ld, const CountGene* __restrict__ genes, const float* __restrict__ lgam,
                              int n, int64_t cell0,
                              double exp_scale, const double* __restrict__ sizef, const RoundKeys rk,
                              int64_t cell_offset, const double* __restrict__ uniforms, int draws,
                              OUT* __restrict__ out, double* __restrict__ lam_out, int* err,
                              int32_t* __restrict__ ovf, unsigned long long ovf_cap, unsigned long long* __restrict__ ovf_count,
                              double nb_phi, double dropout_p) {
  extern __shared__ int32_t s_counts[];
  const int64_t c = cell0 + blockIdx.x;   // local slot
  const uint64_t gid = static_cast<uint64_t>(cell_offset + c);
  const double sf = sizef[c];
  const T* row = state + c * ld;
  const bool sf_ok = FAST && sf >= 0.05 && sf <= 8.0;
  const float sf_f = static_cast<float>(sf);
  const double esl = exp_scale * 1.4426950408889634;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const CountGene cg = genes[t];
    const T x_raw = row[cg.state_idx];
    const int o = cg.gene;
    SCR_CHECK(err, o >= 0 && o < n && cg.state_idx >= 0 && cg.state_idx < ld, 108);
    const double x = static_cast<double>(x_raw);
    int cnt = -1;
    if (FAST && sf_ok) {
      const float t2 = static_cast<float>((x + cg.shift) * esl);
      if (t2 > -64.0f && t2 < 12.0f) {
        const float lam_f = ptx_ex2(t2) * sf_f;
        if (lam_f < 9.999f) cnt = poisson_small_f32(lam_f, static_cast<uint32_t>(o), gid, rk);
#if SCR_COUNTS_PTRS_F32
        else if (lam_f > 10.001f && lam_f < kPtrsLamMax) cnt = poisson_ptrs_f32(lam_f, static_cast<uint32_t>(o), gid, rk, lgam);
#endif
      }
    }
    if (cnt < 0) {
      double lam = exp(exp_scale * (x + cg.shift)) * sf;
      if (lam < 0.0) lam = 0.0;
      UniformSource u;
      u.supplied = uniforms;
      u.sup_base = (c * n + o) * static_cast<int64_t>(draws);
      u.sup_count = draws;
      u.used = 0;
      u.c0 = static_cast<uint32_t>(o);
      u.c2 = static_cast<uint32_t>(gid);
      u.c3 = static_cast<uint32_t>(gid >> 32);
      u.rk = &rk; u.call = 0; u.have = 0; u.err = err;
      u.w0 = u.w1 = u.w2 = u.w3 = 0;
      if (lam_out) lam_out[static_cast<int64_t>(blockIdx.x) * n + o] = lam;
      // extensions, default off (FAST launches never have them): dropout uniform first, then the Gamma factor of the
      // negative-binomial draw, then the Poisson sampler -- the order oracle_counts_ext_batch consumes uniforms in
      bool drop = false;
      if (dropout_p > 0.0) drop = u.next() < dropout_p;
      if (nb_phi > 0.0 && lam > 0.0) lam *= gamma_mt(1.0 / nb_phi, nb_phi, u);
      cnt = poisson_draw(lam, u);
      if (drop) cnt = 0;
    }
    s_counts[o] = cnt;
  }
  __syncthreads();
  OUT* orow = out + static_cast<int64_t>(blockIdx.x) * n;
  constexpr int kEscape = CountLimit<OUT>::kEscape;
  for (int o = threadIdx.x; o < n; o += blockDim.x) {
    const int32_t v = s_counts[o];
    if (sizeof(OUT) < 4 && v >= kEscape) {
      const unsigned long long k = atomicAdd(ovf_count, 1ULL);
      if (k < ovf_cap) {
        ovf[3 * k] = static_cast<int32_t>(c);
        ovf[3 * k + 1] = o;
        ovf[3 * k + 2] = v;
      }
      orow[o] = static_cast<OUT>(kEscape);
    } else {
      orow[o] = static_cast<OUT>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// debug / parity aids
// ------------------------------------------------------------------------------------------
__global__ void debug_philox_kernel(int64_t count, const uint32_t* __restrict__ ctr, uint32_t k0, uint32_t k1,
                                    uint32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  uint32_t w[4];
  philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1, w);
  for (int j = 0; j < 4; ++j) out[4 * i + j] = w[j];
}

// the noise step_kernel adds for (global cell, step): one thread per (block, lane)
__global__ void debug_noise_kernel(int nb, uint32_t gid_lo, uint32_t gid_hi, uint32_t step, float rscale,
                                   uint32_t k0, uint32_t k1, const int* __restrict__ perm, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb * 32) return;
  const int b = t >> 5, lane = t & 31;
  uint32_t w[4];
  float a[4];
#if SCR_NOISE16
  // stream v2 (step_kernel.cuh): one call per lane pair; the lane of the cell id's parity takes words 0, 1
  philox4x32_10(static_cast<uint32_t>(b * 16 + (lane >> 1)), step, gid_lo, gid_hi, k0, k1, w);
  const bool partner_half = ((static_cast<uint32_t>(lane) ^ gid_lo) & 1u) != 0;
  box_muller16(partner_half ? w[2] : w[0], rscale, a[0], a[1]);
  box_muller16(partner_half ? w[3] : w[1], rscale, a[2], a[3]);
#else
  philox4x32_10(static_cast<uint32_t>(b * 32 + lane), step, gid_lo, gid_hi, k0, k1, w);
  box_muller(w[0], w[1], rscale, a[0], a[1]);
  box_muller(w[2], w[3], rscale, a[2], a[3]);
#endif
  for (int j = 0; j < 4; ++j) {
    const int o = perm[b * 128 + j * 32 + lane];
    if (o >= 0) out[o] = static_cast<double>(a[j]);
  }
}

}  // namespace scr
