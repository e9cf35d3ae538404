// Persistent time-stepping kernel of the MatriSeq hot path (sm_100a).
//
// Replaces the per-cell Python loops findCellNextState x developCell
// (reference MatriSeq.py:683-736, 503-537) and, in PROBE mode, the sampled-cell convergence loop of
// findOptimalIterations / findOptimalAlpha (MatriSeq.py:650-669, 456-482).
//
// Design (DESIGN.md "Stepper"):
//   * a work item is a GROUP of CPG cells (4 x fp32 or 2 x fp64) whose states sit interleaved in
//     shared memory as x[gene][cell] -- 16 bytes per gene -- for ALL timesteps of the phase;
//   * WQ warps own one group; lane (blk, lane) owns genes blk*128 + j*32 + lane, j = 0..3, so one
//     Philox4x32-10 call yields the four normals of one cell and every smem access is a
//     conflict-free 128-bit row;
//   * W is a sliced-ELL operator over permuted genes (long rows first, regulators next, rest by
//     in-degree) so that 32-row tiles are homogeneous; rows longer than `cap` spill into fixed
//     length chunks evaluated by all lanes first (phase A) and added by the owner (phase B);
//   * only the leading `npp` genes (those any row can read) are ping-ponged between two buffers;
//     the rest are updated in place, which is what lets several groups share one SM;
//   * operator tables + per-node proj/const are staged once per CTA with a bulk async copy
//     (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier;
//   * everything a 128-gene block needs to get going is ONE 16-byte record at a compile-time shared-memory
//     offset: {first gene | flags, slot count, generic pointer to the block's first operator entry}; the flags carry
//     what used to be per-block comparisons (ping-ponged? long rows? last regulator block of the warp? bias?), the
//     pointer already encodes whether the entries are read from the staged part of the operator or from global memory;
//   * CTAs are persistent and pull groups (sorted by step count, longest first) from an atomic
//     counter; inside a step a group synchronises through three mbarriers (regulator values written / reads of
//     the old buffer finished / long-row partials in place: arrive early, wait late), between groups and in PROBE
//     mode through its own named barrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scr {

// Bounds-checked builds (-DSCR_DEBUG_CHECKS=1; compute-sanitizer is closed on the GPU pool, profiles/r2_sanitizer.txt): every
// index the kernels form from tables or items is range-checked on the device; a violation records its code in the context's
// debug word and the ABI call returns SCR_E_CUDA.  Off by default: the checks cost instructions in the block loop.
#ifndef SCR_DEBUG_CHECKS
#define SCR_DEBUG_CHECKS 0
#endif
#if SCR_DEBUG_CHECKS
#define SCR_CHECK(word, cond, code) do { if (!(cond)) atomicMax((word), (code)); } while (0)
#else
#define SCR_CHECK(word, cond, code) ((void)0)
#endif

// ------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ptx_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the hint (ns) elapses; without
// the hint the default limit is short and the loop below spins, stealing issue slots from other warps
#ifndef SCR_TRYWAIT_HINT
#define SCR_TRYWAIT_HINT 1   // 0: try_wait without a suspend-time hint (A/B measurement)
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if SCR_TRYWAIT_HINT
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// one arrival (release at CTA scope); the waiters' try_wait is the matching acquire
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrival `seq` (0-based step index) of a barrier that `count` warps arrive on per step; with SCR_LAST_ARRIVER only the
// warp that completes the count touches the mbarrier (acq_rel atomic: its arrival publishes the other warps' writes)
#ifndef SCR_LAST_ARRIVER
#define SCR_LAST_ARRIVER 0
#endif
__device__ __forceinline__ void group_arrive(uint64_t* bar, uint32_t* cnt, uint32_t seq, uint32_t count) {
#if SCR_LAST_ARRIVER
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
  if (old + 1u == (seq + 1u) * count) mbar_arrive(bar);
#else
  (void)cnt; (void)seq; (void)count;
  mbar_arrive(bar);
#endif
}
// 128-bit load from a shared-space address (LDS.128 with the table offset folded into the immediate)
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds_pc(uint32_t addr, float2*) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_pc(uint32_t addr, double2*) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Checked bit-for-bit against oracle/philox_ref.py.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// same function with the ten round keys precomputed on the host (kernel parameters live in the
// constant bank, so each XOR takes its key as an immediate-like operand: no key arithmetic at all)
struct RoundKeys { uint32_t k[20]; };
inline void make_round_keys(uint32_t k0, uint32_t k1, RoundKeys& rk) {
  for (int r = 0; r < 10; ++r) {
    rk.k[2 * r] = k0;
    rk.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
#ifndef SCR_PHILOX_ROUNDS
#define SCR_PHILOX_ROUNDS 10   // measurement builds only (profiles/r1_stepper_experiments.txt); every test assumes 10
#endif
// rounds [R0, SCR_PHILOX_ROUNDS) on a state that has already been through rounds [0, R0)
template <int R0>
__device__ __forceinline__ void philox_rounds_from(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   const RoundKeys& rk, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = R0; r < SCR_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const RoundKeys& rk, uint32_t (&out)[4]) {
  philox_rounds_from<0>(c0, c1, c2, c3, rk, out);
}
// Per-cell invariants of round 0 for the stepper's counter layout (c0 = 32*block + lane, c1 = step, c2:c3 = global
// cell id): the product with c2 and the XOR of c3 with the round key do not depend on block, lane or step, so they are
// computed once per cell group (CellRound0 in shared memory) instead of once per 128-gene block.
struct CellRound0 { uint32_t hi_p1_k0, lo_p1, c3_k1; };
__device__ __forceinline__ CellRound0 philox_cell_round0(uint32_t gid_lo, uint32_t gid_hi, const RoundKeys& rk) {
  const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * gid_lo;
  return {static_cast<uint32_t>(p1 >> 32) ^ rk.k[0], static_cast<uint32_t>(p1), gid_hi ^ rk.k[1]};
}
// the same call as philox4x32_10_rk(c0, step, gid_lo, gid_hi): p0 = 0xD2511F53 * c0 is shared by the cells of a group
__device__ __forceinline__ void philox4x32_10_cell(uint64_t p0, uint32_t step, uint32_t hi_p1_k0, uint32_t lo_p1,
                                                   uint32_t c3_k1, const RoundKeys& rk, uint32_t (&out)[4]) {
  philox_rounds_from<1>(hi_p1_k0 ^ step, lo_p1, static_cast<uint32_t>(p0 >> 32) ^ c3_k1, static_cast<uint32_t>(p0), rk, out);
}

// two N(0, sigma^2) values from two words: u1 = 2 - f(wa) in (0,1], theta = 2*pi*f(wb) - 3*pi,
// f(w) = float in [1,2) built from the top 23 bits.  rscale = -2*ln(2)*sigma^2.
__device__ __forceinline__ void box_muller_polar(uint32_t wa, uint32_t wb, float rscale, float& r, float& c, float& sn) {
  const float f1 = __uint_as_float(0x3f800000u | (wa >> 9));
  const float f2 = __uint_as_float(0x3f800000u | (wb >> 9));
  r = ptx_sqrt(fabsf(ptx_lg2(2.0f - f1) * rscale));   // |.|: never NaN if lg2(1-eps) rounds positive
  const float th = fmaf(f2, 6.28318530717958647692f, -9.42477796076937971538f);
  c = ptx_cos(th);
  sn = ptx_sin(th);
}
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float rscale, float& n_cos, float& n_sin) {
  float r, c, sn;
  box_muller_polar(wa, wb, rscale, r, c, sn);
  n_cos = r * c;
  n_sin = r * sn;
}

// ------------------------------------------------------------------------------------------
// Noise stream v2 of the sparse stepper (SCR_NOISE16, default): ONE 32-bit Philox word yields two normals.  The top 16
// bits k give the radius, u1 = 1 - k / 65536 in (0, 1] (largest radius sqrt(2 ln 65536) = 4.71 sigma), the low 16 bits
// m the angle, theta = 2 pi (m + 1/2) / 65536 - pi.  Both are unpacked with one integer instruction each through the
// 2^23 + x float trick (0x4b000000 | x is the float 2^23 + x for x < 2^23) and one exact FMA:
//   u1    = fma(2^23 + k, -2^-16, 129)                        (exact: 129 - 128 - k / 65536)
//   theta = fma(2^23 + m, 2 pi / 65536, -(2^23 * 2 pi / 65536) - pi + pi / 65536)
// A Philox4x32-10 call therefore carries EIGHT normals: the four genes (j = 0..3) a lane owns in a 128-gene block for
// one cell, for BOTH lanes of a lane pair.  Counter (16 * block + lane / 2, step, global cell id lo, hi); the lane whose
// parity equals the parity of the global cell id takes words 0, 1, its partner words 2, 3; word 0 / 2 -> genes j = 0
// (cos) and 1 (sin), word 1 / 3 -> genes j = 2, 3.  The stepper's fast path lets each lane compute the calls of the two
// cells of its own parity (the host places even cell ids in slots 0, 1 and odd ones in slots 2, 3 of a group) and
// passes the partner's half on with one shuffle per word: half the Philox work of stream v1 (4 words -> 4 normals).
// Restated in oracle/philox_ref.py (box_muller16); scr_debug_noise dumps it.
// ------------------------------------------------------------------------------------------
#ifndef SCR_NOISE16
#define SCR_NOISE16 1   // 0: stream v1 (one call per lane and cell, 23-bit radius / angle words) -- A/B builds only
#endif
constexpr float kBm16AngleScale = 9.58738019107841e-05f;           // float(2 pi / 65536)
constexpr float kBm16AngleBias = -807.3892822265625f;              // float(-(2^23 * double(kBm16AngleScale)) - pi + pi / 65536)
__device__ __forceinline__ void bm16_unpack(uint32_t w, float& u1, float& th) {
  const float fk = __uint_as_float((w >> 16) + 0x4b000000u);
  const float fm = __uint_as_float((w & 0xffffu) | 0x4b000000u);
  u1 = fmaf(fk, -1.52587890625e-05f, 129.0f);
  th = fmaf(fm, kBm16AngleScale, kBm16AngleBias);
}
// radius and (cos, sin) of one word; rscale = -2 ln(2) sigma^2
__device__ __forceinline__ void box_muller16_polar(uint32_t w, float rscale, float& r, float& c, float& sn) {
  float u1, th;
  bm16_unpack(w, u1, th);
  r = ptx_sqrt(fabsf(ptx_lg2(u1) * rscale));
  c = ptx_cos(th);
  sn = ptx_sin(th);
}
// two N(0, sigma^2) values from one word
__device__ __forceinline__ void box_muller16(uint32_t w, float rscale, float& n_cos, float& n_sin) {
  float r, c, sn;
  box_muller16_polar(w, rscale, r, c, sn);
  n_cos = r * c;
  n_sin = r * sn;
}

// accurate tanh.  fp32: 1 - 2/(1+e^{2z}) with ex2/rcp approximations, evaluated on z itself (no |z| / copysign pair:
// the form is exact in the limits -- ex2 -> +inf gives rcp -> 0 -> +1, ex2 -> 0 gives -1 -- and its absolute error is
// ~2.5e-7 on either side, entering the state scaled by dt; DESIGN.md "Precision").  SCR_TANH_SYMMETRIC=1 restores the
// odd-symmetric |z| form (one LOP3 more per element).  fp64: CUDA libm.
#ifndef SCR_TANH_SYMMETRIC
#define SCR_TANH_SYMMETRIC 0
#endif
__device__ __forceinline__ float tanh_acc(float z) {
  if (SCR_TANH_SYMMETRIC) {
    const float e = ptx_ex2(fabsf(z) * 2.88539008177792681472f);  // 2*log2(e)
    return copysignf(fmaf(-2.0f, ptx_rcp(e + 1.0f), 1.0f), z);
  }
  const float e = ptx_ex2(z * 2.88539008177792681472f);
  return fmaf(-2.0f, ptx_rcp(e + 1.0f), 1.0f);
}
__device__ __forceinline__ double tanh_acc(double z) { return tanh(z); }

// drive + r*u.  fp32: one FFMA.  fp64: the float product first, so the value added is exactly the
// float the noise dump (scr_debug_noise) reports.
__device__ __forceinline__ float add_noise(float drive, float r, float u) { return fmaf(r, u, drive); }
__device__ __forceinline__ double add_noise(double drive, float r, float u) { return drive + static_cast<double>(r * u); }

// One element of developCell (MS:529-535): x' = proj * (x + dt (tanh z - x)) + const.
// fp32: with r = 1/(1 + 2^m), m = K z, tanh z = 1 - 2 r and the Euler blend folds into two FMAs,
//   x + dt (tanh z - x) = fma(r, -2 dt, fma(x, 1 - dt, dt)),
// the second of which does not wait for the MUFU chain.  Exact in the limits (2^m -> inf gives r -> 0, 2^m -> 0 gives
// r = 1); absolute error of the tanh part ~2.5e-7 on either side, entering the state scaled by dt (DESIGN.md
// "Precision").  Every float32 path of the sparse stepper (packed fast path, ragged tail, probe, supplied noise) goes
// through this one formula, so results do not depend on how cells are grouped.
__device__ __forceinline__ float next_state(float m, float x, float dt, float proj, float cnst) {
  const float r = ptx_rcp(ptx_ex2(m) + 1.0f);
  return fmaf(proj, fmaf(r, -2.0f * dt, fmaf(x, 1.0f - dt, dt)), cnst);
}
// fp64: libm tanh on the unscaled drive, the reference's order of operations
__device__ __forceinline__ double next_state(double z, double x, double dt, double proj, double cnst) {
  return fma(proj, fma(dt, tanh(z) - x, x), cnst);
}
// supplied noise arrives in the reference's units: scale it like the drive
__device__ __forceinline__ float add_supplied(float drive, float nz) { return fmaf(nz, 2.88539008177792681472f, drive); }
__device__ __forceinline__ double add_supplied(double drive, double nz) { return drive + nz; }


// ------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot, two IEEE fp32 results,
// bit-identical to the scalar instructions.  Pairs are (cell c, cell c+1) of one gene row, which sit
// in adjacent registers after the LDS.128 of the state row.
// ------------------------------------------------------------------------------------------
#ifndef SCR_PACKED_F32X2
#define SCR_PACKED_F32X2 1
#endif
#ifndef SCR_SPLIT_BARRIER
#define SCR_SPLIT_BARRIER 1
#endif
#ifndef SCR_CHUNK_UNROLL
#define SCR_CHUNK_UNROLL 1    // long-row chunks of the default length (8) are evaluated with all loads in flight
#endif
#ifndef SCR_EARLY_READ_ARRIVE
// A/B candidate for the next round (default off, bit-identical by construction): the "reads of the old buffer finished"
// arrival is made right after the last block's loads instead of at the end of the step, i.e. before that block's
// Philox / Box-Muller / tanh work.  Arrivals at step ends are what wakes the warps already parked on the next step's
// barrier (profiles/r1_final_step_kernel_ncu_summary.txt: 33 wake-ups per warp and step).
#define SCR_EARLY_READ_ARRIVE 0
#endif
#ifndef SCR_LAST_ARRIVER
// A/B candidate: the warps of a group count their step-barrier arrivals in a shared-memory word and only the LAST one
// performs the mbarrier arrival (barrier count 1), so a group makes 3 mbarrier events per step instead of ~18 -- every
// mbarrier arrival wakes all warps parked in a try_wait on the SM (profiles/r1_final_step_kernel_ncu_summary.txt).
#define SCR_LAST_ARRIVER 0
#endif
#ifndef SCR_MERGE_RARE
#define SCR_MERGE_RARE 1
#endif
#ifndef SCR_CHUNK_ARRIVE_ALL
#define SCR_CHUNK_ARRIVE_ALL 0   // 1: every warp of the group arrives on the chunk barrier (the earlier behaviour)
#endif
#ifndef SCR_CHUNK_SPREAD
#define SCR_CHUNK_SPREAD 0    // experiment: chunk v evaluated by lane v / wq of warp v % wq (measured slower)
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Box-Muller for the same word pair of two cells at once: radius pair and (cos, sin) pairs
__device__ __forceinline__ void box_muller_polar2(uint32_t wa0, uint32_t wb0, uint32_t wa1, uint32_t wb1, f32x2 rscale2,
                                                  f32x2& r, f32x2& c, f32x2& sn) {
  const f32x2 f1 = pk2(__uint_as_float(0x3f800000u | (wa0 >> 9)), __uint_as_float(0x3f800000u | (wa1 >> 9)));
  const f32x2 f2 = pk2(__uint_as_float(0x3f800000u | (wb0 >> 9)), __uint_as_float(0x3f800000u | (wb1 >> 9)));
  float u0, u1;
  upk2(fma2(f1, pk2(-1.0f, -1.0f), pk2(2.0f, 2.0f)), u0, u1);                       // 2 - f1 in (0, 1]
  float t0, t1;
  upk2(mul2(pk2(ptx_lg2(u0), ptx_lg2(u1)), rscale2), t0, t1);
  r = pk2(ptx_sqrt(fabsf(t0)), ptx_sqrt(fabsf(t1)));
  float h0, h1;
  upk2(fma2(f2, pk2(6.28318530717958647692f, 6.28318530717958647692f), pk2(-9.42477796076937971538f, -9.42477796076937971538f)), h0, h1);
  c = pk2(ptx_cos(h0), ptx_cos(h1));
  sn = pk2(ptx_sin(h0), ptx_sin(h1));
}
// stream v2: the same word position of two cells at once -> radius pair and (cos, sin) pairs
__device__ __forceinline__ void box_muller16_polar2(uint32_t w0, uint32_t w1, f32x2 rscale2, f32x2& r, f32x2& c, f32x2& sn) {
  const f32x2 fk = pk2(__uint_as_float((w0 >> 16) + 0x4b000000u), __uint_as_float((w1 >> 16) + 0x4b000000u));
  const f32x2 fm = pk2(__uint_as_float((w0 & 0xffffu) | 0x4b000000u), __uint_as_float((w1 & 0xffffu) | 0x4b000000u));
  float u0, u1;
  upk2(fma2(fk, pk2(-1.52587890625e-05f, -1.52587890625e-05f), pk2(129.0f, 129.0f)), u0, u1);   // 1 - k / 65536 in (0, 1]
  float t0, t1;
  upk2(mul2(pk2(ptx_lg2(u0), ptx_lg2(u1)), rscale2), t0, t1);
  r = pk2(ptx_sqrt(fabsf(t0)), ptx_sqrt(fabsf(t1)));
  float h0, h1;
  upk2(fma2(fm, pk2(kBm16AngleScale, kBm16AngleScale), pk2(kBm16AngleBias, kBm16AngleBias)), h0, h1);
  c = pk2(ptx_cos(h0), ptx_cos(h1));
  sn = pk2(ptx_sin(h0), ptx_sin(h1));
}
// ------------------------------------------------------------------------------------------
// types
// ------------------------------------------------------------------------------------------
template <typename T> struct Traits;
template <> struct Traits<float> {
  static constexpr int CPG = 4;
  // the float32 stepper carries the drive as m = K z, K = 2 log2(e), the argument of ex2 in tanh z = 1 - 2/(1 + 2^m):
  // operator entries, the mean-level bias and the noise radius are scaled by K on the host (scr_api.cu)
  static constexpr double kDriveScale = 2.88539008177792681472;
  struct __align__(8) WEnt { float v; int off; };
  using PC = float2;
};
template <> struct Traits<double> {
  static constexpr int CPG = 2;
  static constexpr double kDriveScale = 1.0;
  struct __align__(16) WEnt { double v; int off; int pad; };
  using PC = double2;
};

// 16-byte row of one gene across the group's cells
template <typename T> struct __align__(16) Row { T v[16 / sizeof(T)]; };


struct __align__(16) Item {   // one cell group
  int cell[4];                // local slots, -1 = empty lane of the group
  int steps[4];               // iterations to run (0 for empty)
  uint32_t base[4];           // iterations already done (noise key)
  int row[4];                 // position in the caller's list (supplied noise / probe outputs)
  double scale;               // PROBE: multiplies W (1/alpha ladder, MatriSeq.py:461)
  int flags;                  // kItemParity: slots 0, 1 hold even and slots 2, 3 odd GLOBAL cell ids (fp64: slot 0 / 1)
  int pad_;
};
constexpr int kItemParity = 1;

template <typename T> struct StepParams {
  // packed operator (global copies; staged to smem by the kernel)
  const unsigned char* tab;   // small tables blob
  const unsigned char* wblob; // tile entries, then long-row chunk entries
  int tab_bytes, w_bytes;     // multiples of 16
  int w_prefix;               // leading bytes of wblob staged in shared memory when the whole blob does not fit (multiple of 128)
  int off_vrow;               // byte offset of chunk entries inside wblob
  int n, n_pad, npp, nb, npp_blocks, nlong_blocks, nv, nv_pad, lv;
  int wq, wq_log2, q;
  // per-phase
  const unsigned char* pc;    // n_pad x PC (proj, const + mu), internal order
  const T* bias;              // n_pad (-mu * rowsum) or nullptr
  T* state;                   // [cells][n_pad]
  int64_t ld;
  const Item* items;
  int nitems;
  int* counter;
  T dt;
  float rscale;               // -2 ln2 sigma^2
  RoundKeys rk;               // Philox round keys of (seed, stream)
  int64_t cell_offset;
  // supplied noise (parity mode): [row][noise_steps][n] in ORIGINAL gene order
  const T* noise;
  int64_t noise_steps;
  const int* perm;            // internal -> original gene (or -1)
  unsigned char* gscratch;    // kModeGlobal: per (CTA, group) state buffers, (n_pad + npp + nv_pad) * 16 bytes each
  int* dbg;                   // SCR_DEBUG_CHECKS builds: code of the first failed bounds check (0 = none)
  int64_t ncells, nrows;      // ... what the indices are checked against: cells of the context, rows of the caller's list
  // probe
  double thresh;
  int max_iter, avg_num;
  int* iters_out;
  double* nd;                 // [k][max_iter]
};


// element-wise update of one 128-gene block for a fully active group of 4 float cells, packed fp32x2: next_state()
// on cell pairs, instruction for instruction (FFMA2 / FADD2 are per-lane IEEE, so the results are bit-identical)
#if SCR_NOISE16
constexpr int kWordsPerCell = 2;   // words a lane consumes per cell and 128-gene block
#else
constexpr int kWordsPerCell = 4;
#endif
__device__ __forceinline__ void fast_update_f32x2(const uint32_t (&w)[4][kWordsPerCell], float rscale, float dt, const float (&acc)[4][4],
                                                  Row<float> (&xo)[4], const float2 (&pcv)[4]) {
  const float omdt = 1.0f - dt, m2dt = -2.0f * dt;
  const f32x2 rscale2 = pk2(rscale, rscale), dt2 = pk2(dt, dt), omdt2 = pk2(omdt, omdt), m2dt2 = pk2(m2dt, m2dt);
  const f32x2 one2 = pk2(1.0f, 1.0f);
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {          // cell pairs (0,1) and (2,3)
    const int c0 = 2 * pr, c1 = 2 * pr + 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {           // word pairs -> genes (2h, 2h+1)
      f32x2 r, cs, sn;
#if SCR_NOISE16
      box_muller16_polar2(w[c0][h], w[c1][h], rscale2, r, cs, sn);
#else
      box_muller_polar2(w[c0][2 * h], w[c0][2 * h + 1], w[c1][2 * h], w[c1][2 * h + 1], rscale2, r, cs, sn);
#endif
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * h + e;
        float m0, m1, s0, s1;
        upk2(fma2(r, e == 0 ? cs : sn, pk2(acc[j][c0], acc[j][c1])), m0, m1);          // m = K (W x) + K sigma xi
        upk2(add2(pk2(ptx_ex2(m0), ptx_ex2(m1)), one2), s0, s1);
        const f32x2 u = fma2(pk2(xo[j].v[c0], xo[j].v[c1]), omdt2, dt2);               // (1 - dt) x + dt
        const f32x2 v = fma2(pk2(ptx_rcp(s0), ptx_rcp(s1)), m2dt2, u);                 // ... - 2 dt / (1 + 2^m)
        upk2(fma2(pk2(pcv[j].x, pcv[j].x), v, pk2(pcv[j].y, pcv[j].y)), xo[j].v[c0], xo[j].v[c1]);
      }
    }
  }
}
// (double contexts never take this path; the overload only keeps the template well-formed)
__device__ __forceinline__ void fast_update_f32x2(const uint32_t (&)[2][kWordsPerCell], float, double, const double (&)[4][2],
                                                  Row<double> (&)[4], const double2 (&)[4]) {}

// Shared-memory layout of a CTA (byte offsets from the start of dynamic shared memory):
//   [0, kRecBytes)                   BlockRec table: one 16-byte record per entry of the warps' block lists
//   [kRecBytes, kFixedTab)           wblk_ptr: list bounds per warp of a group
//   [kFixedTab, + n_pad * PC)        proj/const pairs of the node           (compile-time start: LDS [reg + imm])
//   then ovfidx (long-row chunk ranges), the staged head of the operator, and the cell groups:
//   group = [kGroupHdr: ctrl words, three mbarriers, CellRound0 words, probe partials][bufA][bufB][chunk partials]
// The tables a 128-gene block needs sit at compile-time offsets so that their addresses cost no instructions.
// Shared memory is the stepper's scarce resource: every byte the tables and headers take is a byte less of the operator's
// staged head, and on the bench network 2 KB less of it cost 20 % (profiles/r2_experiments.txt) -- so the on-chip modes keep
// the 128-record table and the probe's window only exists in PROBE kernels.
constexpr int kMaxBlocksOnChip = 128;           // n_pad <= 16384 with the groups in shared memory (they stop fitting near n = 7 000)
constexpr int kMaxBlocks = 256;                 // n_pad <= 32768 with the groups in global memory (kModeGlobal)
__host__ __device__ constexpr int rec_bytes(bool large) { return (large ? kMaxBlocks : kMaxBlocksOnChip) * 16; }
__host__ __device__ constexpr int fixed_tab(bool large) { return rec_bytes(large) + 128; }
// ctrl 16 | mbarriers 24 (+8 pad) | CellRound0 48 | steps 16 | cells 16 | step bases 16 | caller rows 16 | probe partials 16 x 16 |
// probe window: the last kProbeRing values of mean|dx|/dt per cell (float64)
constexpr int kProbeRing = 32;
constexpr int kHdrBar = 16, kHdrR0 = 48, kHdrSteps = 96, kHdrCells = 112, kHdrBase = 128, kHdrRows = 144, kHdrRed = 160,
              kHdrRing = kHdrRed + 256;
__host__ __device__ constexpr int group_hdr(bool probe) { return probe ? kHdrRing + 4 * kProbeRing * 8 : kHdrRing; }
// block record flags
constexpr int kBlkPP = 1, kBlkLong = 2, kBlkLastReg = 4, kBlkFirstPP = 8, kBlkBias = 16, kBlkFlagMask = 127;

__host__ __device__ inline size_t group_state_bytes(int n_pad, int npp, int nv_pad) {   // bufA | bufB | chunk partials
  return static_cast<size_t>(n_pad + npp + nv_pad) * 16;
}
__host__ __device__ inline size_t group_smem_bytes(int n_pad, int npp, int nv_pad, bool probe) {
  return group_state_bytes(n_pad, npp, nv_pad) + group_hdr(probe);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// MODE: where the operator and the cell groups live --
//   kModeL1     groups in shared memory, operator read through L1 (its head staged in shared memory when room is left);
//   kModeSmem   groups and the whole operator in shared memory;
//   kModeGlobal LARGE n (one cell group does not fit in shared memory: n above ~7 000 genes in float32): the groups' state
//               buffers live in a global scratch area (L1 / L2 resident) and proj/const are read from global memory; only
//               the tables and the group headers stay on chip.  Same code, same results, far slower -- the fallback that
//               lets the library take any n the reference takes (up to kMaxBlocks * 128 genes).
constexpr int kModeL1 = 0, kModeSmem = 1, kModeGlobal = 2;

template <typename T, int MODE, bool PROBE, bool SUPPLIED>
__global__ void __launch_bounds__(512, 1) step_kernel(const StepParams<T> p) {
  using TR = Traits<T>;
  using WEnt = typename TR::WEnt;
  using PC = typename TR::PC;
  using RowT = Row<T>;
  constexpr int CPG = TR::CPG;
  constexpr bool W_SMEM = MODE == kModeSmem, GSTATE = MODE == kModeGlobal;
  constexpr int kRecBytes = rec_bytes(GSTATE), kFixedTab = fixed_tab(GSTATE), kGroupHdr = group_hdr(PROBE);

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_pc = GSTATE ? const_cast<unsigned char*>(p.pc) : smem + kFixedTab;
  unsigned char* s_var = GSTATE ? smem + kFixedTab : s_pc + static_cast<size_t>(p.n_pad) * sizeof(PC);   // ovfidx
  const int var_bytes = p.tab_bytes - kFixedTab;
  unsigned char* s_w = s_var + var_bytes;
  const int wpre = W_SMEM ? p.w_bytes : (GSTATE ? 0 : p.w_prefix);   // bytes of the operator held in shared memory
  unsigned char* s_groups = s_w + wpre;
  const size_t gstate_bytes = group_state_bytes(p.n_pad, p.npp, p.nv_pad);
  const size_t gbytes = GSTATE ? static_cast<size_t>(kGroupHdr) : gstate_bytes + kGroupHdr;   // per group in shared memory
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_groups + static_cast<size_t>(p.q) * gbytes);

  const int warp = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
  const int grp = warp >> p.wq_log2, wq = warp & (p.wq - 1);
  const int qthreads = p.wq * 32, qtid = wq * 32 + lane_id, bar_id = 1 + grp;
  unsigned char* ghdr = s_groups + static_cast<size_t>(grp) * gbytes;

  // ---- stage operator tables and the node's proj/const with the bulk-copy engine ----
  if (threadIdx.x == 0) mbar_init(s_mbar, 1);
  if (SCR_SPLIT_BARRIER && !PROBE && qtid == 0) {
    uint64_t* b0 = reinterpret_cast<uint64_t*>(ghdr + kHdrBar);
    const int chunk_warps = (SCR_CHUNK_ARRIVE_ALL || SCR_CHUNK_SPREAD) ? p.wq : max(1, min(p.wq, p.nv_pad / 32));   // bar_a: one arrival per chunk-evaluating warp
    mbar_init(b0, SCR_LAST_ARRIVER ? 1 : p.wq);
    mbar_init(b0 + 1, SCR_LAST_ARRIVER ? 1 : p.wq);
    mbar_init(b0 + 2, SCR_LAST_ARRIVER ? 1 : chunk_warps);
    uint32_t* cnt0 = reinterpret_cast<uint32_t*>(b0 + 3);   // arrival counters of the three barriers (the third is ctrl word 2)
    cnt0[0] = cnt0[1] = 0;
    reinterpret_cast<volatile int*>(ghdr)[2] = 0;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t pc_bytes = static_cast<uint32_t>(p.n_pad * sizeof(PC));
    mbar_expect_tx(s_mbar, p.tab_bytes + (GSTATE ? 0u : pc_bytes) + wpre);
    bulk_g2s(smem, p.tab, kFixedTab, s_mbar);
    if (!GSTATE) bulk_g2s(s_pc, p.pc, pc_bytes, s_mbar);
    if (var_bytes > 0) bulk_g2s(s_var, p.tab + kFixedTab, var_bytes, s_mbar);
    if (wpre > 0) bulk_g2s(s_w, p.wblob, wpre, s_mbar);
  }
  mbar_wait(s_mbar, 0);

  // Block records arrive as {first gene, first slot, end slot, flags}; each CTA rewrites them once into the form the block
  // loop consumes: {first gene | flags, slot count, generic pointer to the block's first entry}.  The pointer already
  // says where the entries are read from (partial residency: blocks whose entries lie inside the staged prefix gather
  // from shared memory, the others from global memory through L1 -- generic loads serve both), and the launch's bias
  // switch becomes a block flag, so neither costs instructions per block.
#ifndef SCR_PC_ASM
#define SCR_PC_ASM 0   // 1: proj/const pairs through ld.shared on the record-table base (measured 0.5 % slower)
#endif
#ifndef SCR_PIN_LOOP_INVARIANTS
#define SCR_PIN_LOOP_INVARIANTS 0   // measured: -7 instructions per block, but 1 % slower (profiles/r1_stepper_experiments.txt)
#endif
  // The lane id and the shared-space address of the record table are used by every 128-gene block; ptxas prefers to
  // re-derive them from special registers (S2R + address arithmetic, five instructions) in every block instead of
  // keeping two registers alive.  A round trip through a volatile shared-memory slot (the still unused state buffer)
  // makes them plain register values.
  uint32_t s_rec = smem_u32(smem);                                           // BlockRec table (shared-space address)
  int lane = lane_id;
  if (SCR_PIN_LOOP_INVARIANTS) {
    volatile uint32_t* pin = reinterpret_cast<volatile uint32_t*>(ghdr + kGroupHdr) + qtid * 2;
    pin[0] = static_cast<uint32_t>(lane_id);
    pin[1] = s_rec;
    lane = static_cast<int>(pin[0]);
    s_rec = pin[1];
  }
  {
    const int nrec = *reinterpret_cast<const int*>(smem + kRecBytes + p.wq * 4);   // wblk_ptr[wq] = number of records
    const int wpre_bytes = W_SMEM ? p.w_bytes : p.w_prefix;
    for (int r = threadIdx.x; r < nrec; r += blockDim.x) {
      const int4 rec = lds_i4(s_rec + r * 16);
      const bool in_smem = W_SMEM || rec.z * static_cast<int>(sizeof(WEnt)) <= wpre_bytes;
      const unsigned long long ptr = reinterpret_cast<unsigned long long>(in_smem ? s_w : p.wblob) +
                                     static_cast<unsigned long long>(rec.y) * sizeof(WEnt);
      int4 out;
      out.x = rec.x | rec.w | (p.bias != nullptr ? kBlkBias : 0);
      out.y = rec.z - rec.y;
      out.z = static_cast<int>(static_cast<uint32_t>(ptr));
      out.w = static_cast<int>(static_cast<uint32_t>(ptr >> 32));
      *reinterpret_cast<int4*>(smem + r * 16) = out;
    }
    __syncthreads();
  }
  const int2* ovfidx = reinterpret_cast<const int2*>(s_var);
  const int* wblk_ptr = reinterpret_cast<const int*>(smem + kRecBytes);
  // the chunk table is read from shared memory when it lies inside the staged part of the operator
  const bool chunks_in_smem = W_SMEM || p.off_vrow + static_cast<int>(sizeof(WEnt)) * p.lv * p.nv_pad <= wpre;
  const WEnt* wchunk = reinterpret_cast<const WEnt*>((chunks_in_smem ? s_w : p.wblob) + p.off_vrow);

  volatile int* ctrl = reinterpret_cast<volatile int*>(ghdr);
  // split step barriers (non-probe runs): three mbarriers per group, one arrival per warp
  //   bar_w: this step's new regulator values are all written     (waited at the start of the next step)
  //   bar_r: every read of this step's `cur` buffer has finished   (waited before the next step's first regulator store)
  //   bar_a: the long-row chunk partials of this step are in place (waited by the owners of long rows)
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(ghdr + kHdrBar);
  uint64_t* bar_r = bar_w + 1;
  uint64_t* bar_a = bar_w + 2;
  uint32_t* cnt_w = reinterpret_cast<uint32_t*>(bar_w + 3);
  uint32_t* cnt_r = cnt_w + 1;
  uint32_t* cnt_a = reinterpret_cast<uint32_t*>(ghdr) + 2;
  const uint32_t n_chunk_warps = (SCR_CHUNK_ARRIVE_ALL || SCR_CHUNK_SPREAD) ? p.wq : max(1, min(p.wq, p.nv_pad / 32));
  uint32_t* cell_r0 = reinterpret_cast<uint32_t*>(ghdr + kHdrR0);   // [3][4] words: CellRound0 per cell
  // per-cell step counts and local slots of the current group: read back where they are needed (ragged tail, write-back)
  // instead of being carried in registers through the block loop
  int* hsteps = reinterpret_cast<int*>(ghdr + kHdrSteps);
  int* hcells = reinterpret_cast<int*>(ghdr + kHdrCells);
  uint32_t* hbase = reinterpret_cast<uint32_t*>(ghdr + kHdrBase);   // iterations done before this phase (Philox counter word c1 = base + s)
  int* hrows = reinterpret_cast<int*>(ghdr + kHdrRows);             // position of each cell in the caller's list
  RowT* red = reinterpret_cast<RowT*>(ghdr + kHdrRed);              // PROBE: per-warp partial sums of |dx|
  double* ring = reinterpret_cast<double*>(ghdr + kHdrRing);        // PROBE: [cell][kProbeRing] window of the convergence trace
  unsigned char* bufA = GSTATE ? p.gscratch + (static_cast<size_t>(blockIdx.x) * p.q + grp) * gstate_bytes : ghdr + kGroupHdr;
  unsigned char* bufB = bufA + static_cast<size_t>(p.n_pad) * 16;
  RowT* ovf = reinterpret_cast<RowT*>(bufB + static_cast<size_t>(p.npp) * 16);

  const int my_b0 = wblk_ptr[wq], my_b1 = wblk_ptr[wq + 1];
  // the warp's list is ordered [regulator blocks without long rows | blocks with long rows | the rest]
  // (operator_pack.h); a warp whose first block is not ping-ponged has no regulator blocks at all
  constexpr bool kSplit = SCR_SPLIT_BARRIER && !PROBE;
  const bool no_reg_blocks = my_b0 >= my_b1 || !(lds_i4(s_rec + my_b0 * 16).x & kBlkPP);
  uint32_t tg = 0;            // steps this group has completed so far (mbarrier phase counter)
  constexpr bool kPacked = !PROBE && !SUPPLIED && SCR_PACKED_F32X2 && sizeof(T) == 4;

  for (;;) {
    if (qtid == 0) ctrl[0] = atomicAdd(p.counter, 1);
    named_bar_sync(bar_id, qthreads);
    const int item_idx = ctrl[0];
    if (item_idx >= p.nitems) break;
    const Item it = p.items[item_idx];
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      SCR_CHECK(p.dbg, it.cell[c] >= -1 && it.cell[c] < p.ncells, 1);
      SCR_CHECK(p.dbg, it.cell[c] < 0 || (it.row[c] >= 0 && it.row[c] < p.nrows), 2);
    }

    // per-cell Philox round-0 invariants (published by the barrier that ends the state load below)
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      if (qtid == c) {
        const uint64_t gid = static_cast<uint64_t>(p.cell_offset + (it.cell[c] >= 0 ? it.cell[c] : 0));
        const CellRound0 r0 = philox_cell_round0(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), p.rk);
        cell_r0[c] = r0.hi_p1_k0;
        cell_r0[4 + c] = r0.lo_p1;
        cell_r0[8 + c] = r0.c3_k1;
        hsteps[c] = it.steps[c];
        hcells[c] = it.cell[c];
        hbase[c] = it.base[c];
        hrows[c] = it.row[c];
      }
    }

    // ---- load the group's states: CPG cell rows -> interleaved x[gene][cell] ----
    for (int g0 = qtid * CPG; g0 < p.n_pad; g0 += qthreads * CPG) {
      RowT r[CPG];
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        if (it.cell[c] >= 0) {
          r[c] = *reinterpret_cast<const RowT*>(p.state + static_cast<int64_t>(it.cell[c]) * p.ld + g0);
        } else {
#pragma unroll
          for (int jj = 0; jj < CPG; ++jj) r[c].v[jj] = T(0);
        }
      }
#pragma unroll
      for (int jj = 0; jj < CPG; ++jj) {
        RowT o;
#pragma unroll
        for (int c = 0; c < CPG; ++c) o.v[c] = r[c].v[jj];
        *reinterpret_cast<RowT*>(bufA + static_cast<size_t>(g0 + jj) * 16) = o;
      }
    }
    named_bar_sync(bar_id, qthreads);

    unsigned char* cur = bufA;   // holds the current values of the ping-ponged genes
    unsigned char* oth = bufB;
    int smax = 0;
#pragma unroll
    for (int c = 0; c < CPG; ++c) smax = max(smax, it.steps[c]);
    int smin = 0x7fffffff;    // steps every cell of the group takes (0 if a slot is empty)
#pragma unroll
    for (int c = 0; c < CPG; ++c) smin = min(smin, it.steps[c]);
    constexpr uint32_t full_mask = (1u << CPG) - 1u;
    uint32_t alive = 0;       // PROBE: bit c set while cell c is still iterating
    int my_iters = 0;         // PROBE: iterations of cell `lane` (lanes 0 .. CPG-1 of the group's first warp keep the books)
    if (PROBE) {
      smax = p.max_iter;
#pragma unroll
      for (int c = 0; c < CPG; ++c) { if (it.cell[c] >= 0) alive |= 1u << c; }
    }
    const T wscale = PROBE ? static_cast<T>(it.scale) : T(1);

#if SCR_NOISE16
    // stream v2: the fast path runs the Philox calls of the two cells of this lane's parity (slots 2q, 2q + 1) and gets
    // the other two cells' words from its partner lane; it needs groups built with parity placement (Item::flags &
    // kItemParity), which is how scr_run_phase builds every group of a free-running float32 phase -- the only launches
    // that compile the fast path (kPacked)
    constexpr bool canonical = true;
    uint32_t sown[2];         // counter word c1 of the current step for the lane's own two cells
    sown[0] = (lane & 1) ? it.base[(2 % CPG)] : it.base[0];
    sown[1] = (lane & 1) ? it.base[(3 % CPG)] : it.base[1 % CPG];
#else
    constexpr bool canonical = true;
    uint32_t sctr[CPG];       // Philox counter word c1 of the current step: iterations done before this phase + s
#pragma unroll
    for (int c = 0; c < CPG; ++c) sctr[c] = it.base[c];
#endif
    for (int s = 0; s < smax; ++s) {
      if (kSplit && s > 0) mbar_wait(bar_w, (tg - 1u) & 1u);   // all of cur is in place
      // -------- phase A: chunks of rows longer than the cap --------
      if (p.nv > 0) {
        // the first nv_pad / 32 warps of the group evaluate the chunks (operator_pack.h charges them for it when it
        // assigns blocks); chunks of the default length run with all their loads in flight
        for (int v0 = 0; v0 < p.nv_pad; v0 += qthreads) {
          const int v = SCR_CHUNK_SPREAD ? v0 + lane * p.wq + wq : v0 + qtid;
          if (v >= p.nv_pad) continue;
          T acc[CPG];
#pragma unroll
          for (int c = 0; c < CPG; ++c) acc[c] = T(0);
          if (SCR_CHUNK_UNROLL && p.lv == 8) {
            WEnt e[8];
            RowT xv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = wchunk[k * p.nv_pad + v];
#pragma unroll
            for (int k = 0; k < 8; ++k) SCR_CHECK(p.dbg, e[k].off >= 0 && e[k].off < p.npp * 16 && (e[k].off & 15) == 0, 3);
#pragma unroll
            for (int k = 0; k < 8; ++k) xv[k] = *reinterpret_cast<const RowT*>(cur + e[k].off);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[c] = fma(e[k].v, xv[k].v[c], acc[c]);
            }
          } else {
            for (int k = 0; k < p.lv; ++k) {
              const WEnt e = wchunk[k * p.nv_pad + v];
              SCR_CHECK(p.dbg, e.off >= 0 && e.off < p.npp * 16 && (e.off & 15) == 0, 3);
              const RowT xv = *reinterpret_cast<const RowT*>(cur + e.off);
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[c] = fma(e.v, xv.v[c], acc[c]);
            }
          }
          RowT o;
#pragma unroll
          for (int c = 0; c < CPG; ++c) o.v[c] = acc[c];
          ovf[v] = o;
        }
        if (kSplit) {
          // only the warps that evaluate chunks (the first nv_pad / 32 of the group) arrive: fewer mbarrier events, and
          // every arrival wakes the warps suspended in a try_wait
          if (SCR_CHUNK_ARRIVE_ALL || SCR_CHUNK_SPREAD || wq * 32 < p.nv_pad) {
            __syncwarp();
            if (lane == 0) group_arrive(bar_a, cnt_a, tg, n_chunk_warps);
          }
        } else {
          named_bar_sync(bar_id, qthreads);
        }
      }
      if (kSplit && no_reg_blocks) {   // a warp without regulator blocks has nothing to publish
        __syncwarp();
        if (lane == 0) group_arrive(bar_w, cnt_w, tg, p.wq);
      }

      T dsum[CPG];
#pragma unroll
      for (int c = 0; c < CPG; ++c) dsum[c] = T(0);
      // -------- phase B: tiles + element-wise update, one 128-gene block at a time --------
      // all cells of the group active (the common case: groups are sorted by step count): no
      // per-cell branches, so the four Philox chains and sixteen tanh's interleave freely
      const bool all_active = PROBE ? (alive == full_mask) : (s < smin);
      for (int bi = my_b0; bi < my_b1; ++bi) {
        const int4 rec = lds_i4(s_rec + bi * 16);   // {first gene | flags, slots, pointer to the first entry}
        const int bflags = rec.x & kBlkFlagMask;
        const int gbase = (rec.x & ~kBlkFlagMask) | lane;   // first gene is a multiple of 128: one LOP3
        SCR_CHECK(p.dbg, gbase >= 0 && gbase + 96 < p.n_pad && rec.y >= 0 && rec.y <= 128 * 4096, 4);
        T acc[4][CPG];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int c = 0; c < CPG; ++c) acc[j][c] = T(0);
        }
        {
          const WEnt* wt = reinterpret_cast<const WEnt*>((static_cast<unsigned long long>(static_cast<uint32_t>(rec.w)) << 32) |
                                                         static_cast<uint32_t>(rec.z)) + lane;
          for (int o = lane; o < rec.y; o += 128, wt += 128) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const WEnt e = wt[j * 32];
              SCR_CHECK(p.dbg, e.off >= 0 && e.off < p.npp * 16 && (e.off & 15) == 0, 5);
              const RowT xv = *reinterpret_cast<const RowT*>(cur + e.off);
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[j][c] = fma(e.v, xv.v[c], acc[j][c]);
            }
          }
        }
#if SCR_MERGE_RARE
        if (bflags & (kBlkBias | kBlkLong)) {   // both rare: one test on the common path
#endif
        if (bflags & kBlkBias) {   // geneMeanLevels != 0 only: -mu * rowsum(W)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const T b0 = p.bias[gbase + j * 32];
#pragma unroll
            for (int c = 0; c < CPG; ++c) acc[j][c] += b0;
          }
        }
        if (bflags & kBlkLong) {
          if (kSplit) mbar_wait(bar_a, tg & 1u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int2 oi = ovfidx[gbase + j * 32];
            SCR_CHECK(p.dbg, oi.x >= 0 && oi.y >= 0 && oi.x + oi.y <= p.nv_pad, 6);
            for (int e = oi.x; e < oi.x + oi.y; ++e) {
              const RowT pv = ovf[e];
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[j][c] += pv.v[c];
            }
          }
        }
#if SCR_MERGE_RARE
        }
#endif
        const bool pp = bflags & kBlkPP;
        // regulator blocks read the current buffer and write the other one; the rest update in place
        const unsigned char* src = (pp ? cur : bufA) + static_cast<size_t>(gbase) * 16;
        unsigned char* dst = (pp ? oth : bufA) + static_cast<size_t>(gbase) * 16;
        RowT xo[4];
        PC pcv[4];
#if SCR_PC_ASM
        const uint32_t pc_addr = s_rec + kFixedTab + static_cast<uint32_t>(gbase) * static_cast<uint32_t>(sizeof(PC));
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xo[j] = *reinterpret_cast<const RowT*>(src + j * 512);
#if SCR_PC_ASM
          pcv[j] = lds_pc(pc_addr + j * 32 * static_cast<uint32_t>(sizeof(PC)), static_cast<PC*>(nullptr));
#else
          pcv[j] = reinterpret_cast<const PC*>(s_pc)[gbase + j * 32];
#endif
        }
        if (kSplit && SCR_EARLY_READ_ARRIVE && bi + 1 == my_b1) {   // the warp's last read of cur in this step is behind it
          __syncwarp();
          if (lane == 0) group_arrive(bar_r, cnt_r, tg, p.wq);
        }
#if SCR_NOISE16
        const uint32_t ctr0 = static_cast<uint32_t>(((rec.x >> 3) & ~15) | (lane >> 1));   // 16 * block + lane pair
#else
        const uint32_t ctr0 = static_cast<uint32_t>(((rec.x >> 2) & ~31) | lane);   // 32 * block + lane
#endif
        if (kPacked && all_active && canonical) {
          uint32_t w[CPG][kWordsPerCell];
          const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * ctr0;
#if SCR_NOISE16
          const bool odd = (lane & 1) != 0;
          const uint32_t* r0own = cell_r0 + (odd ? 2 : 0);
          const uint2 hp = *reinterpret_cast<const uint2*>(r0own);
          const uint2 lp = *reinterpret_cast<const uint2*>(r0own + 4);
          const uint2 ck = *reinterpret_cast<const uint2*>(r0own + 8);
          uint32_t own[2][4], got[2][2];
          philox4x32_10_cell(p0, sown[0], hp.x, lp.x, ck.x, p.rk, own[0]);
          philox4x32_10_cell(p0, sown[1], hp.y, lp.y, ck.y, p.rk, own[1]);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int h = 0; h < 2; ++h) got[e][h] = __shfl_xor_sync(0xffffffffu, own[e][2 + h], 1);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              w[e % CPG][h] = odd ? got[e][h] : own[e][h];              // slots 0, 1: even cell ids
              w[(2 + e) % CPG][h] = odd ? own[e][h] : got[e][h];        // slots 2, 3: odd cell ids
            }
          }
#else
          uint32_t hp[4], lp[4], ck[4];
          *reinterpret_cast<uint4*>(hp) = *reinterpret_cast<const uint4*>(cell_r0);
          *reinterpret_cast<uint4*>(lp) = *reinterpret_cast<const uint4*>(cell_r0 + 4);
          *reinterpret_cast<uint4*>(ck) = *reinterpret_cast<const uint4*>(cell_r0 + 8);
#pragma unroll
          for (int c = 0; c < CPG; ++c) philox4x32_10_cell(p0, sctr[c], hp[c % 4], lp[c % 4], ck[c % 4], p.rk, w[c]);
#endif
          fast_update_f32x2(w, p.rscale, p.dt, acc, xo, pcv);
        } else if (all_active && !SUPPLIED) {
          uint32_t w[CPG][4];
          const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * ctr0;
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
#if SCR_NOISE16
            philox4x32_10_cell(p0, hbase[c] + static_cast<uint32_t>(s), cell_r0[c], cell_r0[4 + c], cell_r0[8 + c], p.rk, w[c]);
            if ((lane ^ (static_cast<int>(p.cell_offset) + hcells[c])) & 1) { w[c][0] = w[c][2]; w[c][1] = w[c][3]; }   // partner half
#else
            philox4x32_10_cell(p0, sctr[c], cell_r0[c], cell_r0[4 + c], cell_r0[8 + c], p.rk, w[c]);
#endif
          }
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
            float r0, c0, s0, r1, c1, s1;
#if SCR_NOISE16
            box_muller16_polar(w[c][0], p.rscale, r0, c0, s0);
            box_muller16_polar(w[c][1], p.rscale, r1, c1, s1);
#else
            box_muller_polar(w[c][0], w[c][1], p.rscale, r0, c0, s0);
            box_muller_polar(w[c][2], w[c][3], p.rscale, r1, c1, s1);
#endif
            const float rr[4] = {r0, r0, r1, r1};
            const float uu[4] = {c0, s0, c1, s1};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const T x = xo[j].v[c];
              T drive = acc[j][c];
              if (PROBE) drive *= wscale;
              const T xn = next_state(add_noise(drive, rr[j], uu[j]), x, p.dt, pcv[j].x, pcv[j].y);
              if (PROBE) dsum[c] += fabs(xn - x);
              xo[j].v[c] = xn;
            }
          }
        } else {
          int og[4];
          if (SUPPLIED) {
#pragma unroll
            for (int j = 0; j < 4; ++j) og[j] = p.perm[gbase + j * 32];
          }
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
            const bool active = PROBE ? ((alive >> c) & 1u) : (s < hsteps[c]);
            if (active) {   // uniform over the group's warps
              T nz[4];
              float rr[4] = {0.f, 0.f, 0.f, 0.f}, uu[4] = {0.f, 0.f, 0.f, 0.f};
              if (SUPPLIED) {
                const int64_t nbase = (static_cast<int64_t>(it.row[c]) * p.noise_steps + s) * p.n;
#pragma unroll
                for (int j = 0; j < 4; ++j) SCR_CHECK(p.dbg, og[j] < p.n && s < p.noise_steps, 7);
#pragma unroll
                for (int j = 0; j < 4; ++j) nz[j] = og[j] >= 0 ? p.noise[nbase + og[j]] : T(0);
              } else {
                uint32_t w[4];
#if SCR_NOISE16
                philox4x32_10_cell(static_cast<uint64_t>(0xD2511F53u) * ctr0, hbase[c] + static_cast<uint32_t>(s), cell_r0[c], cell_r0[4 + c],
                                   cell_r0[8 + c], p.rk, w);
                const bool partner_half = ((lane ^ (static_cast<int>(p.cell_offset) + hcells[c])) & 1) != 0;
                box_muller16_polar(partner_half ? w[2] : w[0], p.rscale, rr[0], uu[0], uu[1]);
                box_muller16_polar(partner_half ? w[3] : w[1], p.rscale, rr[2], uu[2], uu[3]);
#else
                philox4x32_10_cell(static_cast<uint64_t>(0xD2511F53u) * ctr0, sctr[c], cell_r0[c], cell_r0[4 + c], cell_r0[8 + c], p.rk, w);
                box_muller_polar(w[0], w[1], p.rscale, rr[0], uu[0], uu[1]);
                box_muller_polar(w[2], w[3], p.rscale, rr[2], uu[2], uu[3]);
#endif
                rr[1] = rr[0];
                rr[3] = rr[2];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const T x = xo[j].v[c];
                T drive = acc[j][c];
                if (PROBE) drive *= wscale;
                const T z = SUPPLIED ? add_supplied(drive, nz[j]) : add_noise(drive, rr[j], uu[j]);
                const T xn = next_state(z, x, p.dt, pcv[j].x, pcv[j].y);
                if (PROBE) dsum[c] += fabs(xn - x);
                xo[j].v[c] = xn;
              }
            }
          }
        }
        // the first regulator store of a step overwrites the buffer the previous step read from
        if (kSplit && s > 0 && (bflags & kBlkFirstPP)) mbar_wait(bar_r, (tg - 1u) & 1u);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<RowT*>(dst + j * 512) = xo[j];
        if (kSplit && (bflags & kBlkLastReg)) {   // this warp's share of the next step's inputs is published
          __syncwarp();
          if (lane == 0) group_arrive(bar_w, cnt_w, tg, p.wq);
        }
      }

      if (PROBE) {
#pragma unroll
        for (int c = 0; c < CPG; ++c) {
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) dsum[c] += __shfl_xor_sync(0xffffffffu, dsum[c], sh);
        }
        if (lane == 0) {
          RowT o;
#pragma unroll
          for (int c = 0; c < CPG; ++c) o.v[c] = dsum[c];
          red[wq] = o;
        }
      }
      if (kSplit) {   // this warp no longer reads cur
        if (!SCR_EARLY_READ_ARRIVE || my_b0 >= my_b1) {
          __syncwarp();
          if (lane == 0) group_arrive(bar_r, cnt_r, tg, p.wq);
        }
        ++tg;
      } else {
        named_bar_sync(bar_id, qthreads);   // end of step: every new value is in place
      }
      { unsigned char* t = cur; cur = oth; oth = t; }
#if SCR_NOISE16
      ++sown[0];
      ++sown[1];
#else
#pragma unroll
      for (int c = 0; c < CPG; ++c) ++sctr[c];
#endif

      if (PROBE) {
        // stopping rule of MatriSeq.py:657-668, one lane per cell: mean of the last avg_num values of mean|dx|/dt, summed in
        // index order from a shared-memory window (the global trace is only written, for the caller)
        if (wq == 0) {
          bool still = false;
          if (lane < CPG && ((alive >> lane) & 1u)) {
            double tot = 0.0;
            for (int w = 0; w < p.wq; ++w) tot += static_cast<double>(red[w].v[lane]);
            double* trace = p.nd + static_cast<int64_t>(hrows[lane]) * p.max_iter;
            const double val = tot / (static_cast<double>(p.n) * static_cast<double>(p.dt));
            trace[s] = val;
            double* win = ring + lane * kProbeRing;
            win[s & (kProbeRing - 1)] = val;
            my_iters = s + 1;
            still = true;
            if (s >= p.avg_num - 1) {
              double acc = 0.0;
              if (p.avg_num <= kProbeRing) {
                for (int i = s - p.avg_num + 1; i <= s; ++i) acc += win[i & (kProbeRing - 1)];
              } else {
                for (int i = s - p.avg_num + 1; i <= s; ++i) acc += trace[i];
              }
              if (acc / p.avg_num <= p.thresh) still = false;
            }
          }
          const uint32_t m = __ballot_sync(0xffffffffu, still);
          if (lane == 0) ctrl[1] = static_cast<int>(m);
        }
        named_bar_sync(bar_id, qthreads);
        alive = static_cast<uint32_t>(ctrl[1]);
        if (alive == 0) break;
      }
    }

    if (kSplit && smax > 0) {   // both complete = full barrier: every value of the last step is visible
      mbar_wait(bar_w, (tg - 1u) & 1u);
      mbar_wait(bar_r, (tg - 1u) & 1u);
      // with the early arrival a warp's in-place stores of its last (non-ping-ponged) block come after its arrival on
      // bar_r and are covered by neither barrier: close the gap before other threads read those rows
      if (SCR_EARLY_READ_ARRIVE) named_bar_sync(bar_id, qthreads);
    }
    if (PROBE) {
      if (wq == 0 && lane < CPG && hcells[lane] >= 0) p.iters_out[hrows[lane]] = my_iters;
    } else {
      // ---- write the group back: interleaved -> CPG cell rows ----
      for (int g0 = qtid * CPG; g0 < p.n_pad; g0 += qthreads * CPG) {
        const unsigned char* src = (g0 < p.npp) ? cur : bufA;
        RowT r[CPG];
#pragma unroll
        for (int jj = 0; jj < CPG; ++jj) {
          const RowT o = *reinterpret_cast<const RowT*>(src + static_cast<size_t>(g0 + jj) * 16);
#pragma unroll
          for (int c = 0; c < CPG; ++c) r[c].v[jj] = o.v[c];
        }
#pragma unroll
        for (int c = 0; c < CPG; ++c) {
          const int cell = hcells[c];
          if (cell >= 0) *reinterpret_cast<RowT*>(p.state + static_cast<int64_t>(cell) * p.ld + g0) = r[c];
        }
      }
    }
    named_bar_sync(bar_id, qthreads);   // smem is reused by the next group
  }
}

}  // namespace scr
