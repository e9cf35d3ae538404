// Dense-network stepper: developCell (reference MatriSeq.py:503-537) for a batch of cells as a split-precision tcgen05
// GEMM with the whole update fused into the epilogue (sm_100a).
//
//   Z[cell, g] = sum_k X[cell, k] * W[g, k]          (M = cells, N = K = genes, both operands K-major)
//   X'[cell, g] = proj[g] * ((1-dt) X + dt tanh(Z + bias[g] + noise)) + const[g]
//
// Three tensor-core products per K slice recover fp32-class accuracy from 11-bit operands (fp32 accumulate in TMEM):
//   X W^T ~= hi(X) hi(W)^T + hi(X) lo(W)^T + lo(X) hi(W)^T,      hi + lo carries ~22 significant bits of each operand
// Two operand encodings of the same scheme (template parameter KIND):
//   kF16  (default)  hi = fp16(s x), lo = fp16(s x - hi) with power-of-two scales (sx = 64 for the state: |x| < 1023;
//                    sw puts max|W| just below 2^15), tcgen05.mma.kind::f16 M128 N* K16 -- the same 11 significant
//                    bits per term as tf32 at twice the tensor rate and half the operand bytes; the accumulator is
//                    multiplied by 1 / (sx sw) (exact) in the epilogue.  lo terms below the fp16 normal range keep an
//                    absolute error <= 2^-25 / s, i.e. <= 2^-31 in x.  A state outside the range raises the context's
//                    error word (the call fails, it never saturates silently).
//   kTF32            hi = the fp32 word itself (the tensor core ignores the low 13 mantissa bits), lo = x - trunc(x),
//                    tcgen05.mma.kind::tf32 M128 N* K8: north_star's 3xTF32 (SCR_DENSE_KIND=tf32).
// A-hi / A-lo of the NEXT step are written by the epilogue of this one, B-hi / B-lo once per network.
//
// Structure (one CTA per SM, 128 x 256 (wide) or 128 x 64 (narrow) outputs per tile, K slices of 128 bytes per row):
//   warp 0      TMA producer: 4 operand tiles per K slice (SWIZZLE_128B), mbarrier ring (2 stages wide, 4 narrow)
//   warp 1      MMA issuer: 12 tcgen05.mma per K slice, TMEM alloc/dealloc
//   warps 2..9  epilogue: two warps per TMEM lane quarter (each half of the tile's columns), tcgen05.ld 16 columns at a
//               time, Philox noise, tanh, Euler, proj/const, stores X' and the split of X'; TMEM accumulators are
//               double-buffered so the epilogue of one (tile, step) overlaps the MMAs of the next
// Two schedules of the same code:
//   one step per launch   CTAs walk over the (cell tile, gene tile) pairs of the working set (large batches: config c4)
//   steps inside          when every tile has its own CTA (tiles <= SMs: config c1, every probe) the launch runs up to
//                         `nsteps` timesteps: the CTAs of one cell tile meet after each step on a counter in global
//                         memory (release add / acquire poll, cross-proxy fences around the TMA reads of what other
//                         CTAs wrote with ordinary stores); no launch per step, no host in the loop
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "step_kernel.cuh"

namespace scr {
namespace dense {

enum { kTF32 = 0, kF16 = 1 };
constexpr int BM = 128, BN = 256;        // BN: the wide output tile, also the padding unit of n
constexpr int ROW_BYTES = 128;           // K extent of one pipeline stage in bytes per operand row = one swizzle atom
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr float kStateScale = 64.f;      // sx of the kF16 encoding
constexpr float kStateLimit = 1023.f;    // |x| sx must stay below fp16's largest finite value
template <int KIND> struct Enc {
  static constexpr int ESIZE = KIND == kF16 ? 2 : 4;
  static constexpr int BK = ROW_BYTES / ESIZE;   // K elements per stage: 64 halves / 32 floats
  static constexpr int UK = 32 / ESIZE;          // K elements per MMA (32 bytes): 16 / 8
};
// Tile shapes.  Wide (128 x 256 outputs, 2-stage 96 KB ring): large cell batches, the tensor pipe is the bound (config c4).
// Narrow (128 x 64 outputs, 4-stage 48 KB ring): small batches -- with 1 000 cells and n = 500 (config c1) the wide shape makes
// only 16 tiles for 148 SMs; the narrow one makes 64.
template <int BN_> struct Tile {
  static constexpr int BN = BN_, STAGES = BN_ >= 256 ? 2 : 4;
  static constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BN_ * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN_;                      // two accumulators (power of two >= 32)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + BN_ * 8 /*pc tile*/ + 256 /*barriers*/;
};
constexpr int BN_NARROW = 64;
// Resident-B tile (128 x 32 outputs, steps-inside schedule only): the CTA's slice of W -- 32 genes x all of K, hi and lo --
// is loaded ONCE per launch and stays in shared memory; per step only the A tiles stream through a ring of `stages`
// 32 KB slots.  Small networks (n_pad <= 1024: config c1) run their whole phase this way: a third fewer bytes per step
// than the narrow tile, and 16 tiles per 512 genes spread the epilogue over twice as many SMs.
constexpr int BN_RES = 32;
constexpr int kMaxStages = 8;
constexpr int kResBarBytes = 256, kResAlign = 1024;
inline int res_b_bytes(int k_blocks) { return k_blocks * 2 * BN_RES * ROW_BYTES; }
inline int res_smem_bytes(int k_blocks, int stages) {
  return stages * 2 * BM * ROW_BYTES + res_b_bytes(k_blocks) + kResAlign + BN_RES * 8 + kResBarBytes;
}

struct Params {
  float* x[2];               // state working set [m_pad][ld], ping-pong by step parity (step s reads x[s & 1])
  void* hi[2];               // kF16: fp16 hi(sx x) [m_pad][ld]; kTF32: unused (A-hi is x itself)
  void* lo[2];               // kF16: fp16 lo; kTF32: fp32 lo
  int ld, m_tiles, n_tiles, k_blocks, n;
  const float2* pc;          // (proj, const + mu) per gene, padded with (0, 0)
  const float* bias;         // -mu * rowsum(W) or nullptr
  const int* row_steps;      // steps each working row takes in this phase (0 = padding row); rows sorted longest first
  const uint32_t* row_base;  // iterations already done (noise key)
  const int64_t* row_gid;    // global cell id (noise key)
  const int* row_src;        // index in the caller's list (supplied noise)
  int stages;                // resident-B tile: depth of the A ring
  int s0, nsteps;            // this launch computes steps s0 .. s0 + nsteps - 1 (nsteps > 1 only with one CTA per tile)
  unsigned* mt_count;        // steps-inside schedule: arrivals per cell tile (zeroed before the launch), else nullptr
  int* err;                  // context error word
  float dt, rscale, acc_scale;   // acc_scale = 1 / (sx sw) (kF16) or 1
  RoundKeys rk;
  const float* noise;        // supplied noise [row_src][noise_steps][n] or nullptr
  int64_t noise_steps;
  // convergence probe (MatriSeq.py:650-669): per-row divisor of the drive (alpha ladder) and the per-(column part, row)
  // partial sums of |x' - x| of each step, summed on the host in a fixed order
  const float* row_scale;    // nullptr = 1
  float* rowdiff;            // [nsteps][2 n_tiles][m_pad], or nullptr
  int m_pad;
  unsigned long long* trace; // SCR_DENSE_TRACE builds only
};
constexpr int kErrDenseRange = 901, kErrDenseStall = 902;
// developer build (-DSCR_DENSE_TRACE=1): CTA 0 stamps %globaltimer at eight points of every step of a steps-inside launch
// into Params::trace[(s - s0) * 8 + point]; scr_api.cu dumps them to the file named by the SCR_DENSE_TRACE variable
#ifndef SCR_DENSE_TRACE
#define SCR_DENSE_TRACE 0
#endif
__device__ __forceinline__ void trace_stamp(const struct Params& p, int s, int point);

__device__ __forceinline__ void trace_stamp(const Params& p, int s, int point) {
#if SCR_DENSE_TRACE
  if (blockIdx.x == 0 && p.trace && p.mt_count && s - p.s0 < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(s - p.s0) * 8 + point] = t;
  }
#else
  (void)p; (void)s; (void)point;
#endif
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// K-major operand tile, rows of 128 bytes = one swizzle atom, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>((8 * ROW_BYTES) >> 4) << 32;    // stride byte offset between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == kF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the fp16 split of one state value: (hi, lo) of sx * x, round to nearest
__device__ __forceinline__ void split_f16(float x, __half& h, __half& l) {
  const float xs = x * kStateScale;
  h = __float2half_rn(xs);
  l = __float2half_rn(xs - __half2float(h));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// steps of tile row block mt that THIS launch computes: [p.s0, step_end)
__device__ __forceinline__ int step_end(const Params& p, int mt) {
  return p.nsteps > 1 ? min(p.s0 + p.nsteps, p.row_steps[mt * BM]) : p.s0 + 1;
}

template <int BN_, int KIND, bool RESB = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_step_kernel(const __grid_constant__ CUtensorMap tm_ahi0, const __grid_constant__ CUtensorMap tm_ahi1,
                  const __grid_constant__ CUtensorMap tm_alo0, const __grid_constant__ CUtensorMap tm_alo1,
                  const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, const Params p) {
  using TL = Tile<BN_>;
  using EN = Enc<KIND>;
  static_assert(!RESB || BN_ == BN_RES, "the resident-B schedule uses the 128 x 32 tile");
  constexpr int BN = TL::BN, A_BYTES = TL::A_BYTES, B_BYTES = TL::B_BYTES;
  constexpr int STAGE_BYTES = RESB ? 2 * A_BYTES : TL::STAGE_BYTES;   // resident B: a ring slot holds A-hi and A-lo only
  constexpr int BK = EN::BK;
  const int STAGES = RESB ? p.stages : TL::STAGES;
  extern __shared__ unsigned char dsmem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* stage_mem = base;
  unsigned char* b_res = base + STAGES * STAGE_BYTES;                  // [k_blocks][hi | lo][BN rows x 128 bytes] (RESB)
  unsigned char* after = b_res + (RESB ? p.k_blocks * 2 * B_BYTES : 0);
  float2* pc_tile = reinterpret_cast<float2*>(after);
  uint64_t* bars = reinterpret_cast<uint64_t*>(after + BN * 8);
  uint64_t* full = bars;                      // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;        // [kMaxStages]
  uint64_t* tfull = bars + 2 * kMaxStages;    // [2]
  uint64_t* tempty = bars + 2 * kMaxStages + 2;   // [2]
  uint64_t* bfull = bars + 2 * kMaxStages + 4;    // resident B loaded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * EPI_WARPS); }
    mbar_init(bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: two accumulators of BN columns (all 512 columns for the wide tile)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TL::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      if constexpr (RESB) {   // this CTA's slice of W, once
        const int nt = blockIdx.x % p.n_tiles;
        mbar_expect_tx(bfull, static_cast<uint32_t>(p.k_blocks) * 2 * B_BYTES);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          tma_load_2d(b_res + kb * 2 * B_BYTES, &tm_bhi, kb * BK, nt * BN, bfull);
          tma_load_2d(b_res + kb * 2 * B_BYTES + B_BYTES, &tm_blo, kb * BK, nt * BN, bfull);
        }
      }
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = t / p.n_tiles, nt = t % p.n_tiles;
        const int s_end = step_end(p, mt);
        for (int s = p.s0; s < s_end; ++s) {
          const CUtensorMap* ta_hi = (s & 1) ? &tm_ahi1 : &tm_ahi0;
          const CUtensorMap* ta_lo = (s & 1) ? &tm_alo1 : &tm_alo0;
          int kb_b = 0;   // K slices whose B tiles are already in flight
          if (p.mt_count && s > p.s0) {
            // W does not depend on the other CTAs: the B tiles of the first ring stages go out before the wait
            int st2 = st;
            uint32_t ph2 = ph;
            for (; !RESB && kb_b < STAGES && kb_b < p.k_blocks; ++kb_b) {
              mbar_wait_parity(&empty[st2], ph2 ^ 1u);
              unsigned char* sm = stage_mem + st2 * STAGE_BYTES;
              mbar_expect_tx(&full[st2], STAGE_BYTES);
              tma_load_2d(sm + 2 * A_BYTES, &tm_bhi, kb_b * BK, nt * BN, &full[st2]);
              tma_load_2d(sm + 2 * A_BYTES + B_BYTES, &tm_blo, kb_b * BK, nt * BN, &full[st2]);
              if (++st2 == STAGES) { st2 = 0; ph2 ^= 1u; }
            }
            // every CTA of this cell tile has stored step s - 1 (and is done reading the buffer step s overwrites)
            const unsigned want = static_cast<unsigned>(p.n_tiles) * static_cast<unsigned>(s - p.s0);
            unsigned seen, spins = 0;
            unsigned long long t0 = 0, t1;
            for (;;) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.mt_count + mt) : "memory");
              if (seen >= want) break;
              if ((++spins & 1023u) == 0) {   // never hang the device: 4 s of wall clock, then the error word
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                if (t1 - t0 > 4000000000ULL) { atomicMax(p.err, kErrDenseStall); break; }
              }
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores of the other CTAs -> our TMA reads
          }
          trace_stamp(p, s, 0);
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            unsigned char* sm = stage_mem + st * STAGE_BYTES;
            if (kb >= kb_b) {
              mbar_wait_parity(&empty[st], ph ^ 1u);
              mbar_expect_tx(&full[st], STAGE_BYTES);
              if constexpr (!RESB) {
                tma_load_2d(sm + 2 * A_BYTES, &tm_bhi, kb * BK, nt * BN, &full[st]);
                tma_load_2d(sm + 2 * A_BYTES + B_BYTES, &tm_blo, kb * BK, nt * BN, &full[st]);
              }
            }
            tma_load_2d(sm, ta_hi, kb * BK, mt * BM, &full[st]);
            tma_load_2d(sm + A_BYTES, ta_lo, kb * BK, mt * BM, &full[st]);
            if (++st == STAGES) { st = 0; ph ^= 1u; }
          }
          trace_stamp(p, s, 1);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: D = f32 (bit 4), A / B format (bits 7, 10: f16 = 0, tf32 = 2), both K-major, N, M
    constexpr uint32_t fmt = KIND == kF16 ? 0u : 2u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);
    int st = 0, ab = 0;
    uint32_t ph = 0, aph = 0;
    if constexpr (RESB) mbar_wait_parity(bfull, 0u);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int s_end = step_end(p, t / p.n_tiles);
      for (int s = p.s0; s < s_end; ++s) {
        mbar_wait_parity(&tempty[ab], aph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + ab * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_parity(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0 && kb == 0) trace_stamp(p, s, 2);
          if (lane == 0) {
            const uint32_t sa = smem_u32(stage_mem + st * STAGE_BYTES);
            const uint64_t ahi = umma_desc(sa), alo = umma_desc(sa + A_BYTES);
            const uint32_t sb = RESB ? smem_u32(b_res + kb * 2 * B_BYTES) : sa + 2 * A_BYTES;
            const uint64_t bhi = umma_desc(sb), blo = umma_desc(sb + B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);   // 32 bytes along K inside the swizzle atom
              umma<KIND>(tmem_d, alo + adv, bhi + adv, idesc, (kb | k) != 0);   // small terms first
              umma<KIND>(tmem_d, ahi + adv, blo + adv, idesc, 1u);
              umma<KIND>(tmem_d, ahi + adv, bhi + adv, idesc, 1u);
            }
            umma_commit(&empty[st]);                       // frees the smem stage when these MMAs retire
            if (kb == p.k_blocks - 1) { umma_commit(&tfull[ab]); trace_stamp(p, s, 3); }   // accumulator complete
          }
          __syncwarp();
          if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
        if (++ab == 2) { ab = 0; aph ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w owns TMEM lanes 32 (w % 4) .. +31 and half of the tile's columns ----------------
    constexpr int ETHREADS = 32 * EPI_WARPS, PART = BN / 2, CHUNKS = PART / 16;
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int etid = threadIdx.x - 64;
    int ab = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      // stage this gene tile's (proj, const) pairs
      named_bar_sync(1, ETHREADS);             // previous tile's readers are done with pc_tile
      for (int i = etid; i < BN; i += ETHREADS) pc_tile[i] = p.pc[nt * BN + i];
      named_bar_sync(1, ETHREADS);
      const int row = mt * BM + row_in_tile;
      const int steps = p.row_steps[row];
      const uint64_t gid = static_cast<uint64_t>(p.row_gid[row]);
      const uint32_t c1_base = p.row_base[row];
      const int64_t roff = static_cast<int64_t>(row) * p.ld + nt * BN + part * PART;
      const float* nbase = p.noise ? p.noise + static_cast<int64_t>(p.row_src[row]) * p.noise_steps * p.n : nullptr;
      const float scale = p.row_scale ? p.row_scale[row] : 1.f;
      const int s_end = step_end(p, mt);
      // Narrow tile: a thread owns 32 columns of one row for the whole launch, so its slice of the state stays in registers
      // from step to step (no re-read; in the fp16 encoding the fp32 copy in global memory is only refreshed by the launch's
      // last step of the tile) and the step's noise is drawn BEFORE the accumulator is waited for -- both off the step-to-step critical path
      // of the steps-inside schedule.
      constexpr bool kRegState = PART <= 32;
      float xr[kRegState ? PART : 1];
      if constexpr (kRegState) {
        const float* xrow0 = p.x[p.s0 & 1] + roff;
#pragma unroll
        for (int v = 0; v < PART / 4; ++v) {
          const float4 q = *reinterpret_cast<const float4*>(xrow0 + v * 4);
          xr[4 * v] = q.x; xr[4 * v + 1] = q.y; xr[4 * v + 2] = q.z; xr[4 * v + 3] = q.w;
        }
      }
      for (int s = p.s0; s < s_end; ++s) {
        const bool active = s < steps;
        const uint32_t c1 = c1_base + static_cast<uint32_t>(s);
        const float* xrow = p.x[s & 1] + roff;
        float* orow = p.x[(s & 1) ^ 1] + roff;
        const float* nrow = nbase ? nbase + static_cast<int64_t>(s) * p.n : nullptr;
        const int gpart = nt * BN + part * PART;
        float dsum = 0.f, amax = 0.f;
        float nzr[kRegState ? PART : 1];
        if constexpr (kRegState) {
          if (active) {
#pragma unroll
            for (int v = 0; v < PART / 4; ++v) {
              if (nrow) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const int g = gpart + 4 * v + e; nzr[4 * v + e] = g < p.n ? nrow[g] : 0.f; }
              } else {
                uint32_t w[4];
                philox4x32_10_rk(static_cast<uint32_t>((gpart >> 2) + v), c1, static_cast<uint32_t>(gid),
                                 static_cast<uint32_t>(gid >> 32), p.rk, w);
                box_muller(w[0], w[1], p.rscale, nzr[4 * v], nzr[4 * v + 1]);
                box_muller(w[2], w[3], p.rscale, nzr[4 * v + 2], nzr[4 * v + 3]);
              }
            }
          }
        }

        mbar_wait_parity(&tfull[ab], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (etid == 0) trace_stamp(p, s, 4);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ab * BN + part * PART;
        const bool store_x = !kRegState || KIND == kTF32 /* x IS the A-hi operand */ || s == s_end - 1;
#pragma unroll
        for (int ch = 0; ch < (kRegState ? CHUNKS : 1); ++ch) {
#pragma unroll 1
          for (int ch2 = 0; ch2 < (kRegState ? 1 : CHUNKS); ++ch2) {
            const int c = kRegState ? ch : ch2;
            uint32_t acc[16];
            tmem_ld16(taddr + c * 16, acc);
            float xo[16], xn[16];
            if constexpr (kRegState) {
#pragma unroll
              for (int i = 0; i < 16; ++i) xo[i] = xr[ch * 16 + i];
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const float4 q = *reinterpret_cast<const float4*>(xrow + c * 16 + v * 4);
                xo[4 * v] = q.x; xo[4 * v + 1] = q.y; xo[4 * v + 2] = q.z; xo[4 * v + 3] = q.w;
              }
            }
            if (active) {
              const int g0 = gpart + c * 16;
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                float nz[4];
                if constexpr (kRegState) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) nz[e] = nzr[ch * 16 + 4 * v + e];
                } else if (nrow) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) { const int g = g0 + 4 * v + e; nz[e] = g < p.n ? nrow[g] : 0.f; }
                } else {
                  uint32_t w[4];
                  philox4x32_10_rk(static_cast<uint32_t>((g0 >> 2) + v), c1, static_cast<uint32_t>(gid),
                                   static_cast<uint32_t>(gid >> 32), p.rk, w);
                  box_muller(w[0], w[1], p.rscale, nz[0], nz[1]);
                  box_muller(w[2], w[3], p.rscale, nz[2], nz[3]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int i = 4 * v + e;
                  float drive = __uint_as_float(acc[i]);
                  if (KIND == kF16) drive *= p.acc_scale;
                  if (p.bias) drive += p.bias[g0 + i];
                  const float z = p.row_scale ? fmaf(drive, scale, nz[e]) : drive + nz[e];
                  const float tt = tanh_acc(z);
                  const float x1 = fmaf(p.dt, tt - xo[i], xo[i]);
                  const float2 pcv = pc_tile[part * PART + c * 16 + i];
                  xn[i] = fmaf(pcv.x, x1, pcv.y);
                  dsum += fabsf(xn[i] - xo[i]);
                  amax = fmaxf(amax, fabsf(xn[i]));
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) xn[i] = xo[i];   // finished (or padding) row: carried forward
            }
            if constexpr (kRegState) {
#pragma unroll
              for (int i = 0; i < 16; ++i) xr[ch * 16 + i] = xn[i];
            }
            if (store_x) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                *reinterpret_cast<float4*>(orow + c * 16 + v * 4) = make_float4(xn[4 * v], xn[4 * v + 1], xn[4 * v + 2], xn[4 * v + 3]);
            }
            if constexpr (KIND == kF16) {
              uint32_t hw[8], lw[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __half h0, l0, h1, l1;
                split_f16(xn[2 * i], h0, l0);
                split_f16(xn[2 * i + 1], h1, l1);
                hw[i] = pack_h2(h0, h1);
                lw[i] = pack_h2(l0, l1);
              }
              uint4* hrow = reinterpret_cast<uint4*>(static_cast<__half*>(p.hi[(s & 1) ^ 1]) + roff + c * 16);
              uint4* lrow = reinterpret_cast<uint4*>(static_cast<__half*>(p.lo[(s & 1) ^ 1]) + roff + c * 16);
              hrow[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); hrow[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
              lrow[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); lrow[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
            } else {
              float* lrow = static_cast<float*>(p.lo[(s & 1) ^ 1]) + roff + c * 16;
#pragma unroll
              for (int v = 0; v < 4; ++v)
                *reinterpret_cast<float4*>(lrow + v * 4) =
                    make_float4(tf32_lo(xn[4 * v]), tf32_lo(xn[4 * v + 1]), tf32_lo(xn[4 * v + 2]), tf32_lo(xn[4 * v + 3]));
            }
          }
        }
        if (KIND == kF16 && amax > kStateLimit) atomicMax(p.err, kErrDenseRange);
        if (p.rowdiff)
          p.rowdiff[(static_cast<int64_t>(s - p.s0) * (2 * p.n_tiles) + 2 * nt + part) * p.m_pad + row] = dsum;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&tempty[ab]);
        if (++ab == 2) { ab = 0; aph ^= 1u; }
        if (p.mt_count) {
          // publish this CTA's columns of step s to the other CTAs of the cell tile
          // (the CTA barrier orders every epilogue thread's stores before thread 0's release, which is cumulative)
          if (etid == 0) trace_stamp(p, s, 5);
          named_bar_sync(1, ETHREADS);
          if (etid == 0) {
            trace_stamp(p, s, 6);
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.mt_count + mt) : "memory");
            trace_stamp(p, s, 7);
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TL::TMEM_COLS) : "memory");
  }
}

// working-set gather: rows of the phase (sorted by step count) -> X (fp32) and its split; padding rows zero
template <int KIND>
__global__ void gather_rows_kernel(const float* __restrict__ state, int64_t ld_state, const int* __restrict__ row_cell,
                                   int m, float* __restrict__ x, void* __restrict__ hi, void* __restrict__ lo, int ld, int* err) {
  const int row = blockIdx.x;
  const int cell = row < m ? row_cell[row] : -1;
  float amax = 0.f;
  for (int g = threadIdx.x; g < ld; g += blockDim.x) {
    const float v = cell >= 0 ? state[static_cast<int64_t>(cell) * ld_state + g] : 0.f;
    const int64_t at = static_cast<int64_t>(row) * ld + g;
    x[at] = v;
    if constexpr (KIND == kF16) {
      __half h, l;
      split_f16(v, h, l);
      static_cast<__half*>(hi)[at] = h;
      static_cast<__half*>(lo)[at] = l;
      amax = fmaxf(amax, fabsf(v));
    } else {
      static_cast<float*>(lo)[at] = tf32_lo(v);
    }
  }
  if (KIND == kF16 && amax > kStateLimit) atomicMax(err, kErrDenseRange);
}
// scatter back: the buffer holding a row's final state depends on the parity of its tile's step count
__global__ void scatter_rows_kernel(float* __restrict__ state, int64_t ld_state, const int* __restrict__ row_cell, int m,
                                    const float* __restrict__ xa, const float* __restrict__ xb,
                                    const int* __restrict__ tile_steps, int ld) {
  const int row = blockIdx.x;
  if (row >= m) return;
  const float* src = (tile_steps[row / BM] & 1) ? xb : xa;
  const int cell = row_cell[row];
  for (int g = threadIdx.x; g < ld; g += blockDim.x)
    state[static_cast<int64_t>(cell) * ld_state + g] = src[static_cast<int64_t>(row) * ld + g];
}

// probe: per-(step, row) sum of the column parts of |x' - x| in index order, in float64 (what the host's stopping rule reads)
__global__ void rowdiff_sum_kernel(const float* __restrict__ parts, int n_parts, int m_pad, int nsteps, int k, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsteps * k) return;
  const int s = i / k, row = i % k;
  double tot = 0.0;
  for (int t = 0; t < n_parts; ++t) tot += static_cast<double>(parts[(static_cast<int64_t>(s) * n_parts + t) * m_pad + row]);
  out[i] = tot;
}

// noise of the dense path for (global cell, step): word v of counter (gene/4) -> gene
__global__ void debug_noise_dense_kernel(int n, uint32_t gid_lo, uint32_t gid_hi, uint32_t step, float rscale,
                                         RoundKeys rk, double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  uint32_t w[4];
  philox4x32_10_rk(static_cast<uint32_t>(q), step, gid_lo, gid_hi, rk, w);
  float a[4];
  box_muller(w[0], w[1], rscale, a[0], a[1]);
  box_muller(w[2], w[3], rscale, a[2], a[3]);
  for (int e = 0; e < 4; ++e)
    if (q * 4 + e < n) out[q * 4 + e] = static_cast<double>(a[e]);
}

}  // namespace dense
}  // namespace scr
