// Dense-network stepper: one timestep of developCell (reference MatriSeq.py:503-537) for a batch of
// cells as a tcgen05 3xTF32 GEMM with the whole update fused into the epilogue (sm_100a).
//
//   Z[cell, g] = sum_k X[cell, k] * W[g, k]          (M = cells, N = K = genes, both operands K-major)
//   X'[cell, g] = proj[g] * ((1-dt) X + dt tanh(Z + bias[g] + noise)) + const[g]
//
// 3xTF32: the tensor core reads fp32 words and ignores the low 13 mantissa bits, so with
// hi(x) = x truncated to tf32 and lo(x) = x - hi(x) (exact in fp32)
//   X W^T ~= hi(X) hi(W)^T + hi(X) lo(W)^T + lo(X) hi(W)^T        (fp32 accumulate in TMEM)
// A-hi is X itself (the hardware truncates), A-lo is written by the previous step's epilogue,
// B-hi / B-lo are prepared once per network.
//
// Structure (one CTA per SM, persistent over (cell-tile, gene-tile) pairs, 128 x 256 outputs each):
//   warp 0      TMA producer: 4 operand tiles per 32-wide k-block (SWIZZLE_128B), 2-stage mbarrier ring
//   warp 1      MMA issuer: 12 tcgen05.mma.kind::tf32 (M128 N256 K8) per k-block, TMEM alloc/dealloc
//   warps 2..5  epilogue: tcgen05.ld 16 columns at a time, Philox noise, tanh, Euler, proj/const,
//               stores X' and lo(X'); TMEM accumulators are double-buffered (2 x 256 columns) so the
//               epilogue of tile i overlaps the MMAs of tile i+1
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "step_kernel.cuh"

namespace scr {
namespace dense {

// K extent of one pipeline stage in floats: 32 (one 128-byte swizzle atom per row) or 16 (64-byte atoms: the same bytes in
// flight cut into twice as many stages, so the TMA producer runs further ahead of the MMA issuer)
#ifndef SCR_DENSE_BK
#define SCR_DENSE_BK 32
#endif
constexpr int BM = 128, BN = 256, BK = SCR_DENSE_BK;   // BN: the wide output tile, also the padding unit of n
static_assert(BK == 32 || BK == 16, "BK must be one 128-byte or one 64-byte swizzle atom");
constexpr int NUM_THREADS = 192;
// Tile shapes.  Wide (128 x 256 outputs, 2-stage 96 KB ring): large cell batches, the tensor pipe is the bound (config c4).
// Narrow (128 x 64 outputs, 4-stage 48 KB ring): small batches -- with 1 000 cells and n = 500 (config c1) the wide shape makes
// only 16 tiles for 148 SMs; the narrow one makes 64, each a quarter of the work, so a timestep's launch is ~3x shorter.
template <int BN_> struct Tile {
  static constexpr int BN = BN_, STAGES = (BN_ >= 256 ? 2 : 4) * (32 / BK);
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN_ * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN_;                      // two accumulators (power of two >= 32)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + BN_ * 8 /*pc tile*/ + 256 /*barriers*/;
};
constexpr int BN_NARROW = 64;

struct Params {
  const float* x_cur;        // [m_pad][ld]
  float* x_next;
  float* xlo_next;
  int ld, m_tiles, n_tiles, k_blocks, n;
  const float2* pc;          // (proj, const + mu) per gene, padded with (0, 0)
  const float* bias;         // -mu * rowsum(W) or nullptr
  const int* row_steps;      // steps each working row takes in this phase (0 = padding row)
  const uint32_t* row_base;  // iterations already done (noise key)
  const int64_t* row_gid;    // global cell id (noise key)
  const int* row_src;        // index in the caller's list (supplied noise)
  int s;                     // local step being computed
  float dt, rscale;
  RoundKeys rk;
  const float* noise;        // supplied noise [row_src][noise_steps][n] or nullptr
  int64_t noise_steps;
  // convergence probe (MatriSeq.py:650-669): per-row divisor of the drive (alpha ladder) and the
  // per-(gene tile, row) partial sums of |x' - x| of this step, summed on the host in a fixed order
  const float* row_scale;    // nullptr = 1
  float* rowdiff;            // [n_tiles][m_pad] for THIS step, or nullptr
  int m_pad;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// K-major operand tile, rows of BK floats = one swizzle atom (128 B or 64 B), 8-row groups 8 * row bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  constexpr uint64_t kRowBytes = BK * 4, kLayout = BK == 32 ? 2 : 4;   // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>((8 * kRowBytes) >> 4) << 32;    // stride byte offset between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
  d |= kLayout << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
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

template <int BN_>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_step_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                  const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, const Params p) {
  using TL = Tile<BN_>;
  constexpr int BN = TL::BN, STAGES = TL::STAGES, A_BYTES = TL::A_BYTES, B_BYTES = TL::B_BYTES, STAGE_BYTES = TL::STAGE_BYTES;
  extern __shared__ unsigned char dsmem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* stage_mem = base;
  float2* pc_tile = reinterpret_cast<float2*>(base + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES + BN * 8);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
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
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = t / p.n_tiles, nt = t % p.n_tiles;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_parity(&empty[st], ph ^ 1u);
          unsigned char* sm = stage_mem + st * STAGE_BYTES;
          mbar_expect_tx(&full[st], STAGE_BYTES);
          tma_load_2d(sm, &tm_ahi, kb * BK, mt * BM, &full[st]);
          tma_load_2d(sm + A_BYTES, &tm_alo, kb * BK, mt * BM, &full[st]);
          tma_load_2d(sm + 2 * A_BYTES, &tm_bhi, kb * BK, nt * BN, &full[st]);
          tma_load_2d(sm + 2 * A_BYTES + B_BYTES, &tm_blo, kb * BK, nt * BN, &full[st]);
          if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N = 256, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);
    int st = 0, ab = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait_parity(&tempty[ab], aph ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + ab * BN;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait_parity(&full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_mem + st * STAGE_BYTES);
          const uint64_t ahi = umma_desc(sa), alo = umma_desc(sa + A_BYTES);
          const uint64_t bhi = umma_desc(sa + 2 * A_BYTES), blo = umma_desc(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);   // 8 tf32 = 32 bytes along K inside the swizzle atom
            umma_tf32(tmem_d, alo + adv, bhi + adv, idesc, (kb | k) != 0);   // small terms first
            umma_tf32(tmem_d, ahi + adv, blo + adv, idesc, 1u);
            umma_tf32(tmem_d, ahi + adv, bhi + adv, idesc, 1u);
          }
          umma_commit(&empty[st]);                       // frees the smem stage when these MMAs retire
          if (kb == p.k_blocks - 1) umma_commit(&tfull[ab]);   // accumulator complete
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
      if (++ab == 2) { ab = 0; aph ^= 1u; }
    }
  } else {
    // ---------------- epilogue: 4 warps, warp (w % 4) owns TMEM lanes 32 (w % 4) .. +31 ----------------
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int etid = threadIdx.x - 64;    // 0..127
    int ab = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      // stage this gene tile's (proj, const) pairs
      named_bar_sync(1, 128);             // previous tile's readers are done with pc_tile
      for (int i = etid; i < BN; i += 128) pc_tile[i] = p.pc[nt * BN + i];
      named_bar_sync(1, 128);
      const int row = mt * BM + row_in_tile;
      const int steps = p.row_steps[row];
      const bool active = p.s < steps;
      const uint64_t gid = static_cast<uint64_t>(p.row_gid[row]);
      const uint32_t c1 = p.row_base[row] + static_cast<uint32_t>(p.s);
      const float* xrow = p.x_cur + static_cast<int64_t>(row) * p.ld + nt * BN;
      float* orow = p.x_next + static_cast<int64_t>(row) * p.ld + nt * BN;
      float* lrow = p.xlo_next + static_cast<int64_t>(row) * p.ld + nt * BN;
      const float* nrow = p.noise ? p.noise + (static_cast<int64_t>(p.row_src[row]) * p.noise_steps + p.s) * p.n : nullptr;
      const float scale = p.row_scale ? p.row_scale[row] : 1.f;
      float dsum = 0.f;

      mbar_wait_parity(&tfull[ab], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ab * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 16; ++ch) {
        uint32_t acc[16];
        tmem_ld16(taddr + ch * 16, acc);
        float xo[16], xn[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 q = *reinterpret_cast<const float4*>(xrow + ch * 16 + v * 4);
          xo[4 * v] = q.x; xo[4 * v + 1] = q.y; xo[4 * v + 2] = q.z; xo[4 * v + 3] = q.w;
        }
        if (active) {
          const int g0 = nt * BN + ch * 16;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            float nz[4];
            if (nrow) {
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
              if (p.bias) drive += p.bias[g0 + i];
              const float z = p.row_scale ? fmaf(drive, scale, nz[e]) : drive + nz[e];
              const float tt = tanh_acc(z);
              const float x1 = fmaf(p.dt, tt - xo[i], xo[i]);
              const float2 pcv = pc_tile[ch * 16 + i];
              xn[i] = fmaf(pcv.x, x1, pcv.y);
              dsum += fabsf(xn[i] - xo[i]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) xn[i] = xo[i];   // finished (or padding) row: carried forward
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float4 q, l;
          q.x = xn[4 * v]; q.y = xn[4 * v + 1]; q.z = xn[4 * v + 2]; q.w = xn[4 * v + 3];
          l.x = q.x - __uint_as_float(__float_as_uint(q.x) & 0xffffe000u);
          l.y = q.y - __uint_as_float(__float_as_uint(q.y) & 0xffffe000u);
          l.z = q.z - __uint_as_float(__float_as_uint(q.z) & 0xffffe000u);
          l.w = q.w - __uint_as_float(__float_as_uint(q.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(orow + ch * 16 + v * 4) = q;
          *reinterpret_cast<float4*>(lrow + ch * 16 + v * 4) = l;
        }
      }
      if (p.rowdiff) p.rowdiff[static_cast<int64_t>(nt) * p.m_pad + row] = dsum;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tempty[ab]);
      if (++ab == 2) { ab = 0; aph ^= 1u; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TL::TMEM_COLS) : "memory");
  }
}

// working-set gather: rows of the phase (sorted by step count) -> X (fp32) and lo(X); padding rows zero
__global__ void gather_rows_kernel(const float* __restrict__ state, int64_t ld_state, const int* __restrict__ row_cell,
                                   int m, float* __restrict__ x, float* __restrict__ xlo, int ld) {
  const int row = blockIdx.x;
  const int cell = row < m ? row_cell[row] : -1;
  for (int g = threadIdx.x; g < ld; g += blockDim.x) {
    const float v = cell >= 0 ? state[static_cast<int64_t>(cell) * ld_state + g] : 0.f;
    x[static_cast<int64_t>(row) * ld + g] = v;
    xlo[static_cast<int64_t>(row) * ld + g] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
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
