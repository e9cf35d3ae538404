// Dense stepper, CTA-pair schedule (large cell batches: config c4): two CTAs of one TPC form a cluster and compute a
// 256-cell x 256-gene tile with tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages only ITS half of both operands --
// its 128 cell rows of A-hi / A-lo and its 128 gene rows of B-hi / B-lo -- so a K slice costs 64 KB of L2 -> shared
// memory traffic per CTA instead of the 96 KB of the single-CTA 128 x 256 tile, and the ring holds three stages instead
// of two.  The single-CTA kernel runs into the L2 -> SM delivery rate (~8.4 TB/s over 148 SMs, profiles/r2_experiments.txt);
// this schedule moves a third fewer bytes for the same products.
//
// Roles per CTA (dense_kernel.cuh has the arithmetic; Params, encodings and the epilogue math are the same):
//   warp 0      TMA producer: its halves of the four operand tiles per K slice, completion counted on the LEADER's
//               (cluster rank 0) full barrier
//   warp 1      leader only: MMA issuer, 12 tcgen05.mma.cta_group::2 per K slice; tcgen05.commit multicast to both CTAs
//               frees the ring slot in both and publishes the accumulator to both epilogues.  Both CTAs' warp 1
//               allocate / free the pair's TMEM columns.
//   warps 2..9  epilogue of the CTA's own 128 rows (its TMEM lanes); the accumulator is handed back with one remote
//               arrival per warp on the leader's barrier
// Every wait is bounded: a protocol error raises the context's error word and the kernel still terminates.
#pragma once
#include "dense_kernel.cuh"

namespace scr {
namespace dense {

constexpr int PAIR_STAGES = 3;
constexpr int PAIR_B_ROWS = BN / 2;                                   // gene rows of B each CTA stages
constexpr int PAIR_A_BYTES = BM * ROW_BYTES, PAIR_B_BYTES = PAIR_B_ROWS * ROW_BYTES;
constexpr int PAIR_STAGE_BYTES = 2 * PAIR_A_BYTES + 2 * PAIR_B_BYTES;   // 64 KB
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + 1024 /*align*/ + BN * 8 /*pc tile*/ + 256 /*barriers*/;
constexpr int kErrDensePair = 903;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose bytes are counted on a barrier of the pair's leader CTA (given as a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  const uint32_t z = 0;
  if constexpr (KIND == kF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(z) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(z) : "memory");
  }
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// bounded wait: the phase, or 4 s of wall clock (%globaltimer; the suspend-time hint of try_wait is only a hint, so tries
// are not counted).  On a timeout the error word is raised, the CTA's flag is set and the caller leaves its role loop:
// nothing more is issued or arrived on (a stray arrival on a completed barrier would be a fault of its own).
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* err) {
  if (*abort_flag) return false;
  if (mbar_try_wait(bar, parity)) return true;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (mbar_try_wait(bar, parity)) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 4000000000ULL || *abort_flag) break;
  }
  *abort_flag = 1;
  atomicMax(err, kErrDensePair);
  return false;
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dense_pair_kernel(const __grid_constant__ CUtensorMap tm_ahi0, const __grid_constant__ CUtensorMap tm_ahi1,
                  const __grid_constant__ CUtensorMap tm_alo0, const __grid_constant__ CUtensorMap tm_alo1,
                  const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, const Params p) {
  using EN = Enc<KIND>;
  constexpr int BK = EN::BK, STAGES = PAIR_STAGES;
  extern __shared__ unsigned char dsmem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* stage_mem = base;
  float2* pc_tile = reinterpret_cast<float2*>(base + STAGES * PAIR_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * PAIR_STAGE_BYTES + BN * 8);
  uint64_t* full = bars;                  // [STAGES]  leader's: bytes of both CTAs
  uint64_t* empty = bars + STAGES;        // [STAGES]  own: slot released (commit multicast)
  uint64_t* tfull = bars + 2 * STAGES;    // [2]       own: accumulator complete (commit multicast)
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]   leader's: accumulator drained by both epilogues
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;          // 256-row blocks of the working set
  const int ntiles = m_pairs * p.n_tiles;            // pair tiles of 256 x 256 outputs
  const int s = p.s0;                                // this launch computes one timestep

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * EPI_WARPS); }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // both CTAs: the pair's TMEM columns (two accumulators of 256 columns in each CTA)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();     // barrier inits and the allocation are visible to the peer before any remote arrival / commit
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs: own halves) ----------------
    if (lane == 0) {
      const CUtensorMap* ta_hi = (s & 1) ? &tm_ahi1 : &tm_ahi0;
      const CUtensorMap* ta_lo = (s & 1) ? &tm_alo1 : &tm_alo0;
      int st = 0;
      uint32_t ph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        const int mp = t / p.n_tiles, nt = t % p.n_tiles;
        const int arow = mp * 2 * BM + static_cast<int>(rank) * BM;
        const int brow = nt * BN + static_cast<int>(rank) * PAIR_B_ROWS;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (!mbar_wait_bounded(&empty[st], ph ^ 1u, abort_flag, p.err)) { t = ntiles; break; }
          unsigned char* sm = stage_mem + st * PAIR_STAGE_BYTES;
          const uint32_t lbar = mapa_u32(smem_u32(&full[st]), 0);
          if (leader) mbar_expect_tx(&full[st], 2 * PAIR_STAGE_BYTES);
          tma_load_2d_pair(sm, ta_hi, kb * BK, arow, lbar);
          tma_load_2d_pair(sm + PAIR_A_BYTES, ta_lo, kb * BK, arow, lbar);
          tma_load_2d_pair(sm + 2 * PAIR_A_BYTES, &tm_bhi, kb * BK, brow, lbar);
          tma_load_2d_pair(sm + 2 * PAIR_A_BYTES + PAIR_B_BYTES, &tm_blo, kb * BK, brow, lbar);
          if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (leader) {
      constexpr uint32_t fmt = KIND == kF16 ? 0u : 2u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((BN >> 3) << 17) | (((2 * BM) >> 4) << 24);   // M = 256
      int st = 0, ab = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        if (!mbar_wait_bounded(&tempty[ab], aph ^ 1u, abort_flag, p.err)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + ab * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (!mbar_wait_bounded(&full[st], ph, abort_flag, p.err)) { t = ntiles; break; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t sa = smem_u32(stage_mem + st * PAIR_STAGE_BYTES);
            const uint64_t ahi = umma_desc(sa), alo = umma_desc(sa + PAIR_A_BYTES);
            const uint64_t bhi = umma_desc(sa + 2 * PAIR_A_BYTES), blo = umma_desc(sa + 2 * PAIR_A_BYTES + PAIR_B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);
              umma_pair<KIND>(tmem_d, alo + adv, bhi + adv, idesc, (kb | k) != 0);   // small terms first
              umma_pair<KIND>(tmem_d, ahi + adv, blo + adv, idesc, 1u);
              umma_pair<KIND>(tmem_d, ahi + adv, bhi + adv, idesc, 1u);
            }
            umma_commit_pair(&empty[st]);
            if (kb == p.k_blocks - 1) umma_commit_pair(&tfull[ab]);
          }
          __syncwarp();
          if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
        if (++ab == 2) { ab = 0; aph ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue: the CTA's own 128 rows of the pair tile ----------------
    constexpr int ETHREADS = 32 * EPI_WARPS, PART = BN / 2, CHUNKS = PART / 16;
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int etid = threadIdx.x - 64;
    int ab = 0;
    uint32_t aph = 0;
    for (int t = pair; t < ntiles; t += npairs) {
      const int mp = t / p.n_tiles, nt = t % p.n_tiles;
      named_bar_sync(1, ETHREADS);
      for (int i = etid; i < BN; i += ETHREADS) pc_tile[i] = p.pc[nt * BN + i];
      named_bar_sync(1, ETHREADS);
      const int row = mp * 2 * BM + static_cast<int>(rank) * BM + row_in_tile;
      const int steps = p.row_steps[row];
      const bool active = s < steps;
      const uint64_t gid = static_cast<uint64_t>(p.row_gid[row]);
      const uint32_t c1 = p.row_base[row] + static_cast<uint32_t>(s);
      const int64_t roff = static_cast<int64_t>(row) * p.ld + nt * BN + part * PART;
      const float* xrow = p.x[s & 1] + roff;
      float* orow = p.x[(s & 1) ^ 1] + roff;
      const float* nrow = p.noise ? p.noise + (static_cast<int64_t>(p.row_src[row]) * p.noise_steps + s) * p.n : nullptr;
      float amax = 0.f;

      mbar_wait_bounded(&tfull[ab], aph, abort_flag, p.err);   // (after a timeout the loop runs on without waiting or arriving:
                                                               //  every thread still meets the CTA's named barriers)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ab * BN + part * PART;
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS; ++ch) {
        uint32_t acc[16];
        tmem_ld16(taddr + ch * 16, acc);
        float xo[16], xn[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 q = *reinterpret_cast<const float4*>(xrow + ch * 16 + v * 4);
          xo[4 * v] = q.x; xo[4 * v + 1] = q.y; xo[4 * v + 2] = q.z; xo[4 * v + 3] = q.w;
        }
        if (active) {
          const int g0 = nt * BN + part * PART + ch * 16;
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
              if (KIND == kF16) drive *= p.acc_scale;
              if (p.bias) drive += p.bias[g0 + i];
              const float tt = tanh_acc(drive + nz[e]);
              const float x1 = fmaf(p.dt, tt - xo[i], xo[i]);
              const float2 pcv = pc_tile[part * PART + ch * 16 + i];
              xn[i] = fmaf(pcv.x, x1, pcv.y);
              amax = fmaxf(amax, fabsf(xn[i]));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) xn[i] = xo[i];   // finished (or padding) row: carried forward
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<float4*>(orow + ch * 16 + v * 4) = make_float4(xn[4 * v], xn[4 * v + 1], xn[4 * v + 2], xn[4 * v + 3]);
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
          uint4* hrow = reinterpret_cast<uint4*>(static_cast<__half*>(p.hi[(s & 1) ^ 1]) + roff + ch * 16);
          uint4* lrow = reinterpret_cast<uint4*>(static_cast<__half*>(p.lo[(s & 1) ^ 1]) + roff + ch * 16);
          hrow[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); hrow[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          lrow[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); lrow[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        } else {
          float* lrow = static_cast<float*>(p.lo[(s & 1) ^ 1]) + roff + ch * 16;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<float4*>(lrow + v * 4) =
                make_float4(tf32_lo(xn[4 * v]), tf32_lo(xn[4 * v + 1]), tf32_lo(xn[4 * v + 2]), tf32_lo(xn[4 * v + 3]));
        }
      }
      if (KIND == kF16 && amax > kStateLimit) atomicMax(p.err, kErrDenseRange);
      // hand the accumulator back: one arrival per warp on the leader's barrier
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0 && !*abort_flag) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[ab]), 0));
      if (++ab == 2) { ab = 0; aph ^= 1u; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();     // the peer may still be arriving on our barriers / the leader's MMAs read our shared memory
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

}  // namespace dense
}  // namespace scr
