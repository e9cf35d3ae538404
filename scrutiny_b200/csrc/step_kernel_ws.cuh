// Warp-specialised variant of the sparse stepper (float32, free-running noise, full 16+16 warp
// configuration).  Same state layout, operator and arithmetic as step_kernel (step_kernel.cuh);
// the difference is WHO does the work:
//
//   warps  0..15  "update" warps (2 groups x 8): blocked-ELL drive, tanh, Euler, projection
//   warps 16..31  "noise" warps: Philox4x32-10 + Box-Muller only; warp 16+i feeds update warp i
//                 through a private 2 KB shared-memory tile (128 genes x 4 cells) and a full/empty
//                 mbarrier pair
//
// The profile of the single-role kernel (profiles/r1_step_kernel_ncu_summary.txt) shows three
// co-limiters near 750-950 cycles per cell-step per SM (issue slots, MUFU, smem wavefronts) but only
// 55 % issue utilisation: with 128 registers per thread only 16 warps fit, too few to hide the
// fixed-latency dependency stalls.  The noise generator is ~45 % of the instructions, touches no
// state and can run ahead, so it moves to its own warps; `setmaxnreg` rebalances the register file
// (noise warps 48, update warps 80 registers per thread: 16*32*(48+80) = 64 K), doubling the
// resident warps per scheduler from 4 to 8.  The update warp reads its noise tile straight INTO the
// accumulators (z = noise + W x), so the hand-off costs four LDS.128 per block and no extra add.
//
// Bit-compatibility: the noise value for (cell, step, gene) is the same float as in step_kernel
// (same Philox counter layout, same Box-Muller); z = noise + sum_k w_k x_k is accumulated in a
// different order than step_kernel's fma(r, u, sum) -- float32 rounding differences only.
#pragma once
#include "step_kernel.cuh"

namespace scr {

constexpr int WS_NOISE_TILE_BYTES = 128 * 16;   // 128 genes x 4 cells x fp32

__host__ __device__ inline size_t ws_extra_smem(int update_warps) {
  return static_cast<size_t>(update_warps) * (WS_NOISE_TILE_BYTES + 16);
}

template <bool W_SMEM>
__global__ void __launch_bounds__(1024, 1) step_kernel_ws(const StepParams<float> p) {
  using T = float;
  using TR = Traits<float>;
  using WEnt = typename TR::WEnt;
  using PC = typename TR::PC;
  using RowT = Row<float>;
  constexpr int CPG = 4;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_tab = smem;
  unsigned char* s_pc = s_tab + p.tab_bytes;
  unsigned char* s_w = s_pc + static_cast<size_t>(p.n_pad) * sizeof(PC);
  unsigned char* s_groups = s_w + (W_SMEM ? p.w_bytes : 0);
  const size_t gbytes = group_smem_bytes(p.n_pad, p.npp, p.nv_pad);
  const int uwarps = p.q * p.wq;                               // update warps (== noise warps)
  unsigned char* s_noise = s_groups + static_cast<size_t>(p.q) * gbytes;
  uint64_t* s_pair = reinterpret_cast<uint64_t*>(s_noise + static_cast<size_t>(uwarps) * WS_NOISE_TILE_BYTES);   // [uwarps][2]
  uint64_t* s_mbar = s_pair + 2 * uwarps;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp >= uwarps;
  const int uw = producer ? warp - uwarps : warp;              // update-warp index this thread serves
  const int grp = uw >> p.wq_log2, wq = uw & (p.wq - 1);
  const int qthreads = p.wq * 32, qtid = wq * 32 + lane;
  const int bar_step = 1 + grp;                                // update warps of a group, per step
  const int bar_item = 1 + p.q + grp;                          // update + noise warps of a group, per item

  if (threadIdx.x == 0) {
    mbar_init(s_mbar, 1);
    for (int i = 0; i < 2 * uwarps; ++i) mbar_init(&s_pair[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t pc_bytes = static_cast<uint32_t>(p.n_pad * sizeof(PC));
    mbar_expect_tx(s_mbar, p.tab_bytes + pc_bytes + (W_SMEM ? p.w_bytes : 0));
    bulk_g2s(s_tab, p.tab, p.tab_bytes, s_mbar);
    bulk_g2s(s_pc, p.pc, pc_bytes, s_mbar);
    if (W_SMEM) bulk_g2s(s_w, p.wblob, p.w_bytes, s_mbar);
  }
  mbar_wait(s_mbar, 0);

  const int* blk_off = reinterpret_cast<const int*>(s_tab + p.off_slice);
  const int2* ovfidx = reinterpret_cast<const int2*>(s_tab + p.off_ovfidx);
  const int* wblk_ptr = reinterpret_cast<const int*>(s_tab + p.off_wblk_ptr);
  const int* wblk = reinterpret_cast<const int*>(s_tab + p.off_wblk);
  const int my_b0 = wblk_ptr[wq], my_b1 = wblk_ptr[wq + 1];

  unsigned char* bufA = s_groups + static_cast<size_t>(grp) * gbytes;
  unsigned char* bufB = bufA + static_cast<size_t>(p.n_pad) * 16;
  RowT* ovf = reinterpret_cast<RowT*>(bufB + static_cast<size_t>(p.npp) * 16);
  volatile int* ctrl = reinterpret_cast<volatile int*>(ovf + p.nv_pad + 8);
  RowT* tile = reinterpret_cast<RowT*>(s_noise + static_cast<size_t>(uw) * WS_NOISE_TILE_BYTES);   // [4][32]
  uint64_t* bar_full = &s_pair[2 * uw];
  uint64_t* bar_empty = &s_pair[2 * uw + 1];

  if (producer) {
    // =========================== noise warps ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    uint32_t ph_empty = 1;   // a fresh barrier lets the first wait on parity 1 through
    for (;;) {
      named_bar_sync(bar_item, 2 * qthreads);
      const int item_idx = ctrl[0];
      if (item_idx >= p.nitems) break;
      const Item it = p.items[item_idx];
      int smax = 0;
      uint32_t gid_lo[CPG], gid_hi[CPG];
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        smax = max(smax, it.steps[c]);
        const uint64_t gid = static_cast<uint64_t>(p.cell_offset + (it.cell[c] >= 0 ? it.cell[c] : 0));
        gid_lo[c] = static_cast<uint32_t>(gid);
        gid_hi[c] = static_cast<uint32_t>(gid >> 32);
      }
      for (int s = 0; s < smax; ++s) {
        for (int bi = my_b0; bi < my_b1; ++bi) {
          const uint32_t ctr0 = static_cast<uint32_t>(wblk[bi] * 32 + lane);
          float nz[CPG][4];
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
            if (s < it.steps[c] && !(p.dbg & 1)) {   // warp-uniform
              uint32_t w[4];
              philox4x32_10_rk(ctr0, it.base[c] + static_cast<uint32_t>(s), gid_lo[c], gid_hi[c], p.rk, w);
              box_muller(w[0], w[1], p.rscale, nz[c][0], nz[c][1]);
              box_muller(w[2], w[3], p.rscale, nz[c][2], nz[c][3]);
            } else {
              nz[c][0] = nz[c][1] = nz[c][2] = nz[c][3] = 0.f;
            }
          }
          mbar_wait(bar_empty, ph_empty);
          ph_empty ^= 1u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            RowT o;
#pragma unroll
            for (int c = 0; c < CPG; ++c) o.v[c] = nz[c][j];
            tile[j * 32 + lane] = o;
          }
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_full)) : "memory");
        }
      }
    }
    return;
  }

  // =========================== update warps ===========================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
  const unsigned char* wbase = W_SMEM ? s_w : p.wblob;
  const WEnt* wtile = reinterpret_cast<const WEnt*>(wbase);
  const WEnt* wchunk = reinterpret_cast<const WEnt*>(wbase + p.off_vrow);
  const PC* pcs = reinterpret_cast<const PC*>(s_pc);
  const bool has_bias = p.bias != nullptr;
  uint32_t ph_full = 0;

  for (;;) {
    if (qtid == 0) ctrl[0] = atomicAdd(p.counter, 1);
    named_bar_sync(bar_item, 2 * qthreads);
    const int item_idx = ctrl[0];
    if (item_idx >= p.nitems) break;
    const Item it = p.items[item_idx];

    for (int g0 = qtid * CPG; g0 < p.n_pad; g0 += qthreads * CPG) {
      RowT r[CPG];
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        if (it.cell[c] >= 0) {
          r[c] = *reinterpret_cast<const RowT*>(p.state + static_cast<int64_t>(it.cell[c]) * p.ld + g0);
        } else {
#pragma unroll
          for (int jj = 0; jj < CPG; ++jj) r[c].v[jj] = 0.f;
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
    named_bar_sync(bar_step, qthreads);

    unsigned char* cur = bufA;
    unsigned char* oth = bufB;
    int smax = 0, smin = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < CPG; ++c) { smax = max(smax, it.steps[c]); smin = min(smin, it.steps[c]); }

    for (int s = 0; s < smax; ++s) {
      if (p.nv > 0) {
        for (int v = qtid; v < p.nv_pad; v += qthreads) {
          float acc[CPG] = {0.f, 0.f, 0.f, 0.f};
          for (int k = 0; k < p.lv; ++k) {
            const WEnt e = wchunk[k * p.nv_pad + v];
            const RowT xv = *reinterpret_cast<const RowT*>(cur + e.off);
#pragma unroll
            for (int c = 0; c < CPG; ++c) acc[c] = fmaf(e.v, xv.v[c], acc[c]);
          }
          RowT o;
#pragma unroll
          for (int c = 0; c < CPG; ++c) o.v[c] = acc[c];
          ovf[v] = o;
        }
        named_bar_sync(bar_step, qthreads);
      }
      const bool all_active = s < smin;
      int bnext = (my_b0 < my_b1) ? wblk[my_b0] : 0;
      for (int bi = my_b0; bi < my_b1; ++bi) {
        const int b = bnext;
        bnext = (bi + 1 < my_b1) ? wblk[bi + 1] : 0;
        const int gbase = b * 128 + lane;
        // the noise tile becomes the accumulator: z = noise + (bias) + W x
        RowT acc[4];
        mbar_wait(bar_full, ph_full);
        ph_full ^= 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = tile[j * 32 + lane];
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_empty)) : "memory");
        if (p.dbg & 2) continue;
        {
          const int o1 = blk_off[b + 1];
          for (int o = blk_off[b] + lane; o < o1; o += 128) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const WEnt e = wtile[o + j * 32];
              const RowT xv = *reinterpret_cast<const RowT*>(cur + e.off);
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[j].v[c] = fmaf(e.v, xv.v[c], acc[j].v[c]);
            }
          }
        }
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float b0 = p.bias[gbase + j * 32];
#pragma unroll
            for (int c = 0; c < CPG; ++c) acc[j].v[c] += b0;
          }
        }
        if (b < p.nlong_blocks) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int2 oi = ovfidx[gbase + j * 32];
            for (int e = oi.x; e < oi.x + oi.y; ++e) {
              const RowT pv = ovf[e];
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc[j].v[c] += pv.v[c];
            }
          }
        }
        const bool pp = b < p.npp_blocks;
        const unsigned char* src = pp ? cur : bufA;
        unsigned char* dst = pp ? oth : bufA;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = gbase + j * 32;
          RowT xo = *reinterpret_cast<const RowT*>(src + static_cast<size_t>(g) * 16);
          const PC pcv = pcs[g];
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
            if (all_active || s < it.steps[c]) {
              const float x = xo.v[c];
              const float t = tanh_acc(acc[j].v[c]);
              const float x1 = fmaf(p.dt, t - x, x);
              xo.v[c] = fmaf(pcv.x, x1, pcv.y);
            }
          }
          *reinterpret_cast<RowT*>(dst + static_cast<size_t>(g) * 16) = xo;
        }
      }
      named_bar_sync(bar_step, qthreads);
      { unsigned char* t = cur; cur = oth; oth = t; }
    }

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
      for (int c = 0; c < CPG; ++c)
        if (it.cell[c] >= 0)
          *reinterpret_cast<RowT*>(p.state + static_cast<int64_t>(it.cell[c]) * p.ld + g0) = r[c];
    }
    named_bar_sync(bar_step, qthreads);
  }
}

}  // namespace scr
