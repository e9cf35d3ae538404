// Data-parallel front end of the reference's RegNetInference module, sm_100a.
//
//   rank_cells_kernel / rank_genes_kernel   RegNetInference.py:69-78  convertToS: scipy 'average' ranks of every
//                                            cell's genes, mapped to 2((r+1)/n)-1, then of every gene's cells
//   pairs_kernel                             RegNetInference.py:397-416 and :764-777  same-type cell pairs with
//                                            0 < pt1 - pt2 <= thresh whose states differ
//
// Ranks are integer work: a tie group occupying sorted positions [lo, hi) has average rank (lo + hi + 1) / 2, so the
// kernels carry the integer key 2r = lo + hi + 1 (<= 2n, uint16) between the two passes and only the last pass
// produces float64.  The first transform 2((r+1)/n)-1 is strictly increasing in r, so ranking the keys in the second
// pass equals ranking the transformed values the reference ranks.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace scr {
namespace regnet {

constexpr int kHistBins = 4096;   // per-cell counting path: values must be integers in [0, kHistBins)

// exclusive prefix sum, in place, of a[0..len) in shared memory; a[len] receives the total.
// `tmp` holds one int per warp.  All threads of the CTA must call.
__device__ inline void block_exclusive_scan(int* a, int len, int* tmp) {
  const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int per = (len + nthreads - 1) / nthreads;
  const int b = min(tid * per, len), e = min(b + per, len);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += a[i];
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) tmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? tmp[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += v;
    }
    if (lane < nwarps) tmp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  int run = incl - sum + (warp > 0 ? tmp[warp - 1] : 0);
  const int total = tmp[nwarps - 1];
  for (int i = b; i < e; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  if (tid == 0) a[len] = total;
  __syncthreads();
}

constexpr int kTileCells = 16;    // cells per CTA of the tiled counting kernel: 128-byte reads, 32-byte key writes
constexpr int kTileBins = 1024;   // its histogram range per cell

// Counting path for count matrices, coalesced: one CTA per tile of kTileCells adjacent cells.  A warp touches two
// gene rows x 16 cells per access (two full 128-byte lines in, two full 32-byte sectors out).  Pass A builds one
// shared-memory histogram per cell, a warp per cell turns it into prefix sums, pass B re-reads the tile (L2-resident:
// 16 x n x 8 bytes per CTA) and writes the keys.  A cell holding a non-integer, a negative or a value >= kTileBins is
// only flagged in `fallback` and left to rank_cells_kernel.
// Shared memory: int hist[kTileCells][kTileBins + 1] | int bad[kTileCells]
__global__ void __launch_bounds__(512) rank_cells_tile_kernel(const double* __restrict__ slab, int64_t cw, int n,
                                                              uint16_t* __restrict__ keys, int64_t cells, int64_t c0,
                                                              int* __restrict__ fallback, int* __restrict__ path_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* hist = reinterpret_cast<int*>(smem_raw);
  int* bad = hist + kTileCells * (kTileBins + 1);
  const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kTileCells;          // first cell of the tile (slab-local)
  const int cl = tid & (kTileCells - 1);                                      // my cell of the tile
  const int r0 = tid / kTileCells, rstep = nthreads / kTileCells;
  const bool live = t0 + cl < cw;
  for (int i = tid; i < kTileCells * (kTileBins + 1); i += nthreads) hist[i] = 0;
  if (tid < kTileCells) bad[tid] = 0;
  __syncthreads();
  int* myh = hist + cl * (kTileBins + 1);
  const double* col = slab + t0 + cl;
  int mybad = 0;
  if (live) {
    for (int g = r0; g < n; g += rstep) {
      const double x = col[static_cast<int64_t>(g) * cw];
      if (x >= 0.0 && x < static_cast<double>(kTileBins) && x == floor(x)) atomicAdd(&myh[static_cast<int>(x)], 1);
      else mybad = 1;
    }
  }
  if (mybad) bad[cl] = 1;
  __syncthreads();
  // warp w: exclusive prefix sums of cell w's histogram (kTileBins / 32 consecutive bins per lane)
  for (int c = warp; c < kTileCells; c += nthreads >> 5) {
    int* h = hist + c * (kTileBins + 1);
    constexpr int PER = kTileBins / 32;
    int sum = 0;
    for (int i = 0; i < PER; ++i) sum += h[lane * PER + i];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int run = incl - sum;
    for (int i = 0; i < PER; ++i) {
      const int v = h[lane * PER + i];
      h[lane * PER + i] = run;
      run += v;
    }
    if (lane == 31) h[kTileBins] = run;
  }
  __syncthreads();
  if (live) {
    if (bad[cl]) {
      if (r0 == 0) fallback[t0 + cl] = 1;
    } else {
      uint16_t* out = keys + c0 + t0 + cl;
      for (int g = r0; g < n; g += rstep) {
        const int v = static_cast<int>(col[static_cast<int64_t>(g) * cw]);
        out[static_cast<int64_t>(g) * cells] = static_cast<uint16_t>(myh[v] + myh[v + 1] + 1);
      }
      if (r0 == 0) { fallback[t0 + cl] = 0; if (path_counts) atomicAdd(&path_counts[0], 1); }
    }
  }
}

// One CTA per cell of the slab (n x cw float64, row-major: column c is the cell).  Writes keys[g * cells + c0 + c].
// With `fallback` != nullptr only the cells flagged by rank_cells_tile_kernel are processed.
// Shared memory: double key[n_pad] | int hist[kHistBins + 1] | int tmp[32] | uint16 idx[n_pad]
__global__ void __launch_bounds__(256) rank_cells_kernel(const double* __restrict__ slab, int64_t cw, int n, int n_pad,
                                                         uint16_t* __restrict__ keys, int64_t cells, int64_t c0,
                                                         int* __restrict__ path_counts, const int* __restrict__ fallback) {
  if (fallback && !fallback[blockIdx.x]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  int* hist = reinterpret_cast<int*>(key + n_pad);
  int* tmp = hist + kHistBins + 1;
  uint16_t* idx = reinterpret_cast<uint16_t*>(tmp + 32);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int64_t c = blockIdx.x;

  int ok = 1;
  for (int g = tid; g < n; g += nthreads) {
    const double x = slab[static_cast<int64_t>(g) * cw + c];
    key[g] = x;
    ok &= (x >= 0.0 && x < static_cast<double>(kHistBins) && x == floor(x)) ? 1 : 0;
  }
  const int all_int = __syncthreads_and(ok);
  uint16_t* out = keys + c0 + c;

  if (all_int) {
    // counting path (count matrices): histogram, prefix sums, rank straight from the bin
    for (int i = tid; i <= kHistBins; i += nthreads) hist[i] = 0;
    __syncthreads();
    for (int g = tid; g < n; g += nthreads) atomicAdd(&hist[static_cast<int>(key[g])], 1);
    __syncthreads();
    block_exclusive_scan(hist, kHistBins, tmp);
    for (int g = tid; g < n; g += nthreads) {
      const int v = static_cast<int>(key[g]);
      out[static_cast<int64_t>(g) * cells] = static_cast<uint16_t>(hist[v] + hist[v + 1] + 1);
    }
    if (tid == 0 && path_counts) atomicAdd(&path_counts[0], 1);
    return;
  }

  // general path: bitonic sort of (value, gene) in shared memory, then the tie group by binary search
  for (int g = n + tid; g < n_pad; g += nthreads) key[g] = CUDART_INF;
  for (int g = tid; g < n_pad; g += nthreads) idx[g] = static_cast<uint16_t>(g);
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n_pad >> 1); t += nthreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool up = (i & k) == 0;
        const double a = key[i], b = key[l];
        if ((a > b) == up) {
          key[i] = b; key[l] = a;
          const uint16_t ia = idx[i]; idx[i] = idx[l]; idx[l] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int p = tid; p < n; p += nthreads) {
    const double v = key[p];
    int lo = 0, hi = n;            // first position with key >= v
    while (lo < hi) { const int m = (lo + hi) >> 1; if (key[m] < v) lo = m + 1; else hi = m; }
    int lo2 = lo, hi2 = n;         // first position with key > v
    while (lo2 < hi2) { const int m = (lo2 + hi2) >> 1; if (key[m] <= v) lo2 = m + 1; else hi2 = m; }
    out[static_cast<int64_t>(idx[p]) * cells] = static_cast<uint16_t>(lo + lo2 + 1);
  }
  if (tid == 0 && path_counts) atomicAdd(&path_counts[1], 1);
}

// One CTA per gene of the chunk: histogram of the row's keys, prefix sums, float64 output row
//   out[c] = 2 * ((r + 1) / cells) - 1,  r = (lo + hi + 1) / 2   (RegNetInference.py:75-76)
// Shared memory: int hist[nbins + 1] | int tmp[32]
__global__ void __launch_bounds__(512) rank_genes_kernel(const uint16_t* __restrict__ keys, int64_t cells, int nbins,
                                                         int64_t g0, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* hist = reinterpret_cast<int*>(smem_raw);
  int* tmp = hist + nbins + 1;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const uint16_t* row = keys + (g0 + blockIdx.x) * cells;
  double* orow = out + static_cast<int64_t>(blockIdx.x) * cells;
  for (int i = tid; i <= nbins; i += nthreads) hist[i] = 0;
  __syncthreads();
  for (int64_t c = tid; c < cells; c += nthreads) atomicAdd(&hist[row[c]], 1);
  __syncthreads();
  block_exclusive_scan(hist, nbins, tmp);
  const double dcells = static_cast<double>(cells);
  for (int64_t c = tid; c < cells; c += nthreads) {
    const int k = row[c];
    const double r = static_cast<double>(static_cast<int64_t>(hist[k]) + hist[k + 1] + 1) * 0.5;
    orow[c] = 2.0 * ((r + 1.0) / dcells) - 1.0;
  }
}

// np.any(S[:, c1] - S[:, c2] != 0)  (RegNetInference.py:411-412)
__device__ inline bool states_differ(const double* __restrict__ S, int64_t ld, int n, int64_t c1, int64_t c2) {
  for (int g = 0; g < n; ++g)
    if (S[g * ld + c1] - S[g * ld + c2] != 0.0) return true;
  return false;
}

// Thread i owns position i of the type-sorted cell list (cell1); it scans the positions of its own type segment in
// ascending order (cell2), so a thread's pairs come out in the reference's order.  FILL = false counts, FILL = true
// writes [cell1, cell2] rows starting at offsets[i].  Pseudotimes of the scanned range are staged in shared memory.
template <bool FILL>
__global__ void __launch_bounds__(256) pairs_kernel(int64_t m, const double* __restrict__ pt, const int64_t* __restrict__ order,
                                                    const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_end,
                                                    double thresh, const double* __restrict__ S, int64_t ld, int n,
                                                    int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                                    int64_t* __restrict__ pairs) {
  constexpr int TILE = 1024;
  __shared__ double s_pt[TILE];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool live = i < m;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  const int64_t last = min(first + blockDim.x, m) - 1;
  const int64_t jlo = seg_begin[first], jhi = seg_end[last];   // positions are grouped by type: monotone bounds
  const int64_t s0 = live ? seg_begin[i] : 0, s1 = live ? seg_end[i] : 0;
  const double p1 = live ? pt[i] : 0.0;
  const int64_t c1 = live ? order[i] : 0;
  int64_t cnt = 0;
  int64_t w = FILL && live ? offsets[i] : 0;
  for (int64_t t0 = jlo; t0 < jhi; t0 += TILE) {
    const int len = static_cast<int>(min(static_cast<int64_t>(TILE), jhi - t0));
    __syncthreads();
    for (int k = threadIdx.x; k < len; k += blockDim.x) s_pt[k] = pt[t0 + k];
    __syncthreads();
    const int a = static_cast<int>(max(s0 - t0, static_cast<int64_t>(0))), b = static_cast<int>(min(s1 - t0, static_cast<int64_t>(len)));
    for (int k = a; k < b; ++k) {
      const double p2 = s_pt[k];
      if (fabs(p1 - p2) <= thresh && p1 > p2) {
        const int64_t c2 = order[t0 + k];
        if (states_differ(S, ld, n, c1, c2)) {
          if (FILL) { pairs[2 * w] = c1; pairs[2 * w + 1] = c2; ++w; }
          else ++cnt;
        }
      }
    }
  }
  if (!FILL && live) counts[i] = cnt;
}

}  // namespace regnet
}  // namespace scr
