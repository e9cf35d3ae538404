// Host-side packing of the regulatory matrix into the operator the stepper reads.
// Pure C++ (no CUDA) so the CPU test-suite can exercise it through scr_debug_host_matvec.
//
// Input: CSR of gc (rows = regulated gene; what developCell multiplies by, MatriSeq.py:530).
// Output (all in the INTERNAL gene order):
//   perm/inv      internal <-> original gene; order = long rows, then regulators (non-zero
//                 out-degree, the only genes a row can read), then the rest; each class by
//                 in-degree descending, so 32-row tiles have near-equal length
//   tiles         blocked ELL: block b = genes [128b, 128b+128); lane l owns rows 128b + 32j + l
//                 (j = 0..3); entry k of that row sits at slice_off[b] + (4k + j)*32 + l, so one
//                 loop trip feeds four independent gathers per lane; at most `cap` entries per row
//   chunks        the part of a row beyond `cap`, cut into pieces of `lv` entries; chunk v entry k
//                 at k*nv_pad + v; ovf_start/ovf_cnt give each long row its chunk range
//   wblk          128-gene blocks assigned to the wq warps of a cell group (LPT on a cost model)
#pragma once
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

namespace scr {

struct PackedOperator {
  int n = 0, n_pad = 0, nb = 0, npp = 0, npp_blocks = 0, nlong = 0, nlong_blocks = 0;
  int nv = 0, nv_pad = 0, lv = 0, cap = 0, wq = 1, wq_log2 = 0;
  int64_t nnz = 0;
  std::vector<int> perm, inv;
  std::vector<int> slice_off;
  std::vector<double> tile_val;
  std::vector<int> tile_col;
  std::vector<double> chunk_val;
  std::vector<int> chunk_col;
  std::vector<int> ovf_start, ovf_cnt;
  std::vector<double> rowsum;   // internal order, length n_pad
  std::vector<int> wblk_ptr, wblk;
  // the same assignment for a group run by 16 warps (convergence probes: few groups, the SM is theirs, and a step's
  // latency is the number of blocks a warp walks); empty when the operator has fewer than 16 blocks
  int wq_wide = 0;
  std::vector<int> wblk_ptr_wide, wblk_wide;
  // shared-memory wavefront model of one step's gathers (LDS.128: a quarter warp per wavefront, 8 x 16-byte bank groups;
  // rows in the same bank group at different addresses serialise), over the quarter-warp reads that have at least one
  // real entry: modelled count and its conflict-free floor
  int64_t gather_wavefronts = 0, gather_wavefronts_floor = 0;
};

inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// Cost of the quarter-warp read sets of a gene order under the LDS.128 bank model.  A read set is the (up to) eight
// state rows the lanes of one quarter warp gather in the same slot: rows 8m .. 8m+7 of the order at tile entry k, or
// eight consecutive chunks of long rows at chunk entry k.  Its cost is the largest number of DISTINCT positions that
// share a bank group (position mod 8); lanes without an entry read the row a real lane of their set reads (a broadcast,
// free) -- position 0 only when the whole set is padding.  The stepper's gathers were 30 % bank conflicts on the bench
// network before this model was used (profiles/r1_final_step_kernel_ncu_summary.txt).
struct BankModel {
  int n = 0, cap = 0, lv = 0;
  const int32_t* rowptr = nullptr;
  const int32_t* col = nullptr;
  std::vector<int> pos;                         // gene -> position
  std::vector<std::vector<int>> sets;           // read set -> column genes (padding lanes counted in `pad`)
  std::vector<int> pad;
  std::vector<std::vector<int>> readers;        // gene -> read sets it appears in
  int cost(int s) const {
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int seen[8][8];
    if (pad[s] > 0 && sets[s].empty()) { seen[0][0] = 0; cnt[0] = 1; }
    for (int g : sets[s]) {
      const int p = pos[g], b = p & 7;
      bool dup = false;
      for (int i = 0; i < cnt[b]; ++i) dup |= seen[b][i] == p;
      if (!dup && cnt[b] < 8) seen[b][cnt[b]++] = p;
    }
    int m = 1;
    for (int b = 0; b < 8; ++b) m = std::max(m, cnt[b]);
    return m;
  }
  void build(const std::vector<int>& order, const std::vector<int>& indeg, int nlong) {
    pos.assign(n, 0);
    for (int i = 0; i < n; ++i) pos[order[i]] = i;
    sets.clear(); pad.clear();
    readers.assign(n, {});
    auto add = [&](int s, int g) { sets[s].push_back(g); readers[g].push_back(s); };
    // tile entries: quarter m = positions [8m, 8m+8), slot k
    for (int m = 0; m * 8 < n; ++m) {
      int len = 0;
      for (int i = m * 8; i < std::min(n, m * 8 + 8); ++i) len = std::max(len, std::min(indeg[order[i]], cap));
      for (int k = 0; k < len; ++k) {
        const int s = static_cast<int>(sets.size());
        sets.emplace_back(); pad.push_back(0);
        for (int i = m * 8; i < m * 8 + 8; ++i) {
          if (i < n && k < std::min(indeg[order[i]], cap)) add(s, col[rowptr[order[i]] + k]);
          else pad[s]++;
        }
      }
    }
    // chunk entries of the long rows (order positions 0 .. nlong-1 in sequence, lv entries per chunk)
    std::vector<std::pair<int, int>> chunks;   // (row gene, first entry)
    for (int i = 0; i < nlong; ++i)
      for (int e = cap; e < indeg[order[i]]; e += lv) chunks.emplace_back(order[i], e);
    for (size_t v0 = 0; v0 < chunks.size(); v0 += 8)
      for (int k = 0; k < lv; ++k) {
        const int s = static_cast<int>(sets.size());
        sets.emplace_back(); pad.push_back(0);
        for (size_t v = v0; v < v0 + 8; ++v) {
          if (v < chunks.size() && chunks[v].second + k < indeg[chunks[v].first]) add(s, col[rowptr[chunks[v].first] + chunks[v].second + k]);
          else pad[s]++;
        }
      }
  }
  int64_t total() const {
    int64_t t = 0;
    for (size_t s = 0; s < sets.size(); ++s) t += cost(static_cast<int>(s));
    return t;
  }
};

// SCR_BANK_OPT (default 6 passes, 0 = off; +5.8 % on the bench network, profiles/r2_experiments.txt): genes are exchanged inside aligned groups of 8 positions
// -- which keeps every block's row set, the class prefixes and the rows of each quarter warp, and only changes the bank
// group a gene's state row falls into -- whenever that lowers the modelled wavefront count.  Long rows stay where they
// are (their order numbers the chunks, i.e. decides which chunks share a quarter warp), so every read set keeps its
// members while positions move.
inline void bank_optimize(BankModel& bm, std::vector<int>& order, const std::vector<int>& klass, int passes) {
  const int n = bm.n;
  for (int ps = 0; ps < passes; ++ps) {
    int64_t gain = 0;
    for (int g0 = 0; g0 < n; g0 += 8)
      for (int a = g0; a < std::min(n, g0 + 8); ++a)
        for (int b = a + 1; b < std::min(n, g0 + 8); ++b) {
          const int ga = order[a], gb = order[b];
          if (klass[ga] != klass[gb] || klass[ga] == 0) continue;
          if (bm.readers[ga].empty() && bm.readers[gb].empty()) continue;
          std::vector<int> touched(bm.readers[ga]);
          touched.insert(touched.end(), bm.readers[gb].begin(), bm.readers[gb].end());
          std::sort(touched.begin(), touched.end());
          touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
          int64_t before = 0, after = 0;
          for (int s : touched) before += bm.cost(s);
          std::swap(bm.pos[ga], bm.pos[gb]);
          for (int s : touched) after += bm.cost(s);
          if (after < before) {
            std::swap(order[a], order[b]);
            gain += before - after;
          } else {
            std::swap(bm.pos[ga], bm.pos[gb]);
          }
        }
    if (gain == 0) break;
  }
}

// Annealed version of the same move set (exchanges inside aligned groups of 8 positions): the pairwise descent above stops
// in a local minimum; random exchanges accepted with probability exp(-delta / T) under a falling temperature get lower.
// Deterministic (fixed-seed LCG): the gene order is a pure function of the network's pattern.
inline void bank_anneal(BankModel& bm, std::vector<int>& order, const std::vector<int>& klass, int64_t proposals) {
  const int n = bm.n;
  if (n < 16) return;
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return static_cast<uint32_t>(rng >> 33); };
  std::vector<int> touched;
  std::vector<int> best_order = order;
  int64_t cur = bm.total(), best = cur;
  const double t0 = 0.6, t1 = 0.02;
  for (int64_t it = 0; it < proposals; ++it) {
    const double T = t0 * std::pow(t1 / t0, static_cast<double>(it) / static_cast<double>(proposals));
    const int g0 = static_cast<int>(next() % static_cast<uint32_t>((n + 7) / 8)) * 8;
    const int a = g0 + static_cast<int>(next() & 7), b = g0 + static_cast<int>(next() & 7);
    if (a == b || a >= n || b >= n) continue;
    const int ga = order[a], gb = order[b];
    if (klass[ga] != klass[gb] || klass[ga] == 0) continue;
    if (bm.readers[ga].empty() && bm.readers[gb].empty()) continue;
    touched.assign(bm.readers[ga].begin(), bm.readers[ga].end());
    touched.insert(touched.end(), bm.readers[gb].begin(), bm.readers[gb].end());
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    int before = 0, after = 0;
    for (int s : touched) before += bm.cost(s);
    std::swap(bm.pos[ga], bm.pos[gb]);
    for (int s : touched) after += bm.cost(s);
    const int delta = after - before;
    bool accept = delta <= 0;
    if (!accept) accept = (next() & 0xffffff) * (1.0 / 16777216.0) < std::exp(-delta / T);
    if (accept) {
      std::swap(order[a], order[b]);
      cur += delta;
      if (cur < best) { best = cur; best_order = order; }
    } else {
      std::swap(bm.pos[ga], bm.pos[gb]);
    }
  }
  order = best_order;
}

// Second stage (round 2, SCR_BANK_OPT2 passes, default 3): exchanges of two same-class genes anywhere inside one 128-gene block.
// That keeps every block's row set (slot counts, class prefixes) but, unlike the exchanges inside an aligned group of 8,
// also changes WHICH rows share a quarter warp, and lets a gene take any of the 8 bank groups.  Incremental evaluation: a
// trial touches the two quarter warps the genes sit in, the quarter warps of the rows that read either gene, and the
// long-row chunk sets that read them.  Long rows keep their places (their order numbers the chunks).
struct BankSearch {
  int n = 0, cap = 0, lv = 0, nlong = 0;
  const int32_t* rowptr = nullptr;
  const int32_t* col = nullptr;
  const std::vector<int>* indeg = nullptr;
  std::vector<int> order, pos;
  std::vector<std::vector<int>> readers;                    // gene -> rows whose first `cap` entries read it
  std::vector<std::pair<int, int>> chunks;                  // (row gene, first entry) in chunk order
  std::vector<std::vector<std::pair<int, int>>> chunk_reads;   // gene -> (chunk set, k) it is read in

  static int max_mult(const int* p, int cnt) {               // most distinct positions sharing a bank group
    int best = cnt > 0 ? 1 : 0;
    for (int i = 0; i < cnt; ++i) {
      int m = 1;
      for (int j = 0; j < i; ++j) if (p[j] == p[i]) { m = 0; break; }        // duplicate address: counted once
      if (!m) continue;
      for (int j = i + 1; j < cnt; ++j) if ((p[j] & 7) == (p[i] & 7) && p[j] != p[i]) {
        bool dup = false;
        for (int q = i + 1; q < j; ++q) dup |= p[q] == p[j];
        if (!dup) ++m;
      }
      best = std::max(best, m);
    }
    return best;
  }
  int quarter_cost(int q) const {
    int len = 0;
    const int p0 = q * 8, p1 = std::min(n, p0 + 8);
    for (int i = p0; i < p1; ++i) len = std::max(len, std::min((*indeg)[order[i]], cap));
    int total = 0;
    for (int k = 0; k < len; ++k) {
      int pp[8], cnt = 0;
      for (int i = p0; i < p1; ++i) {
        const int r = order[i];
        if (k < std::min((*indeg)[r], cap)) pp[cnt++] = pos[col[rowptr[r] + k]];
      }
      total += std::max(1, max_mult(pp, cnt));
    }
    return total;
  }
  int chunk_set_cost(int s, int k) const {
    int pp[8], cnt = 0;
    for (size_t v = static_cast<size_t>(s) * 8; v < static_cast<size_t>(s) * 8 + 8 && v < chunks.size(); ++v)
      if (chunks[v].second + k < (*indeg)[chunks[v].first]) pp[cnt++] = pos[col[rowptr[chunks[v].first] + chunks[v].second + k]];
    return std::max(1, max_mult(pp, cnt));
  }
  void init(const std::vector<int>& ord, int nl) {
    order = ord;
    nlong = nl;
    pos.assign(n, 0);
    for (int i = 0; i < n; ++i) pos[order[i]] = i;
    readers.assign(n, {});
    for (int r = 0; r < n; ++r)
      for (int k = 0; k < std::min((*indeg)[r], cap); ++k) readers[col[rowptr[r] + k]].push_back(r);
    for (auto& v : readers) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    chunks.clear();
    for (int i = 0; i < nlong; ++i)
      for (int e = cap; e < (*indeg)[order[i]]; e += lv) chunks.emplace_back(order[i], e);
    chunk_reads.assign(n, {});
    for (size_t v = 0; v < chunks.size(); ++v)
      for (int k = 0; k < lv && chunks[v].second + k < (*indeg)[chunks[v].first]; ++k)
        chunk_reads[col[rowptr[chunks[v].first] + chunks[v].second + k]].emplace_back(static_cast<int>(v / 8), k);
    for (auto& v : chunk_reads) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
  }
  int64_t run(const std::vector<int>& klass, int passes) {
    int64_t gained = 0;
    std::vector<int> qs;
    std::vector<std::pair<int, int>> cs;
    for (int ps = 0; ps < passes; ++ps) {
      int64_t gain = 0;
      for (int b0 = 0; b0 < n; b0 += 128) {
        const int b1 = std::min(n, b0 + 128);
        for (int pa = b0; pa < b1; ++pa)
          for (int pb = pa + 1; pb < b1; ++pb) {
            const int ga = order[pa], gb = order[pb];
            if (klass[ga] != klass[gb] || klass[ga] == 0) continue;
            if ((pa >> 3) == (pb >> 3) && readers[ga].empty() && readers[gb].empty() && chunk_reads[ga].empty() && chunk_reads[gb].empty()) continue;
            if (readers[ga].size() + readers[gb].size() > 600) continue;      // hubs: a trial would touch most of the operator
            qs.clear(); cs.clear();
            qs.push_back(pa >> 3); qs.push_back(pb >> 3);
            for (int r : readers[ga]) qs.push_back(pos[r] >> 3);
            for (int r : readers[gb]) qs.push_back(pos[r] >> 3);
            std::sort(qs.begin(), qs.end());
            qs.erase(std::unique(qs.begin(), qs.end()), qs.end());
            cs.insert(cs.end(), chunk_reads[ga].begin(), chunk_reads[ga].end());
            cs.insert(cs.end(), chunk_reads[gb].begin(), chunk_reads[gb].end());
            std::sort(cs.begin(), cs.end());
            cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
            int before = 0, after = 0;
            for (int q : qs) before += quarter_cost(q);
            for (auto& c : cs) before += chunk_set_cost(c.first, c.second);
            std::swap(order[pa], order[pb]);
            pos[ga] = pb; pos[gb] = pa;
            // the rows' own quarters may have changed: a reader that IS ga or gb now sits elsewhere, both quarters are in qs
            for (int q : qs) after += quarter_cost(q);
            for (auto& c : cs) after += chunk_set_cost(c.first, c.second);
            if (after < before) {
              gain += before - after;
            } else {
              std::swap(order[pa], order[pb]);
              pos[ga] = pa; pos[gb] = pb;
            }
          }
      }
      gained += gain;
      if (gain == 0) break;
    }
    return gained;
  }
};

// The bank-aware order is a pure function of the sparsity pattern (and of the search settings), and callers set the same
// pattern again and again (the alpha search rescales the values, bench.py's end-to-end pass re-uploads the network): the
// result is kept per process, keyed by a hash of the pattern.
struct OrderCache {
  std::mutex mu;
  std::map<uint64_t, std::vector<int>> orders;
  static OrderCache& get() { static OrderCache c; return c; }
  static uint64_t key(int n, int64_t nnz, const int32_t* rowptr, const int32_t* col, int a, int b, int c, int d) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 0x100000001b3ull; h ^= h >> 29; };
    mix(static_cast<uint64_t>(n)); mix(static_cast<uint64_t>(nnz)); mix(a); mix(b); mix(c); mix(d);
    for (int i = 0; i <= n; ++i) mix(static_cast<uint32_t>(rowptr[i]));
    for (int64_t e = 0; e < nnz; ++e) mix(static_cast<uint32_t>(col[e]));
    return h;
  }
};

// Entry order inside the rows under the FINAL gene order (positions fixed): which of a row's entries sits in which slot is
// free, and the eight entries one LDS.128 gathers for a quarter warp should fall into eight different bank groups (or be
// the same row).  Two deterministic descents over exchanges of two entries of one row, accepted when the affected read
// sets get cheaper under BankModel's cost:
//   tiles   the first min(in-degree, cap) entries of the eight rows of a quarter warp, slot against slot;
//   chunks  the entries beyond the cap of a long row, chunk slot against chunk slot (a hub target with 300 regulators is
//           38 chunks: its eight-chunk groups read 8 x 8 entries that can be dealt out one bank group per read).
inline void entry_bank_pass(int n, const int32_t* rowptr, int32_t* col, double* val, const std::vector<int>& order,
                            const std::vector<int>& indeg, int cap, int lv, int nlong) {
  std::vector<int> pos(n);
  for (int i = 0; i < n; ++i) pos[order[i]] = i;
  auto mult = [&](const int* p, int cnt) {   // most distinct positions sharing a bank group (BankModel::cost)
    int best = cnt > 0 ? 1 : 0;
    for (int a = 0; a < cnt; ++a) {
      int same = 1;
      bool first = true;
      for (int b = 0; b < a; ++b) if (p[b] == p[a]) first = false;
      if (!first) continue;
      for (int b = a + 1; b < cnt; ++b) {
        if ((p[b] & 7) != (p[a] & 7) || p[b] == p[a]) continue;
        bool dup = false;
        for (int c = a + 1; c < b; ++c) dup |= p[c] == p[b];
        if (!dup) ++same;
      }
      best = std::max(best, same);
    }
    return best;
  };
  auto swap_entries = [&](int64_t a, int64_t b) { std::swap(col[a], col[b]); std::swap(val[a], val[b]); };
  // ---- tiles ----
  for (int m = 0; m * 8 < n; ++m) {
    const int i0 = m * 8, i1 = std::min(n, i0 + 8);
    auto slot_cost = [&](int k) {
      int p[8], cnt = 0;
      for (int i = i0; i < i1; ++i) {
        const int r = order[i];
        if (k < std::min(indeg[r], cap)) p[cnt++] = pos[col[rowptr[r] + k]];
      }
      return mult(p, cnt);
    };
    for (int pass = 0; pass < 4; ++pass) {
      bool improved = false;
      for (int i = i0; i < i1; ++i) {
        const int r = order[i], d = std::min(indeg[r], cap);
        for (int k1 = 0; k1 < d; ++k1)
          for (int k2 = k1 + 1; k2 < d; ++k2) {
            const int before = slot_cost(k1) + slot_cost(k2);
            if (before <= 2) continue;
            swap_entries(rowptr[r] + k1, rowptr[r] + k2);
            if (slot_cost(k1) + slot_cost(k2) < before) improved = true;
            else swap_entries(rowptr[r] + k1, rowptr[r] + k2);
          }
      }
      if (!improved) break;
    }
  }
  // ---- chunks of the long rows (order positions 0 .. nlong-1) ----
  std::vector<std::pair<int, int>> chunks;   // (row gene, first entry)
  for (int i = 0; i < nlong; ++i)
    for (int e = cap; e < indeg[order[i]]; e += lv) chunks.emplace_back(order[i], e);
  const int nch = static_cast<int>(chunks.size());
  auto set_cost = [&](int g, int k) {        // read set: chunks 8g .. 8g+7 at slot k
    int p[8], cnt = 0;
    for (int v = g * 8; v < std::min(nch, g * 8 + 8); ++v)
      if (chunks[v].second + k < indeg[chunks[v].first]) p[cnt++] = pos[col[rowptr[chunks[v].first] + chunks[v].second + k]];
    return mult(p, cnt);
  };
  std::vector<int> first_chunk(n, -1);
  for (int v = nch - 1; v >= 0; --v) first_chunk[chunks[v].first] = v;
  for (int i = 0; i < nlong; ++i) {
    const int r = order[i], extra = indeg[r] - cap;
    if (extra < 2 || first_chunk[r] < 0) continue;
    auto where = [&](int e, int& g, int& k) { const int v = first_chunk[r] + e / lv; g = v / 8; k = e % lv; };
    for (int pass = 0; pass < 3; ++pass) {
      bool improved = false;
      for (int a = 0; a < extra; ++a) {
        int ga, ka;
        where(a, ga, ka);
        if (set_cost(ga, ka) <= 1) continue;
        for (int b = 0; b < extra; ++b) {
          int gb, kb;
          where(b, gb, kb);
          if (b == a || (ga == gb && ka == kb)) continue;
          const int before = set_cost(ga, ka) + set_cost(gb, kb);
          swap_entries(rowptr[r] + cap + a, rowptr[r] + cap + b);
          if (set_cost(ga, ka) + set_cost(gb, kb) < before) { improved = true; if (set_cost(ga, ka) <= 1) break; }
          else swap_entries(rowptr[r] + cap + a, rowptr[r] + cap + b);
        }
      }
      if (!improved) break;
    }
  }
}

// identity_order: keep the caller's gene order and no row cap (dense networks: every row is long and
// every gene is read, so sorting buys nothing and the GEMM path needs the natural order);
// pad_to: n_pad is rounded up to a multiple of this (128 for the sparse path, 256 for the GEMM tiles)
inline bool pack_operator(int n, int64_t nnz, const int32_t* rowptr, const int32_t* col, const double* val,
                          PackedOperator& op, std::string& err, bool identity_order = false, int pad_to = 128) {
  if (n <= 0 || nnz < 0 || !rowptr || (nnz > 0 && (!col || !val))) { err = "bad CSR arguments"; return false; }
  if (rowptr[0] != 0 || rowptr[n] != nnz) { err = "rowptr does not span nnz"; return false; }
  for (int r = 0; r < n; ++r)
    if (rowptr[r + 1] < rowptr[r]) { err = "rowptr not monotone"; return false; }
  for (int64_t e = 0; e < nnz; ++e)
    if (col[e] < 0 || col[e] >= n) { err = "column index out of range"; return false; }

  op = PackedOperator();
  op.n = n;
  op.nnz = nnz;
  op.n_pad = (n + pad_to - 1) / pad_to * pad_to;
  op.nb = op.n_pad / 128;

  std::vector<int> indeg(n), outdeg(n, 0);
  for (int r = 0; r < n; ++r) indeg[r] = rowptr[r + 1] - rowptr[r];
  for (int64_t e = 0; e < nnz; ++e) outdeg[col[e]]++;

  // Entry order inside a row (SCR_ROW_SORT, default on).  The k-th entries of the eight rows of a quarter warp are gathered
  // by ONE LDS.128: lanes that read the SAME state row cost one wavefront together (a broadcast).  Regulatory networks have
  // hubs -- a few regulators with hundreds of targets -- so every row lists its regulators hub first (by out-degree), and
  // the gene order below puts rows with the same leading regulators next to each other: whole quarter warps then read one
  // or two rows per slot instead of eight scattered ones.
  std::vector<int> hub_rank(n, 0);
  std::vector<int32_t> col_sorted;
  std::vector<double> val_sorted;
  const bool row_sort = !identity_order && nnz > 0 && env_int("SCR_ROW_SORT", 1) != 0;
  if (row_sort) {
    std::vector<int> by_out(n);
    std::iota(by_out.begin(), by_out.end(), 0);
    std::stable_sort(by_out.begin(), by_out.end(), [&](int a, int b) { return outdeg[a] > outdeg[b]; });
    for (int i = 0; i < n; ++i) hub_rank[by_out[i]] = i;
    col_sorted.assign(col, col + nnz);
    val_sorted.assign(val, val + nnz);
    std::vector<int> idx;
    for (int r = 0; r < n; ++r) {
      const int d = indeg[r];
      if (d < 2) continue;
      idx.resize(d);
      std::iota(idx.begin(), idx.end(), 0);
      const int32_t* c0 = col + rowptr[r];
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return hub_rank[c0[a]] < hub_rank[c0[b]]; });
      for (int k = 0; k < d; ++k) {
        col_sorted[rowptr[r] + k] = c0[idx[k]];
        val_sorted[rowptr[r] + k] = val[rowptr[r] + idx[k]];
      }
    }
    col = col_sorted.data();
    val = val_sorted.data();
  }

  const double mean_in = static_cast<double>(nnz) / n;
  int cap = std::max(8, static_cast<int>(2.0 * mean_in + 0.999));
  cap = env_int("SCR_ROW_CAP", cap);
  if (cap < 1) cap = 1;
  if (identity_order) {
    cap = 1;
    for (int r = 0; r < n; ++r) cap = std::max(cap, rowptr[r + 1] - rowptr[r]);
  }
  op.cap = cap;
  op.lv = std::max(1, env_int("SCR_CHUNK_LEN", std::min(cap, 8)));

  // ---- gene order ----
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  auto cls = [&](int g) { return identity_order ? 1 : (indeg[g] > cap ? 0 : (outdeg[g] > 0 ? 1 : 2)); };
  if (!identity_order)
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const int ca = cls(a), cb = cls(b);
      if (ca != cb) return ca < cb;
      if (indeg[a] != indeg[b]) return indeg[a] > indeg[b];
      if (!row_sort || ca == 0) return false;
      // same class, same in-degree: by the regulators they read, hub first (rows sharing them become neighbours)
      const int32_t *pa = col + rowptr[a], *pb = col + rowptr[b];
      for (int k = 0; k < indeg[a]; ++k)
        if (pa[k] != pb[k]) return hub_rank[pa[k]] < hub_rank[pb[k]];
      return false;
    });
  if (!identity_order && nnz > 0) {
    BankModel bm;
    bm.n = n; bm.cap = cap; bm.lv = op.lv; bm.rowptr = rowptr; bm.col = col;
    int nlong = 0;
    std::vector<int> klass(n);
    for (int g = 0; g < n; ++g) { klass[g] = cls(g); nlong += klass[g] == 0; }
    const int bank_passes = env_int("SCR_BANK_OPT", 6);
    // annealing effort in proposals per non-zero (default 40: ~0.2 s once per pattern at n = 3000; +1.0 / +1.5 / +1.7 % on the
    // three bench networks, profiles/r2_experiments.txt); the block-level search stays opt-in (+2 % on TRRUST for 0.6 s)
    const int anneal = env_int("SCR_BANK_ANNEAL", 40), passes2 = env_int("SCR_BANK_OPT2", 0);
    const uint64_t okey = OrderCache::key(n, nnz, rowptr, col, bank_passes, anneal, passes2, cap * 64 + op.lv);
    bool cached = false;
    {
      std::lock_guard<std::mutex> lock(OrderCache::get().mu);
      auto it = OrderCache::get().orders.find(okey);
      if (it != OrderCache::get().orders.end() && static_cast<int>(it->second.size()) == n) { order = it->second; cached = true; }
    }
    // the search is for the sparse biological networks (mean in-degree 1.5 - 3); denser operators (mean in-degree > 8) are
    // left in plain order: an exchange would touch hundreds of read sets (seconds of packing), and their gather is bound by
    // operator traffic rather than by bank conflicts
    const bool searchable = nnz <= static_cast<int64_t>(n) * 8;
    if (bank_passes > 0 && !cached && searchable) {
      // the rows of a quarter warp are the members of its 8-group whatever their order inside it, but WHICH slot a
      // row's k-th entry shares with its neighbours' k-th entries is fixed: rebuild the read sets after the exchange
      bm.build(order, indeg, nlong);
      bank_optimize(bm, order, klass, bank_passes == 1 ? 6 : bank_passes);
      if (anneal > 0) {
        // work per proposal ~ the read sets of two genes ~ 2 nnz / n: keep the whole search near 5 M set evaluations
        const int64_t budget = 5000000 / std::max<int64_t>(1, 2 * nnz / std::max(n, 1));
        bank_anneal(bm, order, klass, std::min<int64_t>(static_cast<int64_t>(anneal) * std::max<int64_t>(nnz, 1), budget));
        bm.build(order, indeg, nlong);
        bank_optimize(bm, order, klass, 6);
      }
      if (passes2 > 0) {
        BankSearch bs;
        bs.n = n; bs.cap = cap; bs.lv = op.lv; bs.rowptr = rowptr; bs.col = col; bs.indeg = &indeg;
        bs.init(order, nlong);
        bs.run(klass, passes2);
        order = bs.order;
      }
      std::lock_guard<std::mutex> lock(OrderCache::get().mu);
      if (OrderCache::get().orders.size() > 64) OrderCache::get().orders.clear();
      OrderCache::get().orders[okey] = order;
    }
    if (row_sort && searchable && env_int("SCR_ENTRY_BANKS", 1) != 0)
      entry_bank_pass(n, rowptr, col_sorted.data(), val_sorted.data(), order, indeg, cap, op.lv, nlong);
    bm.build(order, indeg, nlong);
    op.gather_wavefronts = bm.total();
    op.gather_wavefronts_floor = static_cast<int64_t>(bm.sets.size());
    if (env_int("SCR_PACK_STATS", 0)) {   // developer aid: modelled cost of the read sets by size
      // (the chunk read sets come last: one per 8 chunks and chunk slot)
      int64_t nchunk_sets = 0;
      {
        int64_t nchunks = 0;
        for (int i = 0; i < nlong; ++i) nchunks += (indeg[order[i]] - cap + op.lv - 1) / op.lv;
        nchunk_sets = (nchunks + 7) / 8 * op.lv;
      }
      for (int part = 0; part < 2; ++part) {
        int64_t cnt[9][9] = {};
        const size_t s_lo = part == 0 ? 0 : bm.sets.size() - nchunk_sets, s_hi = part == 0 ? bm.sets.size() - nchunk_sets : bm.sets.size();
        for (size_t s2 = s_lo; s2 < s_hi; ++s2) cnt[std::min<size_t>(8, bm.sets[s2].size())][std::min(8, bm.cost(static_cast<int>(s2)))]++;
        for (int sz = 0; sz <= 8; ++sz) {
          std::fprintf(stderr, "%s read sets with %d entries:", part == 0 ? "tile " : "chunk", sz);
          for (int c = 1; c <= 8; ++c) std::fprintf(stderr, " cost%d=%lld", c, static_cast<long long>(cnt[sz][c]));
          std::fprintf(stderr, "\n");
        }
      }
    }
  }
  op.perm.assign(op.n_pad, -1);
  op.inv.assign(n, -1);
  int n_read = 0;
  for (int i = 0; i < n; ++i) {
    op.perm[i] = order[i];
    op.inv[order[i]] = i;
    if (cls(order[i]) == 0) op.nlong++;
    if (cls(order[i]) <= 1) n_read++;
  }
  op.npp = std::min(op.n_pad, (n_read + 127) / 128 * 128);
  op.npp_blocks = op.npp / 128;
  op.nlong_blocks = (op.nlong + 127) / 128;

  // ---- tiles ----
  op.slice_off.assign(op.nb + 1, 0);
  for (int b = 0; b < op.nb; ++b) {
    int len = 0;
    for (int l = 0; l < 128; ++l) {
      const int i = b * 128 + l;
      if (i < n) len = std::max(len, std::min(indeg[op.perm[i]], cap));
    }
    op.slice_off[b + 1] = op.slice_off[b] + 128 * len;
  }
  const int nslots = op.slice_off[op.nb];
  op.tile_val.assign(nslots, 0.0);
  op.tile_col.assign(nslots, 0);
  op.rowsum.assign(op.n_pad, 0.0);

  // ---- chunks of long rows ----
  const int ovf_genes = op.nlong_blocks * 128;
  op.ovf_start.assign(ovf_genes, 0);
  op.ovf_cnt.assign(ovf_genes, 0);
  int nv = 0;
  for (int i = 0; i < op.nlong; ++i) {
    const int extra = indeg[op.perm[i]] - cap;
    const int pieces = (extra + op.lv - 1) / op.lv;
    op.ovf_start[i] = nv;
    op.ovf_cnt[i] = pieces;
    nv += pieces;
  }
  op.nv = nv;
  op.nv_pad = (nv + 31) / 32 * 32;
  op.chunk_val.assign(static_cast<size_t>(op.lv) * op.nv_pad, 0.0);
  op.chunk_col.assign(static_cast<size_t>(op.lv) * op.nv_pad, 0);

  for (int i = 0; i < n; ++i) {
    const int r = op.perm[i];
    const int blk = i / 128, j = (i % 128) / 32, lane = i % 32;
    double sum = 0.0;
    for (int k = 0; k < indeg[r]; ++k) {
      const int64_t e = rowptr[r] + k;
      const int ci = op.inv[col[e]];
      sum += val[e];
      if (k < cap) {
        const int slot = op.slice_off[blk] + (4 * k + j) * 32 + lane;
        op.tile_val[slot] = val[e];
        op.tile_col[slot] = ci;
      } else {
        const int kk = k - cap;
        const int v = op.ovf_start[i] + kk / op.lv;
        const size_t slot = static_cast<size_t>(kk % op.lv) * op.nv_pad + v;
        op.chunk_val[slot] = val[e];
        op.chunk_col[slot] = ci;
      }
    }
    op.rowsum[i] = sum;
  }
  // padding lanes gather the row a real lane of their quarter warp gathers in the same slot (same address: a broadcast,
  // no extra shared-memory wavefront) instead of row 0, which could collide with a real entry of bank group 0
  {
    std::vector<char> real_tile(nslots, 0), real_chunk(op.chunk_val.size(), 0);
    for (int i = 0; i < n; ++i) {
      const int r = op.perm[i];
      const int blk = i / 128, j = (i % 128) / 32, lane = i % 32;
      for (int k = 0; k < indeg[r]; ++k) {
        if (k < cap) real_tile[op.slice_off[blk] + (4 * k + j) * 32 + lane] = 1;
        else real_chunk[static_cast<size_t>((k - cap) % op.lv) * op.nv_pad + op.ovf_start[i] + (k - cap) / op.lv] = 1;
      }
    }
    auto fix = [](std::vector<int>& colv, const std::vector<char>& real, size_t base) {   // one aligned group of 8 lanes
      int donor = -1;
      for (size_t l = 0; l < 8; ++l) if (real[base + l]) { donor = colv[base + l]; break; }
      if (donor < 0) return;
      for (size_t l = 0; l < 8; ++l) if (!real[base + l]) colv[base + l] = donor;
    };
    for (size_t base = 0; base + 8 <= static_cast<size_t>(nslots); base += 8) fix(op.tile_col, real_tile, base);
    for (size_t base = 0; base + 8 <= op.chunk_col.size(); base += 8) fix(op.chunk_col, real_chunk, base);
  }
  // every column any row reads must live in the ping-ponged prefix
  for (int slot = 0; slot < nslots; ++slot)
    if (op.tile_val[slot] != 0.0 && op.tile_col[slot] >= op.npp) { err = "internal: column outside ping-pong prefix"; return false; }

  // ---- warps per cell group and block -> warp assignment ----
  int wq = 1;
  while (wq * 2 <= std::min(8, op.nb)) wq *= 2;
  const int wq_env = env_int("SCR_WARPS", 0);
  if (wq_env == 1 || wq_env == 2 || wq_env == 4 || wq_env == 8 || wq_env == 16) wq = wq_env;
  op.wq = wq;
  op.wq_log2 = 0;
  while ((1 << op.wq_log2) < wq) op.wq_log2++;

  std::vector<double> cost(op.nb, 0.0);
  for (int b = 0; b < op.nb; ++b) {
    for (int j = 0; j < 4; ++j)
      if ((b * 4 + j) * 32 < n) cost[b] += 140.0;               // element-wise work of a live 32-row tile
    cost[b] += 26.0 * (op.slice_off[b + 1] - op.slice_off[b]) / 128;   // one trip = 4 gathers + 16 FMAs
    if (b < op.nlong_blocks)
      for (int i = b * 128; i < b * 128 + 128; ++i) cost[b] += 0.2 * op.ovf_cnt[i];
  }
  std::vector<int> by_cost(op.nb);
  std::iota(by_cost.begin(), by_cost.end(), 0);
  std::stable_sort(by_cost.begin(), by_cost.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  // LPT of the blocks onto `warps` warps -> (list bounds, blocks in list order)
  auto assign = [&](int warps, std::vector<int>& ptr, std::vector<int>& blk) {
    std::vector<double> load(warps, 0.0);
    // warps 0 .. nv_pad/32-1 evaluate the long-row chunks at the start of every step (step_kernel.cuh phase A):
    // charge them for it so that LPT hands them lighter blocks
    const double chunk_cost = env_int("SCR_CHUNK_COST", 60);
    for (int v0 = 0; v0 < op.nv_pad; v0 += 32) load[(v0 / 32) % warps] += chunk_cost;
    std::vector<std::vector<int>> lists(warps);
    for (int b : by_cost) {
      int w = 0;
      for (int k = 1; k < warps; ++k)
        if (load[k] < load[w] - 1e-9 || (std::abs(load[k] - load[w]) <= 1e-9 && lists[k].size() < lists[w].size())) w = k;
      lists[w].push_back(b);
      load[w] += cost[b];
    }
    ptr.assign(warps + 1, 0);
    blk.clear();
    for (int w = 0; w < warps; ++w) {
      // order the stepper's split barriers rely on: regulator blocks without long rows, blocks with long
      // rows (their chunk partials arrive late), then blocks nobody reads (step_kernel.cuh)
      auto kls = [&](int b) { return b < op.nlong_blocks ? 1 : (b < op.npp_blocks ? 0 : 2); };
      std::sort(lists[w].begin(), lists[w].end(), [&](int a, int b) { return kls(a) != kls(b) ? kls(a) < kls(b) : a < b; });
      for (int b : lists[w]) blk.push_back(b);
      ptr[w + 1] = static_cast<int>(blk.size());
    }
  };
  assign(wq, op.wblk_ptr, op.wblk);
  if (wq == 8 && op.nb >= 16 && env_int("SCR_PROBE_WIDE", 1)) {
    op.wq_wide = 16;
    assign(16, op.wblk_ptr_wide, op.wblk_wide);
  }
  return true;
}

// y = W x walking the packed structure exactly as the kernel does (tiles, then chunks)
inline void host_matvec(const PackedOperator& op, const double* x, double* y) {
  std::vector<double> xi(op.n_pad, 0.0), yi(op.n_pad, 0.0), part(std::max(op.nv_pad, 1), 0.0);
  for (int i = 0; i < op.n_pad; ++i)
    if (op.perm[i] >= 0) xi[i] = x[op.perm[i]];
  for (int v = 0; v < op.nv_pad; ++v) {
    double acc = 0.0;
    for (int k = 0; k < op.lv; ++k) {
      const size_t slot = static_cast<size_t>(k) * op.nv_pad + v;
      acc += op.chunk_val[slot] * xi[op.chunk_col[slot]];
    }
    part[v] = acc;
  }
  for (int w = 0; w < op.wq; ++w)
    for (int bi = op.wblk_ptr[w]; bi < op.wblk_ptr[w + 1]; ++bi) {
      const int b = op.wblk[bi];
      for (int j = 0; j < 4; ++j)
        for (int lane = 0; lane < 32; ++lane) {
          const int i = b * 128 + j * 32 + lane;
          double acc = 0.0;
          for (int o = op.slice_off[b] + lane; o < op.slice_off[b + 1]; o += 128) acc += op.tile_val[o + j * 32] * xi[op.tile_col[o + j * 32]];
          if (b < op.nlong_blocks)
            for (int e = op.ovf_start[i]; e < op.ovf_start[i] + op.ovf_cnt[i]; ++e) acc += part[e];
          yi[i] += acc;   // += so a block visited twice (a packing bug) shows up
        }
    }
  for (int i = 0; i < op.n_pad; ++i)
    if (op.perm[i] >= 0) y[op.perm[i]] = yi[i];
}

}  // namespace scr
