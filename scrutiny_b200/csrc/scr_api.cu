// C ABI of scrutiny_b200 (see include/scrutiny_b200.h for the contract and reference citations).
#include "../../include/scrutiny_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "aux_kernels.cuh"
#include "dense_kernel.cuh"
#include "dense_pair_kernel.cuh"
#include "operator_pack.h"
#include "regnet_kernels.cuh"
#include "step_kernel.cuh"

using scr::Item;
using scr::PackedOperator;

namespace {
// NVTX range over one ABI call; the label carries what a lineage node's phase looks like (cells, longest cell)
struct NvtxRange {
  explicit NvtxRange(const char* name, long long a = -1, long long b = -1) {
    char buf[96];
    if (a >= 0 && b >= 0) std::snprintf(buf, sizeof(buf), "%s m=%lld max=%lld", name, a, b);
    else if (a >= 0) std::snprintf(buf, sizeof(buf), "%s m=%lld", name, a);
    else std::snprintf(buf, sizeof(buf), "%s", name);
    nvtxRangePushA(buf);
  }
  ~NvtxRange() { nvtxRangePop(); }
};
std::string g_create_error;
constexpr int kSmemDefault = 232448;   // 227 KB opt-in limit of sm_100
}

struct scr_ctx {
  int device = -1;
  int prec = SCR_F32;
  bool host_only = false;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr};
  mutable std::string err;
  int sm_count = 148, smem_optin = kSmemDefault;

  // operator
  bool has_net = false;
  PackedOperator op;
  unsigned char *d_tab = nullptr, *d_w = nullptr;
  unsigned char* d_tab_wide = nullptr;   // the 16-warp block assignment (probe launches), same size as d_tab
  std::vector<unsigned char> h_tab;   // host copy of the table blob (scr_debug_block_records)
  int tab_bytes = 0, w_bytes = 0, off_vrow = 0;
  int *d_perm = nullptr, *d_inv = nullptr;
  void *d_pc = nullptr, *d_bias = nullptr;
  int q = 1;
  bool w_smem = false;
  bool gstate = false;                       // large n: cell groups in a global scratch area (step_kernel.cuh kModeGlobal)
  size_t gbytes_probe = 0;                   // shared memory of one cell group in a PROBE launch (header with the convergence window)
  unsigned char* d_gscratch = nullptr;
  size_t gscratch_cap = 0;
  size_t gbytes = 0, fixed_bytes = 0;

  // dense (tcgen05 split-precision GEMM) path
  bool dense_mode = false;
  int dense_kind = scr::dense::kF16;           // operand encoding (dense_kernel.cuh): fp16 split (default) or 3xTF32
  float dense_acc_scale = 1.f;                 // 1 / (sx sw) of the fp16 encoding
  void *d_whi = nullptr, *d_wlo = nullptr;     // [n_pad][n_pad] hi / lo splits of W (fp16 or fp32)
  CUtensorMap tm_bhi, tm_blo;                  // B operands, boxes of the wide tile (256 genes x 128 bytes)
  CUtensorMap tm_bhi_n, tm_blo_n;              // ... of the narrow tile (64 genes x 128 bytes)
  CUtensorMap tm_bhi_r, tm_blo_r;              // ... of the resident-B tile (32 genes x 128 bytes)
  CUtensorMap tm_bhi_p, tm_blo_p;              // ... of the CTA-pair schedule (each CTA's 128 genes x 128 bytes)
  bool dense_attr_set = false;
  // working set [rows_cap][n_pad], ping-pong: state (fp32) and its operand split (hi only in the fp16 encoding)
  float* d_x[2] = {nullptr, nullptr};
  void *d_xhi[2] = {nullptr, nullptr}, *d_xlo[2] = {nullptr, nullptr};
  int64_t rows_cap = 0;
  int *d_row_cell = nullptr, *d_row_steps = nullptr, *d_row_src = nullptr, *d_tile_steps = nullptr;
  uint32_t* d_row_base = nullptr;
  int64_t* d_row_gid = nullptr;
  unsigned* d_mt_count = nullptr;              // steps-inside schedule: arrival counters per cell tile
  size_t mt_count_cap = 0;
  int64_t dense_launches = 0, sparse_launches = 0;

  // cells
  void* d_state = nullptr;
  int64_t cells = 0, cell_offset = 0;

  // scratch
  Item* d_items = nullptr;
  Item* h_items = nullptr;       // pinned staging of the same capacity: run_phase builds the groups in place
  size_t items_cap = 0;
  int* d_counter = nullptr;
  int* d_err = nullptr;
  double* d_stage = nullptr;
  size_t stage_elems = 0;
  // transfer of (genes, cells) float64 host matrices (scr_set_states / scr_get_states): two pinned host slabs and two
  // device slabs, so that the host-side strided copy of one chunk runs under the DMA + (un)pack kernel of the other
  double *h_xfer[2] = {nullptr, nullptr}, *d_xfer[2] = {nullptr, nullptr};
  size_t xfer_elems = 0;
  cudaEvent_t ev_x[2] = {nullptr, nullptr};
  // read-out scratch, kept between scr_counts calls (cudaMalloc / cudaFree of the 256 MB staging buffers
  // synchronise the device and cost more than the kernel itself)
  double* d_csize = nullptr;
  scr::CountGene* d_ctab = nullptr;          // genes of the read-out sorted by shift: {gene, state index, shift}
  float* d_lgam = nullptr;                   // lgamma(k + 1), k < scr::kLgamTable (float32 PTRS screening)
  int32_t* d_cbuf[2] = {nullptr, nullptr};
  size_t ctab_cap = 0, csize_cap = 0, cbuf_cap[2] = {0, 0};
  int32_t* d_covf = nullptr;                 // compact read-out: overflow triplets {local cell, gene, count}
  unsigned long long* d_covf_count = nullptr;
  size_t covf_cap = 0;
  bool counts_smem_optin = false;
  std::vector<double> user_bias;            // scr_set_drive_bias: additive drive term per gene (original order), empty = none
  bool readout_only = false;                 // scr_set_genes: no network, only the state / read-out entry points work

  // RegNetInference front end
  double rank_kernel_ms = 0.0;
  int64_t rank_counting = 0, rank_sorted = 0;

  // probe scratch, kept between calls (a lineage run probes every node: cudaMalloc / cudaFree synchronise the device)
  int* d_piters = nullptr;
  double* d_pnd = nullptr;
  size_t piters_cap = 0, pnd_cap = 0;
  unsigned char* d_parena = nullptr;         // dense probe: one arena for its private working set
  double* h_psum = nullptr;                  // dense probe: pinned landing buffer of a chunk's per-(step, cell) sums
  size_t h_psum_cap = 0;
  size_t parena_cap = 0;
  // scr_probe_multi: probes of sibling lineage nodes run side by side, each on its own stream with its own per-launch
  // buffers (work items, proj/const, bias, work-queue counter, outputs); installed into the fields above for the duration
  // of one probe_t call (defer_sync: launch only, the caller synchronises and collects)
  struct ProbeSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Item* d_items = nullptr; Item* h_items = nullptr; size_t items_cap = 0;
    void *d_pc = nullptr, *d_bias = nullptr;
    int* d_counter = nullptr;
    int* d_piters = nullptr; double* d_pnd = nullptr; size_t piters_cap = 0, pnd_cap = 0;
  };
  static constexpr int kProbeSlots = 8;
  ProbeSlot probe_slots[kProbeSlots];
  bool defer_sync = false;
  int probe_share = 1;
  // stepper launch configuration per kernel variant [probe][supplied]: the dynamic shared-memory opt-in and the
  // occupancy query are made once per size, not per launch
  struct LaunchCfg { size_t smem = 0; int threads = 0, per_sm = 0, mode = -1; };
  LaunchCfg launch_cfg[2][2];

  // stats
  double total_probe_ms = 0.0;
  int64_t total_probe_launches = 0;
  int64_t launches = 0;
  double last_ms = -1.0, total_step_ms = 0.0;
  int64_t total_step_launches = 0;

  size_t esize() const { return prec == SCR_F64 ? 8 : 4; }
  int cpg() const { return prec == SCR_F64 ? 2 : 4; }
};

namespace {

int fail(const scr_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}
int fail_cuda(const scr_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, SCR_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call); \
  } while (0)
#define NEED_GPU()                                                                      \
  do {                                                                                  \
    if (!ctx) return SCR_E_ARG;                                                         \
    if (ctx->host_only) return fail(ctx, SCR_E_CUDA, "host-only context: no CUDA device (there is no CPU fallback)"); \
    CK(cudaSetDevice(ctx->device));                                                     \
  } while (0)

int align16(int x) { return (x + 15) & ~15; }

// Upload from pageable host memory, ordered before everything the context launches afterwards.  A plain cudaMemcpy is NOT
// enough here: it returns once the data is staged, the DMA into device memory may still be in flight, and the context's
// streams are non-blocking (they do not synchronise with the legacy default stream the copy runs on) -- a kernel launched
// right after could read the old bytes.  The copy is therefore issued on the context's stream and waited for, which also
// orders it before the work of the context's other streams.
cudaError_t upload_sync(scr_ctx* ctx, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  return e;
}

// table blob for one block-to-warp assignment (wq warps, list bounds wblk_ptr, blocks wblk)
void build_tab(const PackedOperator& op, int wq, const std::vector<int>& wblk_ptr, const std::vector<int>& wblk,
               std::vector<unsigned char>& tab, bool large) {
  // fixed part (step_kernel.cuh "Shared-memory layout"): block records in the order of the warps' lists, list bounds;
  // variable part: chunk ranges of the long rows
  const int off_wblk_ptr = scr::rec_bytes(large), off_ovfidx = scr::fixed_tab(large);
  const int total = align16(off_ovfidx + static_cast<int>(op.ovf_start.size()) * 8);
  tab.assign(total, 0);
  for (int w = 0; w < wq; ++w) {
    int reg_end = wblk_ptr[w];
    while (reg_end < wblk_ptr[w + 1] && wblk[reg_end] < op.npp_blocks) ++reg_end;
    for (int bi = wblk_ptr[w]; bi < wblk_ptr[w + 1]; ++bi) {
      const int b = wblk[bi];
      const bool pp = b < op.npp_blocks;
      int flags = 0;
      if (pp) flags |= scr::kBlkPP;
      if (b < op.nlong_blocks) flags |= scr::kBlkLong;
      if (bi == reg_end - 1) flags |= scr::kBlkLastReg;
      if (bi == wblk_ptr[w] && pp) flags |= scr::kBlkFirstPP;
      const int4 rec = make_int4(b * 128, op.slice_off[b], op.slice_off[b + 1], flags);
      std::memcpy(tab.data() + static_cast<size_t>(bi) * 16, &rec, 16);
    }
  }
  for (size_t i = 0; i < op.ovf_start.size(); ++i) {
    int2 v = make_int2(op.ovf_start[i], op.ovf_cnt[i]);
    std::memcpy(tab.data() + off_ovfidx + i * 8, &v, 8);
  }
  std::memcpy(tab.data() + off_wblk_ptr, wblk_ptr.data(), (wq + 1) * 4);
}

template <typename T>
void build_blobs(const PackedOperator& op, std::vector<unsigned char>& tab, std::vector<unsigned char>& w,
                 int& off_vrow, bool large) {
  using WEnt = typename scr::Traits<T>::WEnt;
  build_tab(op, op.wq, op.wblk_ptr, op.wblk, tab, large);
  const size_t nslots = op.tile_val.size(), nch = op.chunk_val.size();
  off_vrow = align16(static_cast<int>(nslots * sizeof(WEnt)));
  const int wtotal = align16(off_vrow + static_cast<int>(nch * sizeof(WEnt)));
  w.assign(std::max(wtotal, 16), 0);
  for (size_t s = 0; s < nslots; ++s) {
    WEnt e{};
    e.v = static_cast<T>(op.tile_val[s] * scr::Traits<T>::kDriveScale);   // float32: drive carried as K z (step_kernel.cuh)
    e.off = op.tile_col[s] * 16;
    std::memcpy(w.data() + s * sizeof(WEnt), &e, sizeof(WEnt));
  }
  for (size_t s = 0; s < nch; ++s) {
    WEnt e{};
    e.v = static_cast<T>(op.chunk_val[s] * scr::Traits<T>::kDriveScale);
    e.off = op.chunk_col[s] * 16;
    std::memcpy(w.data() + off_vrow + s * sizeof(WEnt), &e, sizeof(WEnt));
  }
}

void free_probe_slots(scr_ctx* ctx);

void free_network(scr_ctx* ctx) {
  if (ctx->host_only) return;
  free_probe_slots(ctx);   // (their proj / const buffers are sized by the network)
  cudaFree(ctx->d_tab); cudaFree(ctx->d_tab_wide); cudaFree(ctx->d_w); cudaFree(ctx->d_perm); cudaFree(ctx->d_inv);
  ctx->d_tab_wide = nullptr;
  cudaFree(ctx->d_pc); cudaFree(ctx->d_bias); cudaFree(ctx->d_whi); cudaFree(ctx->d_wlo);
  ctx->d_tab = ctx->d_w = nullptr; ctx->d_perm = ctx->d_inv = nullptr; ctx->d_pc = ctx->d_bias = nullptr;
  ctx->d_whi = ctx->d_wlo = nullptr;
}

void free_dense_work(scr_ctx* ctx) {
  for (int i = 0; i < 2; ++i) {
    cudaFree(ctx->d_x[i]); cudaFree(ctx->d_xhi[i]); cudaFree(ctx->d_xlo[i]);
    ctx->d_x[i] = nullptr; ctx->d_xhi[i] = ctx->d_xlo[i] = nullptr;
  }
  cudaFree(ctx->d_row_cell); cudaFree(ctx->d_row_steps); cudaFree(ctx->d_row_src); cudaFree(ctx->d_tile_steps);
  cudaFree(ctx->d_row_base); cudaFree(ctx->d_row_gid);
  ctx->d_row_cell = ctx->d_row_steps = ctx->d_row_src = ctx->d_tile_steps = nullptr;
  ctx->d_row_base = nullptr; ctx->d_row_gid = nullptr;
  ctx->rows_cap = 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
// matrix [rows][cols] of fp32 (esize 4) or fp16 (esize 2) elements, row pitch ld elements, boxes of box_rows x 128 bytes,
// 128-byte swizzle
int make_tmap(scr_ctx* ctx, CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int esize) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(ctx, SCR_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esize};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(scr::dense::ROW_BYTES / esize), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, SCR_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")");
  return SCR_OK;
}

int choose_config(scr_ctx* ctx) {
  const PackedOperator& op = ctx->op;
  const size_t pc_bytes = static_cast<size_t>(op.n_pad) * 2 * ctx->esize();
  ctx->gbytes = scr::group_smem_bytes(op.n_pad, op.npp, op.nv_pad, false);
  ctx->gbytes_probe = scr::group_smem_bytes(op.n_pad, op.npp, op.nv_pad, true);
  ctx->fixed_bytes = ctx->tab_bytes + pc_bytes + 16;
  const long limit = ctx->smem_optin;
  const long q_glob = (limit - static_cast<long>(ctx->fixed_bytes)) / static_cast<long>(ctx->gbytes);
  const long q_smem = (limit - static_cast<long>(ctx->fixed_bytes) - ctx->w_bytes) / static_cast<long>(ctx->gbytes_probe);
  if (ctx->gstate) {
    // one cell group does not fit in shared memory (decided in scr_set_network_csr, wants_global_state): the groups' state
    // buffers go to global memory, shared memory keeps the tables and the group headers only (slow path, any n up to
    // kMaxBlocks * 128)
    ctx->w_smem = false;
    ctx->fixed_bytes = ctx->tab_bytes + 16;
    if (static_cast<long>(ctx->fixed_bytes) + 2 * scr::group_hdr(true) > limit)
      return fail(ctx, SCR_E_RESOURCE, "n too large: the operator tables do not fit in shared memory");
    ctx->q = std::max(1, std::min(2, 16 / op.wq));
    return SCR_OK;
  }
  if (q_glob < 1) return fail(ctx, SCR_E_RESOURCE, "internal: cell group does not fit in shared memory");
  const int maxq = std::max(1, std::min(15, 16 / op.wq));
  bool w_smem = q_smem >= 1 && q_smem >= std::min<long>(std::min<long>(q_glob, maxq), 2);
  const int force = scr::env_int("SCR_W_SMEM", -1);
  if (force == 0) w_smem = false;
  if (force == 1 && q_smem >= 1) w_smem = true;
  ctx->w_smem = w_smem;
  int q = static_cast<int>(std::min<long>(w_smem ? q_smem : q_glob, maxq));
  const int q_env = scr::env_int("SCR_GROUPS", 0);
  if (q_env >= 1) q = std::min(q, q_env);
  ctx->q = std::max(1, q);
  return SCR_OK;
}

// When the operator does not fit in shared memory next to q cell groups, the leading part of it that the rest of the
// SM's shared memory holds is staged there and the remainder is served by L1 (step_kernel.cuh "partial residency").
// SCR_W_PREFIX_KB caps the staged part (0 disables it).
int w_prefix_bytes(const scr_ctx* ctx, int q, bool probe = false) {
  if (ctx->w_smem || ctx->gstate) return 0;
  long room = static_cast<long>(ctx->smem_optin) - static_cast<long>(ctx->fixed_bytes) -
              static_cast<long>(q) * static_cast<long>(probe ? ctx->gbytes_probe : ctx->gbytes);
  const int cap_kb = scr::env_int("SCR_W_PREFIX_KB", -1);
  if (cap_kb >= 0) room = std::min<long>(room, static_cast<long>(cap_kb) * 1024);
  room = std::min<long>(room, ctx->w_bytes);
  int prefix = room >= 128 ? static_cast<int>(room / 128 * 128) : 0;
  // Any prefix beyond ~1.5 KB moves the SM to its largest shared-memory carve-out and leaves 28 KB of L1, where the part
  // of the operator that is NOT staged has to stay resident next to the state traffic: measured on the bench network
  // (profiles/r2_experiments.txt), a rest of <= 20 KB runs at full speed, 24 KB at -20 %, while no prefix at all (60 KB of
  // L1 for a 54 KB operator) is as fast as the full one.  So: stage the head only if the rest then fits L1; an operator that
  // fits the larger L1 whole is left there; one that fits neither keeps the prefix (fewer trips to L2).
  if (cap_kb < 0 && prefix > 0 && ctx->w_bytes - prefix > 20 * 1024 && ctx->w_bytes <= 56 * 1024) prefix = 0;
  return prefix;
}

template <typename T, bool PROBE, bool SUPPLIED>
int launch_step_t(scr_ctx* ctx, scr::StepParams<T>& p, int nitems) {
  // few groups: spread them over SMs instead of stacking them in one CTA
  int q = std::min(ctx->q, std::max(1, (nitems + ctx->sm_count - 1) / ctx->sm_count));
  if (PROBE && q == 1 && ctx->d_tab_wide) {   // few probe groups: 16 warps per group, half the blocks per warp and step
    p.tab = ctx->d_tab_wide;
    p.wq = 16;
    p.wq_log2 = 4;
  }
  const size_t per_group = ctx->gstate ? static_cast<size_t>(scr::group_hdr(PROBE)) : (PROBE ? ctx->gbytes_probe : ctx->gbytes);
  if (!ctx->gstate)   // a probe group carries a larger header: fewer of them may fit
    while (q > 1 && ctx->fixed_bytes + (ctx->w_smem ? ctx->w_bytes : 0) + static_cast<size_t>(q) * per_group > static_cast<size_t>(ctx->smem_optin)) --q;
  p.q = q;
  p.w_prefix = w_prefix_bytes(ctx, q, PROBE);
  const size_t smem = ctx->fixed_bytes + (ctx->w_smem ? ctx->w_bytes : 0) + static_cast<size_t>(q) * per_group + p.w_prefix;
  const int threads = q * p.wq * 32;
  auto kern = ctx->gstate ? scr::step_kernel<T, scr::kModeGlobal, PROBE, SUPPLIED>
                          : (ctx->w_smem ? scr::step_kernel<T, scr::kModeSmem, PROBE, SUPPLIED> : scr::step_kernel<T, scr::kModeL1, PROBE, SUPPLIED>);
  scr_ctx::LaunchCfg& cfg = ctx->launch_cfg[PROBE ? 1 : 0][SUPPLIED ? 1 : 0];
  const int mode = ctx->gstate ? scr::kModeGlobal : (ctx->w_smem ? scr::kModeSmem : scr::kModeL1);
  if (cfg.smem != smem || cfg.threads != threads || cfg.mode != mode) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    cfg.smem = smem; cfg.threads = threads; cfg.per_sm = occ; cfg.mode = mode;
  }
  const int per_sm = cfg.per_sm;
  if (per_sm < 1) return fail(ctx, SCR_E_RESOURCE, "stepper does not fit on an SM");
  const int grid = std::max(1, std::min((nitems + q - 1) / q, ctx->sm_count * per_sm));
  if (ctx->gstate) {
    const size_t need = static_cast<size_t>(grid) * q * scr::group_state_bytes(ctx->op.n_pad, ctx->op.npp, ctx->op.nv_pad);
    if (need > ctx->gscratch_cap) {
      cudaFree(ctx->d_gscratch);
      ctx->d_gscratch = nullptr;
      ctx->gscratch_cap = 0;
      CK(cudaMalloc(&ctx->d_gscratch, need));
      ctx->gscratch_cap = need;
    }
    p.gscratch = ctx->d_gscratch;
  }
  CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
#if SCR_DEBUG_CHECKS
  CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  p.dbg = ctx->d_err;
  p.ncells = ctx->cells;
  if (p.nrows == 0) p.nrows = 0x7fffffffffffLL;
#endif
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  kern<<<grid, threads, smem, ctx->stream>>>(p);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->launches++;
  if (ctx->defer_sync) return SCR_OK;   // scr_probe_multi: the caller synchronises, times and collects
  CK(cudaStreamSynchronize(ctx->stream));
#if SCR_DEBUG_CHECKS
  {
    int code = 0;
    CK(cudaMemcpy(&code, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (code) return fail(ctx, SCR_E_CUDA, "stepper bounds check " + std::to_string(code) + " failed (SCR_DEBUG_CHECKS build)");
  }
#endif
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_ms = ms;
  if (!PROBE) { ctx->total_step_ms += ms; ctx->total_step_launches++; ctx->sparse_launches++; }
  else { ctx->total_probe_ms += ms; ctx->total_probe_launches++; }
  return SCR_OK;
}

template <typename T>
int launch_step(scr_ctx* ctx, scr::StepParams<T>& p, int nitems, bool probe, bool supplied) {
  if (probe) return supplied ? launch_step_t<T, true, true>(ctx, p, nitems) : launch_step_t<T, true, false>(ctx, p, nitems);
  return supplied ? launch_step_t<T, false, true>(ctx, p, nitems) : launch_step_t<T, false, false>(ctx, p, nitems);
}

template <typename T>
void fill_common(scr_ctx* ctx, scr::StepParams<T>& p) {
  const PackedOperator& op = ctx->op;
  std::memset(&p, 0, sizeof(p));
  p.tab = ctx->d_tab; p.wblob = ctx->d_w; p.tab_bytes = ctx->tab_bytes; p.w_bytes = ctx->w_bytes;
  p.off_vrow = ctx->off_vrow;
  p.n = op.n; p.n_pad = op.n_pad; p.npp = op.npp; p.nb = op.nb; p.npp_blocks = op.npp_blocks;
  p.nlong_blocks = op.nlong_blocks; p.nv = op.nv; p.nv_pad = op.nv_pad; p.lv = op.lv;
  p.wq = op.wq; p.wq_log2 = op.wq_log2; p.q = ctx->q;
  p.pc = static_cast<const unsigned char*>(ctx->d_pc);
  p.state = static_cast<T*>(ctx->d_state);
  p.ld = op.n_pad;
  p.counter = ctx->d_counter;
  p.cell_offset = ctx->cell_offset;
  p.perm = ctx->d_perm;
}

// proj/const (+mu) and the -mu*rowsum bias of one node, internal order; drive_scale = Traits<T>::kDriveScale for the
// sparse stepper (its drive is carried scaled), 1 for the GEMM stepper
template <typename T>
int upload_node(scr_ctx* ctx, const double* proj, const double* cnst, double mu, scr::StepParams<T>& p, double drive_scale) {
  const PackedOperator& op = ctx->op;
  std::vector<T> pc(static_cast<size_t>(op.n_pad) * 2, T(0));
  for (int i = 0; i < op.n; ++i) {
    pc[2 * i] = static_cast<T>(proj[op.perm[i]]);
    pc[2 * i + 1] = static_cast<T>(cnst[op.perm[i]] + mu);
  }
  CK(cudaMemcpyAsync(ctx->d_pc, pc.data(), pc.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  p.bias = nullptr;
  if (mu != 0.0 || !ctx->user_bias.empty()) {
    std::vector<T> bias(op.n_pad, T(0));
    for (int i = 0; i < op.n; ++i)
      bias[i] = static_cast<T>((-mu * op.rowsum[i] + (ctx->user_bias.empty() ? 0.0 : ctx->user_bias[op.perm[i]])) * drive_scale);
    CK(cudaMemcpyAsync(ctx->d_bias, bias.data(), bias.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    p.bias = static_cast<const T*>(ctx->d_bias);
  }
  // (copies from pageable host memory return once the source has been staged: the vectors may go out of scope)
  return SCR_OK;
}

template <typename P>
int ensure_buf(scr_ctx* ctx, P*& ptr, size_t& cap, size_t elems) {
  if (elems <= cap) return SCR_OK;
  cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  CK(cudaMalloc(&ptr, elems * sizeof(P)));
  cap = elems;
  return SCR_OK;
}

int ensure_items(scr_ctx* ctx, size_t count) {
  if (count <= ctx->items_cap) return SCR_OK;
  cudaFree(ctx->d_items);
  cudaFreeHost(ctx->h_items);
  ctx->d_items = nullptr;
  ctx->h_items = nullptr;
  ctx->items_cap = 0;
  const size_t cap = std::max<size_t>(count + count / 8, 1024);
  CK(cudaMalloc(&ctx->d_items, cap * sizeof(Item)));
  CK(cudaHostAlloc(&ctx->h_items, cap * sizeof(Item), cudaHostAllocDefault));
  ctx->items_cap = cap;
  return SCR_OK;
}

template <typename T>
int upload_noise(scr_ctx* ctx, const double* noise, size_t count, T** d_noise) {
  std::vector<T> h(count);
  for (size_t i = 0; i < count; ++i) h[i] = static_cast<T>(noise[i]);
  CK(cudaMalloc(d_noise, std::max<size_t>(count, 1) * sizeof(T)));
  CK(upload_sync(ctx, *d_noise, h.data(), count * sizeof(T)));
  return SCR_OK;
}

template <typename T>
int run_phase_t(scr_ctx* ctx, int64_t m, const int64_t* cell_ids, const int32_t* steps, const int64_t* step_base,
                const double* proj, const double* cnst, double dt, double sigma, double mu, uint64_t seed,
                const double* noise, int64_t noise_steps) {
  constexpr int CPG = scr::Traits<T>::CPG, H = CPG / 2;
  // Longest first (counting sort on the step count), then groups of CPG cells, built straight in the pinned staging
  // buffer.  Parity placement (step_kernel.cuh, noise stream v2): cells with an even GLOBAL id fill slots [0, H) of the
  // groups, odd ones slots [H, CPG), each parity class in its own sorted sequence -- position k of a class goes to slot
  // (k % H) of its half of group (k / H).  Each (class, bucket) writes sequentially, so this is a ~1000-way streaming
  // scatter.  Large phases split the cell list over a few host threads (per-thread histograms -> per-thread start
  // positions), which keeps the sequential order exactly.
  const int nthr = m >= (1 << 16) ? std::max(1, std::min(scr::env_int("SCR_HOST_THREADS", 4), 16)) : 1;
  auto span = [&](int t) { return std::make_pair(m * t / nthr, m * (t + 1) / nthr); };
  auto parallel = [&](auto&& fn) {
    std::vector<std::thread> th;
    int started = 1;
    try {
      for (int t = 1; t < nthr; ++t) { th.emplace_back(fn, t); ++started; }
    } catch (...) {}                                     // no exception crosses the C ABI: unstarted spans run here
    fn(0);
    for (int t = started; t < nthr; ++t) fn(t);
    for (auto& x : th) x.join();
  };
  std::vector<int> t_smax(nthr, 0), t_bad(nthr, 0);
  // "each cell at most once" (a duplicate would race on its state row): one bit per local slot, set atomically
  std::vector<uint64_t> seen((static_cast<size_t>(ctx->cells) + 63) / 64, 0);
  parallel([&](int t) {
    int mx = 0, bad = 0;
    for (int64_t i = span(t).first; i < span(t).second; ++i) {
      if (cell_ids[i] < 0 || cell_ids[i] >= ctx->cells) {
        bad |= 1;
      } else {
        const uint64_t bit = 1ULL << (cell_ids[i] & 63);
        if (__atomic_fetch_or(&seen[static_cast<size_t>(cell_ids[i]) >> 6], bit, __ATOMIC_RELAXED) & bit) bad |= 4;
      }
      if (steps[i] < 0) bad |= 2;
      mx = std::max(mx, steps[i]);
    }
    t_smax[t] = mx;
    t_bad[t] = bad;
  });
  int smax = 0, bad = 0;
  for (int t = 0; t < nthr; ++t) { smax = std::max(smax, t_smax[t]); bad |= t_bad[t]; }
  if (bad & 1) return fail(ctx, SCR_E_ARG, "cell id out of range");
  if (bad & 2) return fail(ctx, SCR_E_ARG, "negative step count");
  if (bad & 4) return fail(ctx, SCR_E_ARG, "cell id listed twice in one phase");
  if (noise && smax > noise_steps) return fail(ctx, SCR_E_ARG, "supplied noise shorter than the longest cell");
  if (smax == 0) return SCR_OK;
  const size_t nbuckets = static_cast<size_t>(smax);                 // bucket b <-> step count smax - b
  const int64_t par0 = ctx->cell_offset & 1;                         // parity of the global id of local slot 0
  auto klass = [&](int64_t i) { return static_cast<size_t>((cell_ids[i] + par0) & 1); };
  std::vector<int64_t> pos(static_cast<size_t>(nthr) * 2 * nbuckets, 0);  // pos[(t * 2 + class) * nbuckets + b]
  parallel([&](int t) {
    int64_t* h = pos.data() + static_cast<size_t>(t) * 2 * nbuckets;
    for (int64_t i = span(t).first; i < span(t).second; ++i) if (steps[i] > 0) h[klass(i) * nbuckets + (smax - steps[i])]++;
  });
  int64_t live[2] = {0, 0};
  for (size_t k = 0; k < 2; ++k)
    for (size_t b = 0; b < nbuckets; ++b)
      for (int t = 0; t < nthr; ++t) {
        int64_t& h = pos[(static_cast<size_t>(t) * 2 + k) * nbuckets + b];
        const int64_t cnt = h;
        h = live[k];
        live[k] += cnt;
      }
  const int64_t nitems = (std::max(live[0], live[1]) + H - 1) / H;
  if (nitems > 0x7fffffff) return fail(ctx, SCR_E_ARG, "too many cell groups");
  int rc = ensure_items(ctx, static_cast<size_t>(nitems));
  if (rc) return rc;
  Item* items = ctx->h_items;
  parallel([&](int t) {
    int64_t* next = pos.data() + static_cast<size_t>(t) * 2 * nbuckets;
    for (int64_t i = span(t).first; i < span(t).second; ++i) {
      if (steps[i] <= 0) continue;
      const size_t kl = klass(i);
      const int64_t k = next[kl * nbuckets + (smax - steps[i])]++;
      Item& it = items[k / H];
      const int c = static_cast<int>(kl) * H + static_cast<int>(k % H);
      it.cell[c] = static_cast<int>(cell_ids[i]);
      it.steps[c] = steps[i];
      it.base[c] = static_cast<uint32_t>(step_base ? step_base[i] : 0);
      it.row[c] = static_cast<int>(i);
    }
  });
  for (int kl = 0; kl < 2; ++kl)
    for (int64_t k = live[kl]; k < nitems * H; ++k) {   // empty slots of this class
      Item& it = items[k / H];
      const int c = kl * H + static_cast<int>(k % H);
      it.cell[c] = -1;
      it.steps[c] = 0;
      it.base[c] = 0;
      it.row[c] = 0;
    }
  parallel([&](int t) {
    for (int64_t g = nitems * t / nthr; g < nitems * (t + 1) / nthr; ++g) {
      items[g].scale = 1.0;
      items[g].flags = scr::kItemParity;
      items[g].pad_ = 0;
      if (CPG < 4)
        for (int c = CPG; c < 4; ++c) { items[g].cell[c] = -1; items[g].steps[c] = 0; items[g].base[c] = 0; items[g].row[c] = 0; }
    }
  });
  CK(cudaMemcpyAsync(ctx->d_items, items, static_cast<size_t>(nitems) * sizeof(Item), cudaMemcpyHostToDevice, ctx->stream));

  scr::StepParams<T> p;
  fill_common(ctx, p);
  rc = upload_node<T>(ctx, proj, cnst, mu, p, scr::Traits<T>::kDriveScale);
  if (rc) return rc;
  p.items = ctx->d_items;
  p.nitems = static_cast<int>(nitems);
  p.nrows = m;
  p.dt = static_cast<T>(dt);
  p.rscale = static_cast<float>(-2.0 * 0.6931471805599453 * sigma * sigma * scr::Traits<T>::kDriveScale * scr::Traits<T>::kDriveScale);   // radius of K sigma xi
  scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(SCR_STREAM_STEP), p.rk);
  T* d_noise = nullptr;
  if (noise) {
    rc = upload_noise<T>(ctx, noise, static_cast<size_t>(m) * noise_steps * ctx->op.n, &d_noise);
    if (rc) return rc;
    p.noise = d_noise;
    p.noise_steps = noise_steps;
  }
  rc = launch_step<T>(ctx, p, static_cast<int>(nitems), false, noise != nullptr);
  cudaFree(d_noise);
  return rc;
}


// ---- dense path: tcgen05 GEMM stepper over the phase's working set (dense_kernel.cuh) ----
// Schedule per launch: when every (cell tile, gene tile) pair can have its own CTA the launch runs all remaining steps
// ("steps inside": small batches -- config c1, every probe -- take one launch per phase instead of one per timestep; narrow
// 128 x 64 tiles if those fit the SM count, else wide 128 x 256 ones); otherwise one launch per timestep of wide tiles.
struct DenseSchedule { bool narrow, inside, resb, pair; int stages; };
DenseSchedule dense_schedule(const scr_ctx* ctx, int m_tiles, int ld) {
  namespace D = scr::dense;
  const int force_narrow = scr::env_int("SCR_DENSE_NARROW", -1), force_inside = scr::env_int("SCR_DENSE_INSIDE", -1);
  const int t_narrow = m_tiles * (ld / D::BN_NARROW), t_wide = m_tiles * (ld / D::BN), t_res = m_tiles * (ld / D::BN_RES);
  DenseSchedule sc;
  sc.resb = false;
  sc.stages = 0;
  sc.pair = false;
  sc.narrow = force_narrow >= 0 ? force_narrow != 0 : t_narrow <= ctx->sm_count;
  sc.inside = (sc.narrow ? t_narrow : t_wide) <= ctx->sm_count && force_inside != 0;
  // resident-B tiles: W's slice of the CTA stays in shared memory for the whole launch (small networks)
  if (sc.inside && force_narrow < 0 && t_res <= ctx->sm_count && scr::env_int("SCR_DENSE_RESB", 1)) {
    const int k_blocks = ld * (ctx->dense_kind == D::kF16 ? 2 : 4) / D::ROW_BYTES;
    int stages = D::kMaxStages;
    while (stages >= 2 && D::res_smem_bytes(k_blocks, stages) > ctx->smem_optin) --stages;
    if (stages >= 2) { sc.resb = true; sc.narrow = false; sc.stages = std::min(stages, k_blocks); }
  }
  // large batches, one launch per timestep of wide tiles: CTA pairs (cta_group::2) halve the W bytes each CTA stages
  sc.pair = !sc.inside && !sc.narrow && scr::env_int("SCR_DENSE_PAIR", 1) != 0;
  return sc;
}
// the working set of one phase / probe: state and operand split, ping-pong, with the A-operand tensor maps
struct DenseWork {
  float* x[2];
  void *hi[2], *lo[2];
  CUtensorMap tm_hi[2], tm_lo[2];
};
int dense_make_maps(scr_ctx* ctx, DenseWork& w, int64_t m_pad, int ld) {
  namespace D = scr::dense;
  const bool f16 = ctx->dense_kind == D::kF16;
  for (int i = 0; i < 2; ++i) {
    int rc = make_tmap(ctx, &w.tm_hi[i], f16 ? w.hi[i] : static_cast<void*>(w.x[i]), m_pad, ld, ld, D::BM, f16 ? 2 : 4);
    if (!rc) rc = make_tmap(ctx, &w.tm_lo[i], w.lo[i], m_pad, ld, ld, D::BM, f16 ? 2 : 4);
    if (rc) return rc;
  }
  return SCR_OK;
}
size_t dense_split_bytes(const scr_ctx* ctx, int64_t m_pad, int ld) {
  return static_cast<size_t>(m_pad) * ld * (ctx->dense_kind == scr::dense::kF16 ? 2 : 4);
}
int dense_set_attrs(scr_ctx* ctx) {
  namespace D = scr::dense;
  if (ctx->dense_attr_set) return SCR_OK;
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN, D::kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::Tile<D::BN>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN_NARROW, D::kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::Tile<D::BN_NARROW>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN, D::kTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::Tile<D::BN>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN_NARROW, D::kTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::Tile<D::BN_NARROW>::SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_pair_kernel<D::kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::PAIR_SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_pair_kernel<D::kTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::PAIR_SMEM_BYTES));
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN_RES, D::kF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
  CK(cudaFuncSetAttribute(D::dense_step_kernel<D::BN_RES, D::kTF32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
  ctx->dense_attr_set = true;
  return SCR_OK;
}
// one launch: steps p.s0 .. p.s0 + p.nsteps - 1 of the first m_tiles cell tiles
int dense_launch(scr_ctx* ctx, DenseSchedule sc, int m_tiles, const DenseWork& w, scr::dense::Params& p) {
  namespace D = scr::dense;
  p.m_tiles = m_tiles;
  p.n_tiles = p.ld / (sc.resb ? D::BN_RES : sc.narrow ? D::BN_NARROW : D::BN);
  p.stages = sc.stages;
  const int tiles = m_tiles * p.n_tiles;
  if (sc.inside) {
    if (tiles > ctx->sm_count) return fail(ctx, SCR_E_STATE, "steps-inside schedule needs one CTA per tile");
    if (static_cast<size_t>(m_tiles) > ctx->mt_count_cap) {
      cudaFree(ctx->d_mt_count);
      ctx->d_mt_count = nullptr;
      ctx->mt_count_cap = 0;
      CK(cudaMalloc(&ctx->d_mt_count, static_cast<size_t>(ctx->sm_count) * sizeof(unsigned)));
      ctx->mt_count_cap = ctx->sm_count;
    }
    CK(cudaMemsetAsync(ctx->d_mt_count, 0, static_cast<size_t>(m_tiles) * sizeof(unsigned), ctx->stream));
    p.mt_count = ctx->d_mt_count;
  } else {
    p.mt_count = nullptr;
    p.nsteps = 1;
  }
  p.err = ctx->d_err;
  p.acc_scale = ctx->dense_acc_scale;
#if SCR_DENSE_TRACE
  {
    static unsigned long long* d_trace = nullptr;
    if (!d_trace) cudaMalloc(&d_trace, 4096 * 8 * sizeof(unsigned long long));
    cudaMemsetAsync(d_trace, 0, 4096 * 8 * sizeof(unsigned long long), ctx->stream);
    p.trace = d_trace;
  }
#endif
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(std::min(tiles, ctx->sm_count)));
  cfg.blockDim = dim3(D::NUM_THREADS);
  cfg.dynamicSmemBytes = sc.resb ? D::res_smem_bytes(p.k_blocks, sc.stages) : sc.narrow ? D::Tile<D::BN_NARROW>::SMEM_BYTES : D::Tile<D::BN>::SMEM_BYTES;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of the CTAs that wait on each other
  attr[0].val.cooperative = sc.inside ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (sc.pair) {   // clusters of two CTAs (compile-time cluster dimensions), one 256 x 256 tile per pair at a time
    const int pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
    cfg.gridDim = dim3(static_cast<unsigned>(2 * std::min(pair_tiles, ctx->sm_count / 2)));
    cfg.dynamicSmemBytes = D::PAIR_SMEM_BYTES;
    cfg.numAttrs = 0;
    const cudaError_t ep = ctx->dense_kind == D::kF16
        ? cudaLaunchKernelEx(&cfg, D::dense_pair_kernel<D::kF16>, w.tm_hi[0], w.tm_hi[1], w.tm_lo[0], w.tm_lo[1], ctx->tm_bhi_p, ctx->tm_blo_p, p)
        : cudaLaunchKernelEx(&cfg, D::dense_pair_kernel<D::kTF32>, w.tm_hi[0], w.tm_hi[1], w.tm_lo[0], w.tm_lo[1], ctx->tm_bhi_p, ctx->tm_blo_p, p);
    if (ep != cudaSuccess) return fail_cuda(ctx, ep, "dense stepper launch (CTA pairs)");
    ctx->launches++;
    ctx->dense_launches++;
    return SCR_OK;
  }
  const CUtensorMap &bhi = sc.resb ? ctx->tm_bhi_r : sc.narrow ? ctx->tm_bhi_n : ctx->tm_bhi;
  const CUtensorMap &blo = sc.resb ? ctx->tm_blo_r : sc.narrow ? ctx->tm_blo_n : ctx->tm_blo;
  cudaError_t e;
#define SCR_DENSE_GO(BN_, KIND_, RESB_) \
  cudaLaunchKernelEx(&cfg, D::dense_step_kernel<BN_, KIND_, RESB_>, w.tm_hi[0], w.tm_hi[1], w.tm_lo[0], w.tm_lo[1], bhi, blo, p)
  if (ctx->dense_kind == D::kF16)
    e = sc.resb ? SCR_DENSE_GO(D::BN_RES, D::kF16, true) : sc.narrow ? SCR_DENSE_GO(D::BN_NARROW, D::kF16, false) : SCR_DENSE_GO(D::BN, D::kF16, false);
  else
    e = sc.resb ? SCR_DENSE_GO(D::BN_RES, D::kTF32, true) : sc.narrow ? SCR_DENSE_GO(D::BN_NARROW, D::kTF32, false) : SCR_DENSE_GO(D::BN, D::kTF32, false);
#undef SCR_DENSE_GO
  if (e != cudaSuccess) return fail_cuda(ctx, e, "dense stepper launch");
  ctx->launches++;
  ctx->dense_launches++;
#if SCR_DENSE_TRACE
  if (const char* path = std::getenv("SCR_DENSE_TRACE")) {   // developer build: per-step time stamps of CTA 0 of this launch
    if (sc.inside && p.trace) {
      const int ns = std::min(p.nsteps, 4096);
      std::vector<unsigned long long> h(static_cast<size_t>(ns) * 8);
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpy(h.data(), p.trace, h.size() * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = std::fopen(path, "a")) {
        std::fprintf(f, "# launch m_tiles=%d n_tiles=%d nsteps=%d\n", m_tiles, p.n_tiles, p.nsteps);
        for (int i = 0; i < ns; ++i) {
          for (int k = 0; k < 8; ++k) std::fprintf(f, "%llu ", h[i * 8 + k]);
          std::fprintf(f, "\n");
        }
        std::fclose(f);
      }
    }
  }
#endif
  return SCR_OK;
}
void dense_gather(scr_ctx* ctx, int64_t m_pad, const int* d_cell, int live, const DenseWork& w, int ld) {
  namespace D = scr::dense;
  if (ctx->dense_kind == D::kF16)
    D::gather_rows_kernel<D::kF16><<<static_cast<unsigned>(m_pad), 256, 0, ctx->stream>>>(
        static_cast<const float*>(ctx->d_state), ld, d_cell, live, w.x[0], w.hi[0], w.lo[0], ld, ctx->d_err);
  else
    D::gather_rows_kernel<D::kTF32><<<static_cast<unsigned>(m_pad), 256, 0, ctx->stream>>>(
        static_cast<const float*>(ctx->d_state), ld, d_cell, live, w.x[0], nullptr, w.lo[0], ld, ctx->d_err);
  ctx->launches++;
}
// the error word after a dense phase / probe has been synchronised
int dense_check_err(scr_ctx* ctx) {
  int code = 0;
  CK(cudaMemcpy(&code, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (code == scr::dense::kErrDenseRange)
    return fail(ctx, SCR_E_RESOURCE, "dense stepper: a state left the fp16 operand range (|x| >= 1023); SCR_DENSE_KIND=tf32 selects the tf32 encoding");
  if (code) return fail(ctx, SCR_E_CUDA, "dense stepper: device error word " + std::to_string(code));
  return SCR_OK;
}

int run_phase_dense(scr_ctx* ctx, int64_t m, const int64_t* cell_ids, const int32_t* steps, const int64_t* step_base,
                    const double* proj, const double* cnst, double dt, double sigma, double mu, uint64_t seed,
                    const double* noise, int64_t noise_steps) {
  namespace D = scr::dense;
  const PackedOperator& op = ctx->op;
  const int n = op.n, ld = op.n_pad;
  int smax = 0;
  for (int64_t i = 0; i < m; ++i) {
    if (cell_ids[i] < 0 || cell_ids[i] >= ctx->cells) return fail(ctx, SCR_E_ARG, "cell id out of range");
    if (steps[i] < 0) return fail(ctx, SCR_E_ARG, "negative step count");
    smax = std::max(smax, steps[i]);
  }
  if (noise && smax > noise_steps) return fail(ctx, SCR_E_ARG, "supplied noise shorter than the longest cell");
  if (smax == 0) return SCR_OK;
  // rows sorted by step count, longest first: tile t is finished after steps[t*128] steps
  std::vector<int64_t> bucket(static_cast<size_t>(smax) + 2, 0);
  for (int64_t i = 0; i < m; ++i) if (steps[i] > 0) bucket[smax - steps[i] + 1]++;
  for (int s = 0; s <= smax; ++s) bucket[s + 1] += bucket[s];
  const int64_t live = bucket[smax];
  std::vector<int64_t> order(live);
  {
    std::vector<int64_t> pos(bucket.begin(), bucket.end() - 1);
    for (int64_t i = 0; i < m; ++i) if (steps[i] > 0) order[pos[smax - steps[i]]++] = i;
  }
  const int m_tiles = static_cast<int>((live + D::BM - 1) / D::BM);
  const int64_t m_pad = static_cast<int64_t>((m_tiles + 1) / 2) * 2 * D::BM;   // whole CTA-pair tiles (256 rows)
  std::vector<int> h_cell(m_pad, -1), h_steps(m_pad, 0), h_src(m_pad, 0), h_tile(m_pad / D::BM, 0);
  std::vector<uint32_t> h_base(m_pad, 0);
  std::vector<int64_t> h_gid(m_pad, 0);
  for (int64_t r = 0; r < live; ++r) {
    const int64_t i = order[r];
    h_cell[r] = static_cast<int>(cell_ids[i]);
    h_steps[r] = steps[i];
    h_src[r] = static_cast<int>(i);
    h_base[r] = static_cast<uint32_t>(step_base ? step_base[i] : 0);
    h_gid[r] = ctx->cell_offset + cell_ids[i];
  }
  for (int t = 0; t < m_tiles; ++t) h_tile[t] = h_steps[static_cast<size_t>(t) * D::BM];
  if (m_pad > ctx->rows_cap) {
    free_dense_work(ctx);
    const size_t bytes = static_cast<size_t>(m_pad) * ld * 4, sbytes = dense_split_bytes(ctx, m_pad, ld);
    for (int i = 0; i < 2; ++i) {
      CK(cudaMalloc(&ctx->d_x[i], bytes));
      CK(cudaMalloc(&ctx->d_xlo[i], sbytes));
      if (ctx->dense_kind == D::kF16) CK(cudaMalloc(&ctx->d_xhi[i], sbytes));
    }
    CK(cudaMalloc(&ctx->d_row_cell, m_pad * 4)); CK(cudaMalloc(&ctx->d_row_steps, m_pad * 4));
    CK(cudaMalloc(&ctx->d_row_src, m_pad * 4)); CK(cudaMalloc(&ctx->d_tile_steps, (m_pad / D::BM) * 4));
    CK(cudaMalloc(&ctx->d_row_base, m_pad * 4)); CK(cudaMalloc(&ctx->d_row_gid, m_pad * 8));
    ctx->rows_cap = m_pad;
  }
  CK(cudaMemcpyAsync(ctx->d_row_cell, h_cell.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_row_steps, h_steps.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_row_src, h_src.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_tile_steps, h_tile.data(), h_tile.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_row_base, h_base.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_row_gid, h_gid.data(), m_pad * 8, cudaMemcpyHostToDevice, ctx->stream));

  scr::StepParams<float> sp;           // reuse the node upload (proj/const/bias in identity order)
  fill_common(ctx, sp);
  int rc = upload_node<float>(ctx, proj, cnst, mu, sp, 1.0);
  if (rc) return rc;

  DenseWork w;
  for (int i = 0; i < 2; ++i) { w.x[i] = ctx->d_x[i]; w.hi[i] = ctx->d_xhi[i]; w.lo[i] = ctx->d_xlo[i]; }
  if ((rc = dense_make_maps(ctx, w, m_pad, ld))) return rc;
  float* d_noise = nullptr;
  struct FreeOnExit { float*& ptr; ~FreeOnExit() { cudaFree(ptr); } } free_noise{d_noise};   // every return below frees it
  if (noise) {
    rc = upload_noise<float>(ctx, noise, static_cast<size_t>(m) * noise_steps * n, &d_noise);
    if (rc) return rc;
  }
  D::Params p;
  std::memset(&p, 0, sizeof(p));
  for (int i = 0; i < 2; ++i) { p.x[i] = w.x[i]; p.hi[i] = w.hi[i]; p.lo[i] = w.lo[i]; }
  p.ld = ld; p.k_blocks = ld * (ctx->dense_kind == D::kF16 ? 2 : 4) / D::ROW_BYTES; p.n = n; p.m_pad = static_cast<int>(m_pad);
  p.pc = static_cast<const float2*>(ctx->d_pc);
  p.bias = sp.bias;
  p.row_steps = ctx->d_row_steps; p.row_base = ctx->d_row_base; p.row_gid = ctx->d_row_gid; p.row_src = ctx->d_row_src;
  p.dt = static_cast<float>(dt);
  p.rscale = static_cast<float>(-2.0 * 0.6931471805599453 * sigma * sigma);
  scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(SCR_STREAM_STEP), p.rk);
  p.noise = d_noise; p.noise_steps = noise_steps;

  if ((rc = dense_set_attrs(ctx))) return rc;
  CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  dense_gather(ctx, m_pad, ctx->d_row_cell, static_cast<int>(live), w, ld);
  CK(cudaGetLastError());
  int active_tiles = m_tiles;
  int64_t launches = 0;
  for (int s = 0; s < smax;) {
    while (active_tiles > 0 && h_tile[active_tiles - 1] <= s) --active_tiles;
    const DenseSchedule sc = dense_schedule(ctx, active_tiles, ld);
    p.s0 = s;
    p.nsteps = sc.inside ? smax - s : 1;   // (tiles leave a steps-inside launch on their own when their rows are done)
    if ((rc = dense_launch(ctx, sc, active_tiles, w, p))) return rc;
    ++launches;
    s += p.nsteps;
  }
  D::scatter_rows_kernel<<<static_cast<unsigned>(m_pad), 256, 0, ctx->stream>>>(
      static_cast<float*>(ctx->d_state), ld, ctx->d_row_cell, static_cast<int>(live), ctx->d_x[0], ctx->d_x[1], ctx->d_tile_steps, ld);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "dense stepper");
  if ((rc = dense_check_err(ctx))) return rc;
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_ms = ms;
  ctx->total_step_ms += ms;
  ctx->total_step_launches += launches;
  return SCR_OK;
}


// ---- dense probe: the sampled cells are stepped max_iter times through the GEMM stepper (copies, the
// states are not written back); each step leaves per-row sums of |x' - x|; the stopping rule of
// MatriSeq.py:657-668 is then applied per cell on the host (a converged cell's later steps are
// simply ignored, which is what stopping it would have produced).  Steps run in chunks so that the
// launch loop ends as soon as every cell has converged.
int probe_dense(scr_ctx* ctx, int32_t k, const int64_t* sample_ids, const double* w_div, const double* proj,
                const double* cnst, double dt, double sigma, double mu, double thresh, int32_t max_iter,
                int32_t avg_num, uint64_t seed, uint32_t stream, const double* noise, int32_t* iters_out,
                double* normdiff_out) {
  namespace D = scr::dense;
  const PackedOperator& op = ctx->op;
  const int n = op.n, ld = op.n_pad;
  const int64_t m_pad = (static_cast<int64_t>(k) + D::BM - 1) / D::BM * D::BM;
  const int m_tiles = static_cast<int>(m_pad / D::BM);
  DenseSchedule sc = dense_schedule(ctx, m_tiles, ld);
  sc.pair = false;   // (the pair kernel has no probe outputs: a probe too large for the steps-inside schedule takes single-CTA wide tiles)
  const int n_parts = 2 * (ld / (sc.resb ? D::BN_RES : sc.narrow ? D::BN_NARROW : D::BN));   // column parts of a row's |x' - x| sum
  std::vector<int> h_cell(m_pad, -1), h_steps(m_pad, 0), h_src(m_pad, 0);
  std::vector<uint32_t> h_base(m_pad, 0);
  std::vector<int64_t> h_gid(m_pad, 0);
  std::vector<float> h_scale(m_pad, 1.f);
  for (int i = 0; i < k; ++i) {
    if (sample_ids[i] < 0 || sample_ids[i] >= ctx->cells) return fail(ctx, SCR_E_ARG, "sample id out of range");
    h_cell[i] = static_cast<int>(sample_ids[i]);
    h_steps[i] = max_iter;
    h_src[i] = i;
    h_gid[i] = ctx->cell_offset + sample_ids[i];
    if (w_div) h_scale[i] = static_cast<float>(1.0 / w_div[i]);
  }
  // every tile steps until the chunk loop stops: its first row carries the tile's step count
  for (int t = 0; t < m_tiles; ++t) h_steps[static_cast<size_t>(t) * D::BM] = max_iter;
  // private working set (the phase working set of scr_run_phase is left alone)
  DenseWork w;
  float *d_scale = nullptr, *d_diff = nullptr, *d_noise = nullptr;
  int *d_cell = nullptr, *d_steps = nullptr, *d_src = nullptr;
  uint32_t* d_base = nullptr;
  int64_t* d_gid = nullptr;
  const int chunk = 64;
  const size_t diff_per_step = static_cast<size_t>(n_parts) * m_pad;
  auto cleanup = [&]() { cudaFree(d_noise); };   // (supplied noise only; everything else lives in the context's probe arena)
#define CKP(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail_cuda(ctx, e__, #call); } } while (0)
  // the working set is carved out of one arena the context keeps between calls (an alpha search probes dozens of times:
  // eleven cudaMalloc / cudaFree pairs per call cost more than a probe's kernels)
  const size_t bytes = static_cast<size_t>(m_pad) * ld * 4, sbytes = dense_split_bytes(ctx, m_pad, ld);
  {
    auto al = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
    const size_t need = 2 * al(bytes) + 4 * al(sbytes) + al(m_pad * 4) * 5 + al(diff_per_step * chunk * 4) + al(m_pad * 8);
    int rcb = ensure_buf(ctx, ctx->d_parena, ctx->parena_cap, need);
    if (rcb) return rcb;
    unsigned char* at = ctx->d_parena;
    auto take = [&](size_t v) { unsigned char* r = at; at += al(v); return r; };
    for (int i = 0; i < 2; ++i) {
      w.x[i] = reinterpret_cast<float*>(take(bytes));
      w.hi[i] = take(sbytes);
      w.lo[i] = take(sbytes);
    }
    d_scale = reinterpret_cast<float*>(take(m_pad * 4));
    d_diff = reinterpret_cast<float*>(take(diff_per_step * chunk * 4));
    d_cell = reinterpret_cast<int*>(take(m_pad * 4));
    d_steps = reinterpret_cast<int*>(take(m_pad * 4));
    d_src = reinterpret_cast<int*>(take(m_pad * 4));
    d_base = reinterpret_cast<uint32_t*>(take(m_pad * 4));
    d_gid = reinterpret_cast<int64_t*>(take(m_pad * 8));
  }
  CKP(cudaMemcpyAsync(d_scale, h_scale.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_cell, h_cell.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_steps, h_steps.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_src, h_src.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_base, h_base.data(), m_pad * 4, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_gid, h_gid.data(), m_pad * 8, cudaMemcpyHostToDevice, ctx->stream));
  scr::StepParams<float> sp;
  fill_common(ctx, sp);
  int rc = upload_node<float>(ctx, proj, cnst, mu, sp, 1.0);
  if (rc) { cleanup(); return rc; }
  if ((rc = dense_make_maps(ctx, w, m_pad, ld))) { cleanup(); return rc; }
  if (noise) {
    rc = upload_noise<float>(ctx, noise, static_cast<size_t>(k) * max_iter * n, &d_noise);
    if (rc) { cleanup(); return rc; }
  }
  D::Params p;
  std::memset(&p, 0, sizeof(p));
  for (int i = 0; i < 2; ++i) { p.x[i] = w.x[i]; p.hi[i] = w.hi[i]; p.lo[i] = w.lo[i]; }
  p.ld = ld; p.k_blocks = ld * (ctx->dense_kind == D::kF16 ? 2 : 4) / D::ROW_BYTES; p.n = n; p.m_pad = static_cast<int>(m_pad);
  p.pc = static_cast<const float2*>(ctx->d_pc);
  p.bias = sp.bias;
  p.row_steps = d_steps; p.row_base = d_base; p.row_gid = d_gid; p.row_src = d_src; p.row_scale = d_scale;
  p.dt = static_cast<float>(dt);
  p.rscale = static_cast<float>(-2.0 * 0.6931471805599453 * sigma * sigma);
  scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ stream, p.rk);
  p.noise = d_noise; p.noise_steps = max_iter;
  if ((rc = dense_set_attrs(ctx))) { cleanup(); return rc; }
  CKP(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  dense_gather(ctx, m_pad, d_cell, k, w, ld);
  CKP(cudaGetLastError());
  std::vector<double> nd(static_cast<size_t>(k) * max_iter, 0.0);
  std::vector<int> iters(k, 0);
  std::vector<char> done(k, 0);
  // per chunk the device sums the column parts of every (step, row) in index order; only those k doubles per step cross PCIe
  const size_t sum_elems = static_cast<size_t>(chunk) * k;
  if ((rc = ensure_buf(ctx, ctx->d_pnd, ctx->pnd_cap, sum_elems))) { cleanup(); return rc; }
  if (ctx->h_psum_cap < sum_elems) {
    cudaFreeHost(ctx->h_psum);
    ctx->h_psum = nullptr;
    ctx->h_psum_cap = 0;
    CKP(cudaMallocHost(&ctx->h_psum, sum_elems * sizeof(double)));
    ctx->h_psum_cap = sum_elems;
  }
  const double* h_sum = ctx->h_psum;
  int ndone = 0;
  for (int s0 = 0; s0 < max_iter && ndone < k; s0 += chunk) {
    const int cnt = std::min(chunk, max_iter - s0);
    if (sc.inside) {
      p.s0 = s0; p.nsteps = cnt; p.rowdiff = d_diff;
      if ((rc = dense_launch(ctx, sc, m_tiles, w, p))) { cleanup(); return rc; }
    } else {
      for (int j = 0; j < cnt; ++j) {
        p.s0 = s0 + j; p.nsteps = 1; p.rowdiff = d_diff + diff_per_step * j;
        if ((rc = dense_launch(ctx, sc, m_tiles, w, p))) { cleanup(); return rc; }
      }
    }
    D::rowdiff_sum_kernel<<<(cnt * k + 127) / 128, 128, 0, ctx->stream>>>(d_diff, n_parts, static_cast<int>(m_pad), cnt, k, ctx->d_pnd);
    CKP(cudaGetLastError());
    ctx->launches++;
    CKP(cudaMemcpyAsync(ctx->h_psum, ctx->d_pnd, static_cast<size_t>(cnt) * k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CKP(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; ++i) {
      if (done[i]) continue;
      double* trace = nd.data() + static_cast<size_t>(i) * max_iter;
      for (int j = 0; j < cnt && !done[i]; ++j) {
        const int s = s0 + j;
        trace[s] = h_sum[static_cast<size_t>(j) * k + i] / (static_cast<double>(n) * static_cast<double>(static_cast<float>(dt)));
        iters[i] = s + 1;
        if (s >= avg_num - 1) {
          double acc = 0.0;
          for (int q = s - avg_num + 1; q <= s; ++q) acc += trace[q];
          if (acc / avg_num <= thresh) { done[i] = 1; ++ndone; }
        }
      }
    }
  }
#undef CKP
  cleanup();
  if ((rc = dense_check_err(ctx))) return rc;
  for (int i = 0; i < k; ++i) iters_out[i] = iters[i];
  if (normdiff_out) std::memcpy(normdiff_out, nd.data(), nd.size() * sizeof(double));
  return SCR_OK;
}

template <typename T>
int probe_t(scr_ctx* ctx, int32_t k, const int64_t* sample_ids, const double* w_div, const double* proj,
            const double* cnst, double dt, double sigma, double mu, double thresh, int32_t max_iter,
            int32_t avg_num, uint64_t seed, uint32_t stream, const double* noise, int32_t* iters_out,
            double* normdiff_out) {
  constexpr int CPG = scr::Traits<T>::CPG;
  // Cells per group: a probe step of a group costs what its ACTIVE cells cost, and a group has an SM to itself as long as
  // there are no more groups than SMs -- so few samples (a node's 50) are spread one cell per group (50 SMs, a quarter of
  // the per-step work each: 8.0 -> ms per 7 probes, profiles/r2_experiments.txt); the alpha ladder's ~1250 samples keep 4.
  // (the divisor check below keeps the documented contract: constant within each run of CPG consecutive samples)
  int per = CPG;
  if (scr::env_int("SCR_PROBE_SPREAD", 1))   // (probes launched side by side share the SMs: probe_share = how many)
    while (per > 1 && (k + per / 2 - 1) / (per / 2) * ctx->probe_share <= ctx->sm_count) per /= 2;
  const int nitems = (k + per - 1) / per;
  std::vector<Item> items(nitems);
  for (int g = 0; g < nitems; ++g) {
    Item& it = items[g];
    std::memset(&it, 0, sizeof(it));
    for (int c = 0; c < 4; ++c) {
      const int i = g * per + c;
      if (c < per && i < k) {
        if (sample_ids[i] < 0 || sample_ids[i] >= ctx->cells) return fail(ctx, SCR_E_ARG, "sample id out of range");
        it.cell[c] = static_cast<int>(sample_ids[i]);
        it.row[c] = i;
      } else {
        it.cell[c] = -1;
      }
    }
    // one divisor per group: the ladder is laid out so that a group never mixes alphas
    const int i0 = g * per;
    it.scale = w_div ? 1.0 / w_div[i0] : 1.0;
    if (w_div)
      for (int c = 1; c < per && i0 + c < k; ++c)
        if (w_div[i0 + c] != w_div[i0]) return fail(ctx, SCR_E_ARG, "w_div must be constant within a cell group");
  }
  if (w_div)
    for (int i = 0; i < k; ++i)
      if (w_div[i] != w_div[i / CPG * CPG]) return fail(ctx, SCR_E_ARG, "w_div must be constant within a cell group");
  int rc = ensure_items(ctx, items.size());
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_items, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice, ctx->stream));

  scr::StepParams<T> p;
  fill_common(ctx, p);
  rc = upload_node<T>(ctx, proj, cnst, mu, p, scr::Traits<T>::kDriveScale);
  if (rc) return rc;
  p.items = ctx->d_items;
  p.nitems = nitems;
  p.nrows = k;
  p.dt = static_cast<T>(dt);
  p.rscale = static_cast<float>(-2.0 * 0.6931471805599453 * sigma * sigma * scr::Traits<T>::kDriveScale * scr::Traits<T>::kDriveScale);   // radius of K sigma xi
  scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ stream, p.rk);
  p.thresh = thresh;
  p.max_iter = max_iter;
  p.avg_num = avg_num;
  T* d_noise = nullptr;
  const size_t nd_count = static_cast<size_t>(k) * max_iter;
  if ((rc = ensure_buf(ctx, ctx->d_piters, ctx->piters_cap, static_cast<size_t>(k)))) return rc;
  if ((rc = ensure_buf(ctx, ctx->d_pnd, ctx->pnd_cap, nd_count))) return rc;
  if (normdiff_out) CK(cudaMemsetAsync(ctx->d_pnd, 0, nd_count * sizeof(double), ctx->stream));   // entries past a cell's last iteration read 0
  p.iters_out = ctx->d_piters;
  p.nd = ctx->d_pnd;
  if (noise) {
    rc = upload_noise<T>(ctx, noise, nd_count * ctx->op.n, &d_noise);
    if (rc) return rc;
    p.noise = d_noise;
    p.noise_steps = max_iter;
  }
  rc = launch_step<T>(ctx, p, nitems, true, noise != nullptr);
  if (rc == SCR_OK && !ctx->defer_sync) {
    cudaError_t e = cudaMemcpy(iters_out, ctx->d_piters, k * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && normdiff_out) e = cudaMemcpy(normdiff_out, ctx->d_pnd, nd_count * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail_cuda(ctx, e, "probe result copy");
  }
  cudaFree(d_noise);
  return rc;
}

// Probes of several lineage nodes at once (siblings: their sampled cells are final as soon as the parent's phase is).
// Each probe is the launch scr_probe would make -- same kernel, same items, same noise keys, hence the same iteration
// counts -- but on its own stream with its own per-launch buffers, so the launches (a node's 50 samples occupy 50 SMs
// for up to max_iter strictly sequential steps) run side by side instead of one after the other.
template <typename T>
int probe_multi_t(scr_ctx* ctx, int32_t nprobes, const int32_t* k, const int64_t* sample_ids, const double* proj,
                  const double* cnst, double dt, double sigma, double mu, double thresh, int32_t max_iter, int32_t avg_num,
                  const uint64_t* seeds, uint32_t stream, int32_t* iters_out) {
  const int n = ctx->op.n;
  // saved per-launch fields of the context
  struct Saved {
    cudaStream_t stream; cudaEvent_t ev0, ev1; Item *d_items, *h_items; size_t items_cap; void *d_pc, *d_bias; int* d_counter;
    int* d_piters; double* d_pnd; size_t piters_cap, pnd_cap;
  } sv{ctx->stream, ctx->ev0, ctx->ev1, ctx->d_items, ctx->h_items, ctx->items_cap, ctx->d_pc, ctx->d_bias, ctx->d_counter,
       ctx->d_piters, ctx->d_pnd, ctx->piters_cap, ctx->pnd_cap};
  auto install = [&](scr_ctx::ProbeSlot& sl) {
    ctx->stream = sl.stream; ctx->ev0 = sl.ev0; ctx->ev1 = sl.ev1; ctx->d_items = sl.d_items; ctx->h_items = sl.h_items;
    ctx->items_cap = sl.items_cap; ctx->d_pc = sl.d_pc; ctx->d_bias = sl.d_bias; ctx->d_counter = sl.d_counter;
    ctx->d_piters = sl.d_piters; ctx->d_pnd = sl.d_pnd; ctx->piters_cap = sl.piters_cap; ctx->pnd_cap = sl.pnd_cap;
  };
  auto take_back = [&](scr_ctx::ProbeSlot& sl) {   // buffers the call may have grown
    sl.d_items = ctx->d_items; sl.h_items = ctx->h_items; sl.items_cap = ctx->items_cap;
    sl.d_piters = ctx->d_piters; sl.d_pnd = ctx->d_pnd; sl.piters_cap = ctx->piters_cap; sl.pnd_cap = ctx->pnd_cap;
  };
  auto restore = [&]() {
    ctx->stream = sv.stream; ctx->ev0 = sv.ev0; ctx->ev1 = sv.ev1; ctx->d_items = sv.d_items; ctx->h_items = sv.h_items;
    ctx->items_cap = sv.items_cap; ctx->d_pc = sv.d_pc; ctx->d_bias = sv.d_bias; ctx->d_counter = sv.d_counter;
    ctx->d_piters = sv.d_piters; ctx->d_pnd = sv.d_pnd; ctx->piters_cap = sv.piters_cap; ctx->pnd_cap = sv.pnd_cap;
    ctx->defer_sync = false;
    ctx->probe_share = 1;
  };
  int rc = SCR_OK;
  int64_t off = 0;
  std::vector<int64_t> offs(nprobes);
  ctx->defer_sync = true;
  ctx->probe_share = nprobes;
  for (int i = 0; i < nprobes && rc == SCR_OK; ++i) {
    scr_ctx::ProbeSlot& sl = ctx->probe_slots[i];
    if (!sl.stream) {
      cudaError_t e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreate(&sl.ev0);
      if (e == cudaSuccess) e = cudaEventCreate(&sl.ev1);
      if (e == cudaSuccess) e = cudaMalloc(&sl.d_counter, sizeof(int));
      if (e == cudaSuccess) e = cudaMalloc(&sl.d_pc, static_cast<size_t>(ctx->op.n_pad) * 2 * ctx->esize());
      if (e == cudaSuccess) e = cudaMalloc(&sl.d_bias, static_cast<size_t>(ctx->op.n_pad) * ctx->esize());
      if (e != cudaSuccess) {   // (a half-built slot must not be found "ready" by the next call)
        restore();
        for (int j = 0; j < i; ++j) cudaStreamSynchronize(ctx->probe_slots[j].stream);
        free_probe_slots(ctx);
        return fail_cuda(ctx, e, "probe slot");
      }
    }
    install(sl);
    offs[i] = off;
    rc = probe_t<T>(ctx, k[i], sample_ids + off, nullptr, proj + static_cast<size_t>(i) * n, cnst + static_cast<size_t>(i) * n, dt,
                    sigma, mu, thresh, max_iter, avg_num, seeds[i], stream, nullptr, iters_out + off, nullptr);
    take_back(sl);
    off += k[i];
  }
  // collect: every launched probe is waited for even after an error
  double ms_max = 0.0;
  for (int i = 0; i < nprobes; ++i) {
    scr_ctx::ProbeSlot& sl = ctx->probe_slots[i];
    if (!sl.stream) continue;
    cudaError_t e = cudaStreamSynchronize(sl.stream);
    if (e != cudaSuccess && rc == SCR_OK) rc = fail_cuda(ctx, e, "probe (multi)");
    if (rc == SCR_OK && sl.d_piters) {
      e = cudaMemcpy(iters_out + offs[i], sl.d_piters, k[i] * sizeof(int), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = fail_cuda(ctx, e, "probe result copy");
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess) ms_max = std::max(ms_max, static_cast<double>(ms));
    }
  }
  restore();
  ctx->total_probe_ms += ms_max;     // the probes overlap: the longest one is what the batch took
  ctx->total_probe_launches += nprobes;
  ctx->last_ms = ms_max;
  return rc;
}

void free_probe_slots(scr_ctx* ctx) {
  for (auto& sl : ctx->probe_slots) {
    if (sl.stream) cudaStreamDestroy(sl.stream);
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    if (sl.ev1) cudaEventDestroy(sl.ev1);
    cudaFree(sl.d_items); cudaFreeHost(sl.h_items); cudaFree(sl.d_pc); cudaFree(sl.d_bias); cudaFree(sl.d_counter);
    cudaFree(sl.d_piters); cudaFree(sl.d_pnd);
    sl = scr_ctx::ProbeSlot();
  }
}

void free_counts_scratch(scr_ctx* ctx) {
  cudaFree(ctx->d_ctab); cudaFree(ctx->d_csize); cudaFree(ctx->d_lgam); cudaFree(ctx->d_cbuf[0]); cudaFree(ctx->d_cbuf[1]);
  cudaFree(ctx->d_covf); cudaFree(ctx->d_covf_count);
  ctx->d_ctab = nullptr; ctx->d_csize = nullptr; ctx->d_lgam = nullptr; ctx->d_cbuf[0] = ctx->d_cbuf[1] = nullptr;
  ctx->d_covf = nullptr; ctx->d_covf_count = nullptr;
  ctx->ctab_cap = ctx->csize_cap = ctx->cbuf_cap[0] = ctx->cbuf_cap[1] = ctx->covf_cap = 0;
}

int ensure_stage(scr_ctx* ctx, size_t elems) {
  if (elems <= ctx->stage_elems) return SCR_OK;
  cudaFree(ctx->d_stage);
  ctx->d_stage = nullptr;
  ctx->stage_elems = 0;
  CK(cudaMalloc(&ctx->d_stage, elems * sizeof(double)));
  ctx->stage_elems = elems;
  return SCR_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int scr_abi_version(void) { return 2; }
int scr_noise_stream_version(void) { return SCR_NOISE16 ? 2 : 1; }

int scr_create(scr_ctx** out, int device_ordinal, int precision) {
  if (!out || (precision != SCR_F32 && precision != SCR_F64)) return fail(nullptr, SCR_E_ARG, "bad arguments");
  scr_ctx* ctx = new scr_ctx();
  ctx->prec = precision;
  if (device_ordinal < 0) {   // packing-only context for the CPU test-suite; every GPU entry point refuses it
    ctx->host_only = true;
    *out = ctx;
    return SCR_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device_ordinal >= ndev) {
    delete ctx;
    return fail(nullptr, SCR_E_CUDA, e != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(e)
                                                       : std::string("device ordinal out of range"));
  }
  ctx->device = device_ordinal;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess) {
    delete ctx;
    return fail_cuda(nullptr, e, "cudaSetDevice");
  }
  if (prop.major < 10) {
    delete ctx;
    return fail(nullptr, SCR_E_CUDA, "scrutiny_b200 is built for sm_100a only");
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_k[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_k[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_c[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_c[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_counter, sizeof(int))) != cudaSuccess || (e = cudaMalloc(&ctx->d_err, sizeof(int))) != cudaSuccess) {
    delete ctx;
    return fail_cuda(nullptr, e, "context setup");
  }
  *out = ctx;
  return SCR_OK;
}

void scr_destroy(scr_ctx* ctx) {
  if (!ctx) return;
  if (!ctx->host_only) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_network(ctx);
    free_dense_work(ctx);
    cudaFree(ctx->d_state); cudaFree(ctx->d_items); cudaFreeHost(ctx->h_items); cudaFree(ctx->d_counter); cudaFree(ctx->d_err); cudaFree(ctx->d_stage);
    cudaFree(ctx->d_piters); cudaFree(ctx->d_pnd); cudaFree(ctx->d_gscratch); cudaFree(ctx->d_parena); cudaFree(ctx->d_mt_count); cudaFreeHost(ctx->h_psum);
    for (int b = 0; b < 2; ++b) { cudaFreeHost(ctx->h_xfer[b]); cudaFree(ctx->d_xfer[b]); if (ctx->ev_x[b]) cudaEventDestroy(ctx->ev_x[b]); }
    free_counts_scratch(ctx);
    free_probe_slots(ctx);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 2; ++i) { if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]); if (ctx->ev_c[i]) cudaEventDestroy(ctx->ev_c[i]); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  }
  delete ctx;
}

const char* scr_last_error(const scr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int scr_sync(scr_ctx* ctx) {
  NEED_GPU();
  CK(cudaStreamSynchronize(ctx->stream));
  return SCR_OK;
}

int scr_set_network_csr(scr_ctx* ctx, int32_t n, int64_t nnz, const int32_t* rowptr, const int32_t* col, const double* val) {
  if (!ctx) return SCR_E_ARG;
  if (ctx->cells > 0 && ctx->has_net && n != ctx->op.n) return fail(ctx, SCR_E_STATE, "cells are allocated for a different n");
  PackedOperator op;
  std::string err;
  // density switch (north_star): sparse networks -> smem-resident SpMM stepper; dense ones -> tcgen05 GEMM
  const double density = n > 0 ? static_cast<double>(nnz) / (static_cast<double>(n) * n) : 0.0;
  const int dense_env = scr::env_int("SCR_DENSE", -1);
  const bool dense = ctx->prec == SCR_F32 && n > 0 && (dense_env == 1 || (dense_env != 0 && density >= 0.05 && n >= 64));
  if (!scr::pack_operator(n, nnz, rowptr, col, val, op, err, dense, dense ? scr::dense::BN : 128)) return fail(ctx, SCR_E_ARG, err);
  if (ctx->cells > 0 && ctx->has_net && dense != ctx->dense_mode)
    return fail(ctx, SCR_E_STATE, "network density class changed while cells are allocated");
  // the gene order must not change under allocated cells (alpha rescaling keeps the pattern, hence the order)
  if (ctx->cells > 0 && ctx->has_net && op.perm != ctx->op.perm)
    return fail(ctx, SCR_E_STATE, "network pattern changed while cells are allocated");
  if (op.nb > scr::kMaxBlocks) return fail(ctx, SCR_E_RESOURCE, "n too large: more than " + std::to_string(scr::kMaxBlocks * 128) + " genes");
  // Does one cell group (with the probe's header) fit in shared memory next to the tables and the node's proj/const?  If
  // not, the groups live in global memory (step_kernel.cuh kModeGlobal) and the tables take the large record layout.
  {
    const size_t tab_small = static_cast<size_t>(align16(scr::fixed_tab(false) + static_cast<int>(op.ovf_start.size()) * 8));
    const size_t need = tab_small + static_cast<size_t>(op.n_pad) * 2 * ctx->esize() + 16 +
                        scr::group_smem_bytes(op.n_pad, op.npp, op.nv_pad, true);
    ctx->gstate = op.nb > scr::kMaxBlocksOnChip || need > static_cast<size_t>(ctx->smem_optin) || scr::env_int("SCR_GLOBAL_STATE", 0) == 1;
  }
  std::vector<unsigned char> tab, w;
  if (ctx->prec == SCR_F64) build_blobs<double>(op, tab, w, ctx->off_vrow, ctx->gstate);
  else build_blobs<float>(op, tab, w, ctx->off_vrow, ctx->gstate);
  ctx->tab_bytes = static_cast<int>(tab.size());
  ctx->w_bytes = static_cast<int>(w.size());
  ctx->h_tab = tab;
  ctx->op = std::move(op);
  ctx->user_bias.clear();
  ctx->has_net = false;
  int rc = choose_config(ctx);
  if (rc) return rc;
  if (!ctx->host_only) {
    CK(cudaSetDevice(ctx->device));
    free_network(ctx);
    const PackedOperator& o = ctx->op;
    CK(cudaMalloc(&ctx->d_tab, tab.size()));
    CK(cudaMalloc(&ctx->d_w, w.size()));
    CK(cudaMalloc(&ctx->d_perm, o.n_pad * sizeof(int)));
    CK(cudaMalloc(&ctx->d_inv, o.n * sizeof(int)));
    CK(cudaMalloc(&ctx->d_pc, static_cast<size_t>(o.n_pad) * 2 * ctx->esize()));
    CK(cudaMalloc(&ctx->d_bias, static_cast<size_t>(o.n_pad) * ctx->esize()));
    CK(upload_sync(ctx, ctx->d_tab, tab.data(), tab.size()));
    if (o.wq_wide == 16) {
      std::vector<unsigned char> tab_wide;
      build_tab(o, 16, o.wblk_ptr_wide, o.wblk_wide, tab_wide, ctx->gstate);
      CK(cudaMalloc(&ctx->d_tab_wide, tab_wide.size()));
      CK(upload_sync(ctx, ctx->d_tab_wide, tab_wide.data(), tab_wide.size()));
    }
    CK(upload_sync(ctx, ctx->d_w, w.data(), w.size()));
    CK(upload_sync(ctx, ctx->d_perm, o.perm.data(), o.n_pad * sizeof(int)));
    CK(upload_sync(ctx, ctx->d_inv, o.inv.data(), o.n * sizeof(int)));
    if (dense) {
      // B operands of the split-precision GEMM (dense_kernel.cuh): hi / lo of W in the chosen encoding
      namespace D = scr::dense;
      const char* kind_env = std::getenv("SCR_DENSE_KIND");
      const int kind = kind_env && std::string(kind_env) == "tf32" ? D::kTF32 : D::kF16;
      if (kind != ctx->dense_kind) free_dense_work(ctx);   // the working set's split buffers are sized per encoding
      ctx->dense_kind = kind;
      const size_t ld = o.n_pad, cnt = ld * ld;
      int rc2 = SCR_OK;
      if (ctx->dense_kind == D::kF16) {
        // power-of-two scale: max |W| lands in [2^14, 2^15), the products with sx x stay far inside fp32
        double wmax = 0.0;
        for (int64_t e = 0; e < nnz; ++e) wmax = std::max(wmax, std::fabs(val[e]));
        const float wmaxf = static_cast<float>(wmax);
        const double sw = wmaxf > 0.f && std::isfinite(wmaxf) ? std::ldexp(1.0, std::max(-100, std::min(100, 14 - std::ilogb(wmaxf)))) : 1.0;
        ctx->dense_acc_scale = static_cast<float>(1.0 / (static_cast<double>(D::kStateScale) * sw));
        std::vector<__half> whi(cnt, __float2half_rn(0.f)), wlo(cnt, __float2half_rn(0.f));
        for (int r = 0; r < n; ++r)
          for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const float ws = static_cast<float>(val[e]) * static_cast<float>(sw);
            const __half h = __float2half_rn(ws);
            whi[r * ld + col[e]] = h;
            wlo[r * ld + col[e]] = __float2half_rn(ws - __half2float(h));
          }
        CK(cudaMalloc(&ctx->d_whi, cnt * 2));
        CK(cudaMalloc(&ctx->d_wlo, cnt * 2));
        CK(upload_sync(ctx, ctx->d_whi, whi.data(), cnt * 2));
        CK(upload_sync(ctx, ctx->d_wlo, wlo.data(), cnt * 2));
      } else {
        ctx->dense_acc_scale = 1.f;
        std::vector<float> whi(cnt, 0.f), wlo(cnt, 0.f);
        for (int r = 0; r < n; ++r)
          for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const float w = static_cast<float>(val[e]);
            uint32_t bits;
            std::memcpy(&bits, &w, 4);
            bits &= 0xffffe000u;
            float hi;
            std::memcpy(&hi, &bits, 4);
            whi[r * ld + col[e]] = w;
            wlo[r * ld + col[e]] = w - hi;
          }
        CK(cudaMalloc(&ctx->d_whi, cnt * 4));
        CK(cudaMalloc(&ctx->d_wlo, cnt * 4));
        CK(upload_sync(ctx, ctx->d_whi, whi.data(), cnt * 4));
        CK(upload_sync(ctx, ctx->d_wlo, wlo.data(), cnt * 4));
      }
      const int es = ctx->dense_kind == D::kF16 ? 2 : 4;
      rc2 = make_tmap(ctx, &ctx->tm_bhi, ctx->d_whi, ld, ld, ld, D::BN, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_blo, ctx->d_wlo, ld, ld, ld, D::BN, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_bhi_n, ctx->d_whi, ld, ld, ld, D::BN_NARROW, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_blo_n, ctx->d_wlo, ld, ld, ld, D::BN_NARROW, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_bhi_r, ctx->d_whi, ld, ld, ld, D::BN_RES, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_blo_r, ctx->d_wlo, ld, ld, ld, D::BN_RES, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_bhi_p, ctx->d_whi, ld, ld, ld, D::PAIR_B_ROWS, es);
      if (!rc2) rc2 = make_tmap(ctx, &ctx->tm_blo_p, ctx->d_wlo, ld, ld, ld, D::PAIR_B_ROWS, es);
      if (rc2) return rc2;
    }
  }
  ctx->dense_mode = dense;
  ctx->readout_only = false;
  ctx->has_net = true;
  return SCR_OK;
}

int scr_set_genes(scr_ctx* ctx, int32_t n) {
  if (!ctx || n <= 0) return fail(ctx, SCR_E_ARG, "bad arguments");
  if (ctx->cells > 0) return fail(ctx, SCR_E_STATE, "cells are already allocated");
  PackedOperator op;
  std::string err;
  std::vector<int32_t> rowptr(static_cast<size_t>(n) + 1, 0);
  if (!scr::pack_operator(n, 0, rowptr.data(), nullptr, nullptr, op, err, true, 128)) return fail(ctx, SCR_E_ARG, err);
  ctx->op = std::move(op);
  ctx->user_bias.clear();
  ctx->has_net = false;
  ctx->dense_mode = false;
  ctx->tab_bytes = ctx->w_bytes = 0;
  if (!ctx->host_only) {
    CK(cudaSetDevice(ctx->device));
    free_network(ctx);
    const PackedOperator& o = ctx->op;
    CK(cudaMalloc(&ctx->d_perm, o.n_pad * sizeof(int)));
    CK(cudaMalloc(&ctx->d_inv, o.n * sizeof(int)));
    CK(upload_sync(ctx, ctx->d_perm, o.perm.data(), o.n_pad * sizeof(int)));
    CK(upload_sync(ctx, ctx->d_inv, o.inv.data(), o.n * sizeof(int)));
  }
  ctx->readout_only = true;
  ctx->has_net = true;
  return SCR_OK;
}

int scr_set_network_dense(scr_ctx* ctx, int32_t n, const double* gc) {
  if (!ctx || n <= 0 || !gc) return fail(ctx, SCR_E_ARG, "bad arguments");
  std::vector<int32_t> rowptr(n + 1, 0), col;
  std::vector<double> val;
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) {
      const double v = gc[static_cast<size_t>(r) * n + c];
      if (v != 0.0) { col.push_back(c); val.push_back(v); }
    }
    rowptr[r + 1] = static_cast<int32_t>(col.size());
  }
  return scr_set_network_csr(ctx, n, static_cast<int64_t>(col.size()), rowptr.data(), col.data(), val.data());
}

int scr_debug_host_matvec(const scr_ctx* ctx, const double* x, double* y) {
  if (!ctx || !x || !y) return SCR_E_ARG;
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  scr::host_matvec(ctx->op, x, y);
  return SCR_OK;
}

int scr_debug_block_records(const scr_ctx* ctx, int32_t* records, int32_t max_records, int32_t* list_bounds, int32_t max_bounds) {
  if (!ctx || !records || !list_bounds) return SCR_E_ARG;
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  const PackedOperator& o = ctx->op;
  if (max_records < o.nb || max_bounds < o.wq + 1) return fail(ctx, SCR_E_ARG, "output buffers too small");
  std::memcpy(records, ctx->h_tab.data(), static_cast<size_t>(o.nb) * 16);
  std::memcpy(list_bounds, ctx->h_tab.data() + scr::rec_bytes(ctx->gstate), static_cast<size_t>(o.wq + 1) * 4);
  return o.nb;
}

int scr_set_drive_bias(scr_ctx* ctx, const double* bias) {
  if (!ctx) return SCR_E_ARG;
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  if (!bias) { ctx->user_bias.clear(); return SCR_OK; }
  ctx->user_bias.assign(bias, bias + ctx->op.n);
  return SCR_OK;
}

int scr_debug_gene_order(const scr_ctx* ctx, int32_t* perm_out, int32_t count) {
  if (!ctx || !perm_out) return SCR_E_ARG;
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  if (count != ctx->op.n_pad) return fail(ctx, SCR_E_ARG, "count must be n_pad");
  for (int i = 0; i < count; ++i) perm_out[i] = ctx->op.perm[i];
  return SCR_OK;
}

int scr_get_layout(const scr_ctx* ctx, int64_t* out, int count) {
  if (!ctx || !out) return SCR_E_ARG;
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  const PackedOperator& o = ctx->op;
  const size_t per_group = ctx->gstate ? static_cast<size_t>(scr::group_hdr(false)) : ctx->gbytes;
  const int64_t v[18] = {o.n_pad, o.npp, static_cast<int64_t>(o.tile_val.size()), o.nv, o.lv, o.cap, o.wq, ctx->q,
                         static_cast<int64_t>(ctx->fixed_bytes + (ctx->w_smem ? ctx->w_bytes : 0) + ctx->q * per_group +
                                              w_prefix_bytes(ctx, ctx->q)),
                         ctx->w_smem ? 1 : 0, ctx->cpg(), ctx->dense_mode ? 1 : 0, w_prefix_bytes(ctx, ctx->q), ctx->w_bytes,
                         o.gather_wavefronts, o.gather_wavefronts_floor, ctx->gstate ? 1 : 0,
                         ctx->dense_mode ? (ctx->dense_kind == scr::dense::kF16 ? 1 : 2) : 0};
  for (int i = 0; i < count && i < 18; ++i) out[i] = v[i];
  return SCR_OK;
}

int scr_alloc_cells(scr_ctx* ctx, int64_t cells, int64_t global_cell_offset) {
  NEED_GPU();
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "set the network before allocating cells");
  if (cells <= 0 || cells > 0x7fffffff) return fail(ctx, SCR_E_ARG, "bad cell count");
  cudaFree(ctx->d_state);
  ctx->d_state = nullptr;
  ctx->cells = 0;
  const size_t bytes = static_cast<size_t>(cells) * ctx->op.n_pad * ctx->esize();
  CK(cudaMalloc(&ctx->d_state, bytes));
  CK(cudaMemsetAsync(ctx->d_state, 0, bytes, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cells = cells;
  ctx->cell_offset = global_cell_offset;
  return SCR_OK;
}

namespace {
int ensure_xfer(scr_ctx* ctx, size_t elems) {
  if (elems <= ctx->xfer_elems) return SCR_OK;
  for (int b = 0; b < 2; ++b) {
    cudaFreeHost(ctx->h_xfer[b]); cudaFree(ctx->d_xfer[b]);
    ctx->h_xfer[b] = ctx->d_xfer[b] = nullptr;
  }
  ctx->xfer_elems = 0;
  for (int b = 0; b < 2; ++b) {
    CK(cudaHostAlloc(&ctx->h_xfer[b], elems * sizeof(double), cudaHostAllocDefault));
    CK(cudaMalloc(&ctx->d_xfer[b], elems * sizeof(double)));
    if (!ctx->ev_x[b]) CK(cudaEventCreateWithFlags(&ctx->ev_x[b], cudaEventDisableTiming));
  }
  ctx->xfer_elems = elems;
  return SCR_OK;
}
// rows [g0, g1) of a chunk: user matrix (row stride ld) <-> contiguous slab [n][cnt]
void copy_rows(bool to_slab, double* slab, double* user, int64_t ld, int64_t cnt, int n) {
  const int nthr = (static_cast<int64_t>(n) * cnt >= (1 << 20)) ? std::max(1, std::min(scr::env_int("SCR_HOST_THREADS", 4), 16)) : 1;
  auto work = [&](int t) {
    const int g0 = static_cast<int>(static_cast<int64_t>(n) * t / nthr), g1 = static_cast<int>(static_cast<int64_t>(n) * (t + 1) / nthr);
    for (int g = g0; g < g1; ++g) {
      if (to_slab) std::memcpy(slab + static_cast<int64_t>(g) * cnt, user + static_cast<int64_t>(g) * ld, cnt * sizeof(double));
      else std::memcpy(user + static_cast<int64_t>(g) * ld, slab + static_cast<int64_t>(g) * cnt, cnt * sizeof(double));
    }
  };
  std::vector<std::thread> th;
  int started = 1;
  try {
    for (int t = 1; t < nthr; ++t) { th.emplace_back(work, t); ++started; }
  } catch (...) {}
  work(0);
  for (int t = started; t < nthr; ++t) work(t);
  for (auto& x : th) x.join();
}
}  // namespace

int scr_set_states(scr_ctx* ctx, int64_t cell0, int64_t count, const double* S, int64_t ld) {
  NEED_GPU();
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (cell0 < 0 || count < 0 || cell0 + count > ctx->cells || !S || ld < count) return fail(ctx, SCR_E_ARG, "bad range");
  if (count == 0) return SCR_OK;
  const int n = ctx->op.n, n_pad = ctx->op.n_pad;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(count, (16LL << 20) / n));   // 128 MB slabs
  int rc = ensure_xfer(ctx, static_cast<size_t>(chunk) * n);
  if (rc) return rc;
  int64_t ci = 0;
  for (int64_t c0 = 0; c0 < count; c0 += chunk, ++ci) {
    const int b = static_cast<int>(ci & 1);
    const int64_t cnt = std::min(chunk, count - c0);
    if (ci >= 2) CK(cudaEventSynchronize(ctx->ev_x[b]));      // the slab's previous chunk has been packed
    copy_rows(true, ctx->h_xfer[b], const_cast<double*>(S) + c0, ld, cnt, n);
    CK(cudaMemcpyAsync(ctx->d_xfer[b], ctx->h_xfer[b], static_cast<size_t>(cnt) * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((n_pad + 127) / 128, static_cast<unsigned>(std::min<int64_t>(cnt, 4096)));
    if (ctx->prec == SCR_F64)
      scr::pack_states_kernel<double><<<grid, 128, 0, ctx->stream>>>(ctx->d_xfer[b], cnt, static_cast<double*>(ctx->d_state), n_pad, cell0 + c0, ctx->d_perm, n_pad);
    else
      scr::pack_states_kernel<float><<<grid, 128, 0, ctx->stream>>>(ctx->d_xfer[b], cnt, static_cast<float*>(ctx->d_state), n_pad, cell0 + c0, ctx->d_perm, n_pad);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_x[b], ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return SCR_OK;
}

int scr_set_states_broadcast(scr_ctx* ctx, const double* x0) {
  NEED_GPU();
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (!x0) return fail(ctx, SCR_E_ARG, "null state");
  const int n = ctx->op.n, n_pad = ctx->op.n_pad;
  int rc = ensure_stage(ctx, n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_stage, x0, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((n_pad + 127) / 128, static_cast<unsigned>(std::min<int64_t>(ctx->cells, 8192)));
  if (ctx->prec == SCR_F64)
    scr::broadcast_state_kernel<double><<<grid, 128, 0, ctx->stream>>>(ctx->d_stage, static_cast<double*>(ctx->d_state), n_pad, ctx->cells, ctx->d_perm, n_pad);
  else
    scr::broadcast_state_kernel<float><<<grid, 128, 0, ctx->stream>>>(ctx->d_stage, static_cast<float*>(ctx->d_state), n_pad, ctx->cells, ctx->d_perm, n_pad);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaStreamSynchronize(ctx->stream));
  return SCR_OK;
}

int scr_get_states(scr_ctx* ctx, int64_t cell0, int64_t count, double* S, int64_t ld) {
  NEED_GPU();
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (cell0 < 0 || count < 0 || cell0 + count > ctx->cells || !S || ld < count) return fail(ctx, SCR_E_ARG, "bad range");
  if (count == 0) return SCR_OK;
  const int n = ctx->op.n, n_pad = ctx->op.n_pad;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(count, (16LL << 20) / n));
  int rc = ensure_xfer(ctx, static_cast<size_t>(chunk) * n);
  if (rc) return rc;
  int64_t ci = 0, prev_c0 = 0, prev_cnt = 0;
  for (int64_t c0 = 0; c0 < count; c0 += chunk, ++ci) {
    const int b = static_cast<int>(ci & 1);
    const int64_t cnt = std::min(chunk, count - c0);
    dim3 grid((n_pad + 127) / 128, static_cast<unsigned>(std::min<int64_t>(cnt, 4096)));
    if (ctx->prec == SCR_F64)
      scr::unpack_states_kernel<double><<<grid, 128, 0, ctx->stream>>>(ctx->d_xfer[b], cnt, static_cast<const double*>(ctx->d_state), n_pad, cell0 + c0, ctx->d_perm, n_pad);
    else
      scr::unpack_states_kernel<float><<<grid, 128, 0, ctx->stream>>>(ctx->d_xfer[b], cnt, static_cast<const float*>(ctx->d_state), n_pad, cell0 + c0, ctx->d_perm, n_pad);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_xfer[b], ctx->d_xfer[b], static_cast<size_t>(cnt) * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_x[b], ctx->stream));
    if (ci >= 1) {   // the previous chunk has arrived in the other slab: scatter it while this one is in flight
      CK(cudaEventSynchronize(ctx->ev_x[b ^ 1]));
      copy_rows(false, ctx->h_xfer[b ^ 1], S + prev_c0, ld, prev_cnt, n);
    }
    prev_c0 = c0;
    prev_cnt = cnt;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  copy_rows(false, ctx->h_xfer[(ci - 1) & 1], S + prev_c0, ld, prev_cnt, n);
  return SCR_OK;
}

int scr_run_phase(scr_ctx* ctx, int64_t m, const int64_t* cell_ids, const int32_t* steps, const int64_t* step_base,
                  const double* proj, const double* cnst, double dt, double sigma, double mu, uint64_t seed,
                  const double* noise, int64_t noise_steps) {
  NEED_GPU();
  NvtxRange nvtx("scr_run_phase (lineage node: members + descendants)", m);
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (m < 0 || (m > 0 && (!cell_ids || !steps)) || !proj || !cnst) return fail(ctx, SCR_E_ARG, "bad arguments");
  if (ctx->readout_only) return fail(ctx, SCR_E_STATE, "read-out context (scr_set_genes): no network to step with");
  if (m == 0) return SCR_OK;
  if (ctx->dense_mode)
    return run_phase_dense(ctx, m, cell_ids, steps, step_base, proj, cnst, dt, sigma, mu, seed, noise, noise_steps);
  if (ctx->prec == SCR_F64)
    return run_phase_t<double>(ctx, m, cell_ids, steps, step_base, proj, cnst, dt, sigma, mu, seed, noise, noise_steps);
  return run_phase_t<float>(ctx, m, cell_ids, steps, step_base, proj, cnst, dt, sigma, mu, seed, noise, noise_steps);
}

int scr_probe(scr_ctx* ctx, int32_t k, const int64_t* sample_ids, const double* w_div, const double* proj,
              const double* cnst, double dt, double sigma, double mu, double thresh, int32_t max_iter, int32_t avg_num,
              uint64_t seed, uint32_t stream, const double* noise, int32_t* iters_out, double* normdiff_out) {
  NEED_GPU();
  NvtxRange nvtx("scr_probe (lineage node: convergence probe)", k, max_iter);
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (k <= 0 || !sample_ids || !proj || !cnst || !iters_out || max_iter <= 0 || avg_num <= 0)
    return fail(ctx, SCR_E_ARG, "bad arguments");
  if (ctx->readout_only) return fail(ctx, SCR_E_STATE, "read-out context (scr_set_genes): no network to step with");
  if (ctx->dense_mode && scr::env_int("SCR_DENSE_PROBE", 1))
    return probe_dense(ctx, k, sample_ids, w_div, proj, cnst, dt, sigma, mu, thresh, max_iter, avg_num, seed, stream, noise, iters_out, normdiff_out);
  if (ctx->prec == SCR_F64)
    return probe_t<double>(ctx, k, sample_ids, w_div, proj, cnst, dt, sigma, mu, thresh, max_iter, avg_num, seed, stream, noise, iters_out, normdiff_out);
  return probe_t<float>(ctx, k, sample_ids, w_div, proj, cnst, dt, sigma, mu, thresh, max_iter, avg_num, seed, stream, noise, iters_out, normdiff_out);
}

int scr_probe_multi(scr_ctx* ctx, int32_t nprobes, const int32_t* k, const int64_t* sample_ids, const double* proj,
                    const double* cnst, double dt, double sigma, double mu, double thresh, int32_t max_iter, int32_t avg_num,
                    const uint64_t* seeds, uint32_t stream, int32_t* iters_out) {
  NEED_GPU();
  NvtxRange nvtx("scr_probe_multi (sibling lineage nodes)", nprobes, max_iter);
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (nprobes <= 0 || !k || !sample_ids || !proj || !cnst || !seeds || !iters_out || max_iter <= 0 || avg_num <= 0)
    return fail(ctx, SCR_E_ARG, "bad arguments");
  if (ctx->readout_only) return fail(ctx, SCR_E_STATE, "read-out context (scr_set_genes): no network to step with");
  for (int i = 0; i < nprobes; ++i)
    if (k[i] <= 0) return fail(ctx, SCR_E_ARG, "empty probe");
  const bool dense = ctx->dense_mode && scr::env_int("SCR_DENSE_PROBE", 1);
  // side by side only where the launches are independent of shared scratch: the sparse stepper with on-chip cell groups
  if (dense || ctx->gstate || nprobes == 1 || nprobes > scr_ctx::kProbeSlots || scr::env_int("SCR_PROBE_MULTI", 1) == 0) {
    int64_t off = 0;
    const int n = ctx->op.n;
    for (int i = 0; i < nprobes; ++i) {
      const int rc = scr_probe(ctx, k[i], sample_ids + off, nullptr, proj + static_cast<size_t>(i) * n, cnst + static_cast<size_t>(i) * n,
                               dt, sigma, mu, thresh, max_iter, avg_num, seeds[i], stream, nullptr, iters_out + off, nullptr);
      if (rc) return rc;
      off += k[i];
    }
    return SCR_OK;
  }
  if (ctx->prec == SCR_F64)
    return probe_multi_t<double>(ctx, nprobes, k, sample_ids, proj, cnst, dt, sigma, mu, thresh, max_iter, avg_num, seeds, stream, iters_out);
  return probe_multi_t<float>(ctx, nprobes, k, sample_ids, proj, cnst, dt, sigma, mu, thresh, max_iter, avg_num, seeds, stream, iters_out);
}

int scr_gene_sums(scr_ctx* ctx, double* sums_out) {
  NEED_GPU();
  NvtxRange nvtx("scr_gene_sums");
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (!sums_out) return fail(ctx, SCR_E_ARG, "null output");
  const int n = ctx->op.n, n_pad = ctx->op.n_pad;
  const int rows = static_cast<int>(std::min<int64_t>(ctx->cells, 256));
  int rc = ensure_stage(ctx, static_cast<size_t>(rows + 1) * n_pad);   // [0, n_pad): sums; then rows x n_pad partials
  if (rc) return rc;
  double* d_partial = ctx->d_stage + n_pad;
  dim3 grid((n_pad + 127) / 128, static_cast<unsigned>(rows));
  if (ctx->prec == SCR_F64)
    scr::gene_sums_kernel<double><<<grid, 128, 0, ctx->stream>>>(static_cast<const double*>(ctx->d_state), n_pad, ctx->cells, n_pad, d_partial);
  else
    scr::gene_sums_kernel<float><<<grid, 128, 0, ctx->stream>>>(static_cast<const float*>(ctx->d_state), n_pad, ctx->cells, n_pad, d_partial);
  CK(cudaGetLastError());
  scr::gene_sums_finish_kernel<<<(n_pad + 127) / 128, 128, 0, ctx->stream>>>(d_partial, rows, n_pad, ctx->d_stage);
  CK(cudaGetLastError());
  ctx->launches += 2;
  std::vector<double> h(n_pad);
  CK(cudaMemcpyAsync(h.data(), ctx->d_stage, n_pad * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) sums_out[ctx->op.perm[i]] = h[i];
  return SCR_OK;
}

namespace {
// shared body of scr_counts_range / scr_counts_compact.  out_dtype: SCR_COUNT_I32 / _U16 / _U8 (bytes per entry 4 / 2 / 1)
int counts_impl(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift, double exp_scale,
                const double* size_factors, uint64_t seed, const double* uniforms, int32_t draws_per_entry,
                void* counts_out, int out_on_device, double* lambda_out, int out_dtype, int64_t* ovf_out,
                int64_t ovf_cap, int64_t* ovf_count_out, double nb_phi = 0.0, double dropout_p = 0.0) {
  NEED_GPU();
  NvtxRange nvtx("scr_counts (read-out)", cell_count);
  if (!(nb_phi >= 0.0) || !(dropout_p >= 0.0) || dropout_p > 1.0) return fail(ctx, SCR_E_ARG, "bad extension parameters");
  if (!ctx->d_state) return fail(ctx, SCR_E_STATE, "cells not allocated");
  if (!shift || !size_factors || !counts_out) return fail(ctx, SCR_E_ARG, "bad arguments");
  if (out_dtype != SCR_COUNT_I32 && out_dtype != SCR_COUNT_U16 && out_dtype != SCR_COUNT_U8) return fail(ctx, SCR_E_ARG, "bad count dtype");
  if (cell_begin < 0 || cell_count < 0 || cell_begin + cell_count > ctx->cells) return fail(ctx, SCR_E_ARG, "cell range out of bounds");
  if (ovf_count_out) *ovf_count_out = 0;
  if (cell_count == 0) return SCR_OK;
  const int n = ctx->op.n;
  const size_t esz = out_dtype == SCR_COUNT_I32 ? 4 : (out_dtype == SCR_COUNT_U16 ? 2 : 1);
  const int64_t cells = cell_count;   // rows of this call: local slots [cell_begin, cell_begin + cells)
  double *d_uni = nullptr, *d_lam = nullptr;   // parity / debug inputs only: allocated per call
  int rc = SCR_OK;
  auto cleanup = [&]() { cudaFree(d_uni); cudaFree(d_lam); };
  // genes sorted by shift, LARGEST first: neighbouring lanes get similar lambda (see counts_kernel), and the partial last
  // trip of a cell's thread loop (n is not a multiple of the CTA size) handles the cheapest genes instead of the rejection
  // sampler's, so the warps of a CTA reach the barrier closer together
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  const bool descending = scr::env_int("SCR_COUNTS_DESCENDING", 1) != 0;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return descending ? shift[a] > shift[b] : shift[a] < shift[b]; });
  std::vector<scr::CountGene> tab(n);
  for (int t = 0; t < n; ++t) tab[t] = scr::CountGene{order[t], ctx->op.inv[order[t]], shift[order[t]]};
  if ((rc = ensure_buf(ctx, ctx->d_ctab, ctx->ctab_cap, static_cast<size_t>(n)))) return rc;
  if ((rc = ensure_buf(ctx, ctx->d_csize, ctx->csize_cap, static_cast<size_t>(cells)))) return rc;
  if (!ctx->d_lgam) {
    std::vector<float> lg(scr::kLgamTable);
    for (int k = 0; k < scr::kLgamTable; ++k) lg[k] = static_cast<float>(std::lgamma(static_cast<double>(k) + 1.0));
    CK(cudaMalloc(&ctx->d_lgam, lg.size() * sizeof(float)));
    CK(upload_sync(ctx, ctx->d_lgam, lg.data(), lg.size() * sizeof(float)));
  }
  const unsigned long long dev_ovf_cap = esz < 4 ? static_cast<unsigned long long>(std::max<int64_t>(ovf_cap, 0)) : 0ULL;
  if (esz < 4) {
    if ((rc = ensure_buf(ctx, ctx->d_covf, ctx->covf_cap, static_cast<size_t>(3 * dev_ovf_cap + 3)))) return rc;
    if (!ctx->d_covf_count) CK(cudaMalloc(&ctx->d_covf_count, sizeof(unsigned long long)));
  }
  double* d_size = ctx->d_csize;
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail_cuda(ctx, e__, #call); } } while (0)
  if (n * sizeof(int32_t) > 48 * 1024 && !ctx->counts_smem_optin) {   // staging row beyond the default 48 KB of dynamic shared memory
    if (n * sizeof(int32_t) > static_cast<size_t>(ctx->smem_optin)) return fail(ctx, SCR_E_RESOURCE, "count read-out: n too large for the per-cell staging row");
#define SCR_OPTIN(T, FAST, OUT) CKC(cudaFuncSetAttribute(scr::counts_kernel<T, FAST, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin))
#define SCR_OPTIN_ALL(OUT) SCR_OPTIN(float, true, OUT); SCR_OPTIN(float, false, OUT); SCR_OPTIN(double, true, OUT); SCR_OPTIN(double, false, OUT)
    SCR_OPTIN_ALL(int32_t); SCR_OPTIN_ALL(uint16_t); SCR_OPTIN_ALL(uint8_t);
#undef SCR_OPTIN_ALL
#undef SCR_OPTIN
    ctx->counts_smem_optin = true;
  }
  CKC(cudaMemcpyAsync(ctx->d_ctab, tab.data(), n * sizeof(scr::CountGene), cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(d_size, size_factors, cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));   // entry r <-> slot cell_begin + r
  CKC(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  if (esz < 4) CKC(cudaMemsetAsync(ctx->d_covf_count, 0, sizeof(unsigned long long), ctx->stream));
  if (uniforms) {
    const size_t cnt = static_cast<size_t>(cells) * n * draws_per_entry;
    CKC(cudaMalloc(&d_uni, cnt * sizeof(double)));
    CKC(cudaMemcpyAsync(d_uni, uniforms, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (lambda_out) CKC(cudaMalloc(&d_lam, static_cast<size_t>(cells) * n * sizeof(double)));
  scr::RoundKeys crk;
  scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(SCR_STREAM_COUNTS), crk);
  // the kernel indexes size factors and supplied uniforms by local cell slot: rebase the range's buffers
  const double* size_by_slot = d_size - cell_begin;
  // free-running stream without a lambda dump: float32 screening of the lambda < 10 sampler (identical integers,
  // aux_kernels.cuh); SCR_COUNTS_EXACT=1 forces the float64 sampler for every entry (A/B tests)
  const bool fast = !uniforms && !lambda_out && nb_phi == 0.0 && dropout_p == 0.0 && scr::env_int("SCR_COUNTS_EXACT", 0) == 0;
  auto launch = [&](int64_t c0, int64_t cnt, void* out, double* lam) -> cudaError_t {
    const double* uni_by_slot = d_uni ? d_uni - cell_begin * n * static_cast<int64_t>(draws_per_entry) : nullptr;
#define SCR_LAUNCH_COUNTS(T, FAST, OUT) \
    scr::counts_kernel<T, FAST, OUT><<<static_cast<unsigned>(cnt), SCR_COUNTS_THREADS, n * sizeof(int32_t), ctx->stream>>>( \
        static_cast<const T*>(ctx->d_state), ctx->op.n_pad, ctx->d_ctab, ctx->d_lgam, n, cell_begin + c0, exp_scale, \
        size_by_slot, crk, ctx->cell_offset, uni_by_slot, draws_per_entry, static_cast<OUT*>(out), lam, ctx->d_err, \
        ctx->d_covf, dev_ovf_cap, ctx->d_covf_count, nb_phi, dropout_p)
#define SCR_LAUNCH_PREC(OUT) \
    if (ctx->prec == SCR_F64) { if (fast) SCR_LAUNCH_COUNTS(double, true, OUT); else SCR_LAUNCH_COUNTS(double, false, OUT); } \
    else                      { if (fast) SCR_LAUNCH_COUNTS(float, true, OUT);  else SCR_LAUNCH_COUNTS(float, false, OUT); }
    if (out_dtype == SCR_COUNT_I32) { SCR_LAUNCH_PREC(int32_t) }
    else if (out_dtype == SCR_COUNT_U16) { SCR_LAUNCH_PREC(uint16_t) }
    else { SCR_LAUNCH_PREC(uint8_t) }
#undef SCR_LAUNCH_PREC
#undef SCR_LAUNCH_COUNTS
    ctx->launches++;
    return cudaGetLastError();
  };
  unsigned char* out_bytes = static_cast<unsigned char*>(counts_out);
  if (out_on_device) {
    CKC(launch(0, cells, counts_out, d_lam));
  } else {
    // double-buffered: the Poisson kernel of chunk i+1 overlaps the D2H copy of chunk i
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(cells, (256LL << 20) / (static_cast<int64_t>(n) * 4)));
    const size_t words = (static_cast<size_t>(chunk) * n * esz + 3) / 4;
    if ((rc = ensure_buf(ctx, ctx->d_cbuf[0], ctx->cbuf_cap[0], words))) { cleanup(); return rc; }
    if (cells > chunk && (rc = ensure_buf(ctx, ctx->d_cbuf[1], ctx->cbuf_cap[1], words))) { cleanup(); return rc; }
    int used[2] = {0, 0};
    int64_t ci = 0;
    for (int64_t c0 = 0; c0 < cells; c0 += chunk, ++ci) {
      const int b = static_cast<int>(ci & 1);
      const int64_t cnt = std::min(chunk, cells - c0);
      if (used[b]) CKC(cudaStreamWaitEvent(ctx->stream, ctx->ev_c[b], 0));
      CKC(launch(c0, cnt, ctx->d_cbuf[b], d_lam ? d_lam + c0 * n : nullptr));
      CKC(cudaEventRecord(ctx->ev_k[b], ctx->stream));
      CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_k[b], 0));
      CKC(cudaMemcpyAsync(out_bytes + static_cast<size_t>(c0) * n * esz, ctx->d_cbuf[b], static_cast<size_t>(cnt) * n * esz,
                          cudaMemcpyDeviceToHost, ctx->copy_stream));
      CKC(cudaEventRecord(ctx->ev_c[b], ctx->copy_stream));
      used[b] = 1;
    }
    CKC(cudaStreamSynchronize(ctx->copy_stream));
  }
  CKC(cudaStreamSynchronize(ctx->stream));
  if (lambda_out) CKC(cudaMemcpy(lambda_out, d_lam, static_cast<size_t>(cells) * n * sizeof(double), cudaMemcpyDeviceToHost));
  int herr = 0;
  CKC(cudaMemcpy(&herr, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  unsigned long long novf = 0;
  if (esz < 4) {
    CKC(cudaMemcpy(&novf, ctx->d_covf_count, sizeof(novf), cudaMemcpyDeviceToHost));
    const unsigned long long take = std::min(novf, dev_ovf_cap);
    if (take > 0 && ovf_out) {
      std::vector<int32_t> h(3 * take);
      CKC(cudaMemcpy(h.data(), ctx->d_covf, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
      // the kernel appends in arrival order: sort by (cell, gene) so that the list is deterministic
      std::vector<size_t> idx(take);
      for (size_t i = 0; i < take; ++i) idx[i] = i;
      std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
        return h[3 * a] != h[3 * b] ? h[3 * a] < h[3 * b] : h[3 * a + 1] < h[3 * b + 1];
      });
      for (size_t i = 0; i < take; ++i) {
        ovf_out[3 * i] = static_cast<int64_t>(h[3 * idx[i]]) - cell_begin;   // row of THIS call's output
        ovf_out[3 * i + 1] = h[3 * idx[i] + 1];
        ovf_out[3 * i + 2] = h[3 * idx[i] + 2];
      }
    }
    if (ovf_count_out) *ovf_count_out = static_cast<int64_t>(novf);
  }
#undef CKC
  cleanup();
  if (herr >= 100) return fail(ctx, SCR_E_CUDA, "count kernel bounds check " + std::to_string(herr) + " failed (SCR_DEBUG_CHECKS build)");
  if (herr) return fail(ctx, SCR_E_DRAWS, "supplied uniforms exhausted (raise draws_per_entry)");
  if (novf > dev_ovf_cap) return fail(ctx, SCR_E_RESOURCE, "overflow list too small: " + std::to_string(novf) + " counts reach the escape code (call again with that capacity)");
  return rc;
}
}  // namespace

int scr_counts_range(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift, double exp_scale,
                     const double* size_factors, uint64_t seed, const double* uniforms, int32_t draws_per_entry,
                     int32_t* counts_out, int out_on_device, double* lambda_out) {
  return counts_impl(ctx, cell_begin, cell_count, shift, exp_scale, size_factors, seed, uniforms, draws_per_entry, counts_out,
                     out_on_device, lambda_out, SCR_COUNT_I32, nullptr, 0, nullptr);
}

int scr_counts_compact(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift, double exp_scale,
                       const double* size_factors, uint64_t seed, int out_dtype, void* counts_out, int out_on_device,
                       int64_t* overflow_out, int64_t overflow_capacity, int64_t* overflow_count) {
  return counts_impl(ctx, cell_begin, cell_count, shift, exp_scale, size_factors, seed, nullptr, 0, counts_out, out_on_device,
                     nullptr, out_dtype, overflow_out, overflow_capacity, overflow_count);
}

int scr_counts_ext(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift, double exp_scale,
                   const double* size_factors, uint64_t seed, double nb_dispersion, double dropout_p, const double* uniforms,
                   int32_t draws_per_entry, int out_dtype, void* counts_out, int out_on_device, double* lambda_out,
                   int64_t* overflow_out, int64_t overflow_capacity, int64_t* overflow_count) {
  return counts_impl(ctx, cell_begin, cell_count, shift, exp_scale, size_factors, seed, uniforms, draws_per_entry, counts_out,
                     out_on_device, lambda_out, out_dtype, overflow_out, overflow_capacity, overflow_count, nb_dispersion, dropout_p);
}

int scr_counts(scr_ctx* ctx, const double* shift, double exp_scale, const double* size_factors, uint64_t seed,
               const double* uniforms, int32_t draws_per_entry, int32_t* counts_out, int out_on_device, double* lambda_out) {
  if (!ctx) return SCR_E_ARG;
  return scr_counts_range(ctx, 0, ctx->cells, shift, exp_scale, size_factors, seed, uniforms, draws_per_entry, counts_out,
                          out_on_device, lambda_out);
}

int scr_debug_philox(scr_ctx* ctx, int64_t count, const uint32_t* ctr4, uint32_t k0, uint32_t k1, uint32_t* out4) {
  NEED_GPU();
  if (count <= 0 || !ctr4 || !out4) return fail(ctx, SCR_E_ARG, "bad arguments");
  uint32_t *d_in = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_in, count * 16));
  CK(cudaMalloc(&d_out, count * 16));
  CK(upload_sync(ctx, d_in, ctr4, count * 16));
  scr::debug_philox_kernel<<<static_cast<unsigned>((count + 127) / 128), 128, 0, ctx->stream>>>(count, d_in, k0, k1, d_out);
  ctx->launches++;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(out4, d_out, count * 16, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "debug_philox");
  return SCR_OK;
}

int scr_debug_noise(scr_ctx* ctx, int64_t global_cell, uint32_t step, double sigma, uint64_t seed, uint32_t stream, double* out) {
  NEED_GPU();
  if (!ctx->has_net) return fail(ctx, SCR_E_STATE, "network not set");
  if (!out) return fail(ctx, SCR_E_ARG, "null output");
  const int n = ctx->op.n, nb = ctx->op.nb;
  int rc = ensure_stage(ctx, n);
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->d_stage, 0, n * sizeof(double), ctx->stream));
  const uint64_t gid = static_cast<uint64_t>(global_cell);
  const float rscale = static_cast<float>(-2.0 * 0.6931471805599453 * sigma * sigma);
  if (ctx->dense_mode) {
    scr::RoundKeys rk;
    scr::make_round_keys(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ stream, rk);
    scr::dense::debug_noise_dense_kernel<<<(n / 4 + 128) / 128, 128, 0, ctx->stream>>>(
        n, static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), step, rscale, rk, ctx->d_stage);
  } else {
    scr::debug_noise_kernel<<<(nb * 32 + 127) / 128, 128, 0, ctx->stream>>>(
        nb, static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), step, rscale, static_cast<uint32_t>(seed),
        static_cast<uint32_t>(seed >> 32) ^ stream, ctx->d_perm, ctx->d_stage);
  }
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(out, ctx->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SCR_OK;
}

int64_t scr_launch_count(const scr_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t scr_debug_path_launches(const scr_ctx* ctx, int which) {
  if (!ctx) return 0;
  return which == 0 ? ctx->sparse_launches : ctx->dense_launches;
}
double scr_last_kernel_ms(const scr_ctx* ctx) { return ctx ? ctx->last_ms : -1.0; }
int scr_probe_kernel_stats(scr_ctx* ctx, double* total_ms, int64_t* launches, int reset) {
  if (!ctx) return SCR_E_ARG;
  if (total_ms) *total_ms = ctx->total_probe_ms;
  if (launches) *launches = ctx->total_probe_launches;
  if (reset) { ctx->total_probe_ms = 0.0; ctx->total_probe_launches = 0; }
  return SCR_OK;
}
int scr_step_kernel_stats(scr_ctx* ctx, double* total_ms, int64_t* launches, int reset) {
  if (!ctx) return SCR_E_ARG;
  if (total_ms) *total_ms = ctx->total_step_ms;
  if (launches) *launches = ctx->total_step_launches;
  if (reset) { ctx->total_step_ms = 0.0; ctx->total_step_launches = 0; }
  return SCR_OK;
}

// ---- RegNetInference front end -------------------------------------------------------------------------

int scr_rank_transform(scr_ctx* ctx, int64_t n, int64_t cells, double* S, int64_t ld) {
  NEED_GPU();
  if (n <= 0 || cells <= 0 || !S || ld < cells) return fail(ctx, SCR_E_ARG, "bad arguments");
  if (n > 8192) return fail(ctx, SCR_E_RESOURCE, "rank transform: n > 8192 does not fit the per-cell shared-memory sort");
  if (cells > (1LL << 31) - 2) return fail(ctx, SCR_E_RESOURCE, "rank transform: too many cells");
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  const int nbins = static_cast<int>(2 * n + 2);
  const size_t smem1 = static_cast<size_t>(n_pad) * 10 + (scr::regnet::kHistBins + 1 + 32) * 4;
  const size_t smem2 = static_cast<size_t>(nbins + 1 + 32) * 4;
  const size_t smem_t = static_cast<size_t>(scr::regnet::kTileCells) * (scr::regnet::kTileBins + 2) * 4;
  CK(cudaFuncSetAttribute(scr::regnet::rank_cells_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_t)));
  CK(cudaFuncSetAttribute(scr::regnet::rank_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)));
  CK(cudaFuncSetAttribute(scr::regnet::rank_genes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));

  // slabs of ~256 MB (two: the copy of one overlaps the kernel of the other); a slab holds whole columns in
  // pass 1 (n x cw) and whole rows in pass 2 (gw x cells)
  const int64_t slab_elems = std::max<int64_t>({(256LL << 20) / 8, n, cells});
  const int64_t cw = std::max<int64_t>(1, std::min<int64_t>(cells, slab_elems / n));
  const int64_t gw = std::max<int64_t>(1, std::min<int64_t>(n, slab_elems / cells));
  uint16_t* d_keys = nullptr;
  double* d_slab[2] = {nullptr, nullptr};
  int* d_paths = nullptr;
  int* d_fb[2] = {nullptr, nullptr};
  cudaEvent_t k0[2] = {nullptr, nullptr}, k1[2] = {nullptr, nullptr};
  auto cleanup = [&]() {
    cudaFree(d_keys); cudaFree(d_slab[0]); cudaFree(d_slab[1]); cudaFree(d_paths); cudaFree(d_fb[0]); cudaFree(d_fb[1]);
    for (int b = 0; b < 2; ++b) { if (k0[b]) cudaEventDestroy(k0[b]); if (k1[b]) cudaEventDestroy(k1[b]); }
  };
#define CKR(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail_cuda(ctx, e__, #call); } } while (0)
  CKR(cudaMalloc(&d_keys, static_cast<size_t>(n) * cells * sizeof(uint16_t)));
  CKR(cudaMalloc(&d_slab[0], static_cast<size_t>(slab_elems) * 8));
  CKR(cudaMalloc(&d_slab[1], static_cast<size_t>(slab_elems) * 8));
  CKR(cudaMalloc(&d_paths, 2 * sizeof(int)));
  CKR(cudaMalloc(&d_fb[0], static_cast<size_t>(cw) * sizeof(int)));
  CKR(cudaMalloc(&d_fb[1], static_cast<size_t>(cw) * sizeof(int)));
  CKR(cudaMemsetAsync(d_paths, 0, 2 * sizeof(int), ctx->stream));
  CKR(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < 2; ++b) { CKR(cudaEventCreate(&k0[b])); CKR(cudaEventCreate(&k1[b])); }
  double kernel_ms = 0.0;
  auto reap = [&](int b) {   // the kernel that last used slab b has been waited for: add its time
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, k0[b], k1[b]) == cudaSuccess) kernel_ms += ms;
  };

  // pass 1: H2D column slab on the copy stream -> per-cell ranks on the compute stream
  int used[2] = {0, 0};
  int64_t ci = 0;
  for (int64_t c0 = 0; c0 < cells; c0 += cw, ++ci) {
    const int b = static_cast<int>(ci & 1);
    const int64_t cnt = std::min(cw, cells - c0);
    if (used[b]) { CKR(cudaEventSynchronize(k1[b])); reap(b); }
    CKR(cudaMemcpy2DAsync(d_slab[b], cnt * 8, S + c0, ld * 8, cnt * 8, n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CKR(cudaEventRecord(ctx->ev_c[b], ctx->copy_stream));
    CKR(cudaStreamWaitEvent(ctx->stream, ctx->ev_c[b], 0));
    CKR(cudaEventRecord(k0[b], ctx->stream));
    // counting path, tiled and coalesced; the cells it cannot take (non-integers, values >= 1024) are flagged
    // and ranked one CTA per cell by the general kernel
    scr::regnet::rank_cells_tile_kernel<<<static_cast<unsigned>((cnt + scr::regnet::kTileCells - 1) / scr::regnet::kTileCells), 512, smem_t, ctx->stream>>>(
        d_slab[b], cnt, static_cast<int>(n), d_keys, cells, c0, d_fb[b], d_paths);
    CKR(cudaGetLastError());
    scr::regnet::rank_cells_kernel<<<static_cast<unsigned>(cnt), 256, smem1, ctx->stream>>>(
        d_slab[b], cnt, static_cast<int>(n), n_pad, d_keys, cells, c0, d_paths, d_fb[b]);
    CKR(cudaGetLastError());
    ctx->launches++;
    CKR(cudaEventRecord(k1[b], ctx->stream));
    ctx->launches++;
    used[b] = 1;
  }
  CKR(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < 2; ++b) if (used[b]) reap(b);

  // pass 2: per-gene ranks into a row slab -> D2H on the copy stream
  used[0] = used[1] = 0;
  ci = 0;
  for (int64_t g0 = 0; g0 < n; g0 += gw, ++ci) {
    const int b = static_cast<int>(ci & 1);
    const int64_t cnt = std::min(gw, n - g0);
    if (used[b]) { CKR(cudaStreamWaitEvent(ctx->stream, ctx->ev_c[b], 0)); CKR(cudaEventSynchronize(k1[b])); reap(b); }
    CKR(cudaEventRecord(k0[b], ctx->stream));
    scr::regnet::rank_genes_kernel<<<static_cast<unsigned>(cnt), 512, smem2, ctx->stream>>>(d_keys, cells, nbins, g0, d_slab[b]);
    CKR(cudaGetLastError());
    CKR(cudaEventRecord(k1[b], ctx->stream));
    ctx->launches++;
    CKR(cudaEventRecord(ctx->ev_k[b], ctx->stream));
    CKR(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_k[b], 0));
    CKR(cudaMemcpy2DAsync(S + g0 * ld, ld * 8, d_slab[b], cells * 8, cells * 8, cnt, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CKR(cudaEventRecord(ctx->ev_c[b], ctx->copy_stream));
    used[b] = 1;
  }
  CKR(cudaStreamSynchronize(ctx->stream));
  CKR(cudaStreamSynchronize(ctx->copy_stream));
  for (int b = 0; b < 2; ++b) if (used[b]) reap(b);
  int paths[2] = {0, 0};
  CKR(cudaMemcpy(paths, d_paths, sizeof(paths), cudaMemcpyDeviceToHost));
#undef CKR
  cleanup();
  ctx->rank_kernel_ms = kernel_ms;
  ctx->rank_counting = paths[0];
  ctx->rank_sorted = paths[1];
  return SCR_OK;
}

int scr_rank_transform_stats(const scr_ctx* ctx, double* kernel_ms, int64_t* counting_cells, int64_t* sorted_cells) {
  if (!ctx) return SCR_E_ARG;
  if (kernel_ms) *kernel_ms = ctx->rank_kernel_ms;
  if (counting_cells) *counting_cells = ctx->rank_counting;
  if (sorted_cells) *sorted_cells = ctx->rank_sorted;
  return SCR_OK;
}

int scr_cell_pairs(scr_ctx* ctx, int64_t n, int64_t cells, const double* S, int64_t ld, const double* pseudotimes,
                   const int64_t* cell_types, double thresh, int64_t* pairs_out, int64_t capacity, int64_t* count_out) {
  NEED_GPU();
  if (n <= 0 || cells <= 0 || !S || ld < cells || !pseudotimes || !cell_types || !count_out)
    return fail(ctx, SCR_E_ARG, "bad arguments");
  // positions grouped by type (ascending), cells ascending inside a type: the reference's loop order
  int64_t tmax = -1;
  for (int64_t c = 0; c < cells; ++c) tmax = std::max(tmax, cell_types[c]);
  *count_out = 0;
  if (tmax < 0) return SCR_OK;
  std::vector<int64_t> start(static_cast<size_t>(tmax) + 2, 0);
  for (int64_t c = 0; c < cells; ++c)
    if (cell_types[c] >= 0) start[cell_types[c] + 1]++;
  for (int64_t t = 0; t <= tmax; ++t) start[t + 1] += start[t];
  const int64_t m = start[tmax + 1];
  std::vector<int64_t> order(m), seg_b(m), seg_e(m), fillp(start.begin(), start.end() - 1);
  std::vector<double> pt(m);
  for (int64_t c = 0; c < cells; ++c) {
    const int64_t t = cell_types[c];
    if (t < 0) continue;
    const int64_t pos = fillp[t]++;
    order[pos] = c;
    pt[pos] = pseudotimes[c];
    seg_b[pos] = start[t];
    seg_e[pos] = start[t + 1];
  }
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t need = static_cast<size_t>(n) * cells * 8 + static_cast<size_t>(m) * 48;
  if (need + (1ULL << 30) > free_b) return fail(ctx, SCR_E_RESOURCE, "cell pairs: state matrix does not fit in device memory");
  double *d_S = nullptr, *d_pt = nullptr;
  int64_t *d_order = nullptr, *d_sb = nullptr, *d_se = nullptr, *d_cnt = nullptr, *d_off = nullptr, *d_pairs = nullptr;
  auto cleanup = [&]() { cudaFree(d_S); cudaFree(d_pt); cudaFree(d_order); cudaFree(d_sb); cudaFree(d_se); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_pairs); };
#define CKP(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail_cuda(ctx, e__, #call); } } while (0)
  CKP(cudaMalloc(&d_S, static_cast<size_t>(n) * cells * 8));
  CKP(cudaMalloc(&d_pt, m * 8)); CKP(cudaMalloc(&d_order, m * 8)); CKP(cudaMalloc(&d_sb, m * 8));
  CKP(cudaMalloc(&d_se, m * 8)); CKP(cudaMalloc(&d_cnt, m * 8)); CKP(cudaMalloc(&d_off, m * 8));
  CKP(cudaMemcpy2DAsync(d_S, cells * 8, S, ld * 8, cells * 8, n, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_pt, pt.data(), m * 8, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_order, order.data(), m * 8, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_sb, seg_b.data(), m * 8, cudaMemcpyHostToDevice, ctx->stream));
  CKP(cudaMemcpyAsync(d_se, seg_e.data(), m * 8, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = static_cast<unsigned>((m + 255) / 256);
  scr::regnet::pairs_kernel<false><<<grid, 256, 0, ctx->stream>>>(m, d_pt, d_order, d_sb, d_se, thresh, d_S, cells, static_cast<int>(n), d_cnt, nullptr, nullptr);
  CKP(cudaGetLastError());
  ctx->launches++;
  std::vector<int64_t> cnt(m), off(m);
  CKP(cudaMemcpyAsync(cnt.data(), d_cnt, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CKP(cudaStreamSynchronize(ctx->stream));
  int64_t total = 0;
  for (int64_t i = 0; i < m; ++i) { off[i] = total; total += cnt[i]; }
  *count_out = total;
  if (pairs_out && capacity >= total && total > 0) {
    CKP(cudaMalloc(&d_pairs, static_cast<size_t>(total) * 16));
    CKP(cudaMemcpyAsync(d_off, off.data(), m * 8, cudaMemcpyHostToDevice, ctx->stream));
    scr::regnet::pairs_kernel<true><<<grid, 256, 0, ctx->stream>>>(m, d_pt, d_order, d_sb, d_se, thresh, d_S, cells, static_cast<int>(n), nullptr, d_off, d_pairs);
    CKP(cudaGetLastError());
    ctx->launches++;
    CKP(cudaMemcpyAsync(pairs_out, d_pairs, static_cast<size_t>(total) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CKP(cudaStreamSynchronize(ctx->stream));
  }
#undef CKP
  cleanup();
  return SCR_OK;
}

}  // extern "C"
