/* scrutiny_b200 -- C ABI of the B200-native MatriSeq simulation hot path.
 *
 * The reference (steffen12/scRutiNy) is pure Python/numpy and has no FFI; its boundary for this
 * path is the module-level function surface of RNAscrutiny/RNAscrutiny/MatriSeq.py (cited "MS").
 * Each entry point below names the reference function it replaces.  The Python mirror of that
 * surface lives in scrutiny_b200/MatriSeq.py and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or a negative SCR_E_*
 *     code and never throws; scr_last_error(ctx) gives the message for the last failure.
 *   - one context per GPU; a context is NOT thread-safe; calls are synchronous on return unless
 *     stated otherwise (the work is enqueued on the context's own CUDA stream).
 *   - host matrices use the reference's layout: float64, (genes, cells) C-order, `ld` = row
 *     stride in elements (MS:340-365 builds them that way).  Device state is cell-major, in the
 *     context's precision (float32 throughput path or float64 parity path), gene order permuted
 *     internally (DESIGN.md "Data layout").
 *   - cells are addressed by LOCAL slot 0..C-1; the global id (cell_offset + slot) keys the
 *     counter-based RNG so results do not depend on how cells are sharded over GPUs.
 *   - there is no CPU fallback: without a CUDA device scr_create fails with SCR_E_CUDA.
 */
#ifndef SCRUTINY_B200_H
#define SCRUTINY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scr_ctx scr_ctx;

enum {
  SCR_OK = 0,
  SCR_E_ARG = -1,      /* bad argument                                           */
  SCR_E_CUDA = -2,     /* CUDA runtime error (message in scr_last_error)          */
  SCR_E_STATE = -3,    /* call order: network / cells not set yet                 */
  SCR_E_RESOURCE = -4, /* problem does not fit the kernel's on-chip budget        */
  SCR_E_DRAWS = -5     /* supplied uniforms exhausted in scr_counts               */
};

enum { SCR_F32 = 0, SCR_F64 = 1 };

/* element type of the compact count output (scr_counts_compact) */
enum { SCR_COUNT_I32 = 0, SCR_COUNT_U16 = 1, SCR_COUNT_U8 = 2 };

/* RNG stream ids (mixed into the Philox key) */
enum { SCR_STREAM_STEP = 0, SCR_STREAM_PROBE = 1, SCR_STREAM_COUNTS = 2, SCR_STREAM_ALPHA = 16 };

int scr_abi_version(void);
/* Free-running noise stream of the sparse stepper (csrc/step_kernel.cuh): 2 = one Philox4x32-10 call per (cell, step,
 * 128-gene block, lane pair) carrying eight normals, each word split into a 16-bit radius and a 16-bit angle uniform
 * (restated in oracle/philox_ref.py: stepper_noise); 1 = the round-1 stream (one call per lane, 23-bit uniforms), only in
 * A/B builds with -DSCR_NOISE16=0. */
int scr_noise_stream_version(void);

/* ---- context ------------------------------------------------------------------------------ */
int scr_create(scr_ctx** out, int device_ordinal, int precision /* SCR_F32 | SCR_F64 */);
void scr_destroy(scr_ctx* ctx);
const char* scr_last_error(const scr_ctx* ctx); /* ctx may be NULL: last create error */
int scr_sync(scr_ctx* ctx);

/* ---- regulatory network: the `gc` argument of developCell (MS:503-504), already divided by
 * alpha (EX:49).  CSR over rows = regulated gene.  Builds the on-chip operator (gene
 * permutation, blocked-ELL tiles, long-row chunks).  Any density is accepted; float32 contexts
 * switch scr_run_phase to the tcgen05 split-precision GEMM stepper (fp16 split by default,
 * SCR_DENSE_KIND=tf32 for 3xTF32) when nnz >= 5 % of n^2 (the sparse operator is still built:
 * float64 contexts use it). */
int scr_set_network_csr(scr_ctx* ctx, int32_t n, int64_t nnz, const int32_t* rowptr,
                        const int32_t* col, const double* val);
/* same from a dense row-major (n, n) float64 matrix (what the reference passes around) */
int scr_set_network_dense(scr_ctx* ctx, int32_t n, const double* gc_rowmajor);
/* Read-out context: n genes and NO network (transformData, MS:816-891, needs none).  Only the state and read-out entry
 * points work afterwards (scr_alloc_cells, scr_set_states*, scr_get_states, scr_gene_sums, scr_counts*); scr_run_phase and
 * scr_probe return SCR_E_STATE.  No on-chip budget applies: any n the count kernel's staging row holds (n <= 58 000). */
int scr_set_genes(scr_ctx* ctx, int32_t n);

/* Additive term of the drive, per gene (n doubles, caller's gene order; NULL clears it): scr_run_phase and scr_probe then
 * evaluate tanh(gc (x - mu) + bias + noise).  This is how a PER-GENE geneMeanLevels vector (developCell broadcasts
 * `x - geneMeanLevels`, MS:530; Matri-seqFinal.py passes np.zeros(n)) reaches the kernels: bias = -gc mu_vec, const + mu_vec,
 * scalar mu = 0 (scrutiny_b200/engine.py does that).  Stays set until cleared or the network is replaced. */
int scr_set_drive_bias(scr_ctx* ctx, const double* bias);

/* Test hook, host only (no GPU work): y = W x computed by walking the SAME packed operator the
 * kernels read (internal order, tiles, chunks), so packing bugs show up in the CPU test-suite. */
int scr_debug_host_matvec(const scr_ctx* ctx, const double* x, double* y);
/* The block records the stepper's warps walk, as the host uploads them (csrc/step_kernel.cuh "Shared-memory layout"):
 * records[4*i .. 4*i+3] = {first gene, first slot, end slot, flags (1 ping-ponged, 2 long rows, 4 last regulator block of
 * its warp, 8 first block of its warp and ping-ponged)} for list position i; list_bounds[w] .. list_bounds[w+1] is the
 * list of warp w of a cell group.  Returns the number of records (= 128-gene blocks) or a negative error code. */
int scr_debug_block_records(const scr_ctx* ctx, int32_t* records, int32_t max_records, int32_t* list_bounds,
                            int32_t max_bounds);
/* The internal gene order (DESIGN.md "Data layout"): perm_out[i] = original gene at internal position i, -1 for padding;
 * count must be n_pad (scr_get_layout out[0]). */
int scr_debug_gene_order(const scr_ctx* ctx, int32_t* perm_out, int32_t count);
/* Layout facts for DESIGN.md / tests: out[0]=n_pad, [1]=ping-pong genes, [2]=tile slots,
 * [3]=long-row chunks, [4]=chunk length, [5]=row cap, [6]=warps per cell group,
 * [7]=cell groups per CTA, [8]=dynamic smem bytes, [9]=operator resident in smem (0/1),
 * [10]=cells per group, [11]=dense (tcgen05 GEMM) stepper selected (0/1), [12]=operator bytes staged in
 * smem when it is only partly resident, [13]=operator bytes in total, [14]=modelled shared-memory wavefronts of one
 * step's gathers (LDS.128 bank model, csrc/operator_pack.h), [15]=their conflict-free floor, [16]=cell groups held in
 * global memory (0/1: the large-n fallback, n above ~7 000 genes in float32; any n up to 32 768 is accepted),
 * [17]=operand encoding of the dense stepper (0 = sparse stepper, 1 = fp16 split, 2 = 3xTF32: csrc/dense_kernel.cuh). */
int scr_get_layout(const scr_ctx* ctx, int64_t* out, int count);

/* ---- cell states: `currentStates` of findAllNextStates (MS:540-554) --------------------------- */
int scr_alloc_cells(scr_ctx* ctx, int64_t cells, int64_t global_cell_offset);
int scr_set_states(scr_ctx* ctx, int64_t cell0, int64_t count, const double* S_genes_by_cells,
                   int64_t ld);
/* every cell <- x0 (generateInitialStates(sameInitialCell=True), MS:360-361) */
int scr_set_states_broadcast(scr_ctx* ctx, const double* x0 /* n */);
int scr_get_states(scr_ctx* ctx, int64_t cell0, int64_t count, double* S_genes_by_cells,
                   int64_t ld);

/* ---- stepping: replaces the two per-cell loops of findAllNextStates (MS:587-605), i.e.
 * findCellNextState (MS:683-736) x developCell (MS:503-537), for one lineage node.
 *   cell_ids[m]  local slots (each at most once);  steps[m]  iterations for that cell (ragged);
 *   step_base[m] iterations the cell has already done (totalCellIterations, keys the noise);
 *   proj, cnst   the node's vectors (n);  dt, sigma, mu = timeStep, noiseDisp, geneMeanLevels.
 *   noise        NULL -> free-running Philox noise keyed (seed, global cell, step, gene);
 *                else float64 (m, noise_steps, n): noise[i][s][g] is ADDED inside tanh at local
 *                step s of cell_ids[i] (already scaled by sigma) -- the parity mode. */
int scr_run_phase(scr_ctx* ctx, int64_t m, const int64_t* cell_ids, const int32_t* steps,
                  const int64_t* step_base, const double* proj, const double* cnst, double dt,
                  double sigma, double mu, uint64_t seed, const double* noise, int64_t noise_steps);

/* ---- convergence probe: the sampled-cell loop of findOptimalIterations (MS:650-669) and of
 * findOptimalAlpha (MS:456-482).  Steps COPIES of the k sampled cells until the mean of the last
 * avg_num values of mean|dx|/dt is <= thresh or max_iter is reached; states are not modified.
 *   w_div[k] or NULL: per-sample divisor of the network (the alpha ladder, MS:461);
 *   iters_out[k]: iterations executed;  normdiff_out (k, max_iter) float64 or NULL: trace. */
int scr_probe(scr_ctx* ctx, int32_t k, const int64_t* sample_ids, const double* w_div,
              const double* proj, const double* cnst, double dt, double sigma, double mu,
              double thresh, int32_t max_iter, int32_t avg_num, uint64_t seed, uint32_t stream,
              const double* noise /* (k, max_iter, n) or NULL */, int32_t* iters_out,
              double* normdiff_out);

/* The same probe for `nprobes` lineage nodes at once (findAllNextStates probes each node right before its phase, MS:583-586;
 * the sampled cells of SIBLING nodes are final as soon as the parent's phase is, so their probes are independent).
 * Identical results to nprobes scr_probe calls (w_div = noise = NULL); the launches run side by side on separate streams
 * where the path allows it (sparse stepper), one after the other otherwise.
 *   k[nprobes] samples per node; sample_ids / iters_out: concatenated; proj, cnst: (nprobes, n); seeds[nprobes]. */
int scr_probe_multi(scr_ctx* ctx, int32_t nprobes, const int32_t* k, const int64_t* sample_ids,
                    const double* proj, const double* cnst, double dt, double sigma, double mu,
                    double thresh, int32_t max_iter, int32_t avg_num, const uint64_t* seeds, uint32_t stream,
                    int32_t* iters_out);

/* ---- read-out: transformData (MS:816-891) --------------------------------------------------- */
/* per-gene SUM over this context's cells (MS:851 needs the mean over ALL cells: all-reduce
 * these n doubles across ranks, divide by the global cell count). */
int scr_gene_sums(scr_ctx* ctx, double* sums_out /* n */);
/* counts[c][g] = Poisson(max(exp(exp_scale*(x[c][g]+shift[g])) * size[c], 0))   (MS:873-885)
 *   uniforms NULL -> Philox uniforms keyed (seed, global cell, gene, draw);
 *            else float64 (cells, n, draws_per_entry) consumed in order per entry.
 *   counts_out: int32 (cells, n) row-major, host or device memory (out_on_device);
 *   lambda_out: float64 (cells, n) HOST buffer or NULL (parity aid).
 *   Rates above 2e9 (or +inf from an overflowing exp) are sampled as 2e9: counts are int32 across this ABI
 *   (numpy returns int64 there; nothing in MatriSeq's parameter range comes near).
 *   With Philox uniforms and no lambda_out, entries with lambda < 10 are first screened in float32 with a guard band
 *   and redone by the float64 sampler when undecided: the integers are those of the float64 sampler in every case
 *   (environment SCR_COUNTS_EXACT=1 switches the screening off, for A/B tests). */
int scr_counts(scr_ctx* ctx, const double* shift /* n */, double exp_scale,
               const double* size_factors /* cells */, uint64_t seed, const double* uniforms,
               int32_t draws_per_entry, int32_t* counts_out, int out_on_device, double* lambda_out);
/* the same for local cells [cell_begin, cell_begin + cell_count) only: size_factors, uniforms, counts_out and
 * lambda_out hold cell_count rows.  Lets a caller hand finished row blocks on (NCCL gather, file writer) while the
 * next block is sampled; scr_counts is this call over all cells. */
int scr_counts_range(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift /* n */,
                     double exp_scale, const double* size_factors /* cell_count */, uint64_t seed,
                     const double* uniforms, int32_t draws_per_entry, int32_t* counts_out, int out_on_device,
                     double* lambda_out);

/* Compact read-out (free-running Philox stream only): the same integers as scr_counts_range, stored as out_dtype
 * (SCR_COUNT_U8 / SCR_COUNT_U16 / SCR_COUNT_I32), row-major (cell_count, n), host or device memory.  A count that does not
 * fit below the type's largest value (255 / 65535) is stored AS that value (escape code) and listed in overflow_out as
 * int64 triplets {row of this call's output, gene, count}, sorted by (row, gene); *overflow_count receives the number of
 * such entries.  If it exceeds overflow_capacity the call fails with SCR_E_RESOURCE after setting *overflow_count (the
 * count matrix itself is complete; call again with that capacity to get the list).  The reference stores float64
 * (MS:891): at config c5 that is 24 GB per 10^6 cells against 3 GB here (mean rate ~2, MS:855-873 defaults). */
int scr_counts_compact(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift /* n */, double exp_scale,
                       const double* size_factors /* cell_count */, uint64_t seed, int out_dtype, void* counts_out,
                       int out_on_device, int64_t* overflow_out, int64_t overflow_capacity, int64_t* overflow_count);

/* EXTENSIONS of the read-out, default off (north_star (4); the reference has only np.random.poisson, MS:884, and a
 * commented-out dropout, MS:887-889).  With nb_dispersion = dropout_p = 0 this is scr_counts_range / scr_counts_compact.
 *   nb_dispersion phi > 0: counts ~ negative binomial with mean lambda and variance lambda + phi lambda^2, drawn as
 *       Poisson(lambda G), G ~ Gamma(1/phi, phi) by Marsaglia-Tsang;
 *   dropout_p in (0, 1]: every entry is independently set to 0 with that probability (the parallel form of the reference's
 *       commented-out "zero a random share of each gene's cells");
 *   per entry the uniforms (supplied, or the Philox stream of scr_counts) are consumed in the order dropout, Gamma, Poisson:
 *   restated in oracle/poisson_oracle.c (oracle_counts_ext_batch), bit-exact given the same uniforms.
 * uniforms / lambda_out as in scr_counts_range (lambda_out receives the rate BEFORE the Gamma factor); out_dtype, overflow_*
 * as in scr_counts_compact (overflow_out may be NULL for SCR_COUNT_I32). */
int scr_counts_ext(scr_ctx* ctx, int64_t cell_begin, int64_t cell_count, const double* shift /* n */, double exp_scale,
                   const double* size_factors /* cell_count */, uint64_t seed, double nb_dispersion, double dropout_p,
                   const double* uniforms, int32_t draws_per_entry, int out_dtype, void* counts_out, int out_on_device,
                   double* lambda_out, int64_t* overflow_out, int64_t overflow_capacity, int64_t* overflow_count);

/* ---- debugging / parity aids ------------------------------------------------------------------ */
/* raw Philox4x32-10 words for counters (c0..c3)[count] under key (k0,k1): device generator check */
int scr_debug_philox(scr_ctx* ctx, int64_t count, const uint32_t* ctr4, uint32_t k0, uint32_t k1,
                     uint32_t* out4);
/* the noise the free-running stepper adds for (global cell, step) -> out[n] float64 */
int scr_debug_noise(scr_ctx* ctx, int64_t global_cell, uint32_t step, double sigma, uint64_t seed,
                    uint32_t stream, double* out /* n */);
/* kernel launches issued by this context so far (bench.py reports gpu_launches) */
int64_t scr_launch_count(const scr_ctx* ctx);
/* launches of a specific stepper so far: which = 0 persistent sparse stepper (one per phase), 1 dense GEMM stepper (one per
 * timestep of a large batch; one per phase or probe chunk when the steps run inside the launch) */
int64_t scr_debug_path_launches(const scr_ctx* ctx, int which);
/* device-time of the last scr_run_phase stepping kernel in milliseconds (CUDA events on the
 * context's stream); < 0 if none yet */
double scr_last_kernel_ms(const scr_ctx* ctx);
/* accumulated device time (ms) and number of scr_run_phase stepping-kernel launches since the
 * last reset; bench.py derives roofline.achieved from these */
int scr_step_kernel_stats(scr_ctx* ctx, double* total_ms, int64_t* launches, int reset);
/* the same for the sparse convergence-probe kernel (scr_probe): small configurations spend a large share of a pass there */
int scr_probe_kernel_stats(scr_ctx* ctx, double* total_ms, int64_t* launches, int reset);


/* ---- RegNetInference front end ("RNI" = RNAscrutiny/RNAscrutiny/RegNetInference.py).
 * These two calls need only a context (no network / cells): they work on the matrices the caller passes. */

/* RNI:69-78 convertToS(n, cells, cellStates): in place.  Every cell's genes are replaced by
 * 2*((r+1)/n)-1 with r = scipy.stats.rankdata 'average' ranks, then every gene's cells likewise with `cells`.
 * S is (n, cells) float64, row stride ld.  Results are bit-identical to the reference (integer rank work;
 * the final float64 expression is evaluated in the reference's operation order).  n <= 8192; finite inputs. */
int scr_rank_transform(scr_ctx* ctx, int64_t n, int64_t cells, double* S_genes_by_cells, int64_t ld);

/* RNI:397-416 (and RNI:764-777 inside getCVCost): all [cell1, cell2] of the same type with
 * 0 < pt[cell1] - pt[cell2] <= thresh whose state columns differ, ordered by type, cell1, cell2 ascending.
 * count_out receives the number of pairs; rows are written to pairs_out (int64 [count][2]) only when
 * pairs_out != NULL and capacity >= count (call once with NULL to size the buffer).  Negative types match
 * nothing, like the reference's range(max(types)+1) loop. */
int scr_cell_pairs(scr_ctx* ctx, int64_t n, int64_t cells, const double* S_genes_by_cells, int64_t ld,
                   const double* pseudotimes /* cells */, const int64_t* cell_types /* cells */, double thresh,
                   int64_t* pairs_out, int64_t capacity, int64_t* count_out);

/* kernels of the last scr_rank_transform: device milliseconds (CUDA events, both passes), cells ranked by the
 * counting path and by the sorting path */
int scr_rank_transform_stats(const scr_ctx* ctx, double* kernel_ms, int64_t* counting_cells, int64_t* sorted_cells);

#ifdef __cplusplus
}
#endif
#endif /* SCRUTINY_B200_H */
