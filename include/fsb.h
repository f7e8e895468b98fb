/*
 * fsb.h -- C ABI of the B200-native flecsolve solve-loop back end ("fsb").
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json:
 * parallel CSR SpMV, vector operations, global reductions, point-Jacobi.
 * Every entry point is plain C (opaque handles, pointers, sizes, doubles);
 * no C++/torch types cross it.  The C++ header layer under
 * flecsolve_b200/include/flecsolve/ is the only intended caller besides
 * tests and bench.py (through ctypes).
 *
 * Each group cites the reference interface it replaces (paths relative to the
 * flecsolve source tree).
 *
 * Conventions
 *  - every function returns 0 (FSB_OK) or a positive fsb_status; the message is
 *    available from fsb_last_error() (thread local).
 *  - one host thread drives one context; handles are not thread safe.
 *  - all device work of a context is ordered on the context's stream, in the
 *    order the calls were made.  Calls may be *deferred*: element-wise
 *    operations, reductions and SpMV are queued and launched as fused kernels
 *    when a result is needed (fsb_red_get / fsb_vec_download / fsb_ctx_sync /
 *    fsb_ctx_flush).  Deferral never changes results element-wise: the fused
 *    kernels evaluate the queued statements in program order per element.
 *  - there is no CPU fallback.  Without a usable CUDA device fsb_ctx_create
 *    fails with FSB_ERR_NOGPU.
 */
#ifndef FSB_H
#define FSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsb_ctx_s * fsb_ctx_t;
typedef struct fsb_vec_s * fsb_vec_t;
typedef struct fsb_parcsr_s * fsb_parcsr_t;
typedef int64_t fsb_token_t; /* handle of a pending reduction result */

enum fsb_status {
	FSB_OK = 0,
	FSB_ERR_CUDA = 1,
	FSB_ERR_ARG = 2,
	FSB_ERR_NCCL = 3,
	FSB_ERR_NOGPU = 4,
	FSB_ERR_STATE = 5
};

enum fsb_option {
	FSB_OPT_FUSION = 0, /* 1 (default): fuse queued statements; 0: one kernel per call */
	FSB_OPT_SPMV_ROWS_PER_CTA = 1, /* tuning knob for matrices created afterwards */
	FSB_OPT_SPMV_THREADS = 2,
	FSB_OPT_TRACE = 3, /* 1: print every launched group signature to stderr */
	FSB_OPT_PROFILE = 4, /* 1: bracket every SpMV kernel with CUDA events (fsb_ctx_profile_read) */
	FSB_OPT_REPRODUCIBLE = 5 /* 1: SpMV row blocks are assigned to CTAs statically, so a fused dot is bitwise
	                            reproducible run to run (default; the cross-rank fold is in rank order);
	                            0: row blocks are claimed dynamically, which tolerates SMs shared with
	                            communication kernels (default when halo and reductions go through NCCL;
	                            results then vary in the last bits) */
	,
	FSB_OPT_JIT = 6 /* 1 (default): a statement group without an ahead-of-time kernel gets one compiled at run time
	                   (NVRTC, ~0.3 s once per group and process) instead of the generic program kernel, which remains
	                   the fallback when libnvrtc is not available; 0 (or FSB_JIT=0 in the environment): never compile */
	,
	FSB_OPT_TIMELINE = 7 /* n > 0: keep a device-side timeline of the next n kernel launches (fsb_ctx_timeline_read); 0: off */
	,
	FSB_OPT_SPECULATE = 9 /* 1: while the host waits for a reduction, the statement group that followed this point of the
	                         program last time is launched AHEAD of the result, with its coefficients left open; when the
	                         host then issues exactly that group, only the coefficients are handed over (a word in mapped
	                         memory) instead of a launch, otherwise the armed kernel is told to leave.  Takes the launch out
	                         of every host round trip of a Krylov iteration.  Default 0: an armed kernel occupies the
	                         device until the next call into this library, so the caller must not block on the device
	                         by other means in between -- the solver drivers switch it on for the duration of a solve.
	                         Setting it to 0 releases whatever is armed. */
	,
	FSB_OPT_SPMV_DICTIONARY = 8 /* 1 (default): y = A x of a matrix with at most 256 distinct values streams a one-byte value
	                               index per nonzero instead of the fp64 value (FSB_INFO_VALUE_DICTIONARY; same doubles, same
	                               order, same bits); 0 (or FSB_SPMV_DICT=0 in the environment): always stream the values */
};

enum fsb_stat {
	FSB_STAT_KERNEL_LAUNCHES = 0, /* kernels of this library launched so far */
	FSB_STAT_FUSED_STATEMENTS = 1, /* statements that shared a launch with another */
	FSB_STAT_HALO_EXCHANGES = 2,
	FSB_STAT_ALLREDUCES = 3,
	FSB_STAT_HOST_SYNCS = 4,
	FSB_STAT_UNMATCHED_GROUPS = 5, /* groups launched through the generic program kernel (no compile-time instantiation) */
	FSB_STAT_WAIT_NS = 6, /* host time spent waiting for reduction results (fsb_red_get / fsb_red_wait) */
	FSB_STAT_FLUSH_NS = 7, /* host time spent turning queued statements into launches */
	FSB_STAT_JIT_GROUPS = 8, /* groups launched through a run-time compiled kernel */
	FSB_STAT_BIND_NS = 9, /* parts of FSB_STAT_FLUSH_NS: matching statement groups with kernels, */
	FSB_STAT_EW_LAUNCH_NS = 10, /* ... the launch calls of element-wise kernels, */
	FSB_STAT_SPMV_LAUNCH_NS = 11, /* ... preparing and launching SpMV kernels */
	FSB_STAT_ARMED_HITS = 13, /* FSB_OPT_SPECULATE: groups that ran in a kernel launched ahead of the host's coefficients */
	FSB_STAT_ARMED_MISSES = 14 /* ... kernels launched ahead that had to be told to leave */
};

const char * fsb_last_error(void);
int fsb_version(void); /* 200: this layout of fsb_red_opts */
/* number of visible CUDA devices (0 when none / no driver) */
int fsb_device_count(void);

/* ---- context -----------------------------------------------------------
 * Replaces the FleCSI runtime objects the reference path leans on:
 * flecsi::scheduler (task issue order), one colour per process
 * (matrices/parcsr.hh:112-119) and the MPI communicator in topo::csr::init
 * (topo/csr.hh:474-480).  rank/nranks = this process' colour / colour count.
 * nccl_unique_id: 128 bytes produced by fsb_nccl_unique_id() on rank 0 and
 * broadcast by the caller (any transport); NULL when nranks == 1.           */
int fsb_nccl_unique_id(void * out128);
int fsb_ctx_create(int device, int rank, int nranks, const void * nccl_unique_id, fsb_ctx_t * out);
/* All ranks as threads of THIS process on ONE device: out[0 .. nranks) are the contexts of ranks 0 .. nranks-1 of one
 * group.  Every collective call (matrix creation, anything that waits for a reduction) must then be made by nranks host
 * threads, one per context, as nranks processes would.  The ranks' kernels share the device (row blocks small enough to
 * be resident together: a kernel that waits for a peer spins on the device), ghost exchange and all-reduce go through
 * the same peer-memory kernels as across GPUs -- for tests and for debugging a sharded run on one GPU, not for speed. */
int fsb_ctx_create_group(int device, int nranks, fsb_ctx_t * out);
int fsb_ctx_destroy(fsb_ctx_t ctx);
int fsb_ctx_flush(fsb_ctx_t ctx); /* launch everything queued; no host wait */
int fsb_ctx_sync(fsb_ctx_t ctx); /* flush + wait for the stream */
void * fsb_ctx_stream(fsb_ctx_t ctx); /* cudaStream_t all kernels are launched on */
int fsb_ctx_rank(fsb_ctx_t ctx);
int fsb_ctx_nranks(fsb_ctx_t ctx);
int fsb_ctx_set_option(fsb_ctx_t ctx, int option, int64_t value);
int fsb_ctx_get_stat(fsb_ctx_t ctx, int stat, int64_t * out);
int fsb_ctx_reset_stats(fsb_ctx_t ctx);
/* overwrite >= bytes of scratch so the L2 holds none of the caller's data */
int fsb_ctx_flush_l2(fsb_ctx_t ctx);
/* CUDA-event timing on the context's stream (the stream the kernels run on): record flushes the
 * queue, then records event `slot` (0..15); elapsed blocks until both events completed.          */
int fsb_ctx_event_record(fsb_ctx_t ctx, int slot);
int fsb_ctx_event_elapsed_ms(fsb_ctx_t ctx, int slot_start, int slot_stop, double * ms);
/* with FSB_OPT_PROFILE on: total device time and count of the SpMV kernels launched since the last
 * read (waits for them); used by bench.py for the roofline of the dominant kernel                 */
int fsb_ctx_profile_read(fsb_ctx_t ctx, double * spmv_ms, int64_t * spmv_launches);
/* same, split by block: index 0 = diag (owned columns) launches, 1 = offd (ghost columns) launches */
int fsb_ctx_profile_read_split(fsb_ctx_t ctx, double * ms2, int64_t * launches2);

/* Device-side timeline (FSB_OPT_TIMELINE): every kernel of this library stamps %globaltimer (ns) into its slot:
 * slot[0] earliest CTA start, [1] latest CTA end, [2] kind (1 SpMV, 2 its off-process launch, 3 fused SpMV + ghost
 * exchange, +8 Jacobi sweep, 16 + k element-wise program of k statements, 32 / 33 explicit halo push / unpack),
 * [3] ghost push published, [4] first CTA has its ghosts, [5] longest wait of a CTA for them (ns), [6] cross-rank all-reduce
 * begins, [7] result published, [8..10] ghost push: acknowledgements seen / stores issued / fenced (latest CTA each);
 * 0 or ~0 where a mark does not apply.  Reads up to max_slots slots of 16 words in
 * launch order, then clears the timeline.  This is how the per-iteration gaps in profiles/ are measured on several
 * ranks, where ncu cannot be used.                                                                              */
int fsb_ctx_timeline_read(fsb_ctx_t ctx, uint64_t * out, int64_t max_slots, int64_t * n_slots);

/* diagnostics: canonicalise the statement list raw[4 n] = {op, z, x, y} (ops as in csrc/program.h, vector ids
 * arbitrary small integers, -1 = unused) and compile its kernel for sm_100a with the run-time compiler.
 * Needs no GPU.  Returns FSB_OK and the cubin size, or FSB_ERR_STATE with the compiler log in `log`.   */
/* diagnostics: canonical form of a statement list (same input convention as above).  info[0] = 1 when an
 * ahead-of-time kernel is registered for it (with device-resident coefficients when `device_coefficients`),
 * info[1..5] = vectors, scalars, reductions, load mask, store mask; canon[6 n] receives the canonical statements
 * {op, z, x, y, a, b}.  Pure host code: no GPU needed.                                                        */
int fsb_debug_program_info(const int32_t * raw, int n, int device_coefficients, int32_t * info6, int32_t * canon);
int fsb_debug_jit_compile(const int32_t * raw, int n, int device_coefficients, int box_layout, int64_t * cubin_bytes,
                          char * log, int log_capacity);

/* ---- vectors -----------------------------------------------------------
 * A vector is the device image of one field on the reference's `cols' index
 * space: n_owned entries followed by n_ghost ghost entries
 * (topo/csr.hh:420-431, 524-526; vectors/data/topo_view.hh:24-71).          */
int fsb_vec_create(fsb_ctx_t ctx, int64_t n_owned, int64_t n_ghost, fsb_vec_t * out);
/* wrap caller-owned device memory of (n_owned + n_ghost) doubles, 16-byte aligned.  Handles that wrap the SAME memory
 * are the same vector to every call (fused statements see each other's writes); handles whose ranges overlap only
 * partly never share a fused launch, and using two of them in ONE call is rejected (FSB_ERR_ARG). */
int fsb_vec_wrap(fsb_ctx_t ctx, double * device_ptr, int64_t n_owned, int64_t n_ghost, fsb_vec_t * out);
int fsb_vec_destroy(fsb_vec_t v);
int64_t fsb_vec_local_size(fsb_vec_t v); /* vec::ops::topo_view::local_size, operations/topo_view.hh:276-280 */
int64_t fsb_vec_ghost_size(fsb_vec_t v);
double * fsb_vec_device_ptr(fsb_vec_t v);
/* host <-> device copies of owned entries [offset, offset+n) (flushes first) */
int fsb_vec_upload(fsb_vec_t v, const double * host, int64_t n, int64_t offset);
int fsb_vec_download(fsb_vec_t v, double * host, int64_t n, int64_t offset);

/* element-wise operations over the owned entries; any operands may alias.
 * Reference: vec::ops::topo_view dispatch (vectors/operations/topo_view.hh:43-210)
 * and the task bodies (vectors/operations/topo_tasks.hh:68-214).  The `_self'
 * alias variants of the reference are all covered by the alias-safe kernels.
 * Arithmetic is evaluated exactly as written there: products and sums are
 * rounded separately (no FMA contraction), so results are bit-identical to a
 * -ffp-contract=off x86-64 build of the reference.                         */
int fsb_vec_copy(fsb_vec_t z, fsb_vec_t x); /* z = x               (copy, topo_tasks.hh:80-84)  */
int fsb_vec_set(fsb_vec_t z, double a); /* z = a               (set_to_scalar, :68-70)      */
int fsb_vec_scale(fsb_vec_t z, double a, fsb_vec_t x); /* z = x * a  (scale/scale_self, :72-78)  */
int fsb_vec_add(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y); /* z = x + y           (:86-92)   */
int fsb_vec_sub(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y); /* z = x - y           (:94-108)  */
int fsb_vec_mul(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y); /* z = x * y           (:110-118) */
int fsb_vec_div(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y); /* z = x / y           (:120-134) */
int fsb_vec_recip(fsb_vec_t z, fsb_vec_t x); /* z = 1 / x                     (:136-142) */
int fsb_vec_linear_sum(fsb_vec_t z, double a, fsb_vec_t x, double b, fsb_vec_t y); /* z = a*x + b*y (:144-172) */
int fsb_vec_axpy(fsb_vec_t z, double a, fsb_vec_t x, fsb_vec_t y); /* z = a*x + y  (:174-191) */
int fsb_vec_axpby(fsb_vec_t z, double a, double b, fsb_vec_t x); /* z = a*x + b*z (:193-198) */
int fsb_vec_abs(fsb_vec_t z, fsb_vec_t x); /* z = |x|                       (:200-206) */
int fsb_vec_add_scalar(fsb_vec_t z, fsb_vec_t x, double a); /* z = x + a         (:208-214) */
/* mt19937(seed) + uniform_real_distribution(0,1) drawn sequentially over the
 * owned entries on the host, same seed on every rank (topo_tasks.hh:305-314) */
int fsb_vec_set_random(fsb_vec_t z, unsigned seed);
/* text dump `<prefix>-<rank>', one value per line (topo_tasks.hh:316-323) */
int fsb_vec_dump(fsb_vec_t x, const char * prefix);

/* global reductions.  Each call queues the reduction and returns a token;
 * fsb_red_get blocks until the (all-rank) value is on the host -- the
 * counterpart of flecsi::future::get() on scheduler().reduce<>()
 * (vectors/operations/topo_view.hh:212-274; bodies topo_tasks.hh:52-66,216-297).
 * A token may be read more than once until FSB_RED_RING newer tokens exist.   */
#define FSB_RED_RING 256
int fsb_vec_dot(fsb_vec_t x, fsb_vec_t y, fsb_token_t * tok); /* sum x*y                       */
int fsb_vec_sumsq(fsb_vec_t x, fsb_token_t * tok); /* sum x*x  (l2_norm_local; sqrt at get in C++) */
int fsb_vec_asum(fsb_vec_t x, fsb_token_t * tok); /* sum |x|  (l1_norm_local)                 */
int fsb_vec_powsum(fsb_vec_t x, int p, fsb_token_t * tok); /* sum pow(x,p) (lp_norm_local)    */
int fsb_vec_amax(fsb_vec_t x, fsb_token_t * tok); /* max |x|  (inf_norm_local)                */
int fsb_vec_min(fsb_vec_t x, fsb_token_t * tok); /* min x    (local_min)                      */
int fsb_vec_max(fsb_vec_t x, fsb_token_t * tok); /* max x    (local_max)                      */
int fsb_vec_global_size(fsb_vec_t x, int64_t * out); /* sum of local sizes (topo_view.hh:271-274) */
int fsb_red_get(fsb_ctx_t ctx, fsb_token_t tok, double * out);
int fsb_red_wait(fsb_ctx_t ctx, fsb_token_t tok);

/* ---- device scalars (solver variants without host reads) ----
 * The reference's Krylov loops read every reduction on the host and feed it back
 * as a coefficient of the next vector call (solvers/cg.hh:98-131: alpha = rho /
 * <p,Ap>, beta = rho' / rho).  These entry points keep that arithmetic on the
 * device: a reduction may deposit its all-rank value in a scalar slot, a
 * coefficient may be  scale * slot[num] / slot[den]  (evaluated with one IEEE
 * division and one multiplication, so scale = +-1 gives the bits the host loop
 * computes), and a reduction may carry the convergence test: once it passes, the
 * context's halt flag is raised on the device and every later element-wise
 * statement is skipped (vectors keep the converged iterate; reductions still
 * deliver their tokens), until fsb_ctx_halt_disarm.  The host can therefore run
 * several iterations ahead and look at residual norms late without changing the
 * iterate it finally returns.  The flag is looked at when a launch begins:
 * statements queued right behind the testing reduction may share its launch and
 * still run (in CG: z = P r and <r,z>, which only touch work vectors); a statement
 * whose coefficient names a slot stored by a reduction of the same run always
 * starts a new launch, and fsb_ctx_flush ends the current one explicitly.
 * Slot 0 is the constant 1.0: {v, 0, 0} is the plain number v.
 * Handles are plain indices of the context (at most 63 live at a time). */
typedef int32_t fsb_scalar_t;
typedef struct fsb_coef {
	double scale;
	fsb_scalar_t num, den;
} fsb_coef;
enum { FSB_HALT_NEVER = 0, FSB_HALT_IF_SQRT_LT = 1, FSB_HALT_IF_LT = 2 };
/* scalar arithmetic carried by a reduction: once its all-rank value is stored, the thread that publishes it evaluates
 * up to FSB_MAX_POST_OPS statements  slot[dst] = slot[a] op slot[b]  in order (IEEE double, one rounding each) -- how
 * single-reduction CG keeps  beta = g'/g,  alpha = g' / (d - beta g'/alpha)  on the device without a launch of its own */
enum { FSB_SOP_ADD = 0, FSB_SOP_SUB = 1, FSB_SOP_MUL = 2, FSB_SOP_DIV = 3, FSB_SOP_COPY = 4 /* slot[dst] = slot[a] */ };
#define FSB_MAX_POST_OPS 8
typedef struct fsb_scalar_op {
	int32_t op; /* FSB_SOP_* */
	fsb_scalar_t dst, a, b;
} fsb_scalar_op;
typedef struct fsb_red_opts {
	fsb_scalar_t store; /* slot that receives the value (0: none) */
	int halt_mode; /* FSB_HALT_*: raise the halt flag when [sqrt](value) < halt_threshold */
	double halt_threshold;
	int32_t n_post; /* scalar statements evaluated after the value is stored (0: none) */
	fsb_scalar_op post[FSB_MAX_POST_OPS];
} fsb_red_opts;
int fsb_scalar_create(fsb_ctx_t ctx, fsb_scalar_t * out);
int fsb_scalar_destroy(fsb_ctx_t ctx, fsb_scalar_t s);
int fsb_scalar_set(fsb_ctx_t ctx, fsb_scalar_t s, double value); /* in program order */
int fsb_scalar_get(fsb_ctx_t ctx, fsb_scalar_t s, double * out); /* launches queued work and waits */
/* z = a x + b y with device-resident coefficients (any aliasing, like fsb_vec_linear_sum) */
int fsb_vec_linear_sum_c(fsb_vec_t z, fsb_coef a, fsb_vec_t x, fsb_coef b, fsb_vec_t y);
/* dot / sum of squares (x == y) whose value also goes to opts->store and drives the halt flag */
int fsb_vec_dot_opts(fsb_vec_t x, fsb_vec_t y, const fsb_red_opts * opts, fsb_token_t * tok);
/* clear the halt flag and start honouring it / read it (optional), clear it and stop honouring it */
int fsb_ctx_halt_arm(fsb_ctx_t ctx);
int fsb_ctx_halt_disarm(fsb_ctx_t ctx, int * was_halted);

/* ---- structured-grid ("narray") vectors and operators ----
 * The reference's example applications keep their fields on a FleCSI narray mesh: a padded N-d
 * array per colour whose vector operations run over dofs() = the interior sub-box, in
 * colexicographic order, x fastest (examples/poisson/mesh.hh:157-161, index_util.hh:33-71);
 * the layers around it hold boundary data that operators read but vector operations never touch.
 * fsb_vec_create_box makes such a field: extents[dim] of the padded array, dofs = [lo, hi) per axis.
 * Every element-wise call and reduction above works on these handles (operands of one call must
 * share the box); fsb_vec_local_size is the number of dofs; upload/download move dofs in dof order;
 * the *_all calls move the whole padded array (boundary data).
 * fsb_parcsr_create_box_stencil assembles the (2 dim + 1)-point operator
 *     (A u)(i) = center u(i) + sum_axis off[axis] (u(i - e_axis) + u(i + e_axis))
 * over the dofs of such a box as CSR whose column indices are storage offsets of the padded array,
 * so boundary layers take part exactly as in the reference's stencil operators
 * (examples/poisson/poisson.cc:44-82); fsb_parcsr_spmv runs it through the same SpMV kernel
 * (fused dot included).
 * fsb_parcsr_create_box_fvm assembles the finite-volume diffusion operator of
 * physics/volume_diffusion/diffusion.hh:84-207,  v = -beta div(b grad u) + alpha vol a u,  from its coefficient fields:
 *     flux_axis(c) = b_axis(c) kface[axis] (u(c + e_axis) - u(c)),   kface[axis] = dA_axis / dx_axis,
 *     v(c) = -beta sum_axis (flux_axis(c) - flux_axis(c - e_axis)) + alpha vol a(c) u(c);
 * a and b_axis (host pointers) are padded arrays with the box's extents, b_axis(c) = coefficient of the face between
 * c and c + e_axis.  Entries: towards c +- e_axis  -(beta (b kface)),  centre  beta sum (b kface) + (alpha vol) a(c).
 * Several ranks: the mesh is cut into slabs along its LAST axis, one per rank in rank order, and every rank passes the
 * extents of ITS padded array (the same across the other axes).  The pad plane of the last axis that faces a
 * neighbouring rank is a ghost plane: it mirrors the neighbour's outermost dof plane and fsb_parcsr_spmv /
 * fsb_parcsr_halo_exchange refresh it over peer memory before the operator reads it (FleCSI's ghost copy of an narray
 * field); all other pad layers hold boundary data.  Both creators are collective then.                               */
int fsb_vec_create_box(fsb_ctx_t ctx, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi, fsb_vec_t * out);
int fsb_vec_box_upload_all(fsb_vec_t v, const double * host);
int fsb_vec_box_download_all(fsb_vec_t v, double * host);
int fsb_parcsr_create_box_stencil(fsb_ctx_t ctx, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi,
                                  double center, const double * off, fsb_parcsr_t * out);
int fsb_parcsr_create_box_fvm(fsb_ctx_t ctx, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi, double beta,
                              double alpha, double vol, const double * kface, const double * a, const double * const * bface,
                              fsb_parcsr_t * out);

/* ---- parallel CSR matrix ------------------------------------------------
 * Device image of mat::parcsr (matrices/parcsr.hh:101-177) on the topology
 * topo::csr (topo/csr.hh): per rank a `diag' CSR over owned columns and an
 * `offd' CSR whose column indices are n_owned + (rank of the global column id
 * in the sorted unique ghost list), plus colmap[ghost] = global id
 * (topo/csr.hh:482-543 color(), :553-618 init_mats()).
 *
 * fsb_parcsr_create does color() + init_mats() + the ghost copy plan
 * (topo/csr.hh:116-187, 277-304) from this rank's rows in global numbering:
 *   row_part[nranks+1]  row/column partition offsets (same for rows and
 *                       columns, as set_block_map gives; matrices/parcsr.hh:170-172)
 *   rowptr[n_local+1], col[nnz] (global column ids, int64), val[nnz] on the host.
 * Collective over all ranks of the context (exchanges send lists).          */
int fsb_parcsr_create(fsb_ctx_t ctx,
                      int64_t n_global,
                      const int64_t * row_part,
                      const int64_t * rowptr,
                      const int64_t * col,
                      const double * val,
                      fsb_parcsr_t * out);
/* synthetic stencil operators generated on the device:
 * kind 7: diag 6 / off -1;  kind 27: diag 26 / off -1;  kind 5: 2-D (nz==1) diag 4 / off -1;
 * kind 107: 7-point with Neumann closure (diagonal = number of neighbours inside the box; config 4's second component).
 * Dirichlet truncation, columns ascending, rows g = i + nx*(j + ny*k) split into equal contiguous blocks like the
 * reference's equal_map (the first n % nranks blocks one row longer; z-slabs of whole planes when nz % nranks == 0).
 * A block must be at least as long as the stencil reaches (one plane; plane + nx + 1 for kind 27).
 * diag_shift is added to the diagonal, scale multiplies every entry.        */
int fsb_parcsr_create_stencil(fsb_ctx_t ctx, int kind, int64_t nx, int64_t ny, int64_t nz,
                              double diag_shift, double scale, fsb_parcsr_t * out);
int fsb_parcsr_destroy(fsb_parcsr_t A);
int64_t fsb_parcsr_local_rows(fsb_parcsr_t A);
int64_t fsb_parcsr_global_rows(fsb_parcsr_t A);
int64_t fsb_parcsr_num_ghosts(fsb_parcsr_t A);
int64_t fsb_parcsr_row_begin(fsb_parcsr_t A);
int64_t fsb_parcsr_local_nnz(fsb_parcsr_t A, int which /* 0 diag, 1 offd */);
/* how the device holds the matrix (diagnostics / tests / bench byte counts):
 *   FSB_INFO_WINDOW_FORMAT  1 when the owned-column block also has the window format (10 B per nonzero: fp64 value +
 *                           16-bit position inside the x segments its row block stages in shared memory)
 *   FSB_INFO_ROW_BLOCKS     row blocks (units of work of the SpMV pipeline) of the owned-column block
 *   FSB_INFO_WINDOW_X       x entries the largest row block stages (window format)
 *   FSB_INFO_FUSED_HALO     1 when y = A x is ONE launch that also does the ghost exchange over peer memory
 *   FSB_INFO_WIDE_OFFSETS   1 when row offsets are 64-bit (local nnz >= 2^31)
 *   FSB_INFO_VALUE_DICTIONARY  number of distinct values of the owned-column block when there are at most 256 of them
 *                           and the block has the window format -- the SpMV then streams 3 B per slot (one-byte value
 *                           index + 16-bit position; rows padded to multiples of 8 slots, FSB_INFO_DICTIONARY_SLOTS)
 *                           instead of 10 B per nonzero; 0 otherwise                                                */
enum fsb_parcsr_info_key {
	FSB_INFO_WINDOW_FORMAT = 0,
	FSB_INFO_ROW_BLOCKS = 1,
	FSB_INFO_WINDOW_X = 2,
	FSB_INFO_FUSED_HALO = 3,
	FSB_INFO_WIDE_OFFSETS = 4,
	FSB_INFO_VALUE_DICTIONARY = 5,
	FSB_INFO_DICTIONARY_SLOTS = 6 /* slots of the dictionary stream: nonzeros + the padding of every row to a multiple of 8 */
};
int64_t fsb_parcsr_info(fsb_parcsr_t A, int key);
/* copy the split representation back to the host (tests: compare with the
 * oracle's color()/init_mats() restatement).  Any pointer may be NULL.      */
int fsb_parcsr_download(fsb_parcsr_t A, int which, int64_t * rowptr, int32_t * col, double * val);
int fsb_parcsr_download_colmap(fsb_parcsr_t A, int64_t * colmap);
/* y = A x: ghost exchange of x (if stale) overlapped with the diag block, then
 * the offd block accumulates: y = diag*x_owned + offd*x   (matrices/parcsr.hh:61-91,
 * matrices/seq.hh:178-194).  Row sums are accumulated in column-index order.  */
int fsb_parcsr_spmv(fsb_parcsr_t A, fsb_vec_t x, fsb_vec_t y);
/* d = 1 / diag(A) over owned rows: the vector the reference's Dinv() operator
 * holds as a diagonal CSR (util/test/mesh.hh:123-140)                        */
int fsb_parcsr_extract_dinv(fsb_parcsr_t A, fsb_vec_t d);
/* weighted point-Jacobi sweeps, mg::bound_jacobi::relax (solvers/mg/jacobi.hh:44-93):
 * nrelax times: tmp = x (incl. ghosts); x[r] = omega/a_rr * (b[r] - sum_{c!=r} a_rc tmp[c]) + (1-omega) tmp[r] */
int fsb_parcsr_jacobi_relax(fsb_parcsr_t A, double omega, int64_t nrelax, fsb_vec_t b, fsb_vec_t x, fsb_vec_t tmp);
/* explicit ghost update of x (the copy plan of topo/csr.hh:237-245); normally implicit in spmv */
int fsb_parcsr_halo_exchange(fsb_parcsr_t A, fsb_vec_t x);

#ifdef __cplusplus
}
#endif
#endif
