// Internal declarations shared by the translation units of libfsb.so.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fsb.h"
#include "program.h"

namespace fsb {
struct xrank_info;
struct red_out;
struct setup_exchange;
struct speculation; // fuser.cu: armed launches (FSB_OPT_SPECULATE)
constexpr int MAX_SCALARS = 64; // device scalars per context; slot 0 is the constant 1.0
}

namespace fsb {

struct error : std::runtime_error {
	int code;
	error(int c, const std::string & m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string & m);

#define FSB_CUDA(call)                                                                      \
	do {                                                                                    \
		cudaError_t e_ = (call);                                                            \
		if (e_ != cudaSuccess)                                                              \
			throw ::fsb::error(FSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + \
			                                     " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
	} while (0)

#define FSB_NCCL(call)                                                                      \
	do {                                                                                    \
		ncclResult_t r_ = (call);                                                           \
		if (r_ != ncclSuccess)                                                              \
			throw ::fsb::error(FSB_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r_) + \
			                                     " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
	} while (0)

#define FSB_REQUIRE(cond, msg)                                    \
	do {                                                          \
		if (!(cond))                                              \
			throw ::fsb::error(FSB_ERR_ARG, std::string(msg));   \
	} while (0)

constexpr int SM_COUNT = 148; // B200
constexpr int EW_THREADS = 256;
constexpr int EW_CTAS_PER_SM = 8;
constexpr int MAX_RED_BLOCKS = 4096; // partial slots per reduction output

// One queued statement (vector operands are handles; scalars are values).
struct pending {
	enum kind_t { EW, RED, SPMV } kind;
	int op; // ew_op for EW/RED
	fsb_vec_s * z = nullptr;
	fsb_vec_s * x = nullptr;
	fsb_vec_s * y = nullptr;
	double a = 0, b = 0;
	int64_t token = -1; // RED
	fsb_parcsr_s * A = nullptr; // SPMV
	// device-resident coefficients: a (b) is multiplied by scalars[a_num] / scalars[a_den] when a_num >= 0
	int a_num = -1, a_den = -1, b_num = -1, b_den = -1;
	// RED: also store the all-rank value into this scalar slot; raise the halt flag when the test passes
	int store = -1;
	int halt_mode = 0; // 0 none, 1 sqrt(value) < halt_thr, 2 value < halt_thr
	double halt_thr = 0;
	int n_post = 0; // RED: scalar statements evaluated once the value is stored (fsb_red_opts::post)
	fsb_scalar_op post[FSB_MAX_POST_OPS] = {};
};

// interior sub-box of a padded array (structured-grid vectors); colexicographic dof order, x fastest
struct box_shape {
	int64_t ext[3] = {1, 1, 1}, lo[3] = {0, 0, 0}, n[3] = {1, 1, 1};
	bool operator==(const box_shape & o) const {
		for (int k = 0; k < 3; ++k)
			if (ext[k] != o.ext[k] || lo[k] != o.lo[k] || n[k] != o.n[k])
				return false;
		return true;
	}
	int64_t dofs() const { return n[0] * n[1] * n[2]; }
	int64_t storage() const { return ext[0] * ext[1] * ext[2]; }
	int64_t origin() const { return lo[0] + ext[0] * (lo[1] + ext[1] * lo[2]); }
};

// host view of a flagged 16-byte word (ew_kernels.cuh: ll_store): two 8-byte halves {32 data bits, 32 flag bits}
struct ll_word {
	volatile uint64_t half[2];
};
inline void host_ll_store(ll_word * w, double v, uint32_t flag) {
	uint64_t bits;
	static_assert(sizeof(bits) == sizeof(v), "double is 64 bits");
	__builtin_memcpy(&bits, &v, 8);
	w->half[0] = (bits & 0xffffffffull) | (static_cast<uint64_t>(flag) << 32);
	w->half[1] = (bits >> 32) | (static_cast<uint64_t>(flag) << 32);
}
inline bool host_ll_load(const ll_word * w, uint32_t flag, double * v) {
	const uint64_t a = w->half[0], b = w->half[1];
	if (static_cast<uint32_t>(a >> 32) != flag || static_cast<uint32_t>(b >> 32) != flag)
		return false;
	const uint64_t bits = (a & 0xffffffffull) | (b << 32);
	__builtin_memcpy(v, &bits, 8);
	return true;
}

} // namespace fsb

struct fsb_ctx_s {
	int device = 0, rank = 0, nranks = 1;
	cudaStream_t stream = nullptr;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
	std::shared_ptr<fsb::setup_exchange> boot; // host-side exchange among the ranks during setup (setup_exchange.h)
	ncclComm_t nccl = nullptr; // reductions + setup, used on `stream`; null when the ranks are threads of one process
	ncclComm_t nccl_halo_comm = nullptr; // ghost exchange, used on `comm_stream`
	ncclComm_t nccl_halo() const { return nccl_halo_comm; }

	// reductions
	double * d_partials = nullptr; // [MAX_RED][MAX_RED_BLOCKS]
	unsigned * d_counter = nullptr;
	unsigned * d_sched = nullptr; // SpMV dynamic scheduler words {next block, finished CTAs}
	double * d_results = nullptr; // [FSB_RED_RING]
	double * h_results = nullptr; // pinned [FSB_RED_RING]: NCCL transport (device-to-host copy + event)
	fsb::ll_word * h_ll = nullptr; // pinned+mapped [FSB_RED_RING] flagged words the reduction kernels publish into
	fsb::ll_word * h_ll_dev = nullptr; // device alias of h_ll
	int64_t next_token = 1;
	// cross-rank reductions over peer memory (nranks > 1, when cudaIpc mapping succeeded)
	double * d_mailbox = nullptr; // this rank's mailbox [ring][nranks]{value, token}
	void * peer_mailbox[8] = {};
	fsb::xrank_info * d_xrank = nullptr; // device copy handed to kernels; nullptr => NCCL path
	int * h_xrank_error = nullptr; // mapped host flag
	// device scalars + halt flag (device-scalar solver variants)
	double * d_scalars = nullptr; // [MAX_SCALARS], [0] == 1.0
	bool scalar_used[64] = {};
	int * d_halt = nullptr;
	bool halt_armed = false;
	std::vector<int> token_op; // per ring slot: nccl op kind of the reduction
	std::vector<cudaEvent_t> token_event; // multi-rank path

	// deferred statements
	std::vector<fsb::pending> queue;
	bool fusion = true;
	bool trace = false;
	bool reproducible = true; // SpMV row blocks statically assigned to CTAs (set from nranks at creation)
	bool jit = true; // statement groups without a compiled instantiation: compile one at run time (jit.cu); FSB_JIT=0 turns it off
	int spmv_rows_per_cta = 0; // 0 = auto
	int spmv_threads = 0;
	bool spmv_dictionary = true; // FSB_OPT_SPMV_DICTIONARY

	cudaEvent_t timers[16] = {};
	bool profile = false;
	std::vector<cudaEvent_t> prof_events; // pairs: [2k] start, [2k+1] stop
	std::vector<char> prof_tag; // per pair: 0 diag block, 1 offd block
	size_t prof_used = 0;

	// device-side timeline (FSB_OPT_TIMELINE): one slot of 8 x 64-bit words per launch, see ew_kernels.cuh
	unsigned long long * d_timeline = nullptr;
	int64_t timeline_cap = 0, timeline_used = 0;
	std::vector<int> timeline_kind; // per used slot

	// L2 flush scratch
	void * d_flush = nullptr;
	size_t flush_bytes = 0;

	int64_t stats[16] = {};
	uint64_t next_vec_id = 1;
	uint64_t next_mat_id = 1;
	std::vector<void *> deferred_free; // in-process rank groups: vector storage released with the context
	bool speculate = false; // FSB_OPT_SPECULATE
	fsb::speculation * spec = nullptr; // created on first use
};

struct fsb_vec_s {
	fsb_ctx_s * ctx = nullptr;
	double * d = nullptr;
	int64_t n_owned = 0, n_ghost = 0;
	bool owns = true;
	bool halo_valid = false; // ghost entries hold the owners' current values ...
	uint64_t halo_for = 0; // ... in the ghost numbering of the matrix with this id
	uint64_t id = 0;
	bool padded = false; // the allocation extends >= 2 doubles past n_owned + n_ghost (bulk copies may over-read one entry)
	bool box = false; // structured-grid vector: n_owned dofs inside `shape`, n_owned + n_ghost == storage
	fsb::box_shape shape;
};

namespace fsb {

// one CSR block (diag or offd) in device memory
struct csr_block {
	int64_t n_rows = 0; // rows described by rowptr
	int64_t nnz = 0;
	bool wide = false; // 64-bit row offsets
	void * rowptr = nullptr; // int32 or int64 [n_rows+1]
	int32_t * col = nullptr; // [nnz padded]
	double * val = nullptr; // [nnz padded]
	// row-block work descriptors for the streaming kernel
	int32_t * blk_row = nullptr; // [n_blk+1] first row of each CTA's row block
	void * blk_desc = nullptr; // [n_blk] per-block descriptors (spmv.cu)
	int n_blk = 0;
	int max_blk_nnz = 0; // max nnz staged by one CTA (smem sizing)
	int max_blk_rows = 0;
	int32_t * row_ids = nullptr; // compressed row list (offd): row index per compressed row, or null
	bool has_giant_rows = false; // some row exceeds a pipeline stage (streamed from global memory by a whole CTA)
	bool has_offd_map = false; // descriptors carry the rows of the off-process block (fused ghost exchange possible)
	// window format (spmv.cu: spmv_window_kernel), null when the block is in the gather format only
	uint16_t * lcol = nullptr; // [nnz] position of x[col] inside the row block's staged x segments (+ diagonal mark)
	uint16_t * rp16 = nullptr; // [n_rows + n_blk] row offsets relative to the block's first nonzero
	void * segs = nullptr; // [n_blk][16] x segments per row block
	int win_xcap = 0; // max x entries staged by one row block
	// value dictionary (window format only): at most 256 distinct values, one byte per nonzero
	// (rows padded to multiples of 8 slots, row blocks to multiples of 16: spmv.cu, build_value_dictionary)
	uint8_t * vidx = nullptr; // [dict_slots] index of the slot's value in vdict
	uint16_t * plcol = nullptr; // [dict_slots] its local column (lcol in the padded order)
	uint16_t * pmeta = nullptr; // [3 n_rows + n_blk] per row block: slot offsets, row lengths, diagonal positions
	std::vector<unsigned long long> dict_keys; // bit patterns of the distinct values (<= 256), ascending; empty: no dictionary
	double * vdict = nullptr; // [256] the distinct values, ascending bit patterns
	int n_dict = 0;
	int max_blk_pnnz = 0; // slots of the largest row block
	int64_t dict_slots = 0;
};

struct neighbour {
	int rank;
	int64_t send_count, recv_count;
	int64_t recv_offset; // offset into ghost region
	int64_t send_offset; // offset into packed send buffer
	int64_t contiguous_start; // >= 0: send list is the range [start, start+count) of x (no pack)
};

} // namespace fsb

struct fsb_parcsr_s {
	fsb_ctx_s * ctx = nullptr;
	uint64_t id = 0; // unique per context: vectors remember whose ghost numbering they hold (a raw pointer could be reused)
	int64_t n_global = 0, n_local = 0, n_ghost = 0, row_begin = 0;
	fsb::csr_block diag, offd;
	int64_t * d_colmap = nullptr;
	std::vector<int64_t> colmap; // host copy
	std::vector<int64_t> row_part;
	std::vector<fsb::neighbour> nbrs;
	int32_t * d_send_idx = nullptr; // packed send lists
	double * d_send_buf = nullptr;
	int64_t send_total = 0;
	bool need_pack = false;
	double * d_dinv = nullptr; // lazily computed 1/diag (jacobi relax)
	// ghost exchange over peer memory (halo.cu); null => NCCL send/recv
	void * halo_p2p = nullptr; // device halo_dev
	unsigned char * halo_block = nullptr; // flags + double-buffered landing area (IPC-exported)
	unsigned * halo_counters = nullptr;
	std::vector<void *> halo_opened; // peers' blocks mapped here
	long long halo_epoch = 0;
	int halo_push_ctas = 1, halo_unpack_ctas = 1;
	// structured-grid operator: rows = dofs of `shape`, columns and row_ids are storage offsets into the
	// padded array, so boundary layers are read like any other entry (one rank only)
	bool box = false;
	fsb::box_shape shape;
	// box operator over several ranks (slabs along the last axis): where the neighbours' planes go inside the padded array
	long long box_split = -1, box_off[2] = {0, 0};
};

namespace fsb {

// timeline slot for the next launch (nullptr when the timeline is off or full); kinds below
enum { TL_KIND_SPMV = 1, TL_KIND_SPMV_OFFD = 2, TL_KIND_SPMV_FUSED = 3, TL_KIND_EW = 16, TL_KIND_HALO_PUSH = 32, TL_KIND_HALO_UNPACK = 33 };
unsigned long long * timeline_slot(fsb_ctx_s * c, int kind);

// Launch with programmatic stream serialisation: the kernel may be scheduled while its predecessor in the stream
// drains; it calls grid_dependency_wait() before it touches anything the predecessor wrote (FSB_PDL=0 turns it off).
bool pdl_enabled();
template<class Kernel, class Args>
void launch_dependent(Kernel kern, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Args & args) {
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl_enabled() ? 1 : 0;
	FSB_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
}

// CUDA loads a kernel at its first use, and the load waits for everything running in the context.  The ranks of an
// in-process group call this before they meet for a launch whose kernels wait for each other (setup_exchange::rendezvous).
template<class Kernel>
void preload_kernel(Kernel kern) {
	cudaFuncAttributes fa;
	FSB_CUDA(cudaFuncGetAttributes(&fa, kern));
}

// queue / fuser (fuser.cu)
void enqueue(fsb_ctx_s * c, const pending & p);
void flush(fsb_ctx_s * c, bool waiting = false); // waiting: called by a host that is about to wait for a reduction
void speculate_after_flush(fsb_ctx_s * c); // wait_token: launch the expected next group ahead of the result (FSB_OPT_SPECULATE)
void speculation_release(fsb_ctx_s * c); // tell an armed kernel to leave (any path that blocks on the device without a flush)
void speculation_destroy(fsb_ctx_s * c);
bool speculation_failed(const fsb_ctx_s * c); // an armed kernel waited ~10 s for the host and gave up
int64_t new_token(fsb_ctx_s * c, int nccl_op);

// kernels (declared here, defined in their .cu)
// one SpMV launch: y (+)= B x with optional fused epilogues
struct spmv_call {
	const double * x = nullptr;
	double * y = nullptr;
	bool x_padded = false; // x may be over-read by one entry (window format needs it for an odd number of columns)
	bool accumulate = false; // y += B x (off-process block as its own launch)
	bool acc_continue = false; // accumulate: the row's running sum continues from y (Jacobi) instead of being added at the end
	const double * dot_u = nullptr; // CTA b writes its partial of sum y_i u_i to partials[partial_offset + b]
	double * partials = nullptr;
	int partial_offset = 0;
	const pending * fold = nullptr; // the queued dot whose partials the last CTA of this launch folds and publishes
	const fsb_parcsr_s * halo = nullptr; // fused ghost exchange over peer memory: push, interior, boundary rows, acknowledge
	long long epoch = 0;
	int jacobi = 0; // 0 product; 1 weighted-Jacobi sweep (x = old iterate, y = new); 2 product without the diagonal entries
	const double * jacobi_b = nullptr;
	double omega = 0;
};
int launch_spmv(fsb_ctx_s * c, const csr_block & B, const spmv_call & call, cudaStream_t s);
void fill_red_out(fsb_ctx_s * c, const pending & red, red_out & r);
bool program_is_registered(const program & p, bool dev);
// run-time compiled program kernels (jit.cu)
struct ew_args;
bool jit_compile(const program & p, bool dev, bool box, std::vector<char> & cubin, std::string & log);
const void * jit_kernel(const program & p, bool dev, bool box, int * resident);
void jit_launch(const void * kernel, int resident, const ew_args & a, int want, cudaStream_t s);
void finish_reduction_nccl(fsb_ctx_s * c, const pending & red);
void finalize_reduction(fsb_ctx_s * c, int n_partials, const pending & red);
void halo_exchange(fsb_parcsr_s * A, fsb_vec_s * x);

void build_blocks(fsb_ctx_s * c, csr_block & B, const std::vector<int64_t> * host_rowptr);
void build_window_format(fsb_ctx_s * c, csr_block & B, int64_t n_cols);
void probe_value_dictionary(fsb_ctx_s * c, csr_block & B); // before build_blocks
void build_value_dictionary(fsb_ctx_s * c, csr_block & B); // after build_window_format
void attach_offd_rows(fsb_ctx_s * c, csr_block & D, const csr_block & O);
void extract_dinv(fsb_parcsr_s * A, double * d);
void halo_p2p_setup(fsb_parcsr_s * A, const std::vector<int64_t> & dest_off);
void halo_p2p_destroy(fsb_parcsr_s * A);
void halo_p2p_push(fsb_parcsr_s * A, fsb_vec_s * x);
void halo_p2p_unpack(fsb_parcsr_s * A, fsb_vec_s * x);
void spmv_group(fsb_ctx_s * c, const pending & sp, const pending * dot);

} // namespace fsb
