// CSR SpMV for sm_100a: TMA-staged streaming kernel.
//
// Reference semantics (matrices/seq.hh:178-194): y[i] = 0; for off in row i, in index
// order: y[i] += val[off] * x[col[off]].  This kernel keeps exactly that accumulation
// order per row (thread-per-row path), so 1-rank results are bit-identical to the
// reference loop compiled without FMA contraction.
//
// Design (HBM-bound, 12 B/nnz of matrix stream against 8 B/row of output):
//  * The matrix arrays (val fp64, col int32) of a block of consecutive rows are
//    contiguous in CSR, so one elected thread pulls them into shared memory with two
//    1-D bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx) -- fully coalesced
//    128-byte HBM bursts regardless of row length, no per-thread load instructions.
//  * NSTAGE-deep ring of shared-memory stages per CTA: the copy of row block i+NSTAGE-1
//    is in flight while block i is being multiplied; CTAs are persistent and walk row
//    blocks with a grid stride, so the number of reduction partials is bounded.
//  * Rows are assigned thread-per-row with consecutive lanes on consecutive rows:
//    stencil/banded gathers x[col] then touch 2-3 L1 lines per warp instruction, and the
//    strided shared-memory reads of val/col are conflict-free for odd row lengths.
//    x itself is served by L1/L2 (reuse distance of a 3-D stencil is two grid planes).
//  * Row blocks with few, long rows switch to warp-per-row; a row longer than a stage
//    is streamed straight from global memory by the whole CTA.
//  * Optional fused epilogue: per-CTA partial of sum_i y_i * u_i (CG's p.Ap, cg.hh:98)
//    and accumulate mode y += A x for the off-process block (parcsr.hh:61-68).
#include <cub/cub.cuh>

#include "ew_kernels.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int SPMV_THREADS = 256;

struct spmv_args {
	const void * rowptr;
	const int32_t * col;
	const double * val;
	const int32_t * blk_row; // [n_blk + 1]
	const int32_t * row_ids; // compressed-row list or nullptr
	const double * x;
	double * y;
	const double * u; // dot operand or nullptr; u == y means sum y_i^2
	double * partials; // one per CTA
	int n_blk;
	int cap; // nnz capacity of one stage (multiple of 4, includes alignment slack)
};

__device__ __forceinline__ uint32_t smem_u32(const void * p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t * bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					 smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

template<class OffT, int NSTAGE, bool ACC, bool DOT, bool ROWLIST>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_stream_kernel(const __grid_constant__ spmv_args a) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t bar[NSTAGE];
	__shared__ double scratch[32];
	// stage layout: NSTAGE x [cap doubles] then NSTAGE x [cap int32]
	double * s_val = reinterpret_cast<double *>(smem);
	int32_t * s_col = reinterpret_cast<int32_t *>(smem + static_cast<size_t>(NSTAGE) * a.cap * sizeof(double));
	const OffT * __restrict__ rowptr = static_cast<const OffT *>(a.rowptr);
	const double * __restrict__ x = a.x;
	const int tid = threadIdx.x;

	if (tid == 0) {
		for (int s = 0; s < NSTAGE; ++s)
			mbar_init(&bar[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	// issue the bulk copies of row block `blk` into stage `s`; returns nothing (thread 0 only)
	auto issue = [&](int blk, int s) {
		const int r0 = a.blk_row[blk], r1 = a.blk_row[blk + 1];
		const long long z0 = static_cast<long long>(rowptr[r0]), z1 = static_cast<long long>(rowptr[r1]);
		const long long za = z0 & ~3LL; // 16-byte aligned start for the int32 stream
		const long long cnt = ((z1 + 3) & ~3LL) - za;
		if (cnt > 0 && cnt <= a.cap) {
			mbar_expect_tx(&bar[s], static_cast<uint32_t>(cnt * 12));
			tma_load_1d(s_val + static_cast<size_t>(s) * a.cap, a.val + za, static_cast<uint32_t>(cnt * 8), &bar[s]);
			tma_load_1d(s_col + static_cast<size_t>(s) * a.cap, a.col + za, static_cast<uint32_t>(cnt * 4), &bar[s]);
		}
		else {
			mbar_expect_tx(&bar[s], 0); // empty or oversize block: nothing staged, complete the phase
		}
	};

	const int first = blockIdx.x, step = gridDim.x;
	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < NSTAGE - 1; ++s) {
			const int blk = first + s * step;
			if (blk < a.n_blk)
				issue(blk, s);
		}
	}

	double dot_acc = 0.0;
	// write one row result; the dot partial of an accumulate pass only counts what this pass added
	auto emit = [&](int r, double sum) {
		const int yr = ROWLIST ? a.row_ids[r] : r;
		double out = sum;
		if constexpr (ACC) {
			const double old = a.y[yr];
			out = __dadd_rn(old, sum);
			if constexpr (DOT)
				dot_acc += (a.u == a.y) ? (out * out - old * old) : sum * a.u[yr];
		}
		else if constexpr (DOT) {
			dot_acc = fma(out, a.u == a.y ? out : a.u[yr], dot_acc);
		}
		a.y[yr] = out;
	};
	int it = 0;
	for (int blk = first; blk < a.n_blk; blk += step, ++it) {
		const int s = it % NSTAGE;
		const uint32_t parity = (it / NSTAGE) & 1;
		if (tid == 0) { // prefetch NSTAGE-1 blocks ahead into the stage freed last iteration
			const int pre = blk + (NSTAGE - 1) * step;
			if (pre < a.n_blk)
				issue(pre, (it + NSTAGE - 1) % NSTAGE);
		}
		const int r0 = a.blk_row[blk], r1 = a.blk_row[blk + 1];
		const long long z0 = static_cast<long long>(rowptr[r0]), z1 = static_cast<long long>(rowptr[r1]);
		const long long za = z0 & ~3LL;
		const long long cnt = ((z1 + 3) & ~3LL) - za;
		const int nrows = r1 - r0;
		mbar_wait(&bar[s], parity);
		const double * sv = s_val + static_cast<size_t>(s) * a.cap;
		const int32_t * sc = s_col + static_cast<size_t>(s) * a.cap;

		if (cnt > a.cap) {
			// a single row longer than a stage: CTA-wide strided pass over global memory
			for (int r = r0; r < r1; ++r) {
				const long long p0 = static_cast<long long>(rowptr[r]), p1 = static_cast<long long>(rowptr[r + 1]);
				double sum = 0.0;
				for (long long p = p0 + tid; p < p1; p += SPMV_THREADS)
					sum = fma(a.val[p], __ldg(&x[a.col[p]]), sum);
				sum = block_fold<0>(sum, scratch);
				if (tid == 0)
					emit(r, sum);
			}
		}
		else if (nrows * 8 >= SPMV_THREADS || nrows >= cnt) {
			// thread per row, consecutive lanes on consecutive rows; in-order accumulation
			for (int r = r0 + tid; r < r1; r += SPMV_THREADS) {
				const int p0 = static_cast<int>(static_cast<long long>(rowptr[r]) - za);
				const int p1 = static_cast<int>(static_cast<long long>(rowptr[r + 1]) - za);
				double sum = 0.0;
				for (int p = p0; p < p1; ++p)
					sum = __dadd_rn(sum, __dmul_rn(sv[p], __ldg(&x[sc[p]])));
				emit(r, sum);
			}
		}
		else {
			// few long rows: warp per row, lanes stride the staged row, shuffle tree
			const int lane = tid & 31, warp = tid >> 5;
			for (int r = r0 + warp; r < r1; r += SPMV_THREADS / 32) {
				const int p0 = static_cast<int>(static_cast<long long>(rowptr[r]) - za);
				const int p1 = static_cast<int>(static_cast<long long>(rowptr[r + 1]) - za);
				double sum = 0.0;
				for (int p = p0 + lane; p < p1; p += 32)
					sum = fma(sv[p], __ldg(&x[sc[p]]), sum);
				sum = warp_fold<0>(sum);
				if (lane == 0)
					emit(r, sum);
			}
		}
		__syncthreads(); // stage s may be overwritten by the prefetch issued next iteration
	}

	if constexpr (DOT) {
		dot_acc = block_fold<0>(dot_acc, scratch);
		if (tid == 0)
			a.partials[blockIdx.x] = dot_acc;
	}
}

// one CTA: fold `n` partials in fixed order and publish like the element-wise epilogue
__global__ void __launch_bounds__(256) fold_partials_kernel(const double * partials, int n, red_out out) {
	__shared__ double scratch[32];
	double t = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		t += partials[i];
	t = block_fold<0>(t, scratch);
	if (threadIdx.x == 0) {
		*out.d_value = t;
		if (out.h_value) {
			*reinterpret_cast<volatile double *>(out.h_value) = t;
			__threadfence_system();
			*reinterpret_cast<volatile long long *>(out.h_flag) = out.token;
		}
	}
}

struct spmv_config {
	int nstage;
	int cap;
	size_t smem;
	int grid;
};

static spmv_config configure(const csr_block & B) {
	spmv_config k;
	k.cap = ((B.max_blk_nnz + 8 + 3) / 4) * 4; // + alignment slack on both ends
	if (k.cap < 64)
		k.cap = 64;
	const size_t stage_bytes = static_cast<size_t>(k.cap) * 12;
	// aim for ~4 stages resident per SM in total: 2 CTAs x 2 stages when they fit
	k.nstage = 2;
	if (stage_bytes * 2 > 110 * 1024)
		k.nstage = 1;
	k.smem = stage_bytes * k.nstage;
	int ctas_per_sm = static_cast<int>((220 * 1024) / (k.smem + 1024));
	if (ctas_per_sm < 1)
		ctas_per_sm = 1;
	if (ctas_per_sm > 4)
		ctas_per_sm = 4;
	k.grid = std::min(B.n_blk, SM_COUNT * ctas_per_sm);
	if (k.grid < 1)
		k.grid = 1;
	return k;
}

int spmv_partial_count(const csr_block & B) { return B.n_blk > 0 ? configure(B).grid : 0; }

template<class OffT, int NSTAGE, bool ACC, bool DOT, bool ROWLIST>
static void launch_variant(const spmv_args & a, const spmv_config & k, cudaStream_t s) {
	auto kern = spmv_stream_kernel<OffT, NSTAGE, ACC, DOT, ROWLIST>;
	static bool attr_set = false;
	if (!attr_set) {
		FSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
		attr_set = true;
	}
	kern<<<k.grid, SPMV_THREADS, k.smem, s>>>(a);
	FSB_CUDA(cudaGetLastError());
}

template<class OffT, int NSTAGE>
static void launch_flags(const spmv_args & a, const spmv_config & k, bool acc, bool dot, bool rowlist, cudaStream_t s) {
	if (rowlist) { // off-process block: always accumulates
		if (dot)
			launch_variant<OffT, NSTAGE, true, true, true>(a, k, s);
		else
			launch_variant<OffT, NSTAGE, true, false, true>(a, k, s);
	}
	else if (acc) {
		if (dot)
			launch_variant<OffT, NSTAGE, true, true, false>(a, k, s);
		else
			launch_variant<OffT, NSTAGE, true, false, false>(a, k, s);
	}
	else {
		if (dot)
			launch_variant<OffT, NSTAGE, false, true, false>(a, k, s);
		else
			launch_variant<OffT, NSTAGE, false, false, false>(a, k, s);
	}
}

// y (+)= B x on stream s.  When dot_u != nullptr, CTA b writes its partial of sum y_i u_i
// to d_partials[partial_offset + b].
void launch_spmv(fsb_ctx_s * c, const csr_block & B, const double * x, double * y, bool accumulate,
                 const double * dot_u, double * d_partials, int partial_offset, cudaStream_t s) {
	if (B.n_blk == 0)
		return;
	const spmv_config k = configure(B);
	FSB_REQUIRE(k.smem <= 225 * 1024, "spmv: row block does not fit shared memory");
	spmv_args a{};
	a.rowptr = B.rowptr;
	a.col = B.col;
	a.val = B.val;
	a.blk_row = B.blk_row;
	a.row_ids = B.row_ids;
	a.x = x;
	a.y = y;
	a.u = dot_u;
	a.partials = d_partials ? d_partials + partial_offset : nullptr;
	a.n_blk = B.n_blk;
	a.cap = k.cap;
	const bool rowlist = B.row_ids != nullptr;
	const bool dot = dot_u != nullptr;
	if (B.wide) {
		if (k.nstage == 2)
			launch_flags<long long, 2>(a, k, accumulate, dot, rowlist, s);
		else
			launch_flags<long long, 1>(a, k, accumulate, dot, rowlist, s);
	}
	else {
		if (k.nstage == 2)
			launch_flags<int, 2>(a, k, accumulate, dot, rowlist, s);
		else
			launch_flags<int, 1>(a, k, accumulate, dot, rowlist, s);
	}
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
}

// fold the SpMV partials [0, n) into the reduction slot of `token`
void finalize_reduction(fsb_ctx_s * c, int n_partials, int64_t token, int /*op_kind*/) {
	const int slot = static_cast<int>(token % FSB_RED_RING);
	red_out r{};
	r.d_value = c->d_results + slot;
	r.token = token;
	if (c->nranks == 1) {
		r.h_value = c->h_results_dev + slot;
		r.h_flag = reinterpret_cast<long long *>(c->h_flags_dev + slot);
	}
	fold_partials_kernel<<<1, 256, 0, c->stream>>>(c->d_partials, n_partials, r);
	FSB_CUDA(cudaGetLastError());
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
}

// Row-block work descriptors.  host_rowptr == nullptr: uniform blocks of `rows_per_blk` rows
// (caller guarantees they fit); otherwise greedy packing by nnz.
void build_blocks(fsb_ctx_s * c, csr_block & B, const std::vector<int64_t> * host_rowptr) {
	constexpr int CAP = 4096 - 8; // nnz staged per row block (48 KB per stage)
	constexpr int ROWS_MAX = 512;
	std::vector<int32_t> blk;
	int max_nnz = 0, max_rows = 0;
	if (B.n_rows == 0) {
		B.n_blk = 0;
		return;
	}
	if (host_rowptr) {
		const std::vector<int64_t> & rp = *host_rowptr;
		int64_t r = 0;
		while (r < B.n_rows) {
			const int64_t start = r;
			int64_t nnz = 0;
			while (r < B.n_rows && r - start < ROWS_MAX && nnz + (rp[r + 1] - rp[r]) <= CAP) {
				nnz += rp[r + 1] - rp[r];
				++r;
			}
			if (r == start) { // one row longer than a stage: its own block, streamed from global
				nnz = 0; // nothing staged
				++r;
			}
			blk.push_back(static_cast<int32_t>(start));
			max_nnz = std::max<int>(max_nnz, static_cast<int>(nnz));
			max_rows = std::max<int>(max_rows, static_cast<int>(r - start));
		}
		blk.push_back(static_cast<int32_t>(B.n_rows));
	}
	else {
		// uniform: caller stored the row width bound in max_blk_nnz (per row).  Aim for
		// thread-per-row with every thread busy (multiples of 256 rows) within <= 96 KB a stage.
		const int width = std::max(1, B.max_blk_nnz);
		int rows = (4096 / width) / SPMV_THREADS * SPMV_THREADS;
		if (rows < SPMV_THREADS)
			rows = SPMV_THREADS;
		if (rows * width > 8192)
			rows = std::max(1, 8192 / width);
		if (c->spmv_rows_per_cta > 0)
			rows = c->spmv_rows_per_cta;
		if (rows > 32)
			rows = rows / 32 * 32;
		for (int64_t r = 0; r < B.n_rows; r += rows)
			blk.push_back(static_cast<int32_t>(r));
		blk.push_back(static_cast<int32_t>(B.n_rows));
		max_nnz = rows * width;
		max_rows = rows;
	}
	B.n_blk = static_cast<int>(blk.size()) - 1;
	B.max_blk_nnz = max_nnz;
	B.max_blk_rows = max_rows;
	FSB_CUDA(cudaMalloc(&B.blk_row, blk.size() * sizeof(int32_t)));
	FSB_CUDA(cudaMemcpyAsync(B.blk_row, blk.data(), blk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	FSB_CUDA(cudaStreamSynchronize(c->stream));
}

} // namespace fsb
