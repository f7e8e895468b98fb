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
//  * NSTAGE-deep ring of shared-memory stages per CTA: the copies of row blocks
//    i+1 .. i+NSTAGE-1 are in flight while block i is being multiplied; CTAs are persistent
//    and walk row blocks with a grid stride, so the number of reduction partials is bounded.
//  * Rows are assigned thread-per-row with consecutive lanes on consecutive rows:
//    stencil/banded gathers x[col] then touch 2-3 L1 lines per warp instruction, and the
//    strided shared-memory reads of val/col are conflict-free for odd row lengths.
//    Gathers are issued in groups of GATHER independent loads before the in-order sum.
//    x itself is served by L1/L2 (reuse distance of a 3-D stencil is two grid planes).
//  * Row blocks with few, long rows switch to warp-per-row; a row longer than a stage
//    is streamed straight from global memory by the whole CTA.
//  * Optional fused epilogue: per-CTA partial of sum_i y_i * u_i (CG's p.Ap, cg.hh:98)
//    and accumulate mode y += A x for the off-process block (parcsr.hh:61-68).
#include <cstdlib>

#include "ew_kernels.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int SPMV_MAX_THREADS = 544; // 512 consumer threads + one producer warp

// one row block = the unit of work of a pipeline stage
struct blk_desc {
	long long z0; // first nonzero
	int r0; // first row
	int nrows;
	int nnz;
	int pad;
};

struct spmv_args {
	const void * rowptr;
	const int32_t * col;
	const double * val;
	const blk_desc * desc; // [n_blk]
	const int32_t * row_ids; // compressed-row list or nullptr
	const double * x;
	double * y;
	const double * u; // dot operand or nullptr; u == y means sum y_i^2
	double * partials; // one per CTA (this launch writes [fold_extra, fold_extra + gridDim.x))
	unsigned * sched; // [0] next unclaimed row block (dynamic scheduling), [1] finished CTAs (last-CTA election)
	red_out result; // where the folded dot goes (token == 0: do not fold in this launch)
	const xrank_info * xr; // cross-rank all-reduce over peer memory, or nullptr
	int fold_extra; // partials written by an earlier launch that take part in the fold
	int n_blk;
	int cap; // nnz capacity of one stage (multiple of 4, includes alignment slack)
	int rcap; // row-offset capacity of one stage
	int chunk; // row blocks claimed per atomic (1..4)
	int static_sched; // 1: claims are handed out round-robin by CTA index (bitwise reproducible dot partials)
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
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// barrier among the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync(int nconsumers) {
	asm volatile("bar.sync 1, %0;" ::"r"(nconsumers) : "memory");
}

// Warp-specialised: warp 0 is the producer (one lane issues descriptor loads and bulk copies and
// runs up to NSTAGE row blocks ahead, throttled by the `empty` barriers); all other warps are
// consumers (wait `full`, multiply thread-per-row out of shared memory, release the stage).
// GATHER: x[col] loads in flight per thread before the in-order sum (8 for short rows; 32 covers a
// whole 27-point row in one round trip to L2 at the price of registers)
template<class OffT, int NSTAGE, bool ACC, bool DOT, bool ROWLIST, int GATHER>
__global__ void __launch_bounds__(GATHER > 8 ? 288 : SPMV_MAX_THREADS, 2) spmv_stream_kernel(const __grid_constant__ spmv_args a) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[NSTAGE], empty[NSTAGE];
	__shared__ blk_desc sdesc[NSTAGE];
	__shared__ double scratch[32];
	constexpr int E = 16 / sizeof(OffT); // row offsets per 16 bytes
	const size_t stage_bytes = static_cast<size_t>(a.cap) * 12 + static_cast<size_t>(a.rcap) * sizeof(OffT);
	const OffT * __restrict__ rowptr = static_cast<const OffT *>(a.rowptr);
	const double * __restrict__ x = a.x;
	const int tid = threadIdx.x;
	const int nconsumers = blockDim.x - 32;

	if (tid == 0) {
		for (int s = 0; s < NSTAGE; ++s) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], nconsumers / 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	double dot_acc = 0.0;

	if (tid < 32) {
		// ------------------------------------------------ producer
		// Row blocks are claimed dynamically, CHUNK (<= 4) consecutive blocks per atomic: a CTA that starts
		// late (e.g. its SM was busy with an NCCL kernel) simply takes fewer blocks, so the launch
		// never degrades into a second wave, and blocks are still handed out in row order (x stays
		// in L2).  Lane l holds the descriptor of block base + l; the next claim and its descriptor
		// loads are issued before the current chunk is staged, so neither the atomic nor the DRAM
		// latency of the descriptors sits between a freed stage and the next bulk copy.
		const int CHUNK = a.chunk; // 1..4, chosen on the host so that every CTA sees several claims
		const int lane = tid;
		int round = 0;
		auto claim = [&]() {
			if (a.static_sched) // fixed row blocks per CTA: the per-CTA dot partials do not depend on timing
				return static_cast<int>(min(static_cast<long long>(a.n_blk),
				                            (static_cast<long long>(round++) * gridDim.x + blockIdx.x) * CHUNK));
			int base = 0;
			if (lane == 0)
				base = static_cast<int>(atomicAdd(&a.sched[0], static_cast<unsigned>(CHUNK)));
			return __shfl_sync(0xffffffffu, base, 0);
		};
		auto fetch = [&](int base) {
			blk_desc d{};
			if (lane < CHUNK && base + lane < a.n_blk)
				d = a.desc[base + lane];
			return d;
		};
		int base = claim();
		blk_desc cur = fetch(base);
		int it = 0;
		while (base < a.n_blk) {
			const int nbase = claim();
			const blk_desc nxt = fetch(nbase);
			const int jmax = min(CHUNK, a.n_blk - base);
			for (int j = 0; j < jmax; ++j, ++it) {
				blk_desc d;
				d.z0 = __shfl_sync(0xffffffffu, cur.z0, j);
				d.r0 = __shfl_sync(0xffffffffu, cur.r0, j);
				d.nrows = __shfl_sync(0xffffffffu, cur.nrows, j);
				d.nnz = __shfl_sync(0xffffffffu, cur.nnz, j);
				d.pad = 0;
				if (lane == 0) {
					const int s = it % NSTAGE;
					if (it >= NSTAGE)
						mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
					sdesc[s] = d;
					unsigned char * sbase = smem + s * stage_bytes;
					const long long za = d.z0 & ~3LL; // 16-byte aligned start of the int32 stream
					const long long cnt = ((d.z0 + d.nnz + 3) & ~3LL) - za;
					const long long ra = d.r0 & ~static_cast<long long>(E - 1);
					const long long rcnt = ((d.r0 + d.nrows + 1 + E - 1) & ~static_cast<long long>(E - 1)) - ra;
					const bool staged = cnt > 0 && cnt <= a.cap;
					uint32_t bytes = static_cast<uint32_t>(rcnt * sizeof(OffT));
					if (staged)
						bytes += static_cast<uint32_t>(cnt * 12);
					mbar_expect_tx(&full[s], bytes); // also publishes sdesc[s] to the waiters
					if (staged) {
						tma_load_1d(sbase, a.val + za, static_cast<uint32_t>(cnt * 8), &full[s]);
						tma_load_1d(sbase + static_cast<size_t>(a.cap) * 8, a.col + za, static_cast<uint32_t>(cnt * 4),
						            &full[s]);
					}
					tma_load_1d(sbase + static_cast<size_t>(a.cap) * 12, rowptr + ra,
					            static_cast<uint32_t>(rcnt * sizeof(OffT)), &full[s]);
				}
				__syncwarp();
			}
			base = nbase;
			cur = nxt;
		}
		if (lane == 0) { // end-of-work marker for the consumers
			const int s = it % NSTAGE;
			if (it >= NSTAGE)
				mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
			sdesc[s].nrows = -1;
			mbar_arrive(&full[s]);
		}
	}
	else {
		// ------------------------------------------------ consumers
		const int ctid = tid - 32;
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

		for (int it = 0;; ++it) {
			const int s = it % NSTAGE;
			mbar_wait(&full[s], (it / NSTAGE) & 1);
			const blk_desc d = sdesc[s];
			if (d.nrows < 0)
				break;
			const unsigned char * base = smem + s * stage_bytes;
			const double * sv = reinterpret_cast<const double *>(base);
			const int32_t * sc = reinterpret_cast<const int32_t *>(base + static_cast<size_t>(a.cap) * 8);
			const OffT * srp = reinterpret_cast<const OffT *>(base + static_cast<size_t>(a.cap) * 12);
			const long long za = d.z0 & ~3LL;
			const int roff = d.r0 & (E - 1); // srp[roff + i] = rowptr[r0 + i]
			const long long cnt = ((d.z0 + d.nnz + 3) & ~3LL) - za;
			const int r0 = d.r0, nrows = d.nrows;

			if (cnt > a.cap) {
				// a single row longer than a stage: all consumers stream it from global memory
				for (int i = 0; i < nrows; ++i) {
					const long long q0 = static_cast<long long>(srp[roff + i]), q1 = static_cast<long long>(srp[roff + i + 1]);
					double sum = 0.0;
					for (long long p = q0 + ctid; p < q1; p += nconsumers)
						sum = fma(a.val[p], __ldg(&x[a.col[p]]), sum);
					sum = warp_fold<0>(sum);
					consumer_sync(nconsumers);
					if ((ctid & 31) == 0)
						scratch[ctid >> 5] = sum;
					consumer_sync(nconsumers);
					if (ctid == 0) {
						double t = 0.0;
						for (int w = 0; w < nconsumers / 32; ++w)
							t += scratch[w];
						emit(r0 + i, t);
					}
				}
			}
			else if (cnt <= static_cast<long long>(nrows) * 64) {
				// short rows (average <= 64 entries): thread per row, consecutive lanes on consecutive
				// rows; in-order accumulation
				for (int i = ctid; i < nrows; i += nconsumers) {
					const int p0 = static_cast<int>(static_cast<long long>(srp[roff + i]) - za);
					const int p1 = static_cast<int>(static_cast<long long>(srp[roff + i + 1]) - za);
					double sum = 0.0;
					for (int p = p0; p < p1; p += GATHER) {
						double prod[GATHER];
#pragma unroll
						for (int k = 0; k < GATHER; ++k)
							if (p + k < p1)
								prod[k] = __dmul_rn(sv[p + k], __ldg(&x[sc[p + k]]));
#pragma unroll
						for (int k = 0; k < GATHER; ++k)
							if (p + k < p1)
								sum = __dadd_rn(sum, prod[k]);
					}
					emit(r0 + i, sum);
				}
			}
			else {
				// few long rows: warp per row, lanes stride the staged row, shuffle tree
				const int lane = ctid & 31, warp = ctid >> 5;
				for (int i = warp; i < nrows; i += nconsumers / 32) {
					const int q0 = static_cast<int>(static_cast<long long>(srp[roff + i]) - za);
					const int q1 = static_cast<int>(static_cast<long long>(srp[roff + i + 1]) - za);
					double sum = 0.0;
					for (int p = q0 + lane; p < q1; p += 32)
						sum = fma(sv[p], __ldg(&x[sc[p]]), sum);
					sum = warp_fold<0>(sum);
					if (lane == 0)
						emit(r0 + i, sum);
				}
			}
			__syncwarp();
			if ((ctid & 31) == 0)
				mbar_arrive(&empty[s]); // this warp is done reading stage s
		}
	}

	// ---- epilogue: the last CTA to finish folds the dot partials and re-arms the scheduler words
	__shared__ bool is_last;
	if constexpr (DOT)
		dot_acc = block_fold<0>(dot_acc, scratch);
	else
		__syncthreads();
	if (tid == 0) {
		if constexpr (DOT)
			a.partials[a.fold_extra + blockIdx.x] = dot_acc;
		__threadfence();
		is_last = atomicAdd(&a.sched[1], 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (is_last) {
		if constexpr (DOT) {
			if (a.result.token != 0) {
				// fold every partial in a fixed order and publish like the element-wise epilogue
				__threadfence();
				const int total = a.fold_extra + static_cast<int>(gridDim.x);
				double t = 0.0;
				for (int i = tid; i < total; i += blockDim.x)
					t += __ldcg(&a.partials[i]);
				t = block_fold<0>(t, scratch);
				if (a.xr)
					t = xrank_allreduce<0>(a.xr, t, a.result.token, scratch);
				if (tid == 0)
					publish(a.result, t);
			}
		}
		if (tid == 0) {
			a.sched[0] = 0u;
			a.sched[1] = 0u;
		}
	}
}

__global__ void build_desc_kernel(const void * rowptr, bool wide, const int32_t * blk_row, int n_blk, blk_desc * out) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blk)
		return;
	const int r0 = blk_row[b], r1 = blk_row[b + 1];
	long long z0, z1;
	if (wide) {
		z0 = static_cast<const long long *>(rowptr)[r0];
		z1 = static_cast<const long long *>(rowptr)[r1];
	}
	else {
		z0 = static_cast<const int *>(rowptr)[r0];
		z1 = static_cast<const int *>(rowptr)[r1];
	}
	blk_desc d;
	d.z0 = z0;
	d.r0 = r0;
	d.nrows = r1 - r0;
	d.nnz = static_cast<int>(z1 - z0);
	d.pad = 0;
	out[b] = d;
}

// one CTA: fold `n` partials in fixed order and publish like the element-wise epilogue
__global__ void __launch_bounds__(256) fold_partials_kernel(const double * partials, int n, red_out out,
                                                            const xrank_info * xr) {
	__shared__ double scratch[32];
	double t = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		t += partials[i];
	t = block_fold<0>(t, scratch);
	if (xr)
		t = xrank_allreduce<0>(xr, t, out.token, scratch);
	if (threadIdx.x == 0)
		publish(out, t);
}

struct spmv_config {
	int nstage, threads, cap, rcap, grid, gather;
	size_t smem;
};

static int env_int(const char * name, int dflt) {
	const char * v = std::getenv(name);
	return v ? std::atoi(v) : dflt;
}

static spmv_config configure(const fsb_ctx_s * c, const csr_block & B) {
	static const int env_stages = env_int("FSB_SPMV_STAGES", 0);
	static const int env_threads = env_int("FSB_SPMV_THREADS", 0);
	static const int env_ctas = env_int("FSB_SPMV_CTAS_PER_SM", 0);
	spmv_config k;
	k.cap = ((B.max_blk_nnz + 8 + 3) / 4) * 4; // + alignment slack on both ends
	if (k.cap < 64)
		k.cap = 64;
	k.rcap = ((B.max_blk_rows + 1 + 8 + 3) / 4) * 4;
	const size_t stage_bytes = static_cast<size_t>(k.cap) * 12 + static_cast<size_t>(k.rcap) * (B.wide ? 8 : 4);
	// consumer threads: one per row of a block, at most 512
	int consumers = env_threads > 0 ? env_threads : (c->spmv_threads > 0 ? c->spmv_threads : 512);
	static const int env_gather = env_int("FSB_SPMV_GATHER", 0);
	// 8 gathers in flight per thread is the measured optimum; the 32-wide variant (a whole 27-point row
	// per round trip) costs registers => resident CTAs and was slower on B200 (profiles/r1_spmv_sweep.txt)
	k.gather = 8;
	if (env_gather > 0)
		k.gather = env_gather > 8 ? 32 : 8;
	consumers = std::min(consumers, k.gather > 8 ? 256 : 512);
	while (consumers > 64 && consumers / 2 >= B.max_blk_rows)
		consumers /= 2;
	k.threads = consumers + 32;
	// Measured on B200 (profiles/r1_spmv_sweep.txt): throughput follows the number of resident
	// consumer threads (>= 1024 per SM saturates HBM); a second stage only pays when it does not
	// cost resident CTAs.
	const size_t budget = 220 * 1024;
	const size_t ctas_for_1024 = (1024 + consumers - 1) / consumers;
	k.nstage = (2 * stage_bytes * ctas_for_1024 <= budget) ? 2 : 1;
	if (env_stages > 0)
		k.nstage = std::min(env_stages, 2);
	while (k.nstage > 1 && stage_bytes * k.nstage > budget)
		--k.nstage;
	k.smem = stage_bytes * k.nstage;
	k.grid = env_ctas; // optional cap on resident CTAs per SM; the launcher asks the occupancy API
	return k;
}

template<class OffT, int NSTAGE, bool ACC, bool DOT, bool ROWLIST, int GATHER>
static int launch_variant(const spmv_args & a, const spmv_config & k, cudaStream_t s) {
	auto kern = spmv_stream_kernel<OffT, NSTAGE, ACC, DOT, ROWLIST, GATHER>;
	// resident CTAs per SM for this (threads, smem): persistent grid = exactly one wave
	static int cached_threads = -1, cached_ctas = 0;
	static size_t cached_smem = 0;
	if (cached_threads != k.threads || cached_smem != k.smem) {
		FSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
		int nb = 0;
		FSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, k.threads, k.smem));
		FSB_REQUIRE(nb >= 1, "spmv: kernel does not fit on an SM");
		cached_threads = k.threads;
		cached_smem = k.smem;
		cached_ctas = nb;
		if (std::getenv("FSB_SPMV_DEBUG"))
			fprintf(stderr, "[fsb] spmv config: threads %d stages %d gather %d smem %zu B cap %d -> %d CTAs/SM\n", k.threads,
			        NSTAGE, GATHER, k.smem, k.cap, nb);
	}
	int ctas = cached_ctas;
	if (k.grid > 0)
		ctas = std::min(ctas, k.grid);
	const int grid = std::max(1, std::min(a.n_blk, SM_COUNT * ctas));
	spmv_args b = a;
	// dynamic claims: CHUNK blocks per atomic, >= 8 claims per CTA, else finest grain.  Static schedule: plain
	// round-robin of single row blocks -- all CTAs sweep one narrow band of rows, which keeps the x planes a
	// stencil row needs in L2/L1 (measured: 6.9 TB/s vs 6.5 with 4-block chunks on 7-pt 256^3)
	b.chunk = a.static_sched ? 1 : std::max(1, std::min(4, a.n_blk / (grid * 8)));
	static const int env_chunk = env_int("FSB_SPMV_CHUNK", 0);
	if (env_chunk > 0)
		b.chunk = std::min(env_chunk, 4);
	kern<<<grid, k.threads, k.smem, s>>>(b);
	FSB_CUDA(cudaGetLastError());
	return grid;
}

template<class OffT, int NSTAGE, int GATHER>
static int launch_flags(const spmv_args & a, const spmv_config & k, bool acc, bool dot, bool rowlist, cudaStream_t s) {
	if (rowlist && !acc) { // structured-grid operator: rows scatter to storage offsets, plain assignment
		if constexpr (GATHER == 8 && sizeof(OffT) == 4) {
			if (dot)
				return launch_variant<OffT, NSTAGE, false, true, true, GATHER>(a, k, s);
			else
				return launch_variant<OffT, NSTAGE, false, false, true, GATHER>(a, k, s);
		}
		else
			throw error(FSB_ERR_STATE, "spmv: row-list assignment is built for int32 offsets, gather 8 only");
	}
	if (rowlist) { // off-process block: always accumulates
		if (dot)
			return launch_variant<OffT, NSTAGE, true, true, true, GATHER>(a, k, s);
		else
			return launch_variant<OffT, NSTAGE, true, false, true, GATHER>(a, k, s);
	}
	else if (acc) {
		if (dot)
			return launch_variant<OffT, NSTAGE, true, true, false, GATHER>(a, k, s);
		else
			return launch_variant<OffT, NSTAGE, true, false, false, GATHER>(a, k, s);
	}
	else {
		if (dot)
			return launch_variant<OffT, NSTAGE, false, true, false, GATHER>(a, k, s);
		else
			return launch_variant<OffT, NSTAGE, false, false, false, GATHER>(a, k, s);
	}
}

template<class OffT>
static int launch_stages(const spmv_args & a, const spmv_config & k, bool acc, bool dot, bool rowlist, cudaStream_t s) {
	if (k.gather > 8)
		return k.nstage >= 2 ? launch_flags<OffT, 2, 32>(a, k, acc, dot, rowlist, s)
		                     : launch_flags<OffT, 1, 32>(a, k, acc, dot, rowlist, s);
	return k.nstage >= 2 ? launch_flags<OffT, 2, 8>(a, k, acc, dot, rowlist, s)
	                     : launch_flags<OffT, 1, 8>(a, k, acc, dot, rowlist, s);
}

// y (+)= B x on stream s.  When dot_u != nullptr, CTA b writes its partial of sum y_i u_i to
// d_partials[partial_offset + b]; with fold_token > 0 the last CTA folds partials [0, partial_offset + grid)
// into that reduction token.  Returns the number of CTAs launched (= partials written).
int launch_spmv(fsb_ctx_s * c, const csr_block & B, const double * x, double * y, bool accumulate,
                const double * dot_u, double * d_partials, int partial_offset, cudaStream_t s, const pending * fold) {
	if (B.n_blk == 0)
		return 0;
	const spmv_config k = configure(c, B);
	FSB_REQUIRE(k.smem <= 225 * 1024, "spmv: row block does not fit shared memory");
	spmv_args a{};
	a.rowptr = B.rowptr;
	a.col = B.col;
	a.val = B.val;
	a.desc = static_cast<const blk_desc *>(B.blk_desc);
	a.row_ids = B.row_ids;
	a.x = x;
	a.y = y;
	a.u = dot_u;
	a.partials = d_partials;
	a.fold_extra = partial_offset;
	a.sched = c->d_sched;
	if (dot_u && fold) {
		fill_red_out(c, *fold, a.result);
		a.xr = c->d_xrank;
	}
	a.n_blk = B.n_blk;
	a.static_sched = c->reproducible ? 1 : 0;
	a.cap = k.cap;
	a.rcap = k.rcap;
	const bool rowlist = B.row_ids != nullptr;
	const bool dot = dot_u != nullptr;
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	const bool prof = c->profile && s == c->stream;
	if (prof) {
		while (c->prof_events.size() < c->prof_used + 2) {
			cudaEvent_t e;
			FSB_CUDA(cudaEventCreate(&e));
			c->prof_events.push_back(e);
		}
		FSB_CUDA(cudaEventRecord(c->prof_events[c->prof_used], s));
		if (c->prof_tag.size() < c->prof_used / 2 + 1)
			c->prof_tag.resize(c->prof_used / 2 + 1);
		c->prof_tag[c->prof_used / 2] = rowlist ? 1 : 0;
	}
	const int grid = B.wide ? launch_stages<long long>(a, k, accumulate, dot, rowlist, s)
	                        : launch_stages<int>(a, k, accumulate, dot, rowlist, s);
	if (prof) {
		FSB_CUDA(cudaEventRecord(c->prof_events[c->prof_used + 1], s));
		c->prof_used += 2;
	}
	return grid;
}

// fold the SpMV partials [0, n) into the reduction slot of `token`
void finalize_reduction(fsb_ctx_s * c, int n_partials, const pending & red) {
	red_out r{};
	fill_red_out(c, red, r);
	fold_partials_kernel<<<1, 256, 0, c->stream>>>(c->d_partials, n_partials, r, c->d_xrank);
	FSB_CUDA(cudaGetLastError());
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
}

// Row-block work descriptors.  host_rowptr == nullptr: uniform blocks (the caller left the row
// width bound in B.max_blk_nnz); otherwise greedy packing by nnz.
void build_blocks(fsb_ctx_s * c, csr_block & B, const std::vector<int64_t> * host_rowptr) {
	constexpr int CAP = 4096 - 8; // nnz staged per row block (48 KB per stage)
	constexpr int ROWS_MAX = 512;
	static const int env_rows = env_int("FSB_SPMV_ROWS", 0);
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
		// one thread per row, about 44 KB of matrix stream per stage, at most 512 rows
		const int width = std::max(1, B.max_blk_nnz);
		int rows = (44 * 1024) / (width * 12);
		rows = std::max(32, std::min(512, rows));
		int pow2 = 32;
		while (pow2 * 2 <= rows)
			pow2 *= 2;
		rows = pow2;
		// small matrices (the off-process block of a slab): one row block per resident CTA, so the
		// whole block is a single latency-bound pass instead of several sequential ones
		int fit = 32;
		while (fit < rows && static_cast<long long>(fit) * SM_COUNT * 2 < B.n_rows)
			fit *= 2;
		rows = std::min(rows, fit);
		if (c->spmv_rows_per_cta > 0)
			rows = c->spmv_rows_per_cta;
		if (env_rows > 0)
			rows = env_rows;
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
	FSB_CUDA(cudaMalloc(&B.blk_desc, static_cast<size_t>(B.n_blk) * sizeof(blk_desc)));
	build_desc_kernel<<<(B.n_blk + 255) / 256, 256, 0, c->stream>>>(B.rowptr, B.wide, B.blk_row, B.n_blk,
	                                                               static_cast<blk_desc *>(B.blk_desc));
	FSB_CUDA(cudaGetLastError());
	FSB_CUDA(cudaStreamSynchronize(c->stream));
}

} // namespace fsb
