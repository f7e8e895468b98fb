// CSR SpMV for sm_100a: TMA-staged streaming kernels.
//
// Reference semantics (matrices/seq.hh:178-194): y[i] = 0; for off in row i, in index
// order: y[i] += val[off] * x[col[off]].  Both kernels keep exactly that accumulation
// order per row (thread-per-row path), so 1-rank results are bit-identical to the
// reference loop compiled without FMA contraction.
//
// Design (HBM-bound: the matrix stream is 80-95 % of the bytes and perfectly contiguous per row block):
//  * Warp-specialised persistent CTAs.  The producer warp pulls the arrays of a block of consecutive rows into
//    shared memory with 1-D bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx) -- fully coalesced HBM bursts
//    regardless of row length, no per-thread load instructions; consumer warps multiply out of shared memory,
//    thread-per-row with consecutive lanes on consecutive rows.
//  * spmv_stream_kernel ("gather" format = plain CSR, 12 B/nnz): x[col] is gathered through L1/L2, GATHER loads in
//    flight per thread.  Right for short rows (7-point: x reuse distance is two planes, gathers hit L1).
//  * spmv_window_kernel ("window" format, 10 B/nnz): at setup every row block gets the list of contiguous pieces of
//    x it reads (<= 16 segments) and a 16-bit local column index per nonzero = position inside those pieces; the
//    producer stages the pieces with bulk copies as well, so the consumers touch shared memory only.  Right for wide
//    stencils (27-point: 9 segments per block, 23 % of the block's matrix bytes, served by L2): no gather latency,
//    no L1 thrash, 2 B/nnz less HBM traffic, 16-bit block-relative row offsets instead of 4/8-byte ones.
//  * Several ranks: ONE launch does the whole parcsr product (matrices/parcsr.hh:61-91).  The consumer warps of the
//    first CTAs push this rank's boundary entries of x into the neighbours' landing areas over NVLink while the
//    pipeline fills; interior row blocks run; row blocks that own rows with off-process entries are scheduled
//    last, wait for the neighbours' `ready` flags and add the off-process part reading the landing area directly;
//    the last CTA folds the fused dot, all-reduces it across ranks and acknowledges the landing area.
//  * Optional fused epilogues: per-CTA partial of sum_i y_i * u_i (CG's p.Ap, cg.hh:98); weighted Jacobi
//    (mg/jacobi.hh:73-89: skip the diagonal while summing, x = w/d (b - lpu) + (1-w) tmp) in the same pass.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <thrust/execution_policy.h>
#include <thrust/functional.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>

#include "fsb_internal.h"
#include "setup_exchange.h"
#include "spmv_common.cuh"

namespace fsb {

constexpr int SPMV_MAX_THREADS = 544; // 512 consumer threads + one producer warp
constexpr int SPMV_MAX_STAGES = 4;
constexpr int ROWS_MAX = 512; // rows per block (shared scratch of the boundary phases is sized for it)
constexpr int WIN_MAXSEG = 16; // x segments per row block (window format)
constexpr int WIN_GAP = 8; // columns closer than this are covered by one segment
constexpr unsigned WIN_NO_DIAG = 0xffffu; // row meta: the row stores no diagonal entry

struct spmv_args {
	const void * rowptr;
	const int32_t * col;
	const double * val;
	const blk_desc * desc; // [n_blk], interior blocks first
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
	int cap; // nnz capacity of one stage (multiple of 8, includes alignment slack)
	int rcap; // row-offset capacity of one stage
	int chunk; // row blocks claimed per atomic (1..4)
	int static_sched; // 1: claims are handed out round-robin by CTA index (bitwise reproducible dot partials)
	int acc_continue; // accumulate variants: 1 = continue the row's running sum (Jacobi), 0 = add a separate sum (SpMV)
	// fused ghost exchange (HALO variants)
	const halo_dev * halo;
	long long epoch;
	int push_parts; // CTAs whose consumer warps take part in the push
	const int32_t * o_rowptr; // off-process block over its compressed rows
	const int32_t * o_col;
	const double * o_val;
	const int32_t * o_rows;
	long long n_owned;
	// weighted Jacobi (JAC variants): x = previous iterate, y = new iterate
	const double * jb;
	double omega;
	// window format
	const uint16_t * lcol;
	const uint16_t * rp16; // per row block: nrows + 1 row offsets relative to its first nonzero, then nrows diagonal positions
	const x_segment * segs; // [n_blk][WIN_MAXSEG] in row order
	int xcap; // doubles of x staged per block (capacity)
	const double * aux; // row-aligned operand of the epilogue (window kernel): the dot operand u, or b of a Jacobi sweep
	int acap; // its capacity per stage when it is staged with the block (aux_mode 1)
	int aux_mode; // 0 none; 1 staged with the block by a bulk copy; 2 u is x itself: read it from the x window at the
	              // diagonal's position (CG's <Ap, p>: no extra bytes at all); 3 plain global load
	// value dictionary (window format of a matrix with <= 256 distinct values): one byte per nonzero instead of eight
	const uint8_t * vidx; // [padded slots] index into vdict, or nullptr
	const double * vdict; // [256]
	const uint16_t * plcol; // [padded slots] local columns in the same padded order
	const uint16_t * pmeta; // per row block: nrows + 1 slot offsets, nrows row lengths, nrows diagonal positions
	unsigned long long * tl; // timeline slot of this launch or nullptr
	setup_exchange * meet; // host side only: rendezvous of the ranks' threads right before the launch (in-process groups)
};

// ------------------------------------------------------------------------------------------------ shared pieces

// row result of a Jacobi sweep: omega/diag * (b - lpu) + (1 - omega) * old   (mg/jacobi.hh:88-89, evaluated as written)
__device__ __forceinline__ double jacobi_value(double omega, double diag, double b, double lpu, double old) {
	const double dinv = __ddiv_rn(1.0, diag);
	return __dadd_rn(__dmul_rn(__dmul_rn(omega, dinv), __dadd_rn(b, -lpu)), __dmul_rn(__dadd_rn(1.0, -omega), old));
}

// what a consumer thread does with a finished row of an interior block
template<bool ACC, bool DOT, bool ROWLIST, int JAC>
__device__ __forceinline__ void emit_row(const spmv_args & a, int r, double sum, double dg, double & dot_acc) {
	const int yr = ROWLIST ? a.row_ids[r] : r;
	if constexpr (JAC == 1) {
		a.y[yr] = jacobi_value(a.omega, dg, a.jb[yr], sum, a.x[yr]);
		return;
	}
	double out = sum;
	if constexpr (ACC) {
		const double old = a.y[yr];
		out = a.acc_continue ? sum : __dadd_rn(old, sum); // continue: the caller started the sum from y[yr]
		if constexpr (DOT)
			dot_acc += (a.u == a.y) ? (out * out - old * old) : sum * a.u[yr];
	}
	else if constexpr (DOT) {
		dot_acc = fma(out, a.u == a.y ? out : a.u[yr], dot_acc);
	}
	a.y[yr] = out;
}

// the same for the window kernel, whose row-aligned operands (u or b, and the old iterate of a Jacobi sweep) come
// from shared memory
template<bool DOT, int JAC>
__device__ __forceinline__ void emit_row_staged(const spmv_args & a, int r, double sum, double dg, double auxv, double xold,
                                                double & dot_acc) {
	if constexpr (JAC == 1) {
		a.y[r] = jacobi_value(a.omega, dg, auxv, sum, xold);
		return;
	}
	if constexpr (DOT)
		dot_acc = fma(sum, a.u == a.y ? sum : auxv, dot_acc);
	a.y[r] = sum;
}

// Rows of a boundary block (HALO): phase 1 left the owned-column part of every row in srow[] (and, for plain SpMV,
// already in y and in the dot partial); here the rows that have off-process entries get them added, reading the
// ghost values straight from this rank's landing area.
template<bool DOT, int JAC>
__device__ __forceinline__ void boundary_tail(const spmv_args & a, const blk_desc & d, double * srow, double * sdg, int ctid,
                                              int nconsumers, bool & ghosts_ready, double & dot_acc) {
	const halo_dev & h = *a.halo;
	consumer_sync(nconsumers); // srow[] of this block is complete
	const unsigned long long w0 = a.tl ? tl_now() : 0ull;
	const unsigned char * land = halo_landing(h.base[h.me], static_cast<int>(a.epoch & 1), h.gmax);
	const unsigned flag = static_cast<unsigned>(a.epoch);
	for (int k = ctid; k < d.ocnt; k += nconsumers) {
		const int r = a.o_rows[d.o0 + k], i = r - d.r0;
		const int q0 = a.o_rowptr[d.o0 + k], q1 = a.o_rowptr[d.o0 + k + 1];
		if constexpr (JAC == 1) { // lpu keeps running over the off-process entries (mg/jacobi.hh:83-86)
			double s = srow[i];
			for (int q = q0; q < q1; ++q)
				s = __dadd_rn(s, __dmul_rn(a.o_val[q], halo_ghost(h, land, a.o_col[q] - a.n_owned, flag)));
			srow[i] = s;
		}
		else { // y = diag x + (offd x): the off-process part is summed on its own, then added (parcsr.hh:61-68)
			double so = 0.0;
			for (int q = q0; q < q1; ++q)
				so = __dadd_rn(so, __dmul_rn(a.o_val[q], halo_ghost(h, land, a.o_col[q] - a.n_owned, flag)));
			const double old = srow[i], out = __dadd_rn(old, so);
			a.y[r] = out;
			if constexpr (DOT)
				dot_acc += (a.u == a.y) ? (out * out - old * old) : so * a.u[r];
		}
	}
	if (a.tl && ctid == 0) {
		if (!ghosts_ready)
			tl_mark_min(a.tl, 4); // the first boundary block has its ghosts
		atomicMax(a.tl + 5, tl_now() - w0); // longest off-process phase of a block, waiting included
	}
	ghosts_ready = true;
	if constexpr (JAC == 1) {
		consumer_sync(nconsumers);
		for (int i = ctid; i < d.nrows; i += nconsumers) {
			const int r = d.r0 + i;
			a.y[r] = jacobi_value(a.omega, sdg[i], a.jb[r], srow[i], a.x[r]);
		}
	}
	consumer_sync(nconsumers); // srow[] may be overwritten by the next block
}

// the last CTA to finish folds the dot partials (and all-reduces them), acknowledges the landing area and re-arms
// the scheduler words
template<bool DOT, bool HALO>
__device__ __forceinline__ void spmv_epilogue(const spmv_args & a, double dot_acc, double * scratch) {
	__shared__ bool is_last;
	const int tid = threadIdx.x;
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
		if constexpr (HALO) {
			if (tid == 0)
				halo_acknowledge(*a.halo, a.epoch); // every CTA is past its reads of the landing area
		}
		if constexpr (DOT) {
			if (a.result.token != 0) {
				__threadfence();
				const int total = a.fold_extra + static_cast<int>(gridDim.x);
				double t = 0.0;
				for (int i = tid; i < total; i += blockDim.x)
					t += __ldcg(&a.partials[i]);
				t = block_fold<0>(t, scratch);
				if (tid == 0)
					tl_mark_min(a.tl, 6);
				if (a.xr)
					t = xrank_allreduce<0>(a.xr, t, a.result.token, scratch);
				if (tid == 0) {
					publish(a.result, t);
					tl_mark_max(a.tl, 7);
				}
			}
		}
		if (tid == 0) {
			a.sched[0] = 0u;
			a.sched[1] = 0u;
		}
	}
	tl_end(a.tl);
}

// next row block of this CTA: round-robin by CTA index (static) or an atomic claim (dynamic)
__device__ __forceinline__ int claim_blocks(const spmv_args & a, int & round, int chunk, int lane) {
	if (a.static_sched)
		return static_cast<int>(
			min(static_cast<long long>(a.n_blk), (static_cast<long long>(round++) * gridDim.x + blockIdx.x) * chunk));
	int base = 0;
	if (lane == 0)
		base = static_cast<int>(atomicAdd(&a.sched[0], static_cast<unsigned>(chunk)));
	return __shfl_sync(0xffffffffu, base, 0);
}

// ------------------------------------------------------------------------------------------------ gather format

// Warp-specialised: warp 0 is the producer (one lane issues descriptor loads and bulk copies and
// runs up to NSTAGE row blocks ahead, throttled by the `empty` barriers); all other warps are
// consumers (wait `full`, multiply thread-per-row out of shared memory, release the stage).
// GATHER: x[col] loads in flight per thread before the in-order sum.
// JAC: 0 plain product; 1 weighted-Jacobi sweep; 2 product without the diagonal entry (first launch of a Jacobi
// sweep whose off-process part comes in a second launch)
template<class OffT, int NSTAGE, bool ACC, bool DOT, bool ROWLIST, bool HALO, int JAC>
__global__ void __launch_bounds__(SPMV_MAX_THREADS, 2) spmv_stream_kernel(const __grid_constant__ spmv_args a) {
	constexpr int GATHER = 8;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[NSTAGE], empty[NSTAGE];
	__shared__ blk_desc sdesc[NSTAGE];
	__shared__ double scratch[32];
	__shared__ double srow[HALO ? ROWS_MAX : 1], sdg[(HALO && JAC == 1) ? ROWS_MAX : 1];
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
	grid_dependency_wait(); // everything above is independent of the previous kernel (programmatic dependent launch)
	tl_begin(a.tl);

	double dot_acc = 0.0;

	if (tid < 32) {
		// ------------------------------------------------ producer
		// Row blocks are claimed CHUNK (<= 4) consecutive blocks at a time.  Lane l holds the descriptor of block
		// base + l; the next claim and its descriptor loads are issued before the current chunk is staged, so
		// neither the atomic nor the DRAM latency of the descriptors sits between a freed stage and the next copy.
		const int CHUNK = a.chunk;
		const int lane = tid;
		int round = 0;
		auto fetch = [&](int base) {
			blk_desc d{};
			if (lane < CHUNK && base + lane < a.n_blk)
				d = a.desc[base + lane];
			return d;
		};
		int base = claim_blocks(a, round, CHUNK, lane);
		blk_desc cur = fetch(base);
		int it = 0;
		while (base < a.n_blk) {
			const int nbase = claim_blocks(a, round, CHUNK, lane);
			const blk_desc nxt = fetch(nbase);
			const int jmax = min(CHUNK, a.n_blk - base);
			for (int j = 0; j < jmax; ++j, ++it) {
				blk_desc d{};
				d.z0 = __shfl_sync(0xffffffffu, cur.z0, j);
				d.r0 = __shfl_sync(0xffffffffu, cur.r0, j);
				d.nrows = __shfl_sync(0xffffffffu, cur.nrows, j);
				d.nnz = __shfl_sync(0xffffffffu, cur.nnz, j);
				d.o0 = __shfl_sync(0xffffffffu, cur.o0, j);
				d.ocnt = __shfl_sync(0xffffffffu, cur.ocnt, j);
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
		bool ghosts_ready = false;
		if constexpr (HALO) {
			// while the first stage is in flight: push this rank's boundary entries of x to the neighbours
			if (static_cast<int>(blockIdx.x) < a.push_parts) {
				halo_push_part(*a.halo, x, a.epoch, ctid, nconsumers, blockIdx.x, a.push_parts,
				               [&] { consumer_sync(nconsumers); }, a.tl ? a.tl + 8 : nullptr);
				if (ctid == 0)
					tl_mark_max(a.tl, 3);
			}
		}

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
			const bool boundary = HALO && d.ocnt > 0;
			// a finished row: interior blocks write it out; boundary blocks park the owned-column part in srow[]
			auto finish = [&](int i, double sum, double dg) {
				if constexpr (HALO) {
					if (boundary) {
						srow[i] = sum;
						if constexpr (JAC == 1)
							sdg[i] = dg;
						else
							emit_row<ACC, DOT, ROWLIST, JAC>(a, r0 + i, sum, dg, dot_acc);
						return;
					}
				}
				emit_row<ACC, DOT, ROWLIST, JAC>(a, r0 + i, sum, dg, dot_acc);
			};

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
						finish(i, t, 0.0);
					}
				}
			}
			else if (cnt <= static_cast<long long>(nrows) * 64) {
				// short rows (average <= 64 entries): thread per row, consecutive lanes on consecutive
				// rows; in-order accumulation
				for (int i = ctid; i < nrows; i += nconsumers) {
					const int p0 = static_cast<int>(static_cast<long long>(srp[roff + i]) - za);
					const int p1 = static_cast<int>(static_cast<long long>(srp[roff + i + 1]) - za);
					double sum = 0.0, dg = 0.0;
					if constexpr (ACC) {
						if (a.acc_continue)
							sum = a.y[ROWLIST ? a.row_ids[r0 + i] : r0 + i];
					}
					for (int p = p0; p < p1; p += GATHER) {
						double prod[GATHER];
						if constexpr (JAC == 0) {
#pragma unroll
							for (int k = 0; k < GATHER; ++k)
								if (p + k < p1)
									prod[k] = __dmul_rn(sv[p + k], __ldg(&x[sc[p + k]]));
#pragma unroll
							for (int k = 0; k < GATHER; ++k)
								if (p + k < p1)
									sum = __dadd_rn(sum, prod[k]);
						}
						else {
							bool keep[GATHER];
#pragma unroll
							for (int k = 0; k < GATHER; ++k) {
								keep[k] = p + k < p1;
								if (keep[k]) {
									const int c = sc[p + k];
									if (c == r0 + i) { // the diagonal (global_id(row) == global_id(col), mg/jacobi.hh:76-79)
										dg = sv[p + k];
										keep[k] = false;
									}
									else
										prod[k] = __dmul_rn(sv[p + k], __ldg(&x[c]));
								}
							}
#pragma unroll
							for (int k = 0; k < GATHER; ++k)
								if (keep[k])
									sum = __dadd_rn(sum, prod[k]);
						}
					}
					finish(i, sum, dg);
				}
			}
			else {
				// few long rows: warp per row, lanes stride the staged row, shuffle tree
				const int lane = ctid & 31, warp = ctid >> 5;
				for (int i = warp; i < nrows; i += nconsumers / 32) {
					const int q0 = static_cast<int>(static_cast<long long>(srp[roff + i]) - za);
					const int q1 = static_cast<int>(static_cast<long long>(srp[roff + i + 1]) - za);
					double sum = 0.0, dg = 0.0;
					int dpos = -1;
					for (int p = q0 + lane; p < q1; p += 32) {
						const int c = sc[p];
						if (JAC != 0 && c == r0 + i) {
							dg = sv[p];
							dpos = p;
							continue;
						}
						sum = fma(sv[p], __ldg(&x[c]), sum);
					}
					sum = warp_fold<0>(sum);
					if constexpr (JAC != 0) { // the last stored diagonal entry wins, like the reference's loop
						int best = dpos;
#pragma unroll
						for (int o = 16; o > 0; o >>= 1)
							best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
						const unsigned who = __ballot_sync(0xffffffffu, dpos == best && best >= 0);
						dg = who ? __shfl_sync(0xffffffffu, dg, __ffs(who) - 1) : 0.0;
					}
					if (lane == 0) {
						if constexpr (ACC) {
							if (a.acc_continue)
								sum += a.y[ROWLIST ? a.row_ids[r0 + i] : r0 + i];
						}
						finish(i, sum, dg);
					}
				}
			}
			__syncwarp();
			if ((ctid & 31) == 0)
				mbar_arrive(&empty[s]); // this warp is done reading stage s
			if constexpr (HALO) {
				if (boundary)
					boundary_tail<DOT, JAC>(a, d, srow, sdg, ctid, nconsumers, ghosts_ready, dot_acc);
			}
		}
	}
	spmv_epilogue<DOT, HALO>(a, dot_acc, scratch);
}

// ------------------------------------------------------------------------------------------------ window format

// Stage layout: [val 8 cap][x window 8 xcap][u or b 8 acap][local columns 2 cap][row offsets 2 rcap].
// Same roles as above; the producer additionally stages the block's x segments, so consumers read shared memory only.
// VD ("value dictionary"): the matrix holds at most 256 distinct values (every constant-coefficient stencil, graph
// Laplacians, ...): the stage carries one byte per nonzero [value index 1 cap] instead of the fp64 value and the
// consumers look the value up in a 2 KB table in shared memory -- 3 B instead of 10 B of HBM traffic per nonzero, the
// same doubles multiplied in the same order.
template<int NSTAGE, bool DOT, bool HALO, int JAC, bool VD>
__global__ void __launch_bounds__(SPMV_MAX_THREADS, 2) spmv_window_kernel(const __grid_constant__ spmv_args a) {
	constexpr int UNROLL = 8;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[NSTAGE], empty[NSTAGE];
	__shared__ blk_desc sdesc[NSTAGE];
	__shared__ double scratch[32];
	__shared__ double srow[HALO ? ROWS_MAX : 1], sdg[(HALO && JAC == 1) ? ROWS_MAX : 1];
	__shared__ double sdict[VD ? 256 : 1];
	const size_t off_x = static_cast<size_t>(a.cap) * (VD ? 1 : 8);
	const size_t off_a = off_x + static_cast<size_t>(a.xcap) * 8;
	const size_t off_c = off_a + static_cast<size_t>(a.acap) * 8;
	const size_t off_r = off_c + static_cast<size_t>(a.cap) * 2;
	const size_t stage_bytes = off_r + static_cast<size_t>(a.rcap) * 2;
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
	if constexpr (VD) {
		if (tid < 256)
			sdict[tid] = a.vdict[tid];
	}
	__syncthreads();
	grid_dependency_wait(); // everything above is independent of the previous kernel (programmatic dependent launch)
	tl_begin(a.tl);

	double dot_acc = 0.0;

	if (tid < 32) {
		// ------------------------------------------------ producer
		// Software pipeline over this CTA's blocks k = 0, 1, ...: the descriptor of block k+1 (one 4-byte word per
		// lane) and its segment list (one segment per lane) are loaded while block k is being staged, so no global
		// load latency sits between a freed stage and its bulk copies.
		const int lane = tid;
		constexpr int DW = sizeof(blk_desc) / 4;
		int round = 0;
		auto load_desc = [&](int b) { return (b < a.n_blk && lane < DW) ? reinterpret_cast<const int *>(a.desc + b)[lane] : 0; };
		auto word = [&](int w, int k) { return __shfl_sync(0xffffffffu, w, k); };
		auto load_segment = [&](int b, int dwords) {
			const int slot = word(dwords, 8); // blk_desc::seg_slot
			x_segment g{0, 0};
			if (b < a.n_blk && lane < WIN_MAXSEG)
				g = a.segs[static_cast<size_t>(slot) + lane];
			return g;
		};
		int blk = claim_blocks(a, round, 1, lane);
		int dw = load_desc(blk);
		x_segment sg = load_segment(blk, dw);
		int it = 0;
		for (; blk < a.n_blk; ++it) {
			const int nblk = claim_blocks(a, round, 1, lane);
			const int ndw = load_desc(nblk);
			blk_desc d{};
			d.z0 = static_cast<long long>(static_cast<unsigned>(word(dw, 0))) | (static_cast<long long>(word(dw, 1)) << 32);
			d.r0 = word(dw, 2);
			d.nrows = word(dw, 3);
			d.nnz = word(dw, 4);
			d.o0 = word(dw, 5);
			d.ocnt = word(dw, 6);
			d.rp0 = word(dw, 7);
			if constexpr (VD) {
				d.zp0 = static_cast<long long>(static_cast<unsigned>(word(dw, 10))) | (static_cast<long long>(word(dw, 11)) << 32);
				d.pnnz = word(dw, 12);
				d.rpp0 = word(dw, 13);
			}
			// position of every segment inside the staged window: exclusive prefix sum of the lengths
			int soff = sg.len;
#pragma unroll
			for (int o = 1; o < WIN_MAXSEG; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, soff, o);
				if (lane >= o)
					soff += t;
			}
			const int xlen = __shfl_sync(0xffffffffu, soff, WIN_MAXSEG - 1);
			soff -= sg.len;
			const int s = it % NSTAGE;
			unsigned char * sbase = smem + s * stage_bytes;
			const long long za = VD ? d.zp0 : (d.z0 & ~7LL); // 16-byte aligned start of the per-nonzero streams
			const long long cnt = VD ? d.pnnz : ((d.z0 + d.nnz + 7) & ~7LL) - za;
			const long long ra = (VD ? d.rpp0 : d.rp0) & ~7LL;
			const long long rcnt = ((static_cast<long long>(VD ? d.rpp0 : d.rp0) + (VD ? 3 : 2) * d.nrows + 1 + 7) & ~7LL) - ra;
			const long long aa = d.r0 & ~1LL; // row-aligned operand: rows [r0, r0 + nrows), 16-byte granular
			const long long acnt = a.aux_mode == 1 ? ((static_cast<long long>(d.r0) + d.nrows + 1) & ~1LL) - aa : 0;
			if (lane == 0) {
				if (it >= NSTAGE)
					mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
				sdesc[s] = d;
				mbar_expect_tx(&full[s], static_cast<uint32_t>(cnt * (VD ? 3 : 10) + rcnt * 2 + (static_cast<long long>(xlen) + acnt) * 8));
				if (acnt > 0)
					tma_load_1d(sbase + off_a, a.aux + aa, static_cast<uint32_t>(acnt * 8), &full[s]);
				if constexpr (VD) {
					tma_load_1d(sbase, a.vidx + za, static_cast<uint32_t>(cnt), &full[s]);
					tma_load_1d(sbase + off_c, a.plcol + za, static_cast<uint32_t>(cnt * 2), &full[s]);
					tma_load_1d(sbase + off_r, a.pmeta + ra, static_cast<uint32_t>(rcnt * 2), &full[s]);
				}
				else {
					tma_load_1d(sbase, a.val + za, static_cast<uint32_t>(cnt * 8), &full[s]);
					tma_load_1d(sbase + off_c, a.lcol + za, static_cast<uint32_t>(cnt * 2), &full[s]);
					tma_load_1d(sbase + off_r, a.rp16 + ra, static_cast<uint32_t>(rcnt * 2), &full[s]);
				}
			}
			__syncwarp();
			if (lane < WIN_MAXSEG && sg.len > 0) // one bulk copy per segment, issued by the lane that holds it
				tma_load_1d(sbase + off_x + static_cast<size_t>(soff) * 8, a.x + sg.start, static_cast<uint32_t>(sg.len) * 8,
				            &full[s]);
			__syncwarp();
			sg = load_segment(nblk, ndw); // the next block's descriptor has arrived by now
			blk = nblk;
			dw = ndw;
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
		bool ghosts_ready = false;
		if constexpr (HALO) {
			if (static_cast<int>(blockIdx.x) < a.push_parts) {
				halo_push_part(*a.halo, a.x, a.epoch, ctid, nconsumers, blockIdx.x, a.push_parts,
				               [&] { consumer_sync(nconsumers); }, a.tl ? a.tl + 8 : nullptr);
				if (ctid == 0)
					tl_mark_max(a.tl, 3);
			}
		}
		for (int it = 0;; ++it) {
			const int s = it % NSTAGE;
			mbar_wait(&full[s], (it / NSTAGE) & 1);
			const blk_desc d = sdesc[s];
			if (d.nrows < 0)
				break;
			const unsigned char * base = smem + s * stage_bytes;
			const double * xs = reinterpret_cast<const double *>(base + off_x);
			const double * sa = reinterpret_cast<const double *>(base + off_a) + (d.r0 & 1);
			const int aux_mode = a.aux_mode;
			const int r0 = d.r0, nrows = d.nrows;
			const bool boundary = HALO && d.ocnt > 0;
			// plain stream: values, local columns and row offsets as the matrix has them
			const long long za = d.z0 & ~7LL;
			const double * sv = reinterpret_cast<const double *>(base) + (d.z0 - za);
			const uint16_t * sl = reinterpret_cast<const uint16_t *>(base + off_c) + (VD ? 0 : d.z0 - za);
			const uint16_t * srp = reinterpret_cast<const uint16_t *>(base + off_r) + ((VD ? d.rpp0 : d.rp0) & 7);
			// dictionary stream: rows start at multiples of 8 slots, so a trip of 8 nonzeros is one 8-byte load of value
			// indices and one 16-byte load of local columns
			const uint8_t * sv8 = base;
			const uint16_t * slen = srp + nrows + 1; // (VD) nonzeros of each row
			const uint16_t * sdk = VD ? slen + nrows : srp + nrows + 1; // position of the diagonal entry inside each row (or WIN_NO_DIAG)
			for (int i = ctid; i < nrows; i += nconsumers) {
				const int p0 = srp[i];
				const int p1 = VD ? p0 + static_cast<int>(slen[i]) : static_cast<int>(srp[i + 1]);
				const unsigned dk = sdk[i];
				const int pd = dk != WIN_NO_DIAG ? p0 + static_cast<int>(dk) : -1; // the row's diagonal entry
				const bool have_diag = pd >= 0;
				double sum = 0.0, dg = 0.0, xold = 0.0;
				if constexpr (VD) {
					for (int p = p0; p < p1; p += UNROLL) {
						const uint2 iv = *reinterpret_cast<const uint2 *>(sv8 + p);
						const uint4 cv = *reinterpret_cast<const uint4 *>(sl + p);
						const unsigned cw[4] = {cv.x, cv.y, cv.z, cv.w};
						double prod[UNROLL];
#pragma unroll
						for (int k = 0; k < UNROLL; ++k) { // padding slots hold index 0 / column 0 and are dropped below
							const unsigned vi = ((k < 4 ? iv.x : iv.y) >> (8 * (k & 3))) & 0xffu;
							const unsigned lc = (cw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
							prod[k] = __dmul_rn(sdict[vi], xs[lc]);
						}
#pragma unroll
						for (int k = 0; k < UNROLL; ++k) {
							const bool keep = p + k < p1 && (JAC == 0 || p + k != pd); // a Jacobi sweep leaves the diagonal out
							const double t = __dadd_rn(sum, prod[k]);
							sum = keep ? t : sum;
						}
					}
				}
				else {
					for (int p = p0; p < p1; p += UNROLL) {
						double prod[UNROLL];
#pragma unroll
						for (int k = 0; k < UNROLL; ++k)
							if (p + k < p1 && (JAC == 0 || p + k != pd)) // a Jacobi sweep leaves the diagonal out of the sum
								prod[k] = __dmul_rn(sv[p + k], xs[sl[p + k]]);
#pragma unroll
						for (int k = 0; k < UNROLL; ++k)
							if (p + k < p1 && (JAC == 0 || p + k != pd))
								sum = __dadd_rn(sum, prod[k]);
					}
				}
				if ((JAC != 0 || aux_mode == 2) && have_diag) { // x[row] sits in the window where the diagonal entry points
					xold = xs[sl[pd]];
					if constexpr (VD)
						dg = sdict[sv8[pd]];
					else
						dg = sv[pd];
				}
				double auxv = 0.0;
				if (aux_mode == 1)
					auxv = sa[i];
				else if (aux_mode == 2)
					auxv = have_diag ? xold : a.x[r0 + i];
				else if (aux_mode == 3)
					auxv = a.aux[r0 + i];
				if constexpr (JAC == 1) {
					if (!have_diag)
						xold = a.x[r0 + i];
				}
				if constexpr (HALO) {
					if (boundary) {
						srow[i] = sum;
						if constexpr (JAC == 1)
							sdg[i] = dg;
						else
							emit_row_staged<DOT, JAC>(a, r0 + i, sum, dg, auxv, xold, dot_acc);
						continue;
					}
				}
				emit_row_staged<DOT, JAC>(a, r0 + i, sum, dg, auxv, xold, dot_acc);
			}
			__syncwarp();
			if ((ctid & 31) == 0)
				mbar_arrive(&empty[s]);
			if constexpr (HALO) {
				if (boundary)
					boundary_tail<DOT, JAC>(a, d, srow, sdg, ctid, nconsumers, ghosts_ready, dot_acc);
			}
		}
	}
	spmv_epilogue<DOT, HALO>(a, dot_acc, scratch);
}

// ------------------------------------------------------------------------------------------------ setup kernels

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
	blk_desc d{};
	d.z0 = z0;
	d.r0 = r0;
	d.nrows = r1 - r0;
	d.nnz = static_cast<int>(z1 - z0);
	d.rp0 = 2 * r0 + b; // window format: the block's 16-bit row meta (nrows + 1 offsets, then nrows diagonal positions)
	d.seg_slot = b * WIN_MAXSEG; // window format: the block's segment slots (unused ones have len 0)
	out[b] = d;
}

// rows [o_rows[o0], ...) of the off-process block that fall into each row block (o_rows ascending)
__global__ void attach_offd_kernel(blk_desc * desc, int n_blk, const int32_t * o_rows, int n_orows) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blk)
		return;
	auto lower = [&](int r) {
		int lo = 0, hi = n_orows;
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (o_rows[mid] < r)
				lo = mid + 1;
			else
				hi = mid;
		}
		return lo;
	};
	const int o0 = lower(desc[b].r0), o1 = lower(desc[b].r0 + desc[b].nrows);
	desc[b].o0 = o0;
	desc[b].ocnt = o1 - o0;
}

// Window format of one row block per CTA: sort the block's column indices, cut them into segments of nearby
// columns, give every nonzero its position inside the staged segments (+ the diagonal mark), and write the
// block-relative 16-bit row offsets.  status[0] |= 1 if some block needs more than WIN_MAXSEG segments or more than
// xcap_limit staged entries (the matrix then stays in the gather format); status[1] = max staged entries.
constexpr int WIN_THREADS = 256, WIN_ITEMS = 16; // 4096 column indices per block
template<class OffT>
__global__ void __launch_bounds__(WIN_THREADS) build_window_kernel(const OffT * __restrict__ rowptr, const int32_t * __restrict__ col,
                                                                   const int32_t * __restrict__ blk_row, int n_cols, int xcap_limit,
                                                                   uint16_t * __restrict__ lcol, uint16_t * __restrict__ rp16,
                                                                   x_segment * __restrict__ segs, int * __restrict__ status) {
	using Sort = cub::BlockRadixSort<int, WIN_THREADS, WIN_ITEMS>;
	using Scan = cub::BlockScan<int, WIN_THREADS>;
	__shared__ union {
		typename Sort::TempStorage sort;
		typename Scan::TempStorage scan;
	} tmp;
	__shared__ int sorted[WIN_THREADS * WIN_ITEMS];
	__shared__ int seg_start[WIN_MAXSEG + 1], seg_end[WIN_MAXSEG + 1], seg_off[WIN_MAXSEG + 1];
	__shared__ int s_nseg;
	const int b = blockIdx.x, tid = threadIdx.x;
	const int r0 = blk_row[b], r1 = blk_row[b + 1];
	const long long z0 = static_cast<long long>(rowptr[r0]);
	const int nnz = static_cast<int>(static_cast<long long>(rowptr[r1]) - z0);
	uint16_t * meta = rp16 + 2 * static_cast<size_t>(r0) + b; // [nrows + 1 offsets | nrows diagonal positions]
	for (int i = tid; i <= r1 - r0; i += WIN_THREADS)
		meta[i] = static_cast<uint16_t>(static_cast<long long>(rowptr[r0 + i]) - z0);
	if (tid < WIN_MAXSEG)
		segs[static_cast<size_t>(b) * WIN_MAXSEG + tid] = x_segment{0, 0};
	if (nnz > WIN_THREADS * WIN_ITEMS) {
		if (tid == 0)
			atomicOr(&status[0], 1);
		return;
	}
	int keys[WIN_ITEMS];
#pragma unroll
	for (int k = 0; k < WIN_ITEMS; ++k) {
		const int i = tid * WIN_ITEMS + k;
		keys[k] = i < nnz ? col[z0 + i] : 0x7fffffff;
	}
	Sort(tmp.sort).Sort(keys);
	__syncthreads();
#pragma unroll
	for (int k = 0; k < WIN_ITEMS; ++k)
		sorted[tid * WIN_ITEMS + k] = keys[k];
	__syncthreads();
	// a segment starts where the gap to the previous (sorted) column exceeds WIN_GAP
	int heads = 0;
	unsigned head_mask = 0;
#pragma unroll
	for (int k = 0; k < WIN_ITEMS; ++k) {
		const int i = tid * WIN_ITEMS + k;
		const bool head = i < nnz && (i == 0 || sorted[i] - sorted[i - 1] > WIN_GAP);
		heads += head;
		head_mask |= static_cast<unsigned>(head) << k;
	}
	int before = 0, total = 0;
	Scan(tmp.scan).ExclusiveSum(heads, before, total);
	if (tid == 0)
		s_nseg = total;
	if (total <= WIN_MAXSEG) {
		int id = before;
#pragma unroll
		for (int k = 0; k < WIN_ITEMS; ++k) {
			const int i = tid * WIN_ITEMS + k;
			if (head_mask & (1u << k)) {
				seg_start[id] = sorted[i] & ~1; // 16-byte aligned
				if (id > 0)
					seg_end[id - 1] = min((sorted[i - 1] + 2) & ~1, (n_cols + 1) & ~1);
				++id;
			}
		}
		if (nnz > 0 && tid == 0)
			seg_end[total - 1] = min((sorted[nnz - 1] + 2) & ~1, (n_cols + 1) & ~1);
	}
	__syncthreads();
	const int nseg = s_nseg;
	if (nseg > WIN_MAXSEG) {
		if (tid == 0)
			atomicOr(&status[0], 1);
		return;
	}
	if (tid == 0) {
		int off = 0;
		for (int s = 0; s < nseg; ++s) {
			seg_off[s] = off;
			off += seg_end[s] - seg_start[s];
			segs[static_cast<size_t>(b) * WIN_MAXSEG + s] = x_segment{seg_start[s], seg_end[s] - seg_start[s]};
		}
		seg_off[nseg] = off;
		atomicMax(&status[1], off);
		if (off > xcap_limit || off > 0xffff)
			atomicOr(&status[0], 1);
	}
	__syncthreads();
	// local column of every nonzero; thread per row, which also notes where the row keeps its diagonal entry (the last
	// one if it is stored more than once, as the reference's loop would use it)
	for (int i = tid; i < r1 - r0; i += WIN_THREADS) {
		const long long q0 = static_cast<long long>(rowptr[r0 + i]), q1 = static_cast<long long>(rowptr[r0 + i + 1]);
		unsigned dk = WIN_NO_DIAG;
		for (long long q = q0; q < q1; ++q) {
			const int c = col[q];
			int s = 0;
			while (s + 1 < nseg && seg_start[s + 1] <= c)
				++s;
			lcol[q] = static_cast<uint16_t>(seg_off[s] + (c - seg_start[s]));
			if (c == r0 + i)
				dk = static_cast<unsigned>(q - q0);
		}
		meta[(r1 - r0) + 1 + i] = static_cast<uint16_t>(dk);
	}
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

// ------------------------------------------------------------------------------------------------ launch

struct spmv_config {
	int nstage, threads, cap, rcap, xcap, acap, grid;
	size_t smem;
};

static int env_int(const char * name, int dflt) {
	const char * v = std::getenv(name);
	return v ? std::atoi(v) : dflt;
}

static spmv_config configure(const fsb_ctx_s * c, const csr_block & B, bool window, bool staged_operand, bool vd = false) {
	static const int env_stages = env_int("FSB_SPMV_STAGES", 0);
	static const int env_threads = env_int("FSB_SPMV_THREADS", 0);
	static const int env_ctas = env_int("FSB_SPMV_CTAS_PER_SM", 0);
	spmv_config k{};
	k.cap = ((B.max_blk_nnz + 16 + 7) / 8) * 8; // + alignment slack on both ends
	if (vd)
		k.cap = B.max_blk_pnnz; // the padded stream: every block starts and ends on a 16-slot boundary
	if (k.cap < 64)
		k.cap = 64;
	k.rcap = (((vd ? 3 : window ? 2 : 1) * B.max_blk_rows + 1 + 16 + 7) / 8) * 8;
	k.xcap = window ? ((B.win_xcap + 7) / 8) * 8 : 0;
	k.acap = (window && staged_operand) ? ((B.max_blk_rows + 2 + 7) / 8) * 8 : 0;
	const size_t stage_bytes = window ? static_cast<size_t>(k.cap) * (vd ? 3 : 10) + static_cast<size_t>(k.xcap + k.acap) * 8 + static_cast<size_t>(k.rcap) * 2
	                                  : static_cast<size_t>(k.cap) * 12 + static_cast<size_t>(k.rcap) * (B.wide ? 8 : 4);
	// consumer threads: one per row of a block, at most 512
	int consumers = env_threads > 0 ? env_threads : (c->spmv_threads > 0 ? c->spmv_threads : 512);
	consumers = std::min(consumers, 512);
	while (consumers > 64 && consumers / 2 >= B.max_blk_rows)
		consumers /= 2;
	// dictionary stream: the pass is latency-bound (bytes in flight), not HBM-bound: short rows run two per thread, which
	// leaves room for a third resident CTA (measured: 0.153 ms against 0.203 ms on 7-pt 256^3)
	if (vd && env_threads <= 0 && c->spmv_threads <= 0 && B.max_blk_rows > 0 && B.max_blk_nnz / B.max_blk_rows < 16)
		while (consumers > 64 && consumers >= B.max_blk_rows)
			consumers /= 2;
	k.threads = consumers + 32;
	// Measured on B200 (profiles/r1_spmv_sweep.txt): throughput follows the number of resident
	// consumer threads (>= 1024 per SM saturates HBM); a second stage only pays when it does not
	// cost resident CTAs.
	const size_t budget = 206 * 1024; // of the 227 KB an SM offers: static scratch (<= 8.5 KB) and 1 KB reserved per CTA come on top
	const size_t ctas_for_1024 = (1024 + consumers - 1) / consumers;
	k.nstage = (2 * stage_bytes * ctas_for_1024 <= budget) ? 2 : 1;
	if (window) // consumers read shared memory only: two stages per CTA, as many CTAs as fit (r2 sweep)
		k.nstage = 2 * stage_bytes * 2 <= budget ? 2 : 1;
	if (env_stages > 0)
		k.nstage = std::min(env_stages, SPMV_MAX_STAGES);
	while (k.nstage > 1 && stage_bytes * k.nstage > budget)
		--k.nstage;
	k.smem = stage_bytes * k.nstage;
	k.grid = env_ctas; // optional cap on resident CTAs per SM; the launcher asks the occupancy API
	return k;
}

// resident CTAs per SM of one kernel instantiation for (threads, smem), cached per kernel (all instantiations share
// one function-pointer type, so the cache is keyed by the pointer)
template<class Kernel>
static int resident_ctas(Kernel kern, const spmv_config & k, const char * what) {
	struct entry {
		int threads = -1, ctas = 0;
		size_t smem = 0;
	};
	static std::unordered_map<const void *, entry> cache;
	static std::mutex guard; // the ranks of an in-process group launch from their own threads
	std::lock_guard<std::mutex> lock(guard);
	entry & e = cache[reinterpret_cast<const void *>(kern)];
	if (e.threads != k.threads || e.smem != k.smem) {
		// 227 KB per CTA in all; the variants with boundary phases keep up to 8.5 KB of it as static scratch
		cudaFuncAttributes fa{};
		FSB_CUDA(cudaFuncGetAttributes(&fa, kern));
		const int dyn_max = 227 * 1024 - static_cast<int>(fa.sharedSizeBytes);
		FSB_REQUIRE(static_cast<long long>(k.smem) <= dyn_max, "spmv: row block does not fit shared memory");
		FSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
		int nb = 0;
		FSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, k.threads, k.smem));
		FSB_REQUIRE(nb >= 1, "spmv: kernel does not fit on an SM");
		e.threads = k.threads;
		e.smem = k.smem;
		e.ctas = nb;
		if (std::getenv("FSB_SPMV_DEBUG"))
			fprintf(stderr, "[fsb] spmv %s config: threads %d stages %d smem %zu B cap %d xcap %d -> %d CTAs/SM\n", what, k.threads,
			        k.nstage, k.smem, k.cap, k.xcap, nb);
	}
	return e.ctas;
}

template<class Kernel>
static int launch_kernel(Kernel kern, const spmv_args & a, const spmv_config & k, cudaStream_t s, const char * what) {
	int ctas = resident_ctas(kern, k, what);
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
	if (b.halo) // every CTA pushes its share of the boundary entries while its first stage is in flight: a few stores and
		b.push_parts = grid; // one system fence each, so no CTA starts its row blocks later than the others
	if (b.meet)
		b.meet->rendezvous(); // after the occupancy queries above, which may load the kernel
	launch_dependent(kern, dim3(grid), dim3(k.threads), k.smem, s, b);
	if (b.meet)
		b.meet->rendezvous(); // ... and nobody goes on before every rank's kernel is launched
	return grid;
}

// gather format: pick the instantiation
template<class OffT, int NSTAGE>
static int launch_stream(const spmv_args & a, const spmv_config & k, const spmv_call & call, bool rowlist, cudaStream_t s) {
	const bool dot = call.dot_u != nullptr, acc = call.accumulate, halo = a.halo != nullptr;
	const int jac = call.jacobi;
#define FSB_SPMV_LAUNCH(ACC, DOT, ROWLIST, HALO, JAC) \
	return launch_kernel(spmv_stream_kernel<OffT, NSTAGE, ACC, DOT, ROWLIST, HALO, JAC>, a, k, s, "gather")
	if (jac != 0) {
		FSB_REQUIRE(!dot && !acc && !rowlist, "spmv: a Jacobi sweep is a plain pass over the owned-column block");
		if (jac == 2) {
			FSB_REQUIRE(!halo, "spmv: jacobi mode 2 is the two-launch path");
			FSB_SPMV_LAUNCH(false, false, false, false, 2);
		}
		if (halo)
			FSB_SPMV_LAUNCH(false, false, false, true, 1);
		FSB_SPMV_LAUNCH(false, false, false, false, 1);
	}
	if (halo) {
		FSB_REQUIRE(!acc && !rowlist, "spmv: the fused ghost exchange runs with the owned-column block");
		if (dot)
			FSB_SPMV_LAUNCH(false, true, false, true, 0);
		FSB_SPMV_LAUNCH(false, false, false, true, 0);
	}
	if (rowlist && !acc) { // structured-grid operator: rows scatter to storage offsets, plain assignment
		if constexpr (sizeof(OffT) == 4) {
			if (dot)
				FSB_SPMV_LAUNCH(false, true, true, false, 0);
			FSB_SPMV_LAUNCH(false, false, true, false, 0);
		}
		else
			throw error(FSB_ERR_STATE, "spmv: row-list assignment is built for int32 offsets only");
	}
	if (rowlist) { // off-process block: always accumulates
		if constexpr (sizeof(OffT) == 4) {
			if (dot)
				FSB_SPMV_LAUNCH(true, true, true, false, 0);
			FSB_SPMV_LAUNCH(true, false, true, false, 0);
		}
		else
			throw error(FSB_ERR_STATE, "spmv: the off-process block uses int32 offsets");
	}
	if (acc) {
		if (dot)
			FSB_SPMV_LAUNCH(true, true, false, false, 0);
		FSB_SPMV_LAUNCH(true, false, false, false, 0);
	}
	if (dot)
		FSB_SPMV_LAUNCH(false, true, false, false, 0);
	FSB_SPMV_LAUNCH(false, false, false, false, 0);
#undef FSB_SPMV_LAUNCH
}

template<int NSTAGE, bool VD>
static int launch_window(const spmv_args & a, const spmv_config & k, const spmv_call & call, cudaStream_t s) {
	const bool dot = call.dot_u != nullptr, halo = a.halo != nullptr;
#define FSB_SPMV_LAUNCH(DOT, HALO, JAC) \
	return launch_kernel(spmv_window_kernel<NSTAGE, DOT, HALO, JAC, VD>, a, k, s, VD ? "window+dictionary" : "window")
	if (call.jacobi == 1) {
		if (halo)
			FSB_SPMV_LAUNCH(false, true, 1);
		FSB_SPMV_LAUNCH(false, false, 1);
	}
	if (call.jacobi == 2)
		FSB_SPMV_LAUNCH(false, false, 2);
	if (halo) {
		if (dot)
			FSB_SPMV_LAUNCH(true, true, 0);
		FSB_SPMV_LAUNCH(false, true, 0);
	}
	if (dot)
		FSB_SPMV_LAUNCH(true, false, 0);
	FSB_SPMV_LAUNCH(false, false, 0);
#undef FSB_SPMV_LAUNCH
}

// y (+)= B x on stream s (see spmv_call in fsb_internal.h).  Returns the number of CTAs launched (= dot partials written).
int launch_spmv(fsb_ctx_s * c, const csr_block & B, const spmv_call & call, cudaStream_t s) {
	if (B.n_blk == 0)
		return 0;
	const bool rowlist = B.row_ids != nullptr;
	// the window format serves plain passes over the owned-column block with x vectors that may be over-read by one entry
	const bool window = B.lcol != nullptr && !rowlist && !call.accumulate && call.x_padded;
	// row-aligned operand of the window kernel's epilogue: u == x comes out of the x window for free; otherwise it is staged
	// with the block unless that costs a resident CTA (then: plain loads)
	const double * aux = !window ? nullptr : (call.jacobi == 1 ? call.jacobi_b : (call.dot_u && call.dot_u != call.y ? call.dot_u : nullptr));
	int aux_mode = aux ? 1 : 0;
	if (aux && call.jacobi == 0 && aux == call.x)
		aux_mode = 2;
	static const bool dict_allowed = env_int("FSB_SPMV_DICT", 1) != 0;
	const bool vd = window && B.vidx != nullptr && dict_allowed && c->spmv_dictionary;
	spmv_config k = configure(c, B, window, aux_mode == 1, vd);
	if (aux_mode == 1) {
		const spmv_config lean = configure(c, B, window, false, vd);
		const size_t per_sm = 227 * 1024;
		if (per_sm / (lean.smem + 1536) > per_sm / (k.smem + 1536) || lean.nstage > k.nstage) {
			k = lean;
			aux_mode = 3;
		}
	}
	FSB_REQUIRE(k.smem <= 225 * 1024, "spmv: row block does not fit shared memory");
	spmv_args a{};
	a.rowptr = B.rowptr;
	a.col = B.col;
	a.val = B.val;
	a.desc = static_cast<const blk_desc *>(B.blk_desc);
	a.row_ids = B.row_ids;
	a.x = call.x;
	a.y = call.y;
	a.u = call.dot_u;
	a.partials = call.partials;
	a.fold_extra = call.partial_offset;
	a.sched = c->d_sched;
	if (call.dot_u && call.fold) {
		fill_red_out(c, *call.fold, a.result);
		a.xr = c->d_xrank;
	}
	a.n_blk = B.n_blk;
	a.static_sched = c->reproducible ? 1 : 0;
	a.cap = k.cap;
	a.rcap = k.rcap;
	a.xcap = k.xcap;
	a.aux = aux;
	a.acap = k.acap;
	a.aux_mode = aux_mode;
	a.acc_continue = call.acc_continue ? 1 : 0;
	a.jb = call.jacobi_b;
	a.omega = call.omega;
	a.lcol = B.lcol;
	a.vidx = vd ? B.vidx : nullptr;
	a.vdict = B.vdict;
	a.plcol = B.plcol;
	a.pmeta = B.pmeta;
	a.rp16 = B.rp16;
	a.segs = static_cast<const x_segment *>(B.segs);
	if (call.halo) {
		const fsb_parcsr_s * A = call.halo;
		FSB_REQUIRE(B.has_offd_map && !B.has_giant_rows, "spmv: matrix is not prepared for the fused ghost exchange");
		a.halo = static_cast<const halo_dev *>(A->halo_p2p);
		a.epoch = call.epoch;
		a.push_parts = A->halo_push_ctas;
		a.o_rowptr = static_cast<const int32_t *>(A->offd.rowptr);
		a.o_col = A->offd.col;
		a.o_val = A->offd.val;
		a.o_rows = A->offd.row_ids;
		a.n_owned = A->n_local;
	}
	a.meet = (c->boot && (call.halo || a.xr)) ? c->boot.get() : nullptr; // the kernel waits for the peers' kernels of this step
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	a.tl = timeline_slot(c, call.halo ? TL_KIND_SPMV_FUSED : (rowlist ? TL_KIND_SPMV_OFFD : TL_KIND_SPMV) + (call.jacobi ? 8 : 0));
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
	int grid;
	if (window && vd)
		grid = k.nstage >= 4   ? launch_window<4, true>(a, k, call, s)
		       : k.nstage == 3 ? launch_window<3, true>(a, k, call, s)
		       : k.nstage == 2 ? launch_window<2, true>(a, k, call, s)
		                       : launch_window<1, true>(a, k, call, s);
	else if (window)
		grid = k.nstage >= 4   ? launch_window<4, false>(a, k, call, s)
		       : k.nstage == 3 ? launch_window<3, false>(a, k, call, s)
		       : k.nstage == 2 ? launch_window<2, false>(a, k, call, s)
		                       : launch_window<1, false>(a, k, call, s);
	else if (B.wide)
		grid = k.nstage >= 2 ? launch_stream<long long, 2>(a, k, call, rowlist, s) : launch_stream<long long, 1>(a, k, call, rowlist, s);
	else
		grid = k.nstage >= 2 ? launch_stream<int, 2>(a, k, call, rowlist, s) : launch_stream<int, 1>(a, k, call, rowlist, s);
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
	if (c->boot && c->d_xrank && c->boot->in_process()) {
		preload_kernel(fold_partials_kernel);
		c->boot->rendezvous();
	}
	fold_partials_kernel<<<1, 256, 0, c->stream>>>(c->d_partials, n_partials, r, c->d_xrank);
	FSB_CUDA(cudaGetLastError());
	if (c->boot && c->d_xrank && c->boot->in_process())
		c->boot->rendezvous();
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
}

// ------------------------------------------------------------------------------------------------ row blocks

// Row-block work descriptors.  host_rowptr == nullptr: uniform blocks (the caller left the row
// width bound in B.max_blk_nnz); otherwise greedy packing by nnz.
void build_blocks(fsb_ctx_s * c, csr_block & B, const std::vector<int64_t> * host_rowptr) {
	constexpr int CAP = 4096 - 16; // nnz staged per row block (48 KB per stage)
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
				B.has_giant_rows = true;
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
		rows = std::max(32, std::min(ROWS_MAX, rows));
		int pow2 = 32;
		while (pow2 * 2 <= rows)
			pow2 *= 2;
		rows = pow2;
		// wide rows are headed for the window format, which runs best with ~17 KB stages, two per CTA and five CTAs per
		// SM (27-point: 64 rows; measured in profiles/r2_spmv_window_sweep.txt); short rows keep 512-row blocks
		if (env_int("FSB_SPMV_WINDOW", 1) != 0 && !B.row_ids) {
			if (!B.dict_keys.empty()) {
				// dictionary stream (3 B per slot, rows padded to 8 slots): a stage is mostly x window, which costs less per
				// row the more rows a block has -- 4096 slots per block (profiles/r2_spmv_dictionary_sweep.txt)
				rows = std::max(32, std::min(ROWS_MAX, 4096 / ((width + 7) / 8 * 8)));
				int p2 = 32;
				while (p2 * 2 <= rows)
					p2 *= 2;
				rows = p2;
			}
			else if (width >= 16)
				rows = std::min(rows, 64);
			else if (rows == ROWS_MAX)
				rows = 448; // two stages of a 7-point block + its x segments, two CTAs per SM, in every kernel variant
		}
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
		rows = std::min(rows, ROWS_MAX);
		while (rows > 1 && static_cast<long long>(rows) * width > CAP)
			rows /= 2;
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

// Window format for the block (spmv_window_kernel): built when every row block's x footprint is a few contiguous
// pieces that cost less shared-memory fill than they save in gathers.  n_cols: entries of x the block may index.
void build_window_format(fsb_ctx_s * c, csr_block & B, int64_t n_cols) {
	static const int env_window = env_int("FSB_SPMV_WINDOW", -1); // 0 never, 1 when it pays (default), 2 whenever possible
	const int mode = env_window >= 0 ? env_window : 1;
	if (mode == 0 || B.n_blk == 0 || B.row_ids || B.has_giant_rows || B.max_blk_nnz > WIN_THREADS * WIN_ITEMS ||
	    B.max_blk_rows > ROWS_MAX || n_cols >= (1LL << 31) - 2)
		return;
	// worth it when the staged x entries cost less than ~60 % of the block's matrix bytes (27-point: 27 %, 7-point: 45 %;
	// both measured faster than the gather kernel, profiles/r2_spmv_window_sweep.txt)
	static const int env_pct = env_int("FSB_SPMV_WINDOW_PCT", 60);
	const int xcap_limit = mode == 2 ? 20000 : static_cast<int>(static_cast<long long>(B.max_blk_nnz) * 10 * env_pct / 100 / 8);
	uint16_t *lcol = nullptr, *rp16 = nullptr;
	x_segment * segs = nullptr;
	int * status = nullptr;
	FSB_CUDA(cudaMalloc(&lcol, (static_cast<size_t>(B.nnz) + 32) * sizeof(uint16_t)));
	FSB_CUDA(cudaMalloc(&rp16, (2 * static_cast<size_t>(B.n_rows) + B.n_blk + 32) * sizeof(uint16_t)));
	FSB_CUDA(cudaMalloc(&segs, static_cast<size_t>(B.n_blk) * WIN_MAXSEG * sizeof(x_segment)));
	FSB_CUDA(cudaMalloc(&status, 2 * sizeof(int)));
	FSB_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int), c->stream));
	if (B.wide)
		build_window_kernel<long long><<<B.n_blk, WIN_THREADS, 0, c->stream>>>(
			static_cast<const long long *>(B.rowptr), B.col, B.blk_row, static_cast<int>(n_cols), xcap_limit, lcol, rp16, segs, status);
	else
		build_window_kernel<int><<<B.n_blk, WIN_THREADS, 0, c->stream>>>(static_cast<const int *>(B.rowptr), B.col, B.blk_row,
		                                                                 static_cast<int>(n_cols), xcap_limit, lcol, rp16, segs,
		                                                                 status);
	FSB_CUDA(cudaGetLastError());
	int h[2] = {0, 0};
	FSB_CUDA(cudaMemcpyAsync(h, status, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(status);
	if (h[0] != 0) { // some block does not fit: the matrix stays in the gather format
		cudaFree(lcol);
		cudaFree(rp16);
		cudaFree(segs);
		return;
	}
	B.lcol = lcol;
	B.rp16 = rp16;
	B.segs = segs;
	B.win_xcap = h[1];
	if (std::getenv("FSB_SPMV_DEBUG"))
		fprintf(stderr, "[fsb] window format: %d blocks, <= %d x entries staged per block (limit %d)\n", B.n_blk, h[1], xcap_limit);
}

// ------------------------------------------------------------------------------------------------ value dictionary

constexpr unsigned DICT_SLOTS = 4096; // open-addressing table of value bit patterns; gives up beyond 256 distinct ones
constexpr unsigned long long DICT_EMPTY = ~0ULL; // (a NaN pattern; a matrix that holds it gets no dictionary)

// state[0] = distinct values seen, state[1] = 1: more than 256 (or the table cannot hold them): no dictionary
__global__ void __launch_bounds__(256) collect_values_kernel(const double * __restrict__ val, long long nnz, unsigned long long * table,
                                                            int * state) {
	volatile int * given_up = state + 1;
	unsigned long long last = DICT_EMPTY;
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	int trips = 0;
	for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += stride) {
		if ((++trips & 63) == 0 && *given_up)
			return;
		const unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(val[i]));
		if (key == last)
			continue;
		last = key;
		if (key == DICT_EMPTY || *given_up) {
			*given_up = 1;
			return;
		}
		unsigned h = static_cast<unsigned>((key * 0x9E3779B97F4A7C15ULL) >> 40) & (DICT_SLOTS - 1);
		for (unsigned probes = 0;; ++probes) {
			if (probes >= DICT_SLOTS) {
				*given_up = 1;
				return;
			}
			const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(table + h);
			if (cur == key)
				break;
			if (cur == DICT_EMPTY) {
				const unsigned long long prev = atomicCAS(table + h, DICT_EMPTY, key);
				if (prev == DICT_EMPTY) {
					if (atomicAdd(state, 1) + 1 > 256)
						*given_up = 1;
					break;
				}
				if (prev == key)
					break;
			}
			h = (h + 1) & (DICT_SLOTS - 1);
		}
	}
}

// slots of every row block in the dictionary stream: rows padded to multiples of 8, blocks to multiples of 16
template<class OffT>
__global__ void padded_sizes_kernel(const OffT * __restrict__ rowptr, const int32_t * __restrict__ blk_row, int n_blk,
                                    long long * __restrict__ slots) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blk)
		return;
	long long t = 0;
	for (int r = blk_row[b]; r < blk_row[b + 1]; ++r)
		t += (static_cast<long long>(rowptr[r + 1]) - static_cast<long long>(rowptr[r]) + 7) & ~7LL;
	slots[b] = (t + 15) & ~15LL;
}

// The dictionary stream of one row block per CTA: value index (position of val[q] among the sorted distinct bit patterns)
// and local column of every nonzero, each row starting at a multiple of 8 slots (padding: index 0, column 0), and the
// block's row meta [nrows + 1 slot offsets | nrows row lengths | nrows diagonal positions].
template<class OffT>
__global__ void __launch_bounds__(256) build_padded_kernel(const OffT * __restrict__ rowptr, const int32_t * __restrict__ blk_row,
                                                          const double * __restrict__ val, const uint16_t * __restrict__ lcol,
                                                          const uint16_t * __restrict__ rp16, const unsigned long long * __restrict__ keys,
                                                          int n_keys, const long long * __restrict__ zp0,
                                                          const long long * __restrict__ slots, blk_desc * __restrict__ desc,
                                                          uint8_t * __restrict__ pidx, uint16_t * __restrict__ plcol,
                                                          uint16_t * __restrict__ pmeta) {
	using Scan = cub::BlockScan<int, 256>;
	__shared__ typename Scan::TempStorage tmp;
	__shared__ unsigned long long sk[256];
	const int b = blockIdx.x, tid = threadIdx.x;
	const int r0 = blk_row[b], nrows = blk_row[b + 1] - r0;
	sk[tid] = tid < n_keys ? keys[tid] : DICT_EMPTY;
	uint16_t * meta = pmeta + 3 * static_cast<size_t>(r0) + b;
	const uint16_t * old_diag = rp16 + 2 * static_cast<size_t>(r0) + b + nrows + 1;
	long long q0[2] = {0, 0};
	int len[2] = {0, 0}, plen[2] = {0, 0}, off[2] = {0, 0};
#pragma unroll
	for (int k = 0; k < 2; ++k) {
		const int i = 2 * tid + k;
		if (i < nrows) {
			q0[k] = static_cast<long long>(rowptr[r0 + i]);
			len[k] = static_cast<int>(static_cast<long long>(rowptr[r0 + i + 1]) - q0[k]);
			plen[k] = (len[k] + 7) & ~7;
		}
	}
	int total = 0;
	Scan(tmp).ExclusiveSum(plen, off, total);
	__syncthreads(); // sk
	const long long z = zp0[b];
#pragma unroll
	for (int k = 0; k < 2; ++k) {
		const int i = 2 * tid + k;
		if (i >= nrows)
			continue;
		meta[i] = static_cast<uint16_t>(off[k]);
		meta[nrows + 1 + i] = static_cast<uint16_t>(len[k]);
		meta[2 * nrows + 1 + i] = old_diag[i];
		for (int j = 0; j < plen[k]; ++j) {
			unsigned vi = 0, lc = 0;
			if (j < len[k]) {
				const unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(val[q0[k] + j]));
				int lo = 0, hi = n_keys - 1;
				while (lo < hi) {
					const int mid = (lo + hi) >> 1;
					if (sk[mid] < key)
						lo = mid + 1;
					else
						hi = mid;
				}
				vi = static_cast<unsigned>(lo);
				lc = lcol[q0[k] + j];
			}
			pidx[z + off[k] + j] = static_cast<uint8_t>(vi);
			plcol[z + off[k] + j] = static_cast<uint16_t>(lc);
		}
	}
	const int padded = static_cast<int>(slots[b]);
	for (int j = total + tid; j < padded; j += 256) {
		pidx[z + j] = 0;
		plcol[z + j] = 0;
	}
	if (tid == 0) {
		meta[nrows] = static_cast<uint16_t>(total);
		desc[b].zp0 = z;
		desc[b].pnnz = padded;
		desc[b].rpp0 = 3 * r0 + b;
	}
}

// Does the block hold at most 256 distinct values?  Called right after the values are on the device and BEFORE the row
// blocks are cut (build_blocks sizes them for the stream the kernel will read): leaves the sorted bit patterns in
// B.dict_keys, or nothing.  Distinct = distinct bit patterns, so -0.0, NaN payloads and denormals survive: the kernel
// multiplies exactly the doubles the matrix holds.
void probe_value_dictionary(fsb_ctx_s * c, csr_block & B) {
	B.dict_keys.clear();
	if (B.nnz == 0 || B.row_ids || !c->spmv_dictionary || env_int("FSB_SPMV_DICT", 1) == 0 || env_int("FSB_SPMV_WINDOW", 1) == 0)
		return;
	unsigned long long * table = nullptr;
	int * state = nullptr;
	FSB_CUDA(cudaMalloc(&table, DICT_SLOTS * sizeof(unsigned long long)));
	FSB_CUDA(cudaMalloc(&state, 2 * sizeof(int)));
	FSB_CUDA(cudaMemsetAsync(table, 0xff, DICT_SLOTS * sizeof(unsigned long long), c->stream));
	FSB_CUDA(cudaMemsetAsync(state, 0, 2 * sizeof(int), c->stream));
	const int grid = static_cast<int>(std::min<long long>((B.nnz + 255) / 256, SM_COUNT * 8LL));
	collect_values_kernel<<<grid, 256, 0, c->stream>>>(B.val, B.nnz, table, state);
	FSB_CUDA(cudaGetLastError());
	std::vector<unsigned long long> h(DICT_SLOTS);
	int hs[2] = {0, 0};
	FSB_CUDA(cudaMemcpyAsync(h.data(), table, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaMemcpyAsync(hs, state, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(state);
	cudaFree(table);
	std::vector<unsigned long long> keys;
	for (unsigned long long k : h)
		if (k != DICT_EMPTY)
			keys.push_back(k);
	if (hs[1] != 0 || keys.empty() || keys.size() > 256)
		return;
	std::sort(keys.begin(), keys.end());
	B.dict_keys = keys;
}

// A block in the window format whose values are at most 256 distinct doubles (probe_value_dictionary) also gets the
// dictionary stream: one byte (value index) + 16-bit local column per nonzero (spmv_window_kernel<..., VD = true>).
// Call after build_window_format and before attach_offd_rows (the descriptors are still in row order).
void build_value_dictionary(fsb_ctx_s * c, csr_block & B) {
	const std::vector<unsigned long long> keys = B.dict_keys;
	if (!B.lcol || B.nnz == 0 || B.max_blk_rows > 512 || keys.empty())
		return;
	unsigned long long * table = nullptr;
	FSB_CUDA(cudaMalloc(&table, 256 * sizeof(unsigned long long)));
	std::vector<double> dict(256, 0.0);
	std::memcpy(dict.data(), keys.data(), keys.size() * sizeof(double));
	FSB_CUDA(cudaMemcpyAsync(table, keys.data(), keys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
	FSB_CUDA(cudaMalloc(&B.vdict, 256 * sizeof(double)));
	FSB_CUDA(cudaMemcpyAsync(B.vdict, dict.data(), 256 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	// where every row block starts in the padded stream
	long long *slots = nullptr, *zp0 = nullptr;
	FSB_CUDA(cudaMalloc(&slots, static_cast<size_t>(B.n_blk) * sizeof(long long)));
	FSB_CUDA(cudaMalloc(&zp0, static_cast<size_t>(B.n_blk) * sizeof(long long)));
	if (B.wide)
		padded_sizes_kernel<long long><<<(B.n_blk + 255) / 256, 256, 0, c->stream>>>(static_cast<const long long *>(B.rowptr), B.blk_row,
		                                                                              B.n_blk, slots);
	else
		padded_sizes_kernel<int><<<(B.n_blk + 255) / 256, 256, 0, c->stream>>>(static_cast<const int *>(B.rowptr), B.blk_row, B.n_blk, slots);
	FSB_CUDA(cudaGetLastError());
	const auto pol = thrust::cuda::par.on(c->stream);
	thrust::exclusive_scan(pol, slots, slots + B.n_blk, zp0);
	const long long most = thrust::reduce(pol, slots, slots + B.n_blk, 0LL, thrust::maximum<long long>());
	long long last[2] = {0, 0};
	FSB_CUDA(cudaMemcpyAsync(&last[0], zp0 + B.n_blk - 1, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaMemcpyAsync(&last[1], slots + B.n_blk - 1, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	const long long total = last[0] + last[1];
	FSB_CUDA(cudaMalloc(&B.vidx, static_cast<size_t>(total) + 64));
	FSB_CUDA(cudaMalloc(&B.plcol, (static_cast<size_t>(total) + 64) * sizeof(uint16_t)));
	FSB_CUDA(cudaMalloc(&B.pmeta, (3 * static_cast<size_t>(B.n_rows) + B.n_blk + 32) * sizeof(uint16_t)));
	auto * desc = static_cast<blk_desc *>(B.blk_desc);
	if (B.wide)
		build_padded_kernel<long long><<<B.n_blk, 256, 0, c->stream>>>(static_cast<const long long *>(B.rowptr), B.blk_row, B.val, B.lcol,
		                                                               B.rp16, table, static_cast<int>(keys.size()), zp0, slots, desc,
		                                                               B.vidx, B.plcol, B.pmeta);
	else
		build_padded_kernel<int><<<B.n_blk, 256, 0, c->stream>>>(static_cast<const int *>(B.rowptr), B.blk_row, B.val, B.lcol, B.rp16, table,
		                                                         static_cast<int>(keys.size()), zp0, slots, desc, B.vidx, B.plcol,
		                                                         B.pmeta);
	FSB_CUDA(cudaGetLastError());
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(table);
	cudaFree(slots);
	cudaFree(zp0);
	B.n_dict = static_cast<int>(keys.size());
	B.max_blk_pnnz = static_cast<int>(most);
	B.dict_slots = total;
	if (std::getenv("FSB_SPMV_DEBUG"))
		fprintf(stderr, "[fsb] value dictionary: %d distinct values; %lld slots for %lld nonzeros (3 B each), <= %d per row block\n",
		        B.n_dict, total, static_cast<long long>(B.nnz), B.max_blk_pnnz);
}

// Fused ghost exchange: tell every row block of the owned-column block which rows of the off-process block it owns,
// and move the blocks that own any behind all interior blocks (they wait for the neighbours' entries).
void attach_offd_rows(fsb_ctx_s * c, csr_block & D, const csr_block & O) {
	D.has_offd_map = false;
	if (D.n_blk == 0 || D.row_ids || D.has_giant_rows || D.max_blk_rows > ROWS_MAX)
		return;
	auto * desc = static_cast<blk_desc *>(D.blk_desc);
	if (O.n_rows > 0) {
		FSB_REQUIRE(!O.wide && O.row_ids, "offd block: int32 offsets over a compressed row list expected");
		attach_offd_kernel<<<(D.n_blk + 255) / 256, 256, 0, c->stream>>>(desc, D.n_blk, O.row_ids, static_cast<int>(O.n_rows));
		FSB_CUDA(cudaGetLastError());
		std::vector<blk_desc> h(static_cast<size_t>(D.n_blk));
		FSB_CUDA(cudaMemcpyAsync(h.data(), desc, h.size() * sizeof(blk_desc), cudaMemcpyDeviceToHost, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		std::stable_partition(h.begin(), h.end(), [](const blk_desc & d) { return d.ocnt == 0; });
		FSB_CUDA(cudaMemcpyAsync(desc, h.data(), h.size() * sizeof(blk_desc), cudaMemcpyHostToDevice, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
	}
	D.has_offd_map = true;
}

} // namespace fsb
