// Parallel CSR matrix: construction (ghost discovery, diag/offd split, halo plan),
// synthetic stencil generators, SpMV orchestration with halo overlap, Jacobi.
//
// Reference: topo::csr::color() / init_mats() (topo/csr.hh:482-618), the ghost copy plan
// (topo/csr.hh:116-187,277-304), mat::parcsr_ops::spmv (matrices/parcsr.hh:61-91),
// mg::bound_jacobi::relax (solvers/mg/jacobi.hh:44-93), Dinv() (util/test/mesh.hh:123-140).
#include <algorithm>
#include <cstdlib>
#include <cub/cub.cuh>
#include <numeric>
#include <thrust/iterator/transform_iterator.h>

#include "fsb_internal.h"
#include "setup_exchange.h"

namespace fsb {

// ------------------------------------------------------------------ helpers

template<class T>
static T * dev_alloc(size_t n) {
	T * p = nullptr;
	FSB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
	return p;
}

template<class T>
static T * dev_upload(fsb_ctx_s * c, const std::vector<T> & h, size_t pad = 0) {
	T * p = dev_alloc<T>(h.size() + pad);
	if (!h.empty())
		FSB_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
	return p;
}

static void free_block(csr_block & B) {
	cudaFree(B.rowptr);
	cudaFree(B.col);
	cudaFree(B.val);
	cudaFree(B.blk_row);
	cudaFree(B.blk_desc);
	cudaFree(B.row_ids);
	cudaFree(B.lcol);
	cudaFree(B.rp16);
	cudaFree(B.segs);
	cudaFree(B.vidx);
	cudaFree(B.plcol);
	cudaFree(B.pmeta);
	cudaFree(B.vdict);
	B = csr_block{};
}

// upload a host CSR block (int64 rowptr) choosing 32- or 64-bit offsets
static void upload_block(fsb_ctx_s * c, csr_block & B, const std::vector<int64_t> & rowptr,
                         const std::vector<int32_t> & col, const std::vector<double> & val,
                         const std::vector<int32_t> * row_ids) {
	B.n_rows = static_cast<int64_t>(rowptr.size()) - 1;
	B.nnz = rowptr.empty() ? 0 : rowptr.back();
	B.wide = B.nnz >= (1LL << 31) - 16;
	if (B.wide) {
		B.rowptr = dev_upload<int64_t>(c, rowptr, 8);
	}
	else {
		std::vector<int32_t> rp32(rowptr.begin(), rowptr.end());
		B.rowptr = dev_upload<int32_t>(c, rp32, 8);
		FSB_CUDA(cudaStreamSynchronize(c->stream)); // rp32 is a temporary
	}
	B.col = dev_upload<int32_t>(c, col, 16);
	B.val = dev_upload<double>(c, val, 16);
	if (row_ids)
		B.row_ids = dev_upload<int32_t>(c, *row_ids);
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	if (B.n_rows > 0) {
		probe_value_dictionary(c, B);
		build_blocks(c, B, &rowptr);
	}
}

// ------------------------------------------------------------------ halo plan (general)

// Exchange the ghost lists: every rank tells each owner which of its entries it needs.
static void build_halo_plan(fsb_parcsr_s * A) {
	fsb_ctx_s * c = A->ctx;
	const int P = c->nranks, me = c->rank;
	A->nbrs.clear();
	if (P == 1)
		return;
	// recv side: ghosts are sorted by global id, owners are contiguous ranges
	std::vector<int64_t> recv_cnt(P, 0), recv_off(P, 0);
	{
		size_t g = 0;
		for (int q = 0; q < P; ++q) {
			recv_off[q] = static_cast<int64_t>(g);
			while (g < A->colmap.size() && A->colmap[g] < A->row_part[q + 1])
				++g;
			recv_cnt[q] = static_cast<int64_t>(g) - recv_off[q];
		}
	}
	// all-gather the P x P count matrix
	std::vector<int64_t> all(static_cast<size_t>(P) * P);
	c->boot->allgather(recv_cnt.data(), all.data(), P * sizeof(int64_t));

	std::vector<int64_t> send_cnt(P, 0), send_off(P, 0);
	int64_t send_total = 0;
	for (int q = 0; q < P; ++q) {
		send_cnt[q] = all[static_cast<size_t>(q) * P + me]; // what q receives from me
		send_off[q] = send_total;
		send_total += send_cnt[q];
	}
	// every rank learns which of its entries the others need: rank q's ghost list is sorted by global id, so the ids it
	// wants from me are one contiguous piece of it (owners are contiguous ranges)
	const std::vector<std::vector<int64_t>> wanted = c->boot->gatherv(A->colmap, P);
	std::vector<int64_t> asked(static_cast<size_t>(send_total));
	for (int q = 0; q < P; ++q) {
		if (q == me || send_cnt[q] == 0)
			continue;
		const std::vector<int64_t> & w = wanted[static_cast<size_t>(q)];
		const auto first = std::lower_bound(w.begin(), w.end(), A->row_part[me]);
		FSB_REQUIRE(w.end() - first >= send_cnt[q], "halo plan: inconsistent ghost counts");
		std::copy(first, first + send_cnt[q], asked.begin() + send_off[q]);
	}

	std::vector<int32_t> send_idx(static_cast<size_t>(send_total));
	bool need_pack = false;
	for (int q = 0; q < P; ++q) {
		if (q == me || (send_cnt[q] == 0 && recv_cnt[q] == 0))
			continue;
		neighbour nb{};
		nb.rank = q;
		nb.send_count = send_cnt[q];
		nb.recv_count = recv_cnt[q];
		nb.recv_offset = recv_off[q];
		nb.send_offset = send_off[q];
		bool contiguous = true;
		for (int64_t k = 0; k < send_cnt[q]; ++k) {
			const int64_t local = asked[send_off[q] + k] - A->row_begin;
			FSB_REQUIRE(local >= 0 && local < A->n_local, "halo plan: peer asked for an entry this rank does not own");
			send_idx[send_off[q] + k] = static_cast<int32_t>(local);
			if (k > 0 && send_idx[send_off[q] + k] != send_idx[send_off[q] + k - 1] + 1)
				contiguous = false;
		}
		nb.contiguous_start = (contiguous && send_cnt[q] > 0) ? send_idx[send_off[q]] : -1;
		if (!contiguous)
			need_pack = true;
		A->nbrs.push_back(nb);
	}
	A->send_total = send_total;
	A->need_pack = need_pack;
	if (need_pack) {
		A->d_send_idx = dev_upload<int32_t>(c, send_idx);
		A->d_send_buf = dev_alloc<double>(static_cast<size_t>(send_total));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
	}
	// where my entries land in each neighbour's ghost interval: ghosts are ordered by owner rank,
	// so it is the number of entries that neighbour receives from lower ranks
	std::vector<int64_t> dest_off;
	for (const neighbour & nb : A->nbrs) {
		int64_t off = 0;
		for (int s = 0; s < me; ++s)
			off += all[static_cast<size_t>(nb.rank) * P + s];
		dest_off.push_back(off);
	}
	halo_p2p_setup(A, dest_off);
}

__global__ void pack_kernel(const double * __restrict__ x, const int32_t * __restrict__ idx, double * __restrict__ buf,
                            long long n) {
	const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
		buf[i] = x[idx[i]];
}

// Ghost update of x on the communication stream; the caller orders streams around it.
static void halo_exchange_async(fsb_parcsr_s * A, fsb_vec_s * x) {
	fsb_ctx_s * c = A->ctx;
	if (!c->nccl)
		throw error(FSB_ERR_STATE, "ghost exchange: no peer-memory path for this matrix and no NCCL communicator (rank group in one process)");
	// x must be complete on the main stream before it is packed/sent
	FSB_CUDA(cudaEventRecord(c->ev_main, c->stream));
	FSB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_main, 0));
	if (A->need_pack) {
		const long long n = A->send_total;
		pack_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, c->comm_stream>>>(x->d, A->d_send_idx, A->d_send_buf, n);
		FSB_CUDA(cudaGetLastError());
		c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	}
	FSB_NCCL(ncclGroupStart());
	for (const neighbour & nb : A->nbrs) {
		if (nb.send_count > 0) {
			const double * src = (A->need_pack && nb.contiguous_start < 0) ? A->d_send_buf + nb.send_offset
			                                                                : x->d + nb.contiguous_start;
			FSB_NCCL(ncclSend(src, nb.send_count, ncclDouble, nb.rank, c->nccl_halo(), c->comm_stream));
		}
		if (nb.recv_count > 0)
			FSB_NCCL(ncclRecv(x->d + x->n_owned + nb.recv_offset, nb.recv_count, ncclDouble, nb.rank, c->nccl_halo(),
			                  c->comm_stream));
	}
	FSB_NCCL(ncclGroupEnd());
	FSB_CUDA(cudaEventRecord(c->ev_comm, c->comm_stream));
	c->stats[FSB_STAT_HALO_EXCHANGES]++;
	x->halo_valid = true;
	x->halo_for = A->id;
}

void halo_exchange(fsb_parcsr_s * A, fsb_vec_s * x) {
	fsb_ctx_s * c = A->ctx;
	flush(c);
	if (c->nranks == 1 || A->nbrs.empty())
		return;
	FSB_REQUIRE(x->n_owned == A->n_local && x->n_ghost >= A->n_ghost, "halo_exchange: vector does not match the matrix");
	FSB_REQUIRE(!A->box || (x->box && x->shape == A->shape), "halo_exchange: vector does not live on the operator's box");
	if (A->halo_p2p) {
		halo_p2p_push(A, x);
		halo_p2p_unpack(A, x);
		return;
	}
	halo_exchange_async(A, x);
	FSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
}

// ------------------------------------------------------------------ SpMV orchestration

// y = A x (+ optional fused dot).  Called by the queue flush.
void spmv_group(fsb_ctx_s * c, const pending & sp, const pending * dot) {
	fsb_parcsr_s * A = sp.A;
	fsb_vec_s *x = sp.x, *y = sp.y;
	spmv_call call;
	call.x = x->d;
	call.y = y->d;
	call.x_padded = x->padded;
	call.partials = c->d_partials;
	if (dot) {
		fsb_vec_s * other = dot->x == y ? dot->y : dot->x;
		call.dot_u = other->d; // other == y gives sum y^2
		call.x_padded = call.x_padded && other->padded; // the window kernel stages u with bulk copies as well
	}
	if (A->box) { // columns and row ids are storage offsets of the padded arrays: one launch
		if (c->nranks > 1 && !A->nbrs.empty() && !(x->halo_valid && x->halo_for == A->id)) {
			// slabs over several ranks: refresh the ghost planes of x (the neighbours' outermost dof planes) first
			halo_p2p_push(A, x);
			halo_p2p_unpack(A, x);
		}
		call.fold = dot;
		launch_spmv(c, A->diag, call, c->stream);
		y->halo_valid = false;
		return;
	}
	const bool multi = c->nranks > 1 && !A->nbrs.empty();
	static const bool debug_skip_halo = std::getenv("FSB_DEBUG_SKIP_HALO") != nullptr; // timing experiments only
	static const bool fused_allowed = !(std::getenv("FSB_FUSED_HALO") && std::atoi(std::getenv("FSB_FUSED_HALO")) == 0);
	if (multi && A->halo_p2p && A->diag.has_offd_map && A->diag.n_blk > 0 && fused_allowed && !debug_skip_halo) {
		// ONE launch: push of this rank's boundary entries, interior rows, boundary rows (wait for the neighbours'
		// entries, add the off-process part from the landing area), dot fold + all-reduce, acknowledge
		call.halo = A;
		call.epoch = ++A->halo_epoch;
		call.fold = dot;
		launch_spmv(c, A->diag, call, c->stream);
		c->stats[FSB_STAT_HALO_EXCHANGES]++;
		y->halo_valid = false;
		return;
	}
	bool waited = true, unpack = false;
	if (multi && !(x->halo_valid && x->halo_for == A->id) && !debug_skip_halo) {
		if (A->halo_p2p) {
			halo_p2p_push(A, x); // peers' pushes land while the diag block below runs
			unpack = true;
		}
		else {
			halo_exchange_async(A, x); // NCCL on the communication stream, overlaps the diag block
			waited = false;
		}
	}
	const bool has_offd = A->offd.n_blk > 0;
	// the launch that runs last folds the dot partials (no separate fold kernel)
	call.fold = has_offd ? nullptr : dot;
	const int np_diag = launch_spmv(c, A->diag, call, c->stream);
	if (unpack)
		halo_p2p_unpack(A, x);
	if (has_offd) {
		if (!waited)
			FSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
		waited = true;
		call.accumulate = true;
		call.partial_offset = np_diag;
		call.fold = dot;
		launch_spmv(c, A->offd, call, c->stream);
	}
	if (!waited) // ghosts were requested but no row uses them: still order the streams
		FSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
	if (dot && A->diag.n_blk == 0 && !has_offd)
		finalize_reduction(c, 0, *dot); // empty local matrix: publish 0
	y->halo_valid = false;
}

// ------------------------------------------------------------------ creation from host CSR

} // namespace fsb

using namespace fsb;

fsb_parcsr_s * fsb_parcsr_create_impl(fsb_ctx_s * c, int64_t n_global, const int64_t * row_part, const int64_t * rowptr,
                                      const int64_t * col, const double * val) {
	flush(c);
	const int P = c->nranks, me = c->rank;
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->id = c->next_mat_id++;
	A->n_global = n_global;
	A->row_part.assign(row_part, row_part + P + 1);
	FSB_REQUIRE(A->row_part[0] == 0 && A->row_part[P] == n_global, "parcsr: row_part must span [0, n_global]");
	A->row_begin = row_part[me];
	const int64_t row_end = row_part[me + 1];
	A->n_local = row_end - A->row_begin;
	const int64_t nnz = rowptr[A->n_local];

	// color(): ghosts = sorted unique off-rank column ids (topo/csr.hh:506-524)
	std::vector<int64_t> ghosts;
	for (int64_t k = 0; k < nnz; ++k) {
		FSB_REQUIRE(col[k] >= 0 && col[k] < n_global, "parcsr: column index out of range");
		if (col[k] < A->row_begin || col[k] >= row_end)
			ghosts.push_back(col[k]);
	}
	std::sort(ghosts.begin(), ghosts.end());
	ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
	A->n_ghost = static_cast<int64_t>(ghosts.size());
	A->colmap = ghosts;
	FSB_REQUIRE(A->n_local + A->n_ghost < (1LL << 31), "parcsr: local column space exceeds int32");

	// init_mats(): split rows, keeping the input order inside each row (topo/csr.hh:583-616)
	std::vector<int64_t> drp(A->n_local + 1, 0), orp_full(A->n_local + 1, 0);
	std::vector<int32_t> dcol, ocol;
	std::vector<double> dval, oval;
	dcol.reserve(nnz);
	dval.reserve(nnz);
	for (int64_t r = 0; r < A->n_local; ++r) {
		for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
			const int64_t cid = col[k];
			if (cid >= A->row_begin && cid < row_end) {
				dcol.push_back(static_cast<int32_t>(cid - A->row_begin));
				dval.push_back(val[k]);
			}
			else {
				const int64_t g = std::lower_bound(ghosts.begin(), ghosts.end(), cid) - ghosts.begin();
				ocol.push_back(static_cast<int32_t>(A->n_local + g));
				oval.push_back(val[k]);
			}
		}
		drp[r + 1] = static_cast<int64_t>(dcol.size());
		orp_full[r + 1] = static_cast<int64_t>(ocol.size());
	}
	// device offd block is stored over the rows that have off-process entries only
	std::vector<int64_t> orp(1, 0);
	std::vector<int32_t> orows;
	for (int64_t r = 0; r < A->n_local; ++r)
		if (orp_full[r + 1] > orp_full[r]) {
			orows.push_back(static_cast<int32_t>(r));
			orp.push_back(orp_full[r + 1]);
		}
	upload_block(c, A->diag, drp, dcol, dval, nullptr);
	build_window_format(c, A->diag, A->n_local);
	build_value_dictionary(c, A->diag);
	if (!orows.empty())
		upload_block(c, A->offd, orp, ocol, oval, &orows);
	if (P > 1)
		attach_offd_rows(c, A->diag, A->offd);
	A->d_colmap = dev_upload<int64_t>(c, A->colmap);
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	build_halo_plan(A);
	return A;
}

// ------------------------------------------------------------------ stencil generators

namespace fsb {

struct grid_desc {
	long long nx, ny, nz;
	long long row_begin, row_end; // owned global rows
	long long lower, upper; // ghost counts below / above
	long long reach; // how far (in row numbers) a stencil reaches: ghost candidates are [row_begin - reach, row_begin) and
	                 // [row_end, row_end + reach), clipped to the grid
	long long lo_base; // first candidate below
	const int * rank_lo; // exclusive scan of the "referenced" flags over the candidates below / above =
	const int * rank_hi; // position of each ghost in the sorted ghost list
	double diag, off;
	int kind; // 5, 7, 27
	int neumann; // 7-point only: Neumann closure (kind 107 of the C ABI)
};

// visit the stencil of global row g in ascending column order
template<class F>
__device__ __forceinline__ void visit_stencil(const grid_desc & G, long long g, F && f) {
	const long long i = g % G.nx, j = (g / G.nx) % G.ny, k = g / (G.nx * G.ny);
	double diag = G.diag;
	if (G.neumann) { // diagonal = (neighbours inside the box + shift) * scale = G.diag - missing * scale, as the oracle
		const int missing = (i == 0) + (i == G.nx - 1) + (j == 0) + (j == G.ny - 1) + (k == 0) + (k == G.nz - 1);
		diag = G.diag + missing * G.off;
	}
	if (G.kind == 27) {
		for (int dk = -1; dk <= 1; ++dk)
			for (int dj = -1; dj <= 1; ++dj)
				for (int di = -1; di <= 1; ++di) {
					const long long ii = i + di, jj = j + dj, kk = k + dk;
					if (ii < 0 || ii >= G.nx || jj < 0 || jj >= G.ny || kk < 0 || kk >= G.nz)
						continue;
					f(ii + G.nx * (jj + G.ny * kk), (di == 0 && dj == 0 && dk == 0) ? G.diag : G.off);
				}
	}
	else {
		if (G.kind == 7 && k > 0)
			f(g - G.nx * G.ny, G.off);
		if (j > 0)
			f(g - G.nx, G.off);
		if (i > 0)
			f(g - 1, G.off);
		f(g, diag);
		if (i < G.nx - 1)
			f(g + 1, G.off);
		if (j < G.ny - 1)
			f(g + G.nx, G.off);
		if (G.kind == 7 && k < G.nz - 1)
			f(g + G.nx * G.ny, G.off);
	}
}

__global__ void stencil_count_kernel(grid_desc G, int * __restrict__ cnt_diag, int * __restrict__ cnt_offd) {
	const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r >= G.row_end - G.row_begin)
		return;
	int d = 0, o = 0;
	visit_stencil(G, G.row_begin + r, [&](long long c, double) {
		if (c >= G.row_begin && c < G.row_end)
			++d;
		else
			++o;
	});
	cnt_diag[r] = d;
	cnt_offd[r] = o;
}

template<class OffT>
__global__ void stencil_fill_diag_kernel(grid_desc G, const OffT * __restrict__ rowptr, int32_t * __restrict__ col,
                                         double * __restrict__ val) {
	const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r >= G.row_end - G.row_begin)
		return;
	long long p = static_cast<long long>(rowptr[r]);
	visit_stencil(G, G.row_begin + r, [&](long long c, double v) {
		if (c >= G.row_begin && c < G.row_end) {
			col[p] = static_cast<int32_t>(c - G.row_begin);
			val[p] = v;
			++p;
		}
	});
}

__global__ void stencil_fill_offd_kernel(grid_desc G, const int32_t * __restrict__ row_ids, long long n_rows,
                                         const int * __restrict__ rowptr, int32_t * __restrict__ col,
                                         double * __restrict__ val) {
	const long long cr = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (cr >= n_rows)
		return;
	const long long n_owned = G.row_end - G.row_begin;
	long long p = rowptr[cr];
	visit_stencil(G, G.row_begin + row_ids[cr], [&](long long c, double v) {
		if (c < G.row_begin) {
			col[p] = static_cast<int32_t>(n_owned + G.rank_lo[c - G.lo_base]);
			val[p] = v;
			++p;
		}
		else if (c >= G.row_end) {
			col[p] = static_cast<int32_t>(n_owned + G.lower + G.rank_hi[c - G.row_end]);
			val[p] = v;
			++p;
		}
	});
}

// which candidates are referenced by an owned row (only rows within `reach` of either end of the block can)
__global__ void stencil_mark_kernel(grid_desc G, int * __restrict__ flag_lo, int * __restrict__ flag_hi) {
	const long long n = G.row_end - G.row_begin;
	long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r >= 2 * G.reach)
		return;
	if (r >= G.reach) { // second half of the threads: the rows at the upper end
		r = n - 2 * G.reach + r;
		if (r < G.reach) // short block: already covered by the first half
			return;
	}
	if (r < 0 || r >= n)
		return;
	visit_stencil(G, G.row_begin + r, [&](long long c, double) {
		if (c < G.row_begin)
			flag_lo[c - G.lo_base] = 1;
		else if (c >= G.row_end)
			flag_hi[c - G.row_end] = 1;
	});
}

__global__ void iota_kernel(int32_t * p, long long n) {
	const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
		p[i] = static_cast<int32_t>(i);
}

__global__ void colmap_kernel(grid_desc G, const int * __restrict__ flag_lo, const int * __restrict__ flag_hi,
                              long long * __restrict__ colmap) {
	const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < G.reach) {
		if (flag_lo[i])
			colmap[G.rank_lo[i]] = G.lo_base + i;
	}
	else if (i < 2 * G.reach) {
		const long long k = i - G.reach;
		if (flag_hi[k])
			colmap[G.lower + G.rank_hi[k]] = G.row_end + k;
	}
}

struct nonzero_flag {
	__host__ __device__ bool operator()(int v) const { return v > 0; }
};

struct to_i64 {
	__host__ __device__ long long operator()(int v) const { return v; }
};

template<class In, class Out>
static void exclusive_scan(fsb_ctx_s * c, In in, Out out, long long n) {
	void * tmp = nullptr;
	size_t bytes = 0;
	FSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, c->stream));
	FSB_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
	FSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, c->stream));
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(tmp);
}

} // namespace fsb

fsb_parcsr_s * fsb_parcsr_create_stencil_impl(fsb_ctx_s * c, int kind, int64_t nx, int64_t ny, int64_t nz,
                                              double diag_shift, double scale) {
	flush(c);
	const bool neumann = kind == 107;
	if (neumann)
		kind = 7;
	FSB_REQUIRE(kind == 5 || kind == 7 || kind == 27, "stencil: kind must be 5, 7, 27 or 107");
	FSB_REQUIRE(nx > 0 && ny > 0 && nz > 0, "stencil: empty grid");
	FSB_REQUIRE(kind != 5 || nz == 1, "stencil: the 5-point operator is 2-D (nz == 1)");
	const int P = c->nranks, me = c->rank;
	const int64_t plane = nx * ny;
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->id = c->next_mat_id++;
	A->n_global = plane * nz;
	// equal contiguous row blocks, the first N % P one row longer: flecsi::util::equal_map as the reference partitions
	// rows and columns (matrices/parcsr.hh:170-172); z-slabs of whole planes whenever nz % P == 0
	A->row_part.resize(P + 1);
	{
		const int64_t base = A->n_global / P, rem = A->n_global % P;
		for (int q = 0; q <= P; ++q)
			A->row_part[q] = base * q + std::min<int64_t>(q, rem);
	}
	const int64_t reach = kind == 27 ? plane + nx + 1 : (kind == 7 ? plane : nx);
	FSB_REQUIRE(P == 1 || A->n_global / P >= reach,
	            "stencil: row blocks thinner than the stencil's reach (ghosts beyond the adjacent ranks): assemble on the host and use fsb_parcsr_create");
	A->row_begin = A->row_part[me];
	const int64_t row_end = A->row_part[me + 1];
	A->n_local = row_end - A->row_begin;
	FSB_REQUIRE(A->n_local + 2 * reach < (1LL << 31), "stencil: local column space exceeds int32");

	grid_desc G{};
	G.nx = nx;
	G.ny = ny;
	G.nz = nz;
	G.row_begin = A->row_begin;
	G.row_end = row_end;
	G.reach = reach;
	G.lo_base = std::max<int64_t>(0, A->row_begin - reach);
	G.kind = kind;
	G.neumann = neumann ? 1 : 0;
	const double center = kind == 27 ? 26.0 : (kind == 7 ? 6.0 : 4.0);
	G.diag = (center + diag_shift) * scale;
	G.off = -1.0 * scale;

	const long long n = A->n_local;
	const int T = 256;
	const int grid = static_cast<int>((n + T - 1) / T);
	// ghost discovery (topo/csr.hh:506-526): the referenced columns outside the block, sorted = flagged candidates
	int *flag_lo = nullptr, *flag_hi = nullptr, *rank_lo = nullptr, *rank_hi = nullptr;
	if (P > 1) {
		flag_lo = dev_alloc<int>(reach + 1);
		flag_hi = dev_alloc<int>(reach + 1);
		rank_lo = dev_alloc<int>(reach + 1);
		rank_hi = dev_alloc<int>(reach + 1);
		FSB_CUDA(cudaMemsetAsync(flag_lo, 0, (reach + 1) * sizeof(int), c->stream));
		FSB_CUDA(cudaMemsetAsync(flag_hi, 0, (reach + 1) * sizeof(int), c->stream));
		stencil_mark_kernel<<<static_cast<int>((2 * reach + T - 1) / T), T, 0, c->stream>>>(G, flag_lo, flag_hi);
		FSB_CUDA(cudaGetLastError());
		exclusive_scan(c, flag_lo, rank_lo, reach + 1);
		exclusive_scan(c, flag_hi, rank_hi, reach + 1);
		int counts[2] = {0, 0};
		FSB_CUDA(cudaMemcpy(&counts[0], rank_lo + reach, sizeof(int), cudaMemcpyDeviceToHost));
		FSB_CUDA(cudaMemcpy(&counts[1], rank_hi + reach, sizeof(int), cudaMemcpyDeviceToHost));
		G.lower = counts[0];
		G.upper = counts[1];
		G.rank_lo = rank_lo;
		G.rank_hi = rank_hi;
	}
	A->n_ghost = G.lower + G.upper;
	int * cnt_d = dev_alloc<int>(n + 1);
	int * cnt_o = dev_alloc<int>(n + 1);
	FSB_CUDA(cudaMemsetAsync(cnt_d + n, 0, sizeof(int), c->stream));
	FSB_CUDA(cudaMemsetAsync(cnt_o + n, 0, sizeof(int), c->stream));
	stencil_count_kernel<<<grid, T, 0, c->stream>>>(G, cnt_d, cnt_o);
	FSB_CUDA(cudaGetLastError());

	// ---- diag block
	csr_block & D = A->diag;
	D.n_rows = n;
	const int width = kind;
	// nnz upper bound decides the offset width before the scan
	D.wide = static_cast<long long>(width) * n >= (1LL << 31) - 16;
	long long nnz_d = 0;
	if (D.wide) {
		long long * rp = dev_alloc<long long>(n + 1 + 8);
		exclusive_scan(c, thrust::make_transform_iterator(cnt_d, to_i64{}), rp, n + 1);
		FSB_CUDA(cudaMemcpy(&nnz_d, rp + n, sizeof(long long), cudaMemcpyDeviceToHost));
		D.rowptr = rp;
	}
	else {
		int * rp = dev_alloc<int>(n + 1 + 8);
		exclusive_scan(c, cnt_d, rp, n + 1);
		int last = 0;
		FSB_CUDA(cudaMemcpy(&last, rp + n, sizeof(int), cudaMemcpyDeviceToHost));
		nnz_d = last;
		D.rowptr = rp;
	}
	D.nnz = nnz_d;
	D.col = dev_alloc<int32_t>(nnz_d + 16);
	D.val = dev_alloc<double>(nnz_d + 16);
	if (D.wide)
		stencil_fill_diag_kernel<long long><<<grid, T, 0, c->stream>>>(G, static_cast<long long *>(D.rowptr), D.col, D.val);
	else
		stencil_fill_diag_kernel<int><<<grid, T, 0, c->stream>>>(G, static_cast<int *>(D.rowptr), D.col, D.val);
	FSB_CUDA(cudaGetLastError());
	D.max_blk_nnz = width; // row width bound for the uniform block builder
	probe_value_dictionary(c, D);
	build_blocks(c, D, nullptr);
	build_window_format(c, D, n);
	build_value_dictionary(c, D);

	// ---- offd block over the rows that touch a neighbouring slab
	if (A->n_ghost > 0) {
		csr_block & O = A->offd;
		int32_t * iota = dev_alloc<int32_t>(n);
		iota_kernel<<<grid, T, 0, c->stream>>>(iota, n);
		int32_t * rows = dev_alloc<int32_t>(n);
		int * cnt_c = dev_alloc<int>(n + 1);
		int * d_num = dev_alloc<int>(1);
		void * tmp = nullptr;
		size_t bytes = 0, bytes2 = 0;
		auto flags = thrust::make_transform_iterator(cnt_o, nonzero_flag{});
		FSB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flags, rows, d_num, static_cast<int>(n), c->stream));
		FSB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes2, cnt_o, flags, cnt_c, d_num, static_cast<int>(n), c->stream));
		bytes = std::max(bytes, bytes2);
		FSB_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
		FSB_CUDA(cub::DeviceSelect::Flagged(tmp, bytes, iota, flags, rows, d_num, static_cast<int>(n), c->stream));
		FSB_CUDA(cub::DeviceSelect::Flagged(tmp, bytes, cnt_o, flags, cnt_c, d_num, static_cast<int>(n), c->stream));
		int n_rows = 0;
		FSB_CUDA(cudaMemcpyAsync(&n_rows, d_num, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		cudaFree(tmp);
		cudaFree(iota);
		cudaFree(d_num);
		FSB_CUDA(cudaMemsetAsync(cnt_c + n_rows, 0, sizeof(int), c->stream));
		int * rp = dev_alloc<int>(n_rows + 1 + 8);
		exclusive_scan(c, cnt_c, rp, static_cast<long long>(n_rows) + 1);
		int nnz_o = 0;
		FSB_CUDA(cudaMemcpy(&nnz_o, rp + n_rows, sizeof(int), cudaMemcpyDeviceToHost));
		cudaFree(cnt_c);
		O.n_rows = n_rows;
		O.nnz = nnz_o;
		O.wide = false;
		O.rowptr = rp;
		O.row_ids = rows;
		O.col = dev_alloc<int32_t>(nnz_o + 16);
		O.val = dev_alloc<double>(nnz_o + 16);
		if (n_rows > 0) {
			stencil_fill_offd_kernel<<<(n_rows + T - 1) / T, T, 0, c->stream>>>(G, rows, n_rows, rp, O.col, O.val);
			FSB_CUDA(cudaGetLastError());
			// a row reaches at most 9 (27-point) / 1 (7- and 5-point) columns on either side; both sides only when the
			// block is thinner than twice the reach
			const int per_side = kind == 27 ? 9 : 1;
			O.max_blk_nnz = (A->n_local < 2 * reach && G.lower && G.upper) ? 2 * per_side : per_side;
			build_blocks(c, O, nullptr);
		}
		// colmap[ghost] = global id, ascending (topo/csr.hh:524-526)
		A->d_colmap = dev_alloc<int64_t>(A->n_ghost);
		colmap_kernel<<<static_cast<int>((2 * reach + T - 1) / T), T, 0, c->stream>>>(G, flag_lo, flag_hi,
		                                                                             reinterpret_cast<long long *>(A->d_colmap));
		FSB_CUDA(cudaGetLastError());
		A->colmap.resize(A->n_ghost);
		FSB_CUDA(cudaMemcpyAsync(A->colmap.data(), A->d_colmap, A->n_ghost * sizeof(int64_t), cudaMemcpyDeviceToHost,
		                         c->stream));
	}
	FSB_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(flag_lo);
	cudaFree(flag_hi);
	cudaFree(rank_lo);
	cudaFree(rank_hi);
	cudaFree(cnt_d);
	cudaFree(cnt_o);
	if (P > 1) { // collective: every rank takes part even if it had no ghosts
		attach_offd_rows(c, A->diag, A->offd);
		build_halo_plan(A); // the copy plan (topo/csr.hh:116-187): who sends which of its entries where; whole planes for z-slabs
	}
	return A;
}

// ------------------------------------------------------------------ structured-grid (narray) operators, SURVEY 8(f) N3

// The operator's box, checked; several ranks: slabs along the LAST axis in rank order, every rank with its own padded array.
static box_shape checked_box(int dim, const int64_t * ext, const int64_t * lo, const int64_t * hi) {
	FSB_REQUIRE(dim >= 1 && dim <= 3, "box operator: dim must be 1, 2 or 3");
	box_shape b;
	for (int k = 0; k < dim; ++k) {
		FSB_REQUIRE(lo[k] >= 1 && hi[k] <= ext[k] - 1 && lo[k] <= hi[k], "box operator: the interior needs one boundary layer per side");
		b.ext[k] = ext[k];
		b.lo[k] = lo[k];
		b.n[k] = hi[k] - lo[k];
	}
	FSB_REQUIRE(b.storage() < (1LL << 31), "box operator: padded array exceeds int32 offsets");
	return b;
}

// Rows follow the dof order (x fastest), entries ascend in storage offset, every neighbour is stored: boundary layers
// hold boundary data (or, towards a neighbouring rank, ghost copies) and are read like any other entry.
// coef(at, axis, side) -> value of the entry towards the lower (side 0) / upper (side 1) neighbour; diag(at) -> centre.
template<class Coef, class Diag>
static fsb_parcsr_s * assemble_box_operator(fsb_ctx_s * c, int dim, const box_shape & b, Coef && coef, Diag && diag) {
	const int64_t n = b.dofs();
	const int64_t stride[3] = {1, b.ext[0], b.ext[0] * b.ext[1]};
	const int width = 2 * dim + 1;
	std::vector<int64_t> rp(static_cast<size_t>(n) + 1);
	std::vector<int32_t> col(static_cast<size_t>(n) * width), rows(static_cast<size_t>(n));
	std::vector<double> val(static_cast<size_t>(n) * width);
	for (int64_t i = 0; i < n; ++i) {
		const int64_t t = i / b.n[0], i0 = i - t * b.n[0], i2 = t / b.n[1], i1 = t - i2 * b.n[1];
		const int64_t at = b.origin() + i0 + stride[1] * i1 + stride[2] * i2;
		size_t k = static_cast<size_t>(i) * width;
		rp[i] = static_cast<int64_t>(k);
		rows[i] = static_cast<int32_t>(at);
		for (int a = dim - 1; a >= 0; --a) {
			col[k] = static_cast<int32_t>(at - stride[a]);
			val[k++] = coef(at, a, 0);
		}
		col[k] = static_cast<int32_t>(at);
		val[k++] = diag(at);
		for (int a = 0; a < dim; ++a) {
			col[k] = static_cast<int32_t>(at + stride[a]);
			val[k++] = coef(at, a, 1);
		}
	}
	rp[n] = n * width;
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->id = c->next_mat_id++;
	A->box = true;
	A->shape = b;
	A->n_global = A->n_local = n;
	A->row_part = {0, n};
	upload_block(c, A->diag, rp, col, val, &rows);
	if (c->nranks > 1) {
		// Slabs along the last axis: the pad plane of that axis facing rank r -+ 1 is a ghost plane -- it mirrors the
		// neighbour's outermost dof plane (whole padded plane, boundary pads of the other axes included) and is refreshed
		// before the operator reads it, FleCSI's ghost copy of an narray field.  Collective.
		const int P = c->nranks, me = c->rank, last = dim - 1;
		const int64_t plane = b.storage() / b.ext[last];
		struct slab {
			long long plane, n_last;
		};
		const std::vector<slab> all = c->boot->gather(slab{static_cast<long long>(plane), static_cast<long long>(b.n[last])}, P);
		long long total = 0;
		for (const slab & s : all) {
			FSB_REQUIRE(s.plane == plane, "box operator: the ranks' slabs differ in the extents across the partitioned axis");
			FSB_REQUIRE(s.n_last >= 1, "box operator: every rank needs at least one dof plane");
			total += s.n_last * (b.dofs() / b.n[last]);
		}
		A->n_global = total;
		std::vector<int64_t> dest_off;
		int64_t landing = 0;
		if (me > 0) { // lower neighbour: gets my first dof plane, fills the pad plane below it
			neighbour nb{};
			nb.rank = me - 1;
			nb.send_count = nb.recv_count = plane;
			nb.contiguous_start = b.lo[last] * plane;
			nb.recv_offset = landing;
			A->nbrs.push_back(nb);
			dest_off.push_back(me - 1 > 0 ? plane : 0); // I am its upper neighbour: second in its landing area if it has a lower one
			A->box_off[0] = (b.lo[last] - 1) * plane;
			landing += plane;
		}
		A->box_split = landing;
		if (me < P - 1) { // upper neighbour: gets my last dof plane, fills the pad plane above it
			neighbour nb{};
			nb.rank = me + 1;
			nb.send_count = nb.recv_count = plane;
			nb.contiguous_start = (b.lo[last] + b.n[last] - 1) * plane;
			nb.recv_offset = landing;
			A->nbrs.push_back(nb);
			dest_off.push_back(0); // I am its lower neighbour: first in its landing area
			A->box_off[1] = (b.lo[last] + b.n[last]) * plane;
			landing += plane;
		}
		A->n_ghost = landing;
		halo_p2p_setup(A, dest_off);
		FSB_REQUIRE(A->halo_p2p || A->nbrs.empty(), "box operator over several ranks needs the peer-memory ghost exchange");
	}
	return A;
}

// The (2 dim + 1)-point stencil  center * u(i) + sum_axis off[axis] * (u(i - e_axis) + u(i + e_axis))  over the interior
// sub-box of a padded array, as the reference's matrix-free stencil does on its narray mesh
// (examples/poisson/mesh.hh:92-134, poisson.cc:44-82).
fsb_parcsr_s * fsb_parcsr_create_box_stencil_impl(fsb_ctx_s * c, int dim, const int64_t * ext, const int64_t * lo,
                                                  const int64_t * hi, double center, const double * off) {
	flush(c);
	const box_shape b = checked_box(dim, ext, lo, hi);
	return assemble_box_operator(
		c, dim, b, [&](int64_t, int a, int) { return off[a]; }, [&](int64_t) { return center; });
}

// The finite-volume diffusion operator of physics/volume_diffusion/diffusion.hh:84-207,
//     v = -beta div(b grad u) + alpha vol a u :
//   flux_axis(c) = b_axis(c) (dA_axis / dx_axis) (u(c + e_axis) - u(c))        (update_flux, :115-160)
//   du(c) = sum_axis flux_axis(c) - flux_axis(c - e_axis)                       (sum_cell_flux, :162-181)
//   v(c)  = -beta du(c) + alpha vol a(c) u(c)                                   (operate, :183-203)
// assembled as CSR: the entry towards c + e_axis is -(beta (b_axis(c) k_axis)), towards c - e_axis
// -(beta (b_axis(c - e_axis) k_axis)), the centre beta * sum_axis (b_axis(c) k_axis + b_axis(c - e_axis) k_axis) + (alpha vol) a(c),
// k_axis = dA_axis / dx_axis = kface[axis].  a and b_axis are padded arrays with the box's extents; b_axis(c) is the
// coefficient of the face between c and c + e_axis (so the faces of the boundary layers take part, as in the reference).
fsb_parcsr_s * fsb_parcsr_create_box_fvm_impl(fsb_ctx_s * c, int dim, const int64_t * ext, const int64_t * lo, const int64_t * hi,
                                              double beta, double alpha, double vol, const double * kface, const double * a,
                                              const double * const * bface) {
	flush(c);
	const box_shape b = checked_box(dim, ext, lo, hi);
	FSB_REQUIRE(kface && a && bface, "box fvm: null coefficient array");
	for (int k = 0; k < dim; ++k)
		FSB_REQUIRE(bface[k], "box fvm: null face coefficient array");
	const int64_t stride[3] = {1, b.ext[0], b.ext[0] * b.ext[1]};
	auto face = [&](int64_t at, int ax) { return bface[ax][at] * kface[ax]; };
	return assemble_box_operator(
		c, dim, b, [&](int64_t at, int ax, int side) { return -(beta * face(side ? at : at - stride[ax], ax)); },
		[&](int64_t at) {
			double s = 0.0;
			for (int ax = 0; ax < dim; ++ax)
				s += face(at, ax) + face(at - stride[ax], ax);
			return beta * s + (alpha * vol) * a[at];
		});
}

void fsb_parcsr_destroy_impl(fsb_parcsr_s * A) {
	flush(A->ctx);
	cudaStreamSynchronize(A->ctx->stream);
	cudaStreamSynchronize(A->ctx->comm_stream);
	halo_p2p_destroy(A);
	free_block(A->diag);
	free_block(A->offd);
	cudaFree(A->d_colmap);
	cudaFree(A->d_send_idx);
	cudaFree(A->d_send_buf);
	cudaFree(A->d_dinv);
	delete A;
}

// ------------------------------------------------------------------ Jacobi

namespace fsb {

// d[r] = 1 / a_rr; rows without a stored diagonal get 1/0 = inf like the reference's `1. / diag`
template<class OffT>
__global__ void extract_dinv_kernel(const OffT * __restrict__ rowptr, const int32_t * __restrict__ col,
                                    const double * __restrict__ val, double * __restrict__ d, long long n) {
	const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r >= n)
		return;
	double a = 0.0;
	for (long long p = static_cast<long long>(rowptr[r]); p < static_cast<long long>(rowptr[r + 1]); ++p)
		if (col[p] == r)
			a = val[p];
	d[r] = __ddiv_rn(1.0, a);
}

void extract_dinv(fsb_parcsr_s * A, double * d) {
	fsb_ctx_s * c = A->ctx;
	const long long n = A->n_local;
	if (n == 0)
		return;
	const int T = 256, grid = static_cast<int>((n + T - 1) / T);
	if (A->diag.wide)
		extract_dinv_kernel<long long>
			<<<grid, T, 0, c->stream>>>(static_cast<const long long *>(A->diag.rowptr), A->diag.col, A->diag.val, d, n);
	else
		extract_dinv_kernel<int><<<grid, T, 0, c->stream>>>(static_cast<const int *>(A->diag.rowptr), A->diag.col,
		                                                    A->diag.val, d, n);
	FSB_CUDA(cudaGetLastError());
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
}

// second half of a Jacobi sweep whose off-process part came from its own launch (NCCL transport): on entry x holds
// lpu = sum_{c != r} a_rc tmp[c], accumulated in the reference's order (mg/jacobi.hh:73-86)
__global__ void jacobi_finalize_kernel(double * __restrict__ x, const double * __restrict__ b,
                                       const double * __restrict__ tmp, const double * __restrict__ dinv, double omega,
                                       long long n) {
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += stride)
		x[r] = __dadd_rn(__dmul_rn(__dmul_rn(omega, dinv[r]), __dadd_rn(b[r], -x[r])), __dmul_rn(__dadd_rn(1.0, -omega), tmp[r]));
}

} // namespace fsb

// mg::bound_jacobi::relax (solvers/mg/jacobi.hh:58-93): per sweep tmp = x (incl. ghosts), then
// x[r] = omega / a_rr * (b[r] - sum_{c != r} a_rc tmp[c]) + (1 - omega) tmp[r].
// One pass over the matrix per sweep: the SpMV kernel skips the diagonal entry while summing (in the reference's order,
// so the result is bit-identical) and applies the update in its epilogue; ghosts of the previous iterate arrive through
// the kernel's own halo exchange.  The copy x -> tmp is a pointer swap when both vectors own their storage.
void fsb_parcsr_jacobi_relax_impl(fsb_parcsr_s * A, double omega, int64_t nrelax, fsb_vec_s * b, fsb_vec_s * x,
                                  fsb_vec_s * tmp) {
	fsb_ctx_s * c = A->ctx;
	flush(c);
	FSB_REQUIRE(x != tmp && b != tmp && x != b, "jacobi_relax: b, x, tmp must be distinct");
	FSB_REQUIRE(x->d != tmp->d && b->d != tmp->d && x->d != b->d, "jacobi_relax: b, x, tmp must not share storage");
	FSB_REQUIRE(x->n_owned == A->n_local && b->n_owned == A->n_local && tmp->n_owned == A->n_local &&
	                x->n_ghost >= A->n_ghost && tmp->n_ghost >= A->n_ghost,
	            "jacobi_relax: vectors do not match the matrix");
	FSB_REQUIRE(!A->diag.has_giant_rows, "jacobi_relax: rows longer than one pipeline stage (4080 entries) are not supported");
	if (A->n_local == 0 && A->nbrs.empty())
		return;
	const bool multi = c->nranks > 1 && !A->nbrs.empty();
	static const bool fused_allowed = !(std::getenv("FSB_FUSED_HALO") && std::atoi(std::getenv("FSB_FUSED_HALO")) == 0);
	const bool fused = multi && A->halo_p2p && A->diag.has_offd_map && A->diag.n_blk > 0 && fused_allowed;
	const bool can_swap = x->owns && tmp->owns && x->n_ghost == tmp->n_ghost && x->padded == tmp->padded;
	for (int64_t s = 0; s < nrelax; ++s) {
		if (!multi || fused) {
			const double * in = x->d;
			double * out = tmp->d;
			if (!can_swap) { // the reference's std::copy, then an in-place sweep reading the copy
				FSB_CUDA(cudaMemcpyAsync(tmp->d, x->d, (x->n_owned + A->n_ghost) * sizeof(double), cudaMemcpyDeviceToDevice,
				                         c->stream));
				in = tmp->d;
				out = x->d;
			}
			spmv_call call;
			call.x = in;
			call.y = out;
			call.x_padded = x->padded && tmp->padded && b->padded;
			call.jacobi = 1;
			call.jacobi_b = b->d;
			call.omega = omega;
			if (fused) {
				call.halo = A;
				call.epoch = ++A->halo_epoch;
				c->stats[FSB_STAT_HALO_EXCHANGES]++;
			}
			launch_spmv(c, A->diag, call, c->stream);
			if (can_swap)
				std::swap(x->d, tmp->d); // x: the new iterate; tmp: the previous one, as the reference leaves it
		}
		else {
			// two launches: ghosts of x through NCCL (or the explicit peer-memory exchange), copy, owned-column part
			// without the diagonal, off-process part continuing the same running sum, update
			if (!A->d_dinv) {
				A->d_dinv = dev_alloc<double>(A->n_local);
				extract_dinv(A, A->d_dinv);
			}
			if (!(x->halo_valid && x->halo_for == A->id)) {
				if (A->halo_p2p) {
					halo_p2p_push(A, x);
					halo_p2p_unpack(A, x);
				}
				else {
					halo_exchange_async(A, x);
					FSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
				}
			}
			FSB_CUDA(cudaMemcpyAsync(tmp->d, x->d, (x->n_owned + A->n_ghost) * sizeof(double), cudaMemcpyDeviceToDevice,
			                         c->stream));
			spmv_call call;
			call.x = tmp->d;
			call.y = x->d;
			call.x_padded = tmp->padded;
			call.jacobi = 2;
			launch_spmv(c, A->diag, call, c->stream);
			if (A->offd.n_blk > 0) {
				spmv_call oc;
				oc.x = tmp->d;
				oc.y = x->d;
				oc.accumulate = true;
				oc.acc_continue = true;
				launch_spmv(c, A->offd, oc, c->stream);
			}
			const long long n = A->n_local;
			if (n > 0) {
				const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, SM_COUNT * 8));
				jacobi_finalize_kernel<<<grid, 256, 0, c->stream>>>(x->d, b->d, tmp->d, A->d_dinv, omega, n);
				FSB_CUDA(cudaGetLastError());
				c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
			}
		}
		x->halo_valid = false;
		tmp->halo_valid = false;
	}
}
