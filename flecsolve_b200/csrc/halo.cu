// Ghost exchange over NVLink peer memory (replaces NCCL send/recv when cudaIpc mapping works).
//
// Reference: the FleCSI copy plan of topo::csr (topo/csr.hh:116-187, 237-245, 277-304): destination =
// the contiguous ghost interval of the `cols` space, sources = (owner colour, shared offset) pairs.
// Here every rank owns a double-buffered landing area that its neighbours have mapped; an exchange is
//   push   : copy my boundary entries straight into each neighbour's landing area (stores over
//            NVLink), system fence, then store the exchange number into the neighbour's `ready` flag;
//   unpack : wait for my `ready` flags, move the landing area into the ghost interval of x, then
//            store the exchange number into each sender's `ack` flag (buffer may be reused).
// Both are small kernels on the compute stream: push runs before the diag-block SpMV, unpack after it,
// so the transfer overlaps the interior rows without a second stream, an NCCL launch or host work.
#include <cstdlib>
#include <cstring>

#include "fsb_internal.h"
#include "setup_exchange.h"
#include "spmv_common.cuh"

namespace fsb {

constexpr int HALO_THREADS = 256;

// explicit exchange (fsb_parcsr_halo_exchange, NCCL-free fallback of the two-launch SpMV): the fused SpMV kernel runs the
// same halo_push_part from its consumer warps instead
__global__ void __launch_bounds__(HALO_THREADS) halo_push_kernel(const halo_dev * hp, const double * __restrict__ x,
                                                                 long long epoch) {
	halo_push_part(*hp, x, epoch, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x, [] { __syncthreads(); });
}

__global__ void __launch_bounds__(HALO_THREADS) halo_unpack_kernel(const halo_dev * hp, double * __restrict__ ghosts,
                                                                   long long epoch) {
	const halo_dev & h = *hp;
	const unsigned char * src = halo_landing(h.base[h.me], static_cast<int>(epoch & 1), h.gmax);
	const unsigned flag = static_cast<unsigned>(epoch);
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < h.n_ghost; i += stride) {
		const long long at = h.box_split < 0 ? i : (i < h.box_split ? h.box_off[0] + i : h.box_off[1] + (i - h.box_split));
		ghosts[at] = halo_ghost(h, src, i, flag); // every entry validates itself
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		if (atomicAdd(&h.counters[1], 1u) == gridDim.x - 1) { // landing area fully consumed: acknowledge
			halo_acknowledge(h, epoch);
			h.counters[1] = 0u;
		}
	}
}

// Collective over the ranks of the context.  dest_off[k]: where my entries land in neighbour k's
// ghost interval.  Leaves A->halo_p2p null (=> NCCL path) if peer mapping is unavailable.
void halo_p2p_setup(fsb_parcsr_s * A, const std::vector<int64_t> & dest_off) {
	fsb_ctx_s * c = A->ctx;
	const int P = c->nranks;
	if (P == 1 || !c->d_xrank || P > 8 || static_cast<int>(A->nbrs.size()) > HALO_MAX_NBR)
		return;
	if (const char * e = std::getenv("FSB_P2P_HALO"))
		if (std::atoi(e) == 0)
			return;
	// capacity = max ghost count over ranks
	long long g = 0;
	for (long long n : c->boot->gather<long long>(A->n_ghost, P))
		g = std::max(g, n);
	const long long gmax = std::max<long long>(g, 1);
	const size_t bytes = 256 + 2 * static_cast<size_t>(gmax) * 16; // two landing buffers of flagged 16-byte words
	unsigned char * block = nullptr;
	FSB_CUDA(cudaMalloc(&block, bytes));
	FSB_CUDA(cudaMemset(block, 0, bytes)); // flags 0: no exchange number matches
	void * peers[8] = {};
	if (!c->boot->share(block, peers)) {
		cudaFree(block);
		return;
	}
	halo_dev h{};
	h.me = c->rank;
	h.nranks = P;
	std::vector<void *> opened(P, nullptr);
	for (int q = 0; q < P; ++q) {
		h.base[q] = static_cast<unsigned char *>(peers[q]);
		if (q != c->rank)
			opened[q] = peers[q];
	}
	h.n_nbr = static_cast<int>(A->nbrs.size());
	for (int k = 0; k < h.n_nbr; ++k) {
		const neighbour & nb = A->nbrs[k];
		h.nbr_rank[k] = nb.rank;
		h.send_count[k] = nb.send_count;
		h.contiguous[k] = nb.contiguous_start >= 0 ? 1 : 0;
		h.send_start[k] = nb.contiguous_start >= 0 ? nb.contiguous_start : nb.send_offset;
		h.dest_off[k] = dest_off[k];
		h.recv_count[k] = nb.recv_count;
	}
	h.send_idx = A->d_send_idx;
	h.n_ghost = A->n_ghost;
	h.box_split = A->box ? A->box_split : -1;
	h.box_off[0] = A->box_off[0];
	h.box_off[1] = A->box_off[1];
	h.gmax = gmax;
	FSB_CUDA(cudaMalloc(&h.counters, 2 * sizeof(unsigned)));
	FSB_CUDA(cudaMemset(h.counters, 0, 2 * sizeof(unsigned)));
	FSB_CUDA(cudaHostGetDevicePointer(&h.error_flag, c->h_xrank_error, 0));
	halo_dev * d_h = nullptr;
	FSB_CUDA(cudaMalloc(&d_h, sizeof(halo_dev)));
	FSB_CUDA(cudaMemcpy(d_h, &h, sizeof(h), cudaMemcpyHostToDevice));
	A->halo_p2p = d_h;
	A->halo_block = block;
	A->halo_counters = h.counters;
	A->halo_opened = opened;
	A->halo_epoch = 0;
	long long total_send = 0;
	for (const neighbour & nb : A->nbrs)
		total_send += nb.send_count;
	A->halo_push_ctas = static_cast<int>(std::max<long long>(1, std::min<long long>(64, (total_send + 2047) / 2048)));
	A->halo_unpack_ctas = static_cast<int>(std::max<long long>(1, std::min<long long>(64, (A->n_ghost + 2047) / 2048)));
}

void halo_p2p_destroy(fsb_parcsr_s * A) {
	if (!A->halo_p2p)
		return;
	for (void * p : A->halo_opened)
		if (p && A->ctx->boot)
			A->ctx->boot->unshare(p);
	cudaFree(A->halo_counters);
	cudaFree(A->halo_block);
	cudaFree(A->halo_p2p);
	A->halo_p2p = nullptr;
}

// start exchange number ++epoch: push my boundary entries.  Runs on the communication stream (after
// everything already queued on the compute stream) so it overlaps the diag-block SpMV.
void halo_p2p_push(fsb_parcsr_s * A, fsb_vec_s * x) {
	fsb_ctx_s * c = A->ctx;
	++A->halo_epoch;
	FSB_CUDA(cudaEventRecord(c->ev_main, c->stream));
	FSB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_main, 0));
	halo_push_kernel<<<A->halo_push_ctas, HALO_THREADS, 0, c->comm_stream>>>(static_cast<const halo_dev *>(A->halo_p2p),
	                                                                          x->d, A->halo_epoch);
	FSB_CUDA(cudaGetLastError());
	FSB_CUDA(cudaEventRecord(c->ev_comm, c->comm_stream));
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	c->stats[FSB_STAT_HALO_EXCHANGES]++;
}

// finish it: ghosts of x become valid (compute stream)
void halo_p2p_unpack(fsb_parcsr_s * A, fsb_vec_s * x) {
	fsb_ctx_s * c = A->ctx;
	if (c->boot && c->boot->in_process()) { // the kernel waits for the neighbours' push kernels
		preload_kernel(halo_unpack_kernel);
		c->boot->rendezvous();
	}
	halo_unpack_kernel<<<A->halo_unpack_ctas, HALO_THREADS, 0, c->stream>>>(static_cast<const halo_dev *>(A->halo_p2p),
	                                                                         A->box ? x->d : x->d + x->n_owned, A->halo_epoch);
	FSB_CUDA(cudaGetLastError());
	if (c->boot && c->boot->in_process())
		c->boot->rendezvous();
	// my own push has long finished by now (the peers waited for it); ordering it before whatever
	// the compute stream does next keeps later writers of x from racing with it
	FSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	x->halo_valid = true;
	x->halo_for = A->id;
}

} // namespace fsb
