// Pieces shared by the SpMV kernels (spmv.cu) and the ghost exchange (halo.cu): PTX wrappers for mbarriers and
// bulk copies, the row-block descriptor, and the device side of the peer-memory halo (push / wait / acknowledge),
// which the fused SpMV kernel executes itself.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ew_kernels.cuh"

namespace fsb {

// ---------------------------------------------------------------- mbarrier / bulk-copy wrappers
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

// ---------------------------------------------------------------- row-block descriptor
// one row block = the unit of work of a pipeline stage
struct blk_desc {
	long long z0; // first nonzero
	int r0; // first row
	int nrows;
	int nnz;
	int o0; // first compressed row of the off-process block that falls into [r0, r0 + nrows)
	int ocnt; // number of such rows (> 0: the block is scheduled after all interior blocks, see attach_offd_rows)
	int rp0; // window format: where the block's 16-bit row offsets start in rp16[]
	int seg_slot; // window format: where the block's WIN_MAXSEG segment slots start in segs[]
	int pad;
	// value-dictionary stream (spmv.cu: build_value_dictionary): every row padded to a multiple of 8 slots
	long long zp0; // first slot of the block (multiple of 16)
	int pnnz; // slots of the block (multiple of 16)
	int rpp0; // where the block's row meta starts in pmeta[]: nrows + 1 slot offsets, nrows lengths, nrows diagonal positions
};
static_assert(sizeof(blk_desc) == 56, "descriptor layout");

// window format: a contiguous piece of x that a row block reads, staged in shared memory by one bulk copy
struct x_segment {
	int start; // first entry (even: bulk copies are 16-byte granular)
	int len; // entries (even)
};

// ---------------------------------------------------------------- peer-memory halo, device side
constexpr int HALO_MAX_NBR = 8;

struct halo_dev {
	int me, nranks, n_nbr;
	int nbr_rank[HALO_MAX_NBR];
	long long send_count[HALO_MAX_NBR]; // entries I send to nbr k
	long long send_start[HALO_MAX_NBR]; // contiguous: first owned index; packed: offset into send_idx
	int contiguous[HALO_MAX_NBR];
	long long dest_off[HALO_MAX_NBR]; // where my entries land in nbr k's ghost interval
	long long recv_count[HALO_MAX_NBR]; // entries nbr k sends me
	const int32_t * send_idx; // packed send lists (general partitions)
	long long n_ghost; // my ghost entries
	// structured-grid (box) vectors: the ghost entries are pad planes inside the padded array, not an interval behind the
	// owned entries: landing entries [0, box_split) go to x[box_off[0] + i], the rest to x[box_off[1] + i - box_split];
	// box_split < 0: contiguous ghost interval
	long long box_split, box_off[2];
	long long gmax; // landing-area capacity (max ghost count over ranks)
	unsigned char * base[8]; // every rank's block: [spare 16 x 8][ack 8 x 8][pad to 256][landing 0][landing 1], 16 B per entry
	unsigned * counters; // local: [1] unpack CTAs done
	int * error_flag;
};

__device__ __forceinline__ volatile long long * halo_ack(unsigned char * base, int dst) {
	return reinterpret_cast<volatile long long *>(base) + 16 + dst;
}
// entry i of landing buffer `buf`: a flagged 16-byte word (ew_kernels.cuh: ll_store) whose flag is the exchange number
__device__ __forceinline__ unsigned char * halo_landing(unsigned char * base, int buf, long long gmax) {
	return base + 256 + static_cast<size_t>(buf) * static_cast<size_t>(gmax) * 16;
}

__device__ __forceinline__ bool spin_until(volatile long long * flag, long long at_least, bool exact) {
	const long long t0 = clock64();
	for (;;) {
		const long long v = *flag;
		if (exact ? v == at_least : v >= at_least)
			return true;
		if (clock64() - t0 > 20000000000LL)
			return false;
	}
}

// Part `part` of `nparts` of exchange number `epoch`: store my boundary entries of x into the neighbours' landing areas
// over NVLink, each as a flagged word carrying the exchange number -- the payload validates itself, so there is no
// fence, no `ready` flag and no last-writer election.  Executed by `nthreads` threads that synchronise through `sync()`
// (a CTA, or the consumer warps of an SpMV CTA).
template<class Sync>
__device__ __forceinline__ void halo_push_part(const halo_dev & h, const double * __restrict__ x, long long epoch, int tid,
                                               int nthreads, int part, int nparts, Sync && sync,
                                               unsigned long long * dbg = nullptr) {
	const int buf = static_cast<int>(epoch & 1);
	const unsigned flag = static_cast<unsigned>(epoch);
	// the landing buffer was last used by exchange epoch-2: its consumer must have acknowledged it
	if (tid < h.n_nbr && h.send_count[tid] > 0) {
		if (!spin_until(halo_ack(h.base[h.me], h.nbr_rank[tid]), epoch - 2, false))
			*reinterpret_cast<volatile int *>(h.error_flag) = 2;
	}
	sync();
	if (dbg && tid == 0)
		atomicMax(dbg + 0, tl_now());
	const long long first = static_cast<long long>(part) * nthreads + tid;
	const long long stride = static_cast<long long>(nparts) * nthreads;
	for (int k = 0; k < h.n_nbr; ++k) {
		unsigned char * dst = halo_landing(h.base[h.nbr_rank[k]], buf, h.gmax) + static_cast<size_t>(h.dest_off[k]) * 16;
		const long long n = h.send_count[k];
		if (h.contiguous[k]) {
			const double * src = x + h.send_start[k];
			for (long long i = first; i < n; i += stride)
				ll_store(dst + i * 16, src[i], flag);
		}
		else {
			const int32_t * idx = h.send_idx + h.send_start[k];
			for (long long i = first; i < n; i += stride)
				ll_store(dst + i * 16, x[idx[i]], flag);
		}
	}
	if (dbg && tid == 0)
		atomicMax(dbg + 1, tl_now());
}

// ghost entry g of exchange `epoch` on this rank (spins until the neighbour's store has landed)
__device__ __forceinline__ double halo_ghost(const halo_dev & h, const unsigned char * landing, long long g, unsigned flag) {
	double v;
	if (!ll_load(landing + g * 16, flag, v))
		*reinterpret_cast<volatile int *>(h.error_flag) = 3;
	return v;
}

// the landing area of exchange `epoch` is fully consumed on this rank: the senders may reuse it (one thread)
__device__ __forceinline__ void halo_acknowledge(const halo_dev & h, long long epoch) {
	for (int k = 0; k < h.n_nbr; ++k)
		if (h.recv_count[k] > 0)
			*halo_ack(h.base[h.nbr_rank[k]], h.me) = epoch;
}

} // namespace fsb
