// Host-side exchange among the ranks of a context while contexts and matrices are set up (never on the data path):
// gathers of small blobs and making a rank's device block addressable by the other ranks' kernels.
//  * nccl_exchange: one process per GPU (production): ncclAllGather through a staging buffer, cudaIpc handles.
//  * local_exchange: all ranks are threads of one process on ONE device (fsb_ctx_create_group; tests and single-GPU
//    debugging of sharded runs): a barrier + shared slots, raw pointers.
#pragma once
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "fsb_internal.h"

namespace fsb {

struct setup_exchange {
	virtual ~setup_exchange() = default;
	// all[q * bytes, (q + 1) * bytes) = what rank q passed as `mine`; collective, host buffers
	virtual void allgather(const void * mine, void * all, size_t bytes) = 0;
	// peers[q] = address under which rank q's block `mine` can be read and written by this rank's kernels (peers[me] = mine);
	// false if some rank could not map some block (all ranks then get false)
	virtual bool share(void * mine, void ** peers) = 0;
	virtual void unshare(void * peer) = 0;
	virtual bool in_process() const { return false; }
	// Called right before AND right after a launch whose kernel waits for the peers' kernels of the same step.  Among the
	// threads of an in-process group no rank launches before all have arrived (with the kernel loaded, preload_kernel),
	// and none goes on before all have launched: calls that wait for the whole context while holding its lock -- memory
	// allocation, the load of a kernel at its first use -- would otherwise wait for a spinning kernel whose partner can
	// not be launched until they return.  Separate processes on separate GPUs need nothing.
	virtual void rendezvous() {}

	template<class T>
	std::vector<T> gather(const T & mine, int nranks) {
		std::vector<T> all(static_cast<size_t>(nranks));
		allgather(&mine, all.data(), sizeof(T));
		return all;
	}
	// variable-length version: every rank gets every rank's list
	template<class T>
	std::vector<std::vector<T>> gatherv(const std::vector<T> & mine, int nranks) {
		const std::vector<long long> counts = gather<long long>(static_cast<long long>(mine.size()), nranks);
		long long most = 1;
		for (long long n : counts)
			most = std::max(most, n);
		std::vector<T> padded(static_cast<size_t>(most)), all(static_cast<size_t>(most) * nranks);
		std::copy(mine.begin(), mine.end(), padded.begin());
		allgather(padded.data(), all.data(), static_cast<size_t>(most) * sizeof(T));
		std::vector<std::vector<T>> out(static_cast<size_t>(nranks));
		for (int q = 0; q < nranks; ++q)
			out[q].assign(all.begin() + static_cast<size_t>(q) * most, all.begin() + static_cast<size_t>(q) * most + counts[q]);
		return out;
	}
};

struct nccl_exchange : setup_exchange {
	fsb_ctx_s * c;
	explicit nccl_exchange(fsb_ctx_s * ctx) : c(ctx) {}
	void allgather(const void * mine, void * all, size_t bytes) override {
		const int P = c->nranks;
		char * d = nullptr;
		FSB_CUDA(cudaMalloc(&d, std::max<size_t>(bytes * P, 16)));
		FSB_CUDA(cudaMemcpyAsync(d + bytes * c->rank, mine, bytes, cudaMemcpyHostToDevice, c->stream));
		FSB_NCCL(ncclAllGather(d + bytes * c->rank, d, bytes, ncclChar, c->nccl, c->stream));
		FSB_CUDA(cudaMemcpyAsync(all, d, bytes * P, cudaMemcpyDeviceToHost, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		cudaFree(d);
	}
	bool share(void * mine, void ** peers) override {
		struct packet {
			cudaIpcMemHandle_t h;
			long long ok;
		};
		packet me{};
		me.ok = cudaIpcGetMemHandle(&me.h, mine) == cudaSuccess ? 1 : 0;
		if (!me.ok)
			cudaGetLastError();
		const std::vector<packet> all = gather(me, c->nranks);
		bool ok = true;
		for (const packet & p : all)
			ok = ok && p.ok;
		for (int q = 0; q < c->nranks; ++q)
			peers[q] = nullptr;
		for (int q = 0; q < c->nranks && ok; ++q) {
			if (q == c->rank) {
				peers[q] = mine;
				continue;
			}
			if (cudaIpcOpenMemHandle(&peers[q], all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
				cudaGetLastError();
				peers[q] = nullptr;
				ok = false;
			}
		}
		// every rank must agree, otherwise some would wait on blocks nobody writes
		const std::vector<long long> oks = gather<long long>(ok ? 1 : 0, c->nranks);
		for (long long v : oks)
			ok = ok && v;
		if (!ok)
			for (int q = 0; q < c->nranks; ++q)
				if (q != c->rank && peers[q]) {
					cudaIpcCloseMemHandle(peers[q]);
					peers[q] = nullptr;
				}
		return ok;
	}
	void unshare(void * peer) override { cudaIpcCloseMemHandle(peer); }
};

// shared by the contexts of one fsb_ctx_create_group call
struct local_group {
	explicit local_group(int n) : nranks(n), slot(static_cast<size_t>(n), nullptr) {}
	int nranks;
	std::mutex m;
	std::condition_variable cv;
	int waiting = 0;
	long long generation = 0;
	std::vector<const void *> slot;
	void barrier() {
		std::unique_lock<std::mutex> lock(m);
		const long long gen = generation;
		if (++waiting == nranks) {
			waiting = 0;
			++generation;
			cv.notify_all();
		}
		else
			cv.wait(lock, [&] { return generation != gen; });
	}
};

struct local_exchange : setup_exchange {
	std::shared_ptr<local_group> g;
	int rank;
	local_exchange(std::shared_ptr<local_group> group, int r) : g(std::move(group)), rank(r) {}
	void allgather(const void * mine, void * all, size_t bytes) override {
		g->slot[static_cast<size_t>(rank)] = mine;
		g->barrier();
		for (int q = 0; q < g->nranks; ++q)
			std::memcpy(static_cast<char *>(all) + bytes * q, g->slot[static_cast<size_t>(q)], bytes);
		g->barrier(); // nobody's `mine` goes away before everyone has copied it
	}
	bool share(void * mine, void ** peers) override { // one address space: the pointers themselves
		allgather(&mine, peers, sizeof(void *));
		return true;
	}
	void unshare(void *) override {}
	bool in_process() const override { return true; }
	void rendezvous() override { g->barrier(); }
};

} // namespace fsb
