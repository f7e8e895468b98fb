// extern "C" surface of libfsb.so: contexts, vectors, reductions, matrix wrappers.
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <random>
#include <thread>

#include "ew_kernels.cuh"
#include "fsb_internal.h"
#include "setup_exchange.h"

using namespace fsb;

fsb_parcsr_s * fsb_parcsr_create_impl(fsb_ctx_s *, int64_t, const int64_t *, const int64_t *, const int64_t *,
                                      const double *);
fsb_parcsr_s * fsb_parcsr_create_stencil_impl(fsb_ctx_s *, int, int64_t, int64_t, int64_t, double, double);
fsb_parcsr_s * fsb_parcsr_create_box_stencil_impl(fsb_ctx_s *, int, const int64_t *, const int64_t *, const int64_t *, double,
                                                  const double *);
fsb_parcsr_s * fsb_parcsr_create_box_fvm_impl(fsb_ctx_s *, int, const int64_t *, const int64_t *, const int64_t *, double, double, double,
                                              const double *, const double *, const double * const *);
void fsb_parcsr_destroy_impl(fsb_parcsr_s *);
void fsb_parcsr_jacobi_relax_impl(fsb_parcsr_s *, double, int64_t, fsb_vec_s *, fsb_vec_s *, fsb_vec_s *);

namespace fsb {
static thread_local std::string g_last_error;
void set_last_error(const std::string & m) { g_last_error = m; }

// Off for the whole process once an in-process rank group exists: there the peers' kernels share the device, and a
// dependent kernel scheduled early sits on SM resources while the kernel it depends on waits for a peer that may need them
// (observed as a time-out of the cross-rank reduction on B200; separate processes on separate GPUs are not affected).
static std::atomic<bool> g_pdl_off{false};
bool pdl_enabled() {
	static const bool on = !(std::getenv("FSB_PDL") && std::atoi(std::getenv("FSB_PDL")) == 0);
	return on && !g_pdl_off.load(std::memory_order_relaxed);
}

// the host only hands out slots (and stamps the kind); the kernels write the times
unsigned long long * timeline_slot(fsb_ctx_s * c, int kind) {
	if (!c->d_timeline || c->timeline_used >= c->timeline_cap)
		return nullptr;
	unsigned long long * slot = c->d_timeline + c->timeline_used * TL_WORDS;
	c->timeline_kind.push_back(kind); // merged into word 2 when the timeline is read: no extra work on the stream
	++c->timeline_used;
	return slot;
}

static void timeline_reset(fsb_ctx_s * c) {
	std::vector<unsigned long long> init(static_cast<size_t>(c->timeline_cap) * TL_WORDS, 0ull);
	for (int64_t i = 0; i < c->timeline_cap; ++i)
		init[i * TL_WORDS + 0] = init[i * TL_WORDS + 4] = init[i * TL_WORDS + 6] = ~0ull;
	FSB_CUDA(cudaMemcpy(c->d_timeline, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
	c->timeline_used = 0;
	c->timeline_kind.clear();
}
} // namespace fsb

// run `body`, translate exceptions into status codes
template<class F>
static int guarded(F && body) noexcept {
	try {
		body();
		return FSB_OK;
	}
	catch (const fsb::error & e) {
		set_last_error(e.what());
		return e.code;
	}
	catch (const std::exception & e) {
		set_last_error(e.what());
		return FSB_ERR_STATE;
	}
	catch (...) {
		set_last_error("unknown error");
		return FSB_ERR_STATE;
	}
}

// Make every rank's reduction mailbox addressable by this rank's kernels (cudaIpc across processes, plain pointers among
// the threads of a rank group) so the kernels can all-reduce scalars with plain stores/loads over NVLink.  Any failure
// leaves d_xrank null => NCCL path.
static void setup_peer_reductions(fsb_ctx_s * c) {
	static_assert(XRANK_RING == FSB_RED_RING, "mailbox ring must match the token ring");
	const int P = c->nranks;
	if (P > XRANK_MAX)
		return;
	const size_t bytes = sizeof(double) * 2 * XRANK_RING * P;
	FSB_CUDA(cudaMalloc(&c->d_mailbox, bytes));
	FSB_CUDA(cudaMemset(c->d_mailbox, 0, bytes));
	void * peers[XRANK_MAX] = {};
	if (!c->boot->share(c->d_mailbox, peers)) {
		if (c->rank == 0)
			fprintf(stderr, "[fsb] peer-memory reductions unavailable (cudaIpc failed); using ncclAllReduce\n");
		return;
	}
	xrank_info info{};
	info.me = c->rank;
	info.nranks = P;
	for (int q = 0; q < P; ++q) {
		info.mailbox[q] = static_cast<double *>(peers[q]);
		if (q != c->rank)
			c->peer_mailbox[q] = peers[q];
	}
	FSB_CUDA(cudaHostAlloc(&c->h_xrank_error, sizeof(int), cudaHostAllocMapped));
	*c->h_xrank_error = 0;
	FSB_CUDA(cudaHostGetDevicePointer(&info.error_flag, c->h_xrank_error, 0));
	FSB_CUDA(cudaMalloc(&c->d_xrank, sizeof(xrank_info)));
	FSB_CUDA(cudaMemcpy(c->d_xrank, &info, sizeof(info), cudaMemcpyHostToDevice));
}

// everything of a context that does not depend on how its ranks talk to each other
static fsb_ctx_s * new_context(int device, int rank, int nranks) {
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		throw fsb::error(FSB_ERR_NOGPU, "no CUDA device available: this library has no CPU fallback");
	}
	FSB_REQUIRE(device >= 0 && device < ndev, "device index out of range");
	FSB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop{};
	FSB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
		throw fsb::error(FSB_ERR_NOGPU, std::string("device is sm_") + std::to_string(prop.major) +
		                                    std::to_string(prop.minor) + "; kernels are built for sm_100a only");
	auto * c = new fsb_ctx_s;
	c->device = device;
	c->rank = rank;
	c->nranks = nranks;
	FSB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	FSB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
	FSB_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
	FSB_CUDA(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
	FSB_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * MAXR * MAX_RED_BLOCKS));
	FSB_CUDA(cudaMalloc(&c->d_counter, sizeof(unsigned)));
	FSB_CUDA(cudaMemset(c->d_counter, 0, sizeof(unsigned)));
	FSB_CUDA(cudaMalloc(&c->d_sched, 2 * sizeof(unsigned)));
	FSB_CUDA(cudaMemset(c->d_sched, 0, 2 * sizeof(unsigned)));
	FSB_CUDA(cudaMalloc(&c->d_results, sizeof(double) * FSB_RED_RING));
	FSB_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * fsb::MAX_SCALARS));
	{
		std::vector<double> init(fsb::MAX_SCALARS, 0.0);
		init[0] = 1.0;
		FSB_CUDA(cudaMemcpy(c->d_scalars, init.data(), sizeof(double) * fsb::MAX_SCALARS, cudaMemcpyHostToDevice));
		c->scalar_used[0] = true;
	}
	FSB_CUDA(cudaMalloc(&c->d_halt, sizeof(int)));
	FSB_CUDA(cudaMemset(c->d_halt, 0, sizeof(int)));
	FSB_CUDA(cudaHostAlloc(&c->h_results, sizeof(double) * FSB_RED_RING, cudaHostAllocDefault));
	FSB_CUDA(cudaHostAlloc(&c->h_ll, sizeof(fsb::ll_word) * FSB_RED_RING, cudaHostAllocMapped));
	std::memset(c->h_results, 0, sizeof(double) * FSB_RED_RING);
	std::memset(c->h_ll, 0, sizeof(fsb::ll_word) * FSB_RED_RING);
	FSB_CUDA(cudaHostGetDevicePointer(&c->h_ll_dev, c->h_ll, 0));
	c->token_op.assign(FSB_RED_RING, 0);
	c->token_event.resize(FSB_RED_RING, nullptr);
	return c;
}

// the part that follows once the ranks can talk (c->boot is set when nranks > 1)
static void finish_context(fsb_ctx_s * c) {
	const int nranks = c->nranks;
	if (nranks > 1 && (c->boot->in_process() || !(std::getenv("FSB_P2P_REDUCE") && std::atoi(std::getenv("FSB_P2P_REDUCE")) == 0)))
		setup_peer_reductions(c);
	FSB_REQUIRE(nranks == 1 || c->nccl || c->d_xrank, "a rank group in one process needs peer-memory reductions");
	// Static row-block schedule (bitwise reproducible, and measured faster: profiles/r1_scaling_final.txt)
	// unless NCCL kernels share the SMs with the SpMV: a CTA that starts late behind a communication
	// kernel would otherwise stretch the one-wave grid into two.
	c->reproducible = nranks == 1 || c->d_xrank != nullptr;
	if (const char * e = std::getenv("FSB_REPRODUCIBLE")) // measurement override
		c->reproducible = std::atoi(e) != 0;
	if (const char * t = std::getenv("FSB_JIT"))
		c->jit = std::atoi(t) != 0;
	if (const char * t = std::getenv("FSB_TRACE"))
		c->trace = std::atoi(t) != 0;
	if (const char * t = std::getenv("FSB_FUSION"))
		c->fusion = std::atoi(t) != 0;
}

extern "C" {

const char * fsb_last_error(void) { return g_last_error.c_str(); }
int fsb_version(void) { return 200; }

int fsb_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int fsb_nccl_unique_id(void * out128) {
	return guarded([&] {
		FSB_REQUIRE(out128, "null output");
		static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
		ncclUniqueId id;
		FSB_NCCL(ncclGetUniqueId(&id));
		std::memcpy(out128, &id, sizeof(id));
	});
}

int fsb_ctx_create(int device, int rank, int nranks, const void * uid, fsb_ctx_t * out) {
	return guarded([&] {
		FSB_REQUIRE(out, "null output");
		FSB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
		FSB_REQUIRE(nranks == 1 || uid, "nccl_unique_id required when nranks > 1");
		fsb_ctx_s * c = new_context(device, rank, nranks);
		if (nranks > 1) {
			ncclUniqueId id;
			std::memcpy(&id, uid, sizeof(id));
			FSB_NCCL(ncclCommInitRank(&c->nccl, nranks, id, rank));
			FSB_NCCL(ncclCommSplit(c->nccl, 0, rank, &c->nccl_halo_comm, nullptr));
			for (auto & e : c->token_event)
				FSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			c->boot = std::make_shared<fsb::nccl_exchange>(c);
		}
		finish_context(c);
		*out = c;
	});
}

int fsb_ctx_create_group(int device, int nranks, fsb_ctx_t * out) {
	return guarded([&] {
		FSB_REQUIRE(out && nranks >= 1 && nranks <= XRANK_MAX, "bad arguments");
		auto group = std::make_shared<fsb::local_group>(nranks);
		if (nranks > 1)
			fsb::g_pdl_off.store(true);
		std::vector<fsb_ctx_s *> made;
		for (int r = 0; r < nranks; ++r) {
			fsb_ctx_s * c = new_context(device, r, nranks);
			if (nranks > 1)
				c->boot = std::make_shared<fsb::local_exchange>(group, r);
			made.push_back(c);
		}
		// the collective part of the setup, one thread per rank as every later collective call needs it
		std::vector<std::thread> threads;
		std::vector<std::string> errors(static_cast<size_t>(nranks));
		for (int r = 0; r < nranks; ++r)
			threads.emplace_back([&, r] {
				try {
					cudaSetDevice(device);
					finish_context(made[static_cast<size_t>(r)]);
				}
				catch (const std::exception & e) {
					errors[static_cast<size_t>(r)] = e.what();
				}
			});
		for (auto & t : threads)
			t.join();
		for (const std::string & e : errors)
			if (!e.empty())
				throw fsb::error(FSB_ERR_STATE, e);
		for (int r = 0; r < nranks; ++r)
			out[r] = made[static_cast<size_t>(r)];
	});
}

int fsb_ctx_destroy(fsb_ctx_t c) {
	return guarded([&] {
		if (!c)
			return;
		flush(c);
		fsb::speculation_release(c);
		cudaStreamSynchronize(c->stream);
		cudaStreamSynchronize(c->comm_stream);
		fsb::speculation_destroy(c);
		for (void * p : c->peer_mailbox)
			if (p && c->boot)
				c->boot->unshare(p);
		cudaFree(c->d_mailbox);
		cudaFree(c->d_xrank);
		if (c->h_xrank_error)
			cudaFreeHost(c->h_xrank_error);
		if (c->nccl_halo_comm)
			ncclCommDestroy(c->nccl_halo_comm);
		if (c->nccl)
			ncclCommDestroy(c->nccl);
		for (auto e : c->token_event)
			if (e)
				cudaEventDestroy(e);
		for (auto e : c->timers)
			if (e)
				cudaEventDestroy(e);
		for (auto e : c->prof_events)
			cudaEventDestroy(e);
		cudaFree(c->d_partials);
		cudaFree(c->d_counter);
		cudaFree(c->d_sched);
		cudaFree(c->d_results);
		cudaFree(c->d_scalars);
		cudaFree(c->d_halt);
		cudaFree(c->d_flush);
		cudaFree(c->d_timeline);
		for (void * p : c->deferred_free)
			cudaFree(p);
		cudaFreeHost(c->h_results);
		cudaFreeHost(c->h_ll);
		cudaEventDestroy(c->ev_main);
		cudaEventDestroy(c->ev_comm);
		cudaStreamDestroy(c->stream);
		cudaStreamDestroy(c->comm_stream);
		delete c;
	});
}

int fsb_ctx_flush(fsb_ctx_t c) {
	return guarded([&] { flush(c); });
}

int fsb_ctx_sync(fsb_ctx_t c) {
	return guarded([&] {
		flush(c);
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->comm_stream));
		c->stats[FSB_STAT_HOST_SYNCS]++;
	});
}

void * fsb_ctx_stream(fsb_ctx_t c) { return c ? c->stream : nullptr; }
int fsb_ctx_rank(fsb_ctx_t c) { return c->rank; }
int fsb_ctx_nranks(fsb_ctx_t c) { return c->nranks; }

int fsb_ctx_set_option(fsb_ctx_t c, int option, int64_t value) {
	return guarded([&] {
		flush(c);
		switch (option) {
		case FSB_OPT_FUSION:
			c->fusion = value != 0;
			break;
		case FSB_OPT_SPMV_ROWS_PER_CTA:
			c->spmv_rows_per_cta = static_cast<int>(value);
			break;
		case FSB_OPT_SPMV_THREADS:
			c->spmv_threads = static_cast<int>(value);
			break;
		case FSB_OPT_TRACE:
			c->trace = value != 0;
			break;
		case FSB_OPT_PROFILE:
			c->profile = value != 0;
			break;
		case FSB_OPT_REPRODUCIBLE:
			c->reproducible = value != 0;
			break;
		case FSB_OPT_JIT:
			c->jit = value != 0;
			break;
		case FSB_OPT_SPMV_DICTIONARY:
			c->spmv_dictionary = value != 0;
			break;
		case FSB_OPT_SPECULATE:
			c->speculate = value != 0;
			if (!c->speculate)
				fsb::speculation_release(c);
			break;
		case FSB_OPT_TIMELINE:
			FSB_CUDA(cudaStreamSynchronize(c->stream));
			cudaFree(c->d_timeline);
			c->d_timeline = nullptr;
			c->timeline_cap = value > 0 ? value : 0;
			if (value > 0) {
				FSB_CUDA(cudaMalloc(&c->d_timeline, static_cast<size_t>(value) * TL_WORDS * sizeof(unsigned long long)));
				timeline_reset(c);
			}
			break;
		default:
			throw fsb::error(FSB_ERR_ARG, "unknown option");
		}
	});
}

int fsb_debug_program_info(const int32_t * raw, int n, int device_coefficients, int32_t * info6, int32_t * canon) {
	return guarded([&] {
		FSB_REQUIRE(raw && info6 && n >= 1 && n <= fsb::MAXS, "bad arguments");
		fsb::raw_stmt rs[fsb::MAXS];
		for (int i = 0; i < n; ++i)
			rs[i] = fsb::raw_stmt{raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]};
		const fsb::canon_result cr = fsb::canonicalize(rs, n);
		FSB_REQUIRE(cr.ok, "statement list exceeds the program limits");
		info6[0] = fsb::program_is_registered(cr.p, device_coefficients != 0) ? 1 : 0;
		info6[1] = cr.p.nv;
		info6[2] = cr.p.ns;
		info6[3] = cr.p.nr;
		info6[4] = static_cast<int32_t>(cr.p.load_mask);
		info6[5] = static_cast<int32_t>(cr.p.store_mask);
		if (canon)
			for (int i = 0; i < n; ++i) {
				const fsb::stmt & t = cr.p.st[i];
				const int v[6] = {t.op, t.z, t.x, t.y, t.a, t.b};
				for (int k = 0; k < 6; ++k)
					canon[6 * i + k] = v[k];
			}
	});
}

int fsb_debug_jit_compile(const int32_t * raw, int n, int device_coefficients, int box_layout, int64_t * cubin_bytes,
                          char * log, int log_capacity) {
	return guarded([&] {
		FSB_REQUIRE(raw && n >= 1 && n <= fsb::MAXS && cubin_bytes, "bad arguments");
		fsb::raw_stmt rs[fsb::MAXS];
		for (int i = 0; i < n; ++i)
			rs[i] = fsb::raw_stmt{raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]};
		const fsb::canon_result cr = fsb::canonicalize(rs, n);
		FSB_REQUIRE(cr.ok, "statement list exceeds the program limits");
		std::vector<char> cubin;
		std::string text;
		const bool ok = fsb::jit_compile(cr.p, device_coefficients != 0, box_layout != 0, cubin, text);
		if (log && log_capacity > 0) {
			std::strncpy(log, text.c_str(), static_cast<size_t>(log_capacity) - 1);
			log[log_capacity - 1] = 0;
		}
		if (!ok)
			throw fsb::error(FSB_ERR_STATE, "run-time compilation failed: " + text.substr(0, 400));
		*cubin_bytes = static_cast<int64_t>(cubin.size());
	});
}

int fsb_ctx_get_stat(fsb_ctx_t c, int stat, int64_t * out) {
	return guarded([&] {
		FSB_REQUIRE(stat >= 0 && stat < 16 && out, "bad stat");
		*out = c->stats[stat];
	});
}

int fsb_ctx_reset_stats(fsb_ctx_t c) {
	return guarded([&] {
		for (auto & s : c->stats)
			s = 0;
	});
}

int fsb_ctx_flush_l2(fsb_ctx_t c) {
	return guarded([&] {
		flush(c);
		const size_t bytes = 256u << 20; // 2x the 126 MB L2
		if (!c->d_flush) {
			FSB_CUDA(cudaMalloc(&c->d_flush, bytes));
			c->flush_bytes = bytes;
		}
		FSB_CUDA(cudaMemsetAsync(c->d_flush, 1, c->flush_bytes, c->stream));
	});
}

int fsb_ctx_event_record(fsb_ctx_t c, int slot) {
	return guarded([&] {
		FSB_REQUIRE(c && slot >= 0 && slot < 16, "event slot out of range");
		flush(c);
		if (!c->timers[slot])
			FSB_CUDA(cudaEventCreate(&c->timers[slot]));
		FSB_CUDA(cudaEventRecord(c->timers[slot], c->stream));
	});
}

int fsb_ctx_event_elapsed_ms(fsb_ctx_t c, int a, int b, double * ms) {
	return guarded([&] {
		FSB_REQUIRE(c && ms && a >= 0 && a < 16 && b >= 0 && b < 16 && c->timers[a] && c->timers[b], "bad event slots");
		FSB_CUDA(cudaEventSynchronize(c->timers[b]));
		float t = 0;
		FSB_CUDA(cudaEventElapsedTime(&t, c->timers[a], c->timers[b]));
		*ms = t;
	});
}

int fsb_ctx_profile_read_split(fsb_ctx_t c, double * ms2, int64_t * launches2) {
	return guarded([&] {
		FSB_REQUIRE(c && ms2 && launches2, "bad arguments");
		flush(c);
		ms2[0] = ms2[1] = 0;
		launches2[0] = launches2[1] = 0;
		for (size_t k = 0; k + 1 < c->prof_used; k += 2) {
			FSB_CUDA(cudaEventSynchronize(c->prof_events[k + 1]));
			float t = 0;
			FSB_CUDA(cudaEventElapsedTime(&t, c->prof_events[k], c->prof_events[k + 1]));
			const int tag = c->prof_tag[k / 2] ? 1 : 0;
			ms2[tag] += t;
			launches2[tag]++;
		}
		c->prof_used = 0;
	});
}

int fsb_ctx_timeline_read(fsb_ctx_t c, uint64_t * out, int64_t max_slots, int64_t * n_slots) {
	return guarded([&] {
		FSB_REQUIRE(c && n_slots, "bad arguments");
		flush(c);
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->comm_stream));
		const int64_t n = std::min<int64_t>(c->timeline_used, max_slots);
		*n_slots = n;
		if (out && n > 0) {
			FSB_CUDA(cudaMemcpy(out, c->d_timeline, static_cast<size_t>(n) * TL_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost));
			for (int64_t i = 0; i < n; ++i)
				out[i * TL_WORDS + 2] = static_cast<uint64_t>(c->timeline_kind[static_cast<size_t>(i)]);
		}
		if (c->d_timeline)
			timeline_reset(c);
	});
}

int fsb_ctx_profile_read(fsb_ctx_t c, double * spmv_ms, int64_t * spmv_launches) {
	double ms2[2];
	int64_t n2[2];
	const int rc = fsb_ctx_profile_read_split(c, ms2, n2);
	if (rc == FSB_OK && spmv_ms && spmv_launches) {
		*spmv_ms = ms2[0] + ms2[1];
		*spmv_launches = n2[0] + n2[1];
	}
	return rc;
}

// ------------------------------------------------------------------ vectors

int fsb_vec_create(fsb_ctx_t c, int64_t n_owned, int64_t n_ghost, fsb_vec_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out && n_owned >= 0 && n_ghost >= 0, "bad arguments");
		fsb::speculation_release(c); // cudaMalloc waits for the device: nothing may sit there waiting for the host
		auto * v = new fsb_vec_s;
		v->ctx = c;
		v->n_owned = n_owned;
		v->n_ghost = n_ghost;
		v->id = c->next_vec_id++;
		// two doubles of slack: 16-byte bulk copies of the SpMV kernels may read one entry past an odd length
		const size_t n = static_cast<size_t>(n_owned + n_ghost) + 2;
		FSB_CUDA(cudaMalloc(&v->d, n * sizeof(double)));
		FSB_CUDA(cudaMemsetAsync(v->d, 0, n * sizeof(double), c->stream));
		v->padded = true;
		*out = v;
	});
}

int fsb_vec_wrap(fsb_ctx_t c, double * p, int64_t n_owned, int64_t n_ghost, fsb_vec_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out && p && n_owned >= 0 && n_ghost >= 0, "bad arguments");
		FSB_REQUIRE(reinterpret_cast<uintptr_t>(p) % 16 == 0, "wrapped device pointer must be 16-byte aligned");
		cudaPointerAttributes at{};
		FSB_CUDA(cudaPointerGetAttributes(&at, p));
		FSB_REQUIRE(at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged, "pointer is not device memory");
		auto * v = new fsb_vec_s;
		v->ctx = c;
		v->d = p;
		v->owns = false;
		v->n_owned = n_owned;
		v->n_ghost = n_ghost;
		v->id = c->next_vec_id++;
		*out = v;
	});
}

int fsb_vec_destroy(fsb_vec_t v) {
	return guarded([&] {
		if (!v)
			return;
		flush(v->ctx);
		if (v->owns) {
			FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
			FSB_CUDA(cudaStreamSynchronize(v->ctx->comm_stream));
			// cudaFree waits for the whole device: among the ranks of an in-process group that would wait for a peer's
			// kernel which may itself be waiting for this rank's next launch -- the memory goes with the context instead
			if (v->ctx->boot && v->ctx->boot->in_process())
				v->ctx->deferred_free.push_back(v->d);
			else
				cudaFree(v->d);
		}
		delete v;
	});
}

int64_t fsb_vec_local_size(fsb_vec_t v) { return v->n_owned; }
int64_t fsb_vec_ghost_size(fsb_vec_t v) { return v->n_ghost; }
double * fsb_vec_device_ptr(fsb_vec_t v) { return v->d; }

// dofs [first, first + n) of a structured-grid vector <-> a packed buffer (colexicographic order)
__global__ void box_copy_kernel(double * __restrict__ field, double * __restrict__ packed, long long first, long long n,
                                long long n0, long long n1, long long E0, long long E01, long long origin, bool to_field) {
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
		const long long i = first + k, t = i / n0, i0 = i - t * n0, i2 = t / n1, i1 = t - i2 * n1;
		const long long at = origin + i0 + E0 * i1 + E01 * i2;
		if (to_field)
			field[at] = packed[k];
		else
			packed[k] = field[at];
	}
}

static void box_copy(fsb_vec_t v, double * d_packed, int64_t first, int64_t n, bool to_field) {
	if (n == 0)
		return;
	const fsb::box_shape & b = v->shape;
	const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
	box_copy_kernel<<<grid, 256, 0, v->ctx->stream>>>(v->d, d_packed, first, n, b.n[0], b.n[1], b.ext[0], b.ext[0] * b.ext[1],
	                                                 b.origin(), to_field);
	FSB_CUDA(cudaGetLastError());
}

int fsb_vec_upload(fsb_vec_t v, const double * host, int64_t n, int64_t offset) {
	return guarded([&] {
		FSB_REQUIRE(v && host && n >= 0 && offset >= 0 && offset + n <= v->n_owned, "upload range");
		flush(v->ctx);
		if (v->box) {
			double * tmp = nullptr;
			FSB_CUDA(cudaMalloc(&tmp, std::max<int64_t>(n, 1) * sizeof(double)));
			FSB_CUDA(cudaMemcpyAsync(tmp, host, n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
			box_copy(v, tmp, offset, n, true);
			FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
			cudaFree(tmp);
			return;
		}
		FSB_CUDA(cudaMemcpyAsync(v->d + offset, host, n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
		v->halo_valid = false;
	});
}

int fsb_vec_download(fsb_vec_t v, double * host, int64_t n, int64_t offset) {
	return guarded([&] {
		FSB_REQUIRE(v && host && n >= 0 && offset >= 0, "download range");
		flush(v->ctx);
		if (v->box) {
			FSB_REQUIRE(offset + n <= v->n_owned, "download range");
			double * tmp = nullptr;
			FSB_CUDA(cudaMalloc(&tmp, std::max<int64_t>(n, 1) * sizeof(double)));
			box_copy(v, tmp, offset, n, false);
			FSB_CUDA(cudaMemcpyAsync(host, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
			FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
			cudaFree(tmp);
			v->ctx->stats[FSB_STAT_HOST_SYNCS]++;
			return;
		}
		FSB_REQUIRE(offset + n <= v->n_owned + v->n_ghost, "download range");
		FSB_CUDA(cudaMemcpyAsync(host, v->d + offset, n * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
		v->ctx->stats[FSB_STAT_HOST_SYNCS]++;
	});
}

// ---- structured-grid ("narray") vectors, SURVEY 8(f) N3 ----

int fsb_vec_create_box(fsb_ctx_t c, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi, fsb_vec_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out && extents && lo && hi && dim >= 1 && dim <= 3, "bad arguments");
		fsb::box_shape b;
		for (int k = 0; k < dim; ++k) {
			FSB_REQUIRE(extents[k] >= 1 && lo[k] >= 0 && lo[k] <= hi[k] && hi[k] <= extents[k], "box: need 0 <= lo <= hi <= extent");
			b.ext[k] = extents[k];
			b.lo[k] = lo[k];
			b.n[k] = hi[k] - lo[k];
		}
		FSB_REQUIRE(b.storage() < (1LL << 31), "box: padded array exceeds int32 offsets");
		fsb::speculation_release(c);
		auto * v = new fsb_vec_s;
		v->ctx = c;
		v->box = true;
		v->shape = b;
		v->n_owned = b.dofs();
		v->n_ghost = b.storage() - b.dofs(); // boundary / padding entries: never touched by vector operations
		v->id = c->next_vec_id++;
		const size_t n = static_cast<size_t>(b.storage());
		FSB_CUDA(cudaMalloc(&v->d, std::max<size_t>(n, 2) * sizeof(double)));
		FSB_CUDA(cudaMemsetAsync(v->d, 0, std::max<size_t>(n, 2) * sizeof(double), c->stream));
		*out = v;
	});
}

// the whole padded array (boundary layers included), e.g. to place Dirichlet data
int fsb_vec_box_upload_all(fsb_vec_t v, const double * host) {
	return guarded([&] {
		FSB_REQUIRE(v && v->box && host, "bad arguments");
		flush(v->ctx);
		FSB_CUDA(cudaMemcpyAsync(v->d, host, v->shape.storage() * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
		v->halo_valid = false; // ghost planes (several ranks) are stale from here on
	});
}
int fsb_vec_box_download_all(fsb_vec_t v, double * host) {
	return guarded([&] {
		FSB_REQUIRE(v && v->box && host, "bad arguments");
		flush(v->ctx);
		FSB_CUDA(cudaMemcpyAsync(host, v->d, v->shape.storage() * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(v->ctx->stream));
	});
}

static void check_same(fsb_vec_t a, fsb_vec_t b) {
	FSB_REQUIRE(a && b, "null vector");
	FSB_REQUIRE(a->ctx == b->ctx, "vectors belong to different contexts");
	FSB_REQUIRE(a->n_owned == b->n_owned, "vector sizes differ");
	FSB_REQUIRE(a->box == b->box && (!a->box || a->shape == b->shape), "vectors live on different index sets");
}

static void check_coef(fsb_ctx_s * c, const fsb_coef & k) {
	FSB_REQUIRE(k.num >= 0 && k.num < fsb::MAX_SCALARS && k.den >= 0 && k.den < fsb::MAX_SCALARS &&
	                c->scalar_used[k.num] && c->scalar_used[k.den],
	            "coefficient names an unknown device scalar");
}

static void push_ew_c(int op, fsb_vec_t z, fsb_vec_t x, fsb_vec_t y, fsb_coef a, fsb_coef b) {
	FSB_REQUIRE(z, "null vector");
	if (x)
		check_same(z, x);
	if (y)
		check_same(z, y);
	check_coef(z->ctx, a);
	check_coef(z->ctx, b);
	// commutative statements: keep an aliased destination in the y operand
	if ((op == OP_LIN2 || op == OP_MUL) && z == x && z != y) {
		std::swap(x, y);
		std::swap(a, b);
	}
	pending p{};
	p.kind = pending::EW;
	p.op = op;
	p.z = z;
	p.x = x;
	p.y = y;
	p.a = a.scale;
	p.b = b.scale;
	if (a.num != 0 || a.den != 0) {
		p.a_num = a.num;
		p.a_den = a.den;
	}
	if (b.num != 0 || b.den != 0) {
		p.b_num = b.num;
		p.b_den = b.den;
	}
	z->halo_valid = false;
	enqueue(z->ctx, p);
}

static void push_ew(int op, fsb_vec_t z, fsb_vec_t x, fsb_vec_t y, double a, double b) {
	push_ew_c(op, z, x, y, fsb_coef{a, 0, 0}, fsb_coef{b, 0, 0});
}

int fsb_vec_copy(fsb_vec_t z, fsb_vec_t x) {
	return guarded([&] {
		if (z == x)
			return;
		push_ew(OP_SCALE, z, x, nullptr, 1.0, 0);
	});
}
int fsb_vec_set(fsb_vec_t z, double a) {
	return guarded([&] { push_ew(OP_SET, z, nullptr, nullptr, a, 0); });
}
int fsb_vec_scale(fsb_vec_t z, double a, fsb_vec_t x) {
	return guarded([&] { push_ew(OP_SCALE, z, x, nullptr, a, 0); });
}
int fsb_vec_add(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_LIN2, z, x, y, 1.0, 1.0); });
}
int fsb_vec_sub(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_LIN2, z, x, y, 1.0, -1.0); });
}
int fsb_vec_mul(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_MUL, z, x, y, 0, 0); });
}
int fsb_vec_div(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_DIV, z, x, y, 0, 0); });
}
int fsb_vec_recip(fsb_vec_t z, fsb_vec_t x) {
	return guarded([&] { push_ew(OP_RECIP, z, x, nullptr, 0, 0); });
}
int fsb_vec_linear_sum(fsb_vec_t z, double a, fsb_vec_t x, double b, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_LIN2, z, x, y, a, b); });
}
int fsb_vec_linear_sum_c(fsb_vec_t z, fsb_coef a, fsb_vec_t x, fsb_coef b, fsb_vec_t y) {
	return guarded([&] { push_ew_c(OP_LIN2, z, x, y, a, b); });
}
int fsb_vec_axpy(fsb_vec_t z, double a, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] { push_ew(OP_LIN2, z, x, y, a, 1.0); });
}
int fsb_vec_axpby(fsb_vec_t z, double a, double b, fsb_vec_t x) {
	return guarded([&] { push_ew(OP_LIN2, z, x, z, a, b); });
}
int fsb_vec_abs(fsb_vec_t z, fsb_vec_t x) {
	return guarded([&] { push_ew(OP_ABS, z, x, nullptr, 0, 0); });
}
int fsb_vec_add_scalar(fsb_vec_t z, fsb_vec_t x, double a) {
	return guarded([&] { push_ew(OP_ADDS, z, x, nullptr, a, 0); });
}

int fsb_vec_set_random(fsb_vec_t z, unsigned seed) {
	return guarded([&] {
		FSB_REQUIRE(z, "null vector");
		std::mt19937 gen(seed);
		std::uniform_real_distribution<double> dis(0., 1.);
		std::vector<double> h(static_cast<size_t>(z->n_owned));
		for (auto & v : h)
			v = dis(gen);
		if (z->box) { // dofs in colexicographic order, like the reference's loop over dofs()
			if (fsb_vec_upload(z, h.data(), static_cast<int64_t>(h.size()), 0) != FSB_OK)
				throw fsb::error(FSB_ERR_CUDA, fsb_last_error());
			return;
		}
		flush(z->ctx);
		if (!h.empty())
			FSB_CUDA(cudaMemcpyAsync(z->d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, z->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(z->ctx->stream));
		z->halo_valid = false;
	});
}

int fsb_vec_dump(fsb_vec_t x, const char * prefix) {
	return guarded([&] {
		FSB_REQUIRE(x && prefix, "bad arguments");
		std::vector<double> h(static_cast<size_t>(x->n_owned));
		flush(x->ctx);
		if (x->box) {
			if (fsb_vec_download(x, h.data(), static_cast<int64_t>(h.size()), 0) != FSB_OK)
				throw fsb::error(FSB_ERR_CUDA, fsb_last_error());
		}
		else if (!h.empty())
			FSB_CUDA(cudaMemcpyAsync(h.data(), x->d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, x->ctx->stream));
		FSB_CUDA(cudaStreamSynchronize(x->ctx->stream));
		std::ofstream f(std::string(prefix) + "-" + std::to_string(x->ctx->rank));
		for (double v : h)
			f << v << '\n';
	});
}

// ------------------------------------------------------------------ reductions

static void push_red(int op, fsb_vec_t x, fsb_vec_t y, double a, fsb_token_t * tok, const fsb_red_opts * o = nullptr) {
	FSB_REQUIRE(x && tok, "bad arguments");
	if (y)
		check_same(x, y);
	fsb_ctx_s * c = x->ctx;
	pending p{};
	p.kind = pending::RED;
	p.op = op;
	p.x = x;
	p.y = y;
	p.a = a;
	if (o) {
		FSB_REQUIRE(o->store >= 0 && o->store < fsb::MAX_SCALARS && c->scalar_used[o->store], "unknown device scalar");
		FSB_REQUIRE(o->halt_mode >= FSB_HALT_NEVER && o->halt_mode <= FSB_HALT_IF_LT, "bad halt mode");
		FSB_REQUIRE(o->halt_mode == FSB_HALT_NEVER || c->halt_armed, "halt test outside fsb_ctx_halt_arm/disarm");
		if (o->store > 0)
			p.store = o->store;
		p.halt_mode = o->halt_mode;
		p.halt_thr = o->halt_threshold;
		FSB_REQUIRE(o->n_post >= 0 && o->n_post <= FSB_MAX_POST_OPS, "too many scalar statements on one reduction");
		p.n_post = o->n_post;
		for (int k = 0; k < o->n_post; ++k) {
			const fsb_scalar_op & s = o->post[k];
			FSB_REQUIRE(s.op >= FSB_SOP_ADD && s.op <= FSB_SOP_COPY, "unknown scalar operation");
			for (fsb_scalar_t id : {s.dst, s.a, s.b})
				FSB_REQUIRE(id >= 0 && id < fsb::MAX_SCALARS && c->scalar_used[id], "scalar statement names an unknown device scalar");
			FSB_REQUIRE(s.dst > 0, "slot 0 is the constant 1.0");
			p.post[k] = s;
		}
	}
	p.token = new_token(c, fold_of(op));
	*tok = p.token;
	enqueue(c, p);
}

int fsb_vec_dot(fsb_vec_t x, fsb_vec_t y, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_DOT, x, y, 0, tok); });
}
int fsb_vec_sumsq(fsb_vec_t x, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_DOT, x, x, 0, tok); });
}
int fsb_vec_dot_opts(fsb_vec_t x, fsb_vec_t y, const fsb_red_opts * o, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_DOT, x, y, 0, tok, o); });
}

// ------------------------------------------------------------------ device scalars

int fsb_scalar_create(fsb_ctx_t c, fsb_scalar_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out, "bad arguments");
		for (int k = 1; k < fsb::MAX_SCALARS; ++k)
			if (!c->scalar_used[k]) {
				c->scalar_used[k] = true;
				*out = k;
				return;
			}
		throw fsb::error(FSB_ERR_STATE, "out of device scalars");
	});
}
int fsb_scalar_destroy(fsb_ctx_t c, fsb_scalar_t s) {
	return guarded([&] {
		FSB_REQUIRE(c && s > 0 && s < fsb::MAX_SCALARS && c->scalar_used[s], "unknown device scalar");
		flush(c); // queued statements may still name it
		c->scalar_used[s] = false;
	});
}
int fsb_scalar_set(fsb_ctx_t c, fsb_scalar_t s, double v) {
	return guarded([&] {
		FSB_REQUIRE(c && s > 0 && s < fsb::MAX_SCALARS && c->scalar_used[s], "unknown device scalar");
		flush(c);
		FSB_CUDA(cudaMemcpyAsync(c->d_scalars + s, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream)); // v is a stack temporary
	});
}
int fsb_scalar_get(fsb_ctx_t c, fsb_scalar_t s, double * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out && s >= 0 && s < fsb::MAX_SCALARS && c->scalar_used[s], "unknown device scalar");
		flush(c);
		FSB_CUDA(cudaMemcpyAsync(out, c->d_scalars + s, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		FSB_CUDA(cudaStreamSynchronize(c->stream));
		c->stats[FSB_STAT_HOST_SYNCS]++;
	});
}
int fsb_ctx_halt_arm(fsb_ctx_t c) {
	return guarded([&] {
		FSB_REQUIRE(c, "null context");
		flush(c);
		FSB_CUDA(cudaMemsetAsync(c->d_halt, 0, sizeof(int), c->stream));
		c->halt_armed = true;
	});
}
int fsb_ctx_halt_disarm(fsb_ctx_t c, int * was_halted) {
	return guarded([&] {
		FSB_REQUIRE(c, "null context");
		flush(c);
		if (was_halted) {
			FSB_CUDA(cudaMemcpyAsync(was_halted, c->d_halt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
			FSB_CUDA(cudaStreamSynchronize(c->stream));
		}
		FSB_CUDA(cudaMemsetAsync(c->d_halt, 0, sizeof(int), c->stream));
		c->halt_armed = false;
	});
}
int fsb_vec_asum(fsb_vec_t x, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_ASUM, x, nullptr, 0, tok); });
}
int fsb_vec_powsum(fsb_vec_t x, int p, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_POWSUM, x, nullptr, static_cast<double>(p), tok); });
}
int fsb_vec_amax(fsb_vec_t x, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_AMAX, x, nullptr, 0, tok); });
}
int fsb_vec_min(fsb_vec_t x, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_MIN, x, nullptr, 0, tok); });
}
int fsb_vec_max(fsb_vec_t x, fsb_token_t * tok) {
	return guarded([&] { push_red(RD_MAX, x, nullptr, 0, tok); });
}

int fsb_vec_global_size(fsb_vec_t x, int64_t * out) {
	return guarded([&] {
		FSB_REQUIRE(x && out, "bad arguments");
		fsb_ctx_s * c = x->ctx;
		if (c->nranks == 1) {
			*out = x->n_owned;
			return;
		}
		flush(c);
		long long v = 0;
		for (long long n : c->boot->gather<long long>(x->n_owned, c->nranks))
			v += n;
		*out = v;
	});
}

static void wait_token(fsb_ctx_t c, fsb_token_t tok) {
	FSB_REQUIRE(c, "null context");
	FSB_REQUIRE(tok > 0 && tok < c->next_token, "unknown reduction token");
	FSB_REQUIRE(tok + FSB_RED_RING > c->next_token - 1, "reduction token expired");
	flush(c, true);
	if (c->speculate)
		fsb::speculate_after_flush(c); // the kernel the host will most likely ask for next goes in behind the one it waits for
	const auto t_begin = std::chrono::steady_clock::now();
	const int slot = static_cast<int>(tok % FSB_RED_RING);
	if (c->nranks > 1 && !c->d_xrank) {
		FSB_CUDA(cudaEventSynchronize(c->token_event[slot]));
	}
	else {
		// the producing kernel publishes value and token as one flagged word in mapped pinned memory
		const fsb::ll_word * w = c->h_ll + slot;
		const uint32_t flag = static_cast<uint32_t>(tok);
		double v;
		unsigned spins = 0;
		while (!fsb::host_ll_load(w, flag, &v)) {
			_mm_pause();
			if ((++spins & 0xfffffu) == 0) { // every ~1M spins make sure the stream is still alive
				cudaError_t q = cudaStreamQuery(c->stream);
				if (q != cudaSuccess && q != cudaErrorNotReady)
					FSB_CUDA(q);
				if (q == cudaSuccess && !fsb::host_ll_load(w, flag, &v))
					throw fsb::error(FSB_ERR_STATE, "reduction result never arrived");
			}
		}
		c->h_results[slot] = v;
		if (fsb::speculation_failed(c))
			throw fsb::error(FSB_ERR_STATE, "a kernel launched ahead of its coefficients (FSB_OPT_SPECULATE) never heard from the host");
		if (c->h_xrank_error && *(volatile int *)c->h_xrank_error) {
			static const char * what[] = {"", "a reduction's all-reduce", "the acknowledgement of a ghost landing buffer",
			                              "ghost entries"};
			const int word = *(volatile int *)c->h_xrank_error, code = word & 15;
			std::string detail;
			if (code == 1) // which reduction, whose value (ew_kernels.cuh: xrank_allreduce)
				detail = " (reduction " + std::to_string(word >> 8) + ", rank " + std::to_string((word >> 4) & 15) + "'s value; this rank has issued " +
				         std::to_string(c->next_token - 1) + ")";
			throw fsb::error(FSB_ERR_STATE, std::string("rank ") + std::to_string(c->rank) + ": a kernel timed out waiting for " +
			                                    what[code >= 1 && code <= 3 ? code : 0] + " from a peer" + detail);
		}
	}
	c->stats[FSB_STAT_HOST_SYNCS]++;
	c->stats[FSB_STAT_WAIT_NS] +=
		std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
}

int fsb_red_wait(fsb_ctx_t c, fsb_token_t tok) {
	return guarded([&] { wait_token(c, tok); });
}

int fsb_red_get(fsb_ctx_t c, fsb_token_t tok, double * out) {
	return guarded([&] {
		FSB_REQUIRE(out, "null output");
		wait_token(c, tok);
		*out = *(volatile double *)(c->h_results + tok % FSB_RED_RING);
	});
}

// ------------------------------------------------------------------ matrix

int fsb_parcsr_create(fsb_ctx_t c, int64_t n_global, const int64_t * row_part, const int64_t * rowptr,
                      const int64_t * col, const double * val, fsb_parcsr_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && row_part && rowptr && out, "bad arguments");
		*out = fsb_parcsr_create_impl(c, n_global, row_part, rowptr, col, val);
	});
}

int fsb_parcsr_create_stencil(fsb_ctx_t c, int kind, int64_t nx, int64_t ny, int64_t nz, double diag_shift,
                              double scale, fsb_parcsr_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && out, "bad arguments");
		*out = fsb_parcsr_create_stencil_impl(c, kind, nx, ny, nz, diag_shift, scale);
	});
}

int fsb_parcsr_create_box_stencil(fsb_ctx_t c, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi,
                                  double center, const double * off, fsb_parcsr_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && extents && lo && hi && off && out, "bad arguments");
		*out = fsb_parcsr_create_box_stencil_impl(c, dim, extents, lo, hi, center, off);
	});
}

int fsb_parcsr_create_box_fvm(fsb_ctx_t c, int dim, const int64_t * extents, const int64_t * lo, const int64_t * hi, double beta,
                              double alpha, double vol, const double * kface, const double * a, const double * const * bface,
                              fsb_parcsr_t * out) {
	return guarded([&] {
		FSB_REQUIRE(c && extents && lo && hi && out, "bad arguments");
		*out = fsb_parcsr_create_box_fvm_impl(c, dim, extents, lo, hi, beta, alpha, vol, kface, a, bface);
	});
}

int fsb_parcsr_destroy(fsb_parcsr_t A) {
	return guarded([&] {
		if (A)
			fsb_parcsr_destroy_impl(A);
	});
}

int64_t fsb_parcsr_local_rows(fsb_parcsr_t A) { return A->n_local; }
int64_t fsb_parcsr_global_rows(fsb_parcsr_t A) { return A->n_global; }
int64_t fsb_parcsr_num_ghosts(fsb_parcsr_t A) { return A->n_ghost; }
int64_t fsb_parcsr_row_begin(fsb_parcsr_t A) { return A->row_begin; }
int64_t fsb_parcsr_local_nnz(fsb_parcsr_t A, int which) { return which == 0 ? A->diag.nnz : A->offd.nnz; }
int64_t fsb_parcsr_info(fsb_parcsr_t A, int key) {
	switch (key) {
	case FSB_INFO_WINDOW_FORMAT: return A->diag.lcol != nullptr;
	case FSB_INFO_ROW_BLOCKS: return A->diag.n_blk;
	case FSB_INFO_WINDOW_X: return A->diag.win_xcap;
	case FSB_INFO_FUSED_HALO: return A->halo_p2p != nullptr && A->diag.has_offd_map && !A->nbrs.empty();
	case FSB_INFO_WIDE_OFFSETS: return A->diag.wide;
	case FSB_INFO_VALUE_DICTIONARY: return A->diag.vidx ? A->diag.n_dict : 0;
	case FSB_INFO_DICTIONARY_SLOTS: return A->diag.vidx ? A->diag.dict_slots : 0;
	default: return -1;
	}
}

int fsb_parcsr_download(fsb_parcsr_t A, int which, int64_t * rowptr, int32_t * col, double * val) {
	return guarded([&] {
		FSB_REQUIRE(A && (which == 0 || which == 1), "bad arguments");
		fsb_ctx_s * c = A->ctx;
		flush(c);
		const csr_block & B = which == 0 ? A->diag : A->offd;
		if (rowptr) {
			// expand to one offset per local row (the offd block is stored over a row subset)
			std::vector<int64_t> rp(static_cast<size_t>(B.n_rows) + 1, 0);
			if (B.n_rows > 0) {
				if (B.wide) {
					FSB_CUDA(cudaMemcpy(rp.data(), B.rowptr, rp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
				}
				else {
					std::vector<int32_t> t(rp.size());
					FSB_CUDA(cudaMemcpy(t.data(), B.rowptr, t.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
					std::copy(t.begin(), t.end(), rp.begin());
				}
			}
			if ((B.row_ids && !A->box) || which == 1) { // (a structured-grid operator's row ids are storage offsets; its rows are the dofs in order)
				std::vector<int32_t> ids(static_cast<size_t>(B.n_rows));
				if (B.n_rows > 0)
					FSB_CUDA(cudaMemcpy(ids.data(), B.row_ids, ids.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
				std::vector<int64_t> cnt(static_cast<size_t>(A->n_local), 0);
				for (int64_t k = 0; k < B.n_rows; ++k)
					cnt[ids[k]] = rp[k + 1] - rp[k];
				rowptr[0] = 0;
				for (int64_t r = 0; r < A->n_local; ++r)
					rowptr[r + 1] = rowptr[r] + cnt[r];
			}
			else {
				std::copy(rp.begin(), rp.end(), rowptr);
			}
		}
		if (col && B.nnz > 0)
			FSB_CUDA(cudaMemcpy(col, B.col, B.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
		if (val && B.nnz > 0)
			FSB_CUDA(cudaMemcpy(val, B.val, B.nnz * sizeof(double), cudaMemcpyDeviceToHost));
	});
}

int fsb_parcsr_download_colmap(fsb_parcsr_t A, int64_t * colmap) {
	return guarded([&] {
		FSB_REQUIRE(A && colmap, "bad arguments");
		std::copy(A->colmap.begin(), A->colmap.end(), colmap);
	});
}

int fsb_parcsr_spmv(fsb_parcsr_t A, fsb_vec_t x, fsb_vec_t y) {
	return guarded([&] {
		FSB_REQUIRE(A && x && y, "null argument");
		FSB_REQUIRE(x != y && x->d != y->d, "spmv: x and y must be different vectors");
		FSB_REQUIRE(x->ctx == A->ctx && y->ctx == A->ctx, "spmv: context mismatch");
		FSB_REQUIRE(x->n_owned == A->n_local && y->n_owned == A->n_local, "spmv: vector length != local rows");
		if (A->box)
			FSB_REQUIRE(x->box && y->box && x->shape == A->shape && y->shape == A->shape,
			            "spmv: a structured-grid operator needs vectors on its own box");
		else
			FSB_REQUIRE(!x->box && !y->box, "spmv: structured-grid vectors need a structured-grid operator");
		FSB_REQUIRE(x->n_ghost >= A->n_ghost, "spmv: x has too few ghost entries for this matrix");
		pending p{};
		p.kind = pending::SPMV;
		p.op = -1;
		p.A = A;
		p.x = x;
		p.y = y;
		p.z = y;
		enqueue(A->ctx, p);
	});
}

int fsb_parcsr_extract_dinv(fsb_parcsr_t A, fsb_vec_t d) {
	return guarded([&] {
		FSB_REQUIRE(A && d && d->n_owned == A->n_local, "extract_dinv: bad arguments");
		FSB_REQUIRE(!A->box, "extract_dinv: not available for structured-grid operators");
		flush(A->ctx);
		extract_dinv(A, d->d);
		d->halo_valid = false;
	});
}

int fsb_parcsr_jacobi_relax(fsb_parcsr_t A, double omega, int64_t nrelax, fsb_vec_t b, fsb_vec_t x, fsb_vec_t tmp) {
	return guarded([&] {
		FSB_REQUIRE(A && b && x && tmp, "null argument");
		FSB_REQUIRE(!A->box, "jacobi_relax: not available for structured-grid operators");
		fsb_parcsr_jacobi_relax_impl(A, omega, nrelax, b, x, tmp);
	});
}

int fsb_parcsr_halo_exchange(fsb_parcsr_t A, fsb_vec_t x) {
	return guarded([&] {
		FSB_REQUIRE(A && x, "null argument");
		FSB_REQUIRE(A->box == x->box, "halo_exchange: a structured-grid operator takes structured-grid vectors (and only those)");
		halo_exchange(A, x);
	});
}

} // extern "C"
