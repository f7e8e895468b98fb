// Fused element-wise / reduction kernels, one instantiation per registered program.
//
// HBM-bound streaming kernels (<= 0.125 flop/B): 128-bit loads and stores, a grid-stride loop
// with EW_CTAS_PER_SM x 148 CTAs so every SM keeps >= 2048 threads of loads in flight,
// warp-shuffle + shared-memory block reduction, and a "last CTA" epilogue that folds the
// per-CTA partials in a fixed order (bitwise reproducible) and publishes the result to
// pinned host memory so the host can poll instead of synchronising the stream.
#pragma once
#ifndef __CUDACC_RTC__ // NVRTC (csrc/jit.cu) has the CUDA built-ins without headers
#include <cuda_runtime.h>
#endif

#include "program.h"

namespace fsb {

constexpr int EW_BLOCK = 256; // threads per CTA of every element-wise kernel

struct red_out {
	double * d_value; // device result slot
	void * h_ll; // mapped host slot (16-byte flagged word, see ll_store) or nullptr (NCCL transport finishes the job)
	long long token;
	// device-scalar solvers (fsb.h, "device scalars"): keep the value on the device for later
	// coefficients, and raise the context's halt flag when the convergence test passes
	double * d_extra; // scalar slot or nullptr
	int * halt; // flag to raise, used when halt_mode != 0
	double halt_thr;
	int halt_mode; // 0 none, 1: sqrt(value) < thr, 2: value < thr
	// scalar statements evaluated by the publishing thread once the value is stored: slots[dst] = slots[a] op slots[b]
	double * slots;
	int n_post;
	unsigned char post[8][4]; // {op, dst, a, b}
};

// 0 add, 1 sub, 2 mul, 3 div, 4 copy (fsb.h: FSB_SOP_*), one IEEE rounding each
__device__ __forceinline__ void run_post_ops(const red_out & r) {
	for (int k = 0; k < r.n_post; ++k) {
		const double a = r.slots[r.post[k][2]], b = r.slots[r.post[k][3]];
		const int op = r.post[k][0];
		r.slots[r.post[k][1]] = op == 0 ? __dadd_rn(a, b) : op == 1 ? __dadd_rn(a, -b) : op == 2 ? __dmul_rn(a, b) : op == 3 ? __ddiv_rn(a, b) : a;
	}
}

// Flagged 16-byte words ("LL" protocol, as NCCL's low-latency path): a double travels as two 8-byte halves
// {32 data bits, 32 flag bits}.  8-byte stores are single-copy atomic over NVLink and PCIe, so a reader that finds the
// expected flag in BOTH halves has the data -- no fence between payload and flag, no separate flag word: the cost of
// a cross-GPU hand-over is one store latency instead of store + system fence (~6 us under load, measured in
// profiles/r2_timeline_*.txt) + flag store.
__device__ __forceinline__ void ll_store(void * dst, double v, unsigned flag) {
	const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(static_cast<unsigned>(bits)), "r"(flag),
	             "r"(static_cast<unsigned>(bits >> 32)), "r"(flag)
	             : "memory");
}
__device__ __forceinline__ bool ll_try_load(const void * src, unsigned flag, double & v) {
	unsigned lo, f0, hi, f1;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
	v = __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo));
	return f0 == flag && f1 == flag;
}
// spin until the word carries `flag`; false (and NaN) after ~10 s: the writer is gone
__device__ __forceinline__ bool ll_load(const void * src, unsigned flag, double & v) {
	if (ll_try_load(src, flag, v))
		return true;
	const long long t0 = clock64();
	while (!ll_try_load(src, flag, v)) {
		if (clock64() - t0 > 20000000000LL) {
			v = __longlong_as_double(0x7ff8000000000000LL);
			return false;
		}
	}
	return true;
}

// Cross-rank all-reduce of one scalar over NVLink peer memory, executed by the CTA that finishes
// a reduction: every rank stores its partial as a flagged word into slot [token % ring][me] of every
// rank's mailbox, waits for the P words of its own mailbox and folds the P values in rank order --
// the same bits on every rank, no NCCL launch, no extra kernel, no fence.  Mailboxes are cudaIpc-mapped.
constexpr int XRANK_MAX = 8;
constexpr int XRANK_RING = 256; // == FSB_RED_RING
struct xrank_info {
	int me, nranks;
	double * mailbox[XRANK_MAX]; // [ring][nranks] flagged 16-byte words
	int * error_flag; // mapped host int: set when a peer never showed up
};

// Device-side timeline (FSB_OPT_TIMELINE): one slot of TL_WORDS 64-bit words per kernel launch, written with
// %globaltimer (ns).  [0] earliest CTA start (atomicMin, initialised to ~0), [1] latest CTA end, [2] kind (host),
// [3] ghost push complete, [4] first CTA has its ghosts, [5] longest wait of a CTA for them (ns), [6] all-reduce begins,
// [7] result published, [8] push: acknowledgements seen, [9] push: stores issued (latest CTA each); rest spare.
constexpr int TL_WORDS = 16;
__device__ __forceinline__ unsigned long long tl_now() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long * tl) {
	if (tl && threadIdx.x == 0)
		atomicMin(tl, tl_now());
}
__device__ __forceinline__ void tl_end(unsigned long long * tl) {
	if (tl && threadIdx.x == 0)
		atomicMax(tl + 1, tl_now());
}
// words 0, 4, 6 are minima (slots are initialised to ~0 there), the others maxima (initialised to 0)
__device__ __forceinline__ void tl_mark_min(unsigned long long * tl, int word) {
	if (tl)
		atomicMin(tl + word, tl_now());
}
__device__ __forceinline__ void tl_mark_max(unsigned long long * tl, int word) {
	if (tl)
		atomicMax(tl + word, tl_now());
}

// Programmatic dependent launch: kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while the previous kernel of the stream is still draining (its launch latency and prologue hide behind
// that tail); this is the point where they wait for the previous kernel's memory to be complete and visible.  Without
// the attribute it returns at once.
__device__ __forceinline__ void grid_dependency_wait() {
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Late-bound coefficients ("armed" launches, fuser.cu): a kernel whose coefficients the host computes from a reduction
// it is still waiting for can be launched AHEAD of that result -- it becomes resident behind its predecessor and waits
// here for a word in mapped host memory instead of for a launch: {sequence number << 2 | verdict}.  GO: the coefficients
// are in late_host::s; ABORT: the host went another way, the kernel leaves without touching anything.  CTA 0 watches
// the host word and republishes it in device memory for the other CTAs (one reader on PCIe, the rest in L2).
constexpr unsigned LATE_GO = 1, LATE_ABORT = 2;
struct late_host { // mapped pinned host memory, written by the host
	unsigned long long word;
	double s[MAXSC];
};
struct late_dev { // device memory, written by CTA 0 (one reader on PCIe; every other CTA reads L2)
	unsigned long long word;
	double s[MAXSC];
};
// thread 0 of every CTA; returns the verdict and, for LATE_GO, leaves the first NS coefficients in sc[] (shared memory)
template<int NS>
__device__ __forceinline__ unsigned late_wait(const late_host * host, late_dev * dev, unsigned seq, bool leader, int * error_flag,
                                              double * sc) {
	const long long t0 = clock64();
	for (;;) {
		unsigned long long w;
		if (leader)
			asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(&host->word) : "memory");
		else
			asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(&dev->word) : "memory");
		if (leader && static_cast<unsigned>(w >> 2) != seq && clock64() - t0 > 20000000000LL) {
			*reinterpret_cast<volatile int *>(error_flag) = 4; // ~10 s: the host is gone; fail loudly instead of hanging
			w = (static_cast<unsigned long long>(seq) << 2) | LATE_ABORT;
		}
		if (static_cast<unsigned>(w >> 2) != seq)
			continue;
		const unsigned verdict = static_cast<unsigned>(w & 3);
		__threadfence(); // the word first, then what it guards
		if (verdict == LATE_GO) {
			// the host stored them before the word (release).  All loads are issued before the first use: the leader's
			// cross PCIe, one latency for the lot; it then passes them on
			double v[NS > 0 ? NS : 1];
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				if (leader)
					asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v[k]) : "l"(&host->s[k]));
				else
					asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v[k]) : "l"(&dev->s[k]));
			}
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				sc[k] = v[k];
				if (leader)
					asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(&dev->s[k]), "d"(v[k]) : "memory");
			}
		}
		if (leader) {
			__threadfence(); // coefficients before the word
			asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(&dev->word), "l"(w) : "memory");
		}
		return verdict;
	}
}

struct ew_args {
	double * v[MAXV];
	double s[MAXSC];
	// coefficient k is s[k] when snum[k] < 0, else s[k] * sdev[snum[k]] / sdev[sden[k]] read at kernel start
	const double * sdev;
	signed char snum[MAXSC], sden[MAXSC];
	const int * halt; // non-null while a device-scalar solver is running: skip all element work once set
	// structured-grid vectors (fsb_vec_create_box): element i of the n dofs is the interior point
	// (i % bn0, (i / bn0) % bn1, i / (bn0 bn1)) of a padded array, stored at boff + i0 + bE0 i1 + bE01 i2.
	// bn0 == 0: plain contiguous vectors.  Box layouts always run through ew_interp_kernel (one kernel for
	// every program), so the compiled instantiations keep their register budgets.
	long long bn0, bn1, bE0, bE01, boff;
	long long n;
	double * partials; // [MAXR][MAX_RED_BLOCKS]
	unsigned * counter;
	int partial_stride;
	const xrank_info * xr; // nullptr on one rank
	unsigned long long * tl; // timeline slot of this launch or nullptr
	// armed launch: coefficients arrive through `late` (mapped host memory) once word == late_seq << 2 | LATE_GO
	const late_host * late; // nullptr: coefficients are s[] above
	late_dev * late_relay; // device copy of word and coefficients for the CTAs other than 0
	int * late_error; // mapped host int, set when the host never answered
	unsigned late_seq;
	red_out r[MAXR];
};

template<int FOLD>
__device__ __forceinline__ double fold(double a, double b) {
	if constexpr (FOLD == 0)
		return a + b;
	else if constexpr (FOLD == 1)
		return fmax(a, b);
	else
		return fmin(a, b);
}

template<int FOLD>
__device__ __forceinline__ double fold_identity() {
	if constexpr (FOLD == 0)
		return 0.0;
	else if constexpr (FOLD == 1)
		return -__longlong_as_double(0x7ff0000000000000LL); // -inf
	else
		return __longlong_as_double(0x7ff0000000000000LL); // +inf
}

template<int FOLD>
__device__ __forceinline__ double warp_fold(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v = fold<FOLD>(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// Fold one value per thread over the CTA; result valid in thread 0. `scratch` holds >= 32 doubles.
template<int FOLD>
__device__ __forceinline__ double block_fold(double v, double * scratch) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v = warp_fold<FOLD>(v);
	__syncthreads(); // scratch may still be read from a previous call
	if (lane == 0)
		scratch[warp] = v;
	__syncthreads();
	const int nwarps = (blockDim.x + 31) >> 5;
	if (warp == 0) {
		v = lane < nwarps ? scratch[lane] : fold_identity<FOLD>();
		v = warp_fold<FOLD>(v);
	}
	return v;
}

// all threads of the CTA call this; `local` is valid in thread 0; result valid in thread 0
template<int FOLD>
__device__ __forceinline__ double xrank_allreduce(const xrank_info * xr, double local, long long token, double * scratch) {
	const int P = xr->nranks, me = xr->me;
	const size_t slot = static_cast<size_t>(token % XRANK_RING);
	const unsigned flag = static_cast<unsigned>(token);
	__syncthreads();
	if (threadIdx.x == 0)
		scratch[0] = local;
	__syncthreads();
	local = scratch[0];
	__syncthreads();
	if (threadIdx.x < P) {
		const int q = threadIdx.x;
		ll_store(xr->mailbox[q] + (slot * P + me) * 2, local, flag);
		double v;
		if (!ll_load(xr->mailbox[me] + (slot * P + q) * 2, flag, v))
			*reinterpret_cast<volatile int *>(xr->error_flag) = 1 + 16 * q + 256 * static_cast<int>(token & 0x3fffff);
		scratch[q] = v;
	}
	__syncthreads();
	double r = fold_identity<FOLD>();
	if (threadIdx.x == 0)
		for (int q = 0; q < P; ++q)
			r = fold<FOLD>(r, scratch[q]);
	return r;
}

// publish a finished reduction: device slot, then the mapped host slot (thread 0 only)
__device__ __forceinline__ void publish(const red_out & r, double t) {
	*r.d_value = t;
	if (r.d_extra)
		*r.d_extra = t;
	if (r.halt_mode != 0 && (r.halt_mode == 1 ? sqrt(t) : t) < r.halt_thr)
		*reinterpret_cast<volatile int *>(r.halt) = 1;
	run_post_ops(r);
	if (r.h_ll) // value and token in one flagged word: the host polls it, no fence needed
		ll_store(r.h_ll, t, static_cast<unsigned>(r.token));
}

template<class PT, int I, class SC>
__device__ __forceinline__ void exec_stmt(double (&v)[MAXV], double (&acc)[MAXR], const SC & sc) {
	constexpr stmt S = PT::value.st[I];
	// products and sums are rounded separately on purpose (see fsb.h): the reference's
	// task bodies (vectors/operations/topo_tasks.hh) compiled for baseline x86-64 do the same.
	if constexpr (S.op == OP_SET)
		v[S.z] = sc[S.a];
	else if constexpr (S.op == OP_SCALE)
		v[S.z] = __dmul_rn(v[S.x], sc[S.a]);
	else if constexpr (S.op == OP_LIN2)
		v[S.z] = __dadd_rn(__dmul_rn(sc[S.a], v[S.x]), __dmul_rn(sc[S.b], v[S.y]));
	else if constexpr (S.op == OP_MUL)
		v[S.z] = __dmul_rn(v[S.x], v[S.y]);
	else if constexpr (S.op == OP_DIV)
		v[S.z] = __ddiv_rn(v[S.x], v[S.y]);
	else if constexpr (S.op == OP_RECIP)
		v[S.z] = __ddiv_rn(1.0, v[S.x]);
	else if constexpr (S.op == OP_ABS)
		v[S.z] = fabs(v[S.x]);
	else if constexpr (S.op == OP_ADDS)
		v[S.z] = __dadd_rn(v[S.x], sc[S.a]);
	else if constexpr (S.op == RD_DOT)
		acc[S.z] = fma(v[S.x], v[S.y], acc[S.z]);
	else if constexpr (S.op == RD_ASUM)
		acc[S.z] += fabs(v[S.x]);
	else if constexpr (S.op == RD_AMAX)
		acc[S.z] = fmax(acc[S.z], fabs(v[S.x]));
	else if constexpr (S.op == RD_MIN)
		acc[S.z] = fmin(acc[S.z], v[S.x]);
	else if constexpr (S.op == RD_MAX)
		acc[S.z] = fmax(acc[S.z], v[S.x]);
	else if constexpr (S.op == RD_POWSUM)
		acc[S.z] += pow(v[S.x], sc[S.a]);
}

template<class PT, class SC>
__device__ __forceinline__ void exec_all(double (&v)[MAXV], double (&acc)[MAXR], const SC & sc) {
	[&]<int... I>(iseq<I...>) {
		(exec_stmt<PT, I>(v, acc, sc), ...);
	}(make_iseq<PT::value.n>{});
}

// the element loops of a program; SC is either the kernel-parameter array (immediate coefficients,
// served by the constant bank) or a register copy with the device-resident coefficients resolved
template<class PT, class SC>
__device__ __forceinline__ void run_elements(const ew_args & a, const SC & sc, double (&acc)[MAXR]) {
	constexpr program P = PT::value;
	const long long n2 = a.n >> 1; // number of double2 packets
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += stride) {
		double lo[MAXV], hi[MAXV];
#pragma unroll
		for (int k = 0; k < P.nv; ++k) {
			if (P.load_mask & (1u << k)) {
				const double2 t = reinterpret_cast<const double2 *>(a.v[k])[i];
				lo[k] = t.x;
				hi[k] = t.y;
			}
		}
		exec_all<PT>(lo, acc, sc);
		exec_all<PT>(hi, acc, sc);
#pragma unroll
		for (int k = 0; k < P.nv; ++k) {
			if (P.store_mask & (1u << k))
				reinterpret_cast<double2 *>(a.v[k])[i] = make_double2(lo[k], hi[k]);
		}
	}
	if ((a.n & 1) && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { // odd tail element
		const long long i = a.n - 1;
		double t[MAXV];
#pragma unroll
		for (int k = 0; k < P.nv; ++k)
			if (P.load_mask & (1u << k))
				t[k] = a.v[k][i];
		exec_all<PT>(t, acc, sc);
#pragma unroll
		for (int k = 0; k < P.nv; ++k)
			if (P.store_mask & (1u << k))
				a.v[k][i] = t[k];
	}
}

// fold kind of reduction output R of program P
template<class PT, int R>
constexpr int red_fold() {
	for (int i = 0; i < PT::value.n; ++i)
		if (is_reduction(PT::value.st[i].op) && PT::value.st[i].z == R)
			return fold_of(PT::value.st[i].op);
	return 0;
}

// the same loop over the interior sub-box of padded arrays (dofs of an narray mesh, reference
// examples/poisson/mesh.hh:157-161 + index_util.hh:33-71): scalar 8-byte accesses, consecutive threads on
// consecutive x, rows of the box need not be 16-byte aligned
template<class PT, class SC>
__device__ __forceinline__ void run_elements_box(const ew_args & a, const SC & sc, double (&acc)[MAXR]) {
	constexpr program P = PT::value;
	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
		const long long t = i / a.bn0, i0 = i - t * a.bn0, i2 = t / a.bn1, i1 = t - i2 * a.bn1;
		const long long at = a.boff + i0 + a.bE0 * i1 + a.bE01 * i2;
		double v[MAXV];
#pragma unroll
		for (int k = 0; k < P.nv; ++k)
			if (P.load_mask & (1u << k))
				v[k] = a.v[k][at];
		exec_all<PT>(v, acc, sc);
#pragma unroll
		for (int k = 0; k < P.nv; ++k)
			if (P.store_mask & (1u << k))
				a.v[k][at] = v[k];
	}
}

// Body of every program kernel.  DEV: some coefficients live in device memory (and the halt flag is
// honoured); BOX: structured-grid layout.  The ahead-of-time instantiations below use BOX = false only (box
// layouts cost registers); the run-time compiler (csrc/jit.cu) instantiates whatever a statement group needs.
template<class PT, bool DEV, bool BOX>
__device__ __forceinline__ void ew_program_body(const ew_args & a) {
	constexpr program P = PT::value;
	grid_dependency_wait();
	tl_begin(a.tl);
	auto run = [&](const auto & sc, double (&acc)[MAXR]) {
		if constexpr (BOX)
			run_elements_box<PT>(a, sc, acc);
		else
			run_elements<PT>(a, sc, acc);
	};
	double acc[MAXR];
	[&]<int... R>(iseq<R...>) {
		((acc[R] = fold_identity<red_fold<PT, R>()>()), ...);
	}(make_iseq<(P.nr > 0 ? P.nr : 0)>{});

	if constexpr (DEV) {
		if (!(a.halt && *a.halt)) { // once halted, vectors stay as they are; reductions publish identities
			double sc[MAXSC];
#pragma unroll
			for (int k = 0; k < P.ns; ++k)
				sc[k] = a.snum[k] < 0 ? a.s[k] : __dmul_rn(a.s[k], __ddiv_rn(a.sdev[a.snum[k]], a.sdev[a.sden[k]]));
			run(sc, acc);
		}
	}
	else if (a.late) { // armed launch: wait for the host's verdict, then take the coefficients it left
		__shared__ unsigned verdict;
		__shared__ double late_sc[MAXSC];
		if (threadIdx.x == 0)
			verdict = late_wait<P.ns>(a.late, a.late_relay, a.late_seq, blockIdx.x == 0, a.late_error, late_sc);
		__syncthreads();
		if (verdict != LATE_GO)
			return; // nothing has been touched
		double sc[MAXSC];
#pragma unroll
		for (int k = 0; k < P.ns; ++k)
			sc[k] = late_sc[k];
		run(sc, acc);
	}
	else
		run(a.s, acc);

	if constexpr (P.nr > 0) {
		__shared__ double scratch[32];
		__shared__ bool is_last;
		[&]<int... R>(iseq<R...>) {
			((acc[R] = block_fold<red_fold<PT, R>()>(acc[R], scratch)), ...);
		}(make_iseq<P.nr>{});
		if (threadIdx.x == 0) {
#pragma unroll
			for (int r = 0; r < P.nr; ++r)
				a.partials[r * a.partial_stride + blockIdx.x] = acc[r];
			__threadfence();
			const unsigned ticket = atomicAdd(a.counter, 1u);
			is_last = (ticket == gridDim.x - 1);
		}
		__syncthreads();
		if (is_last) {
			__threadfence();
			// fold every reduction's partials (thread 0 ends up with the rank-local values) ...
			__shared__ double local[MAXR], gathered[MAXR][XRANK_MAX];
			[&]<int... R>(iseq<R...>) {
				(([&] {
					 constexpr int F = red_fold<PT, R>();
					 double t = fold_identity<F>();
					 for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x)
						 t = fold<F>(t, __ldcg(&a.partials[R * a.partial_stride + b]));
					 t = block_fold<F>(t, scratch);
					 if (threadIdx.x == 0)
						 local[R] = t;
				 }()),
				 ...);
			}(make_iseq<P.nr>{});
			__syncthreads();
			if (threadIdx.x == 0)
				tl_mark_min(a.tl, 6);
			if (a.xr) {
				// ... exchange ALL of them with the other ranks at once: thread (R, q) stores reduction R's value into
				// rank q's mailbox and picks up q's value from this rank's, so k reductions cost one round trip
				const int np = a.xr->nranks, me = a.xr->me;
				if (static_cast<int>(threadIdx.x) < np * P.nr) {
					const int R = threadIdx.x / np, q = threadIdx.x % np;
					const long long token = a.r[R].token;
					const size_t slot = static_cast<size_t>(token % XRANK_RING);
					ll_store(a.xr->mailbox[q] + (slot * np + me) * 2, local[R], static_cast<unsigned>(token));
					double v;
					if (!ll_load(a.xr->mailbox[me] + (slot * np + q) * 2, static_cast<unsigned>(token), v))
						*reinterpret_cast<volatile int *>(a.xr->error_flag) = 1 + 16 * q + 256 * static_cast<int>(token & 0x3fffff);
					gathered[R][q] = v;
				}
				__syncthreads();
			}
			if (threadIdx.x == 0) {
				[&]<int... R>(iseq<R...>) {
					(([&] {
						 constexpr int F = red_fold<PT, R>();
						 double t = local[R];
						 if (a.xr) { // rank order: the same bits on every rank
							 t = fold_identity<F>();
							 for (int q = 0; q < a.xr->nranks; ++q)
								 t = fold<F>(t, gathered[R][q]);
						 }
						 publish(a.r[R], t);
					 }()),
					 ...);
				}(make_iseq<P.nr>{});
				tl_mark_max(a.tl, 7);
				*a.counter = 0u;
			}
		}
	}
	tl_end(a.tl);
}

template<class PT, bool DEV = false>
__global__ void __launch_bounds__(EW_BLOCK) ew_program_kernel(const __grid_constant__ ew_args a) {
	ew_program_body<PT, DEV, false>(a);
}

// ---------------------------------------------------------------------------------------------
// Generic program kernel: the same semantics as ew_program_kernel for ANY canonical program, with
// the program passed at run time.  Slots are indexed dynamically, so they live in local memory
// (L1-resident) instead of registers: slower per element than a registered instantiation, but one
// pass over HBM instead of one pass per statement.  Used for statement groups that have no
// compile-time instantiation (fuser.cu).
struct interp_args {
	ew_args a;
	program p;
};

__device__ __forceinline__ void interp_exec(const program & P, const double * sc, double * v, double * acc) {
	for (int i = 0; i < P.n; ++i) {
		const stmt S = P.st[i];
		switch (S.op) {
		case OP_SET: v[S.z] = sc[S.a]; break;
		case OP_SCALE: v[S.z] = __dmul_rn(v[S.x], sc[S.a]); break;
		case OP_LIN2: v[S.z] = __dadd_rn(__dmul_rn(sc[S.a], v[S.x]), __dmul_rn(sc[S.b], v[S.y])); break;
		case OP_MUL: v[S.z] = __dmul_rn(v[S.x], v[S.y]); break;
		case OP_DIV: v[S.z] = __ddiv_rn(v[S.x], v[S.y]); break;
		case OP_RECIP: v[S.z] = __ddiv_rn(1.0, v[S.x]); break;
		case OP_ABS: v[S.z] = fabs(v[S.x]); break;
		case OP_ADDS: v[S.z] = __dadd_rn(v[S.x], sc[S.a]); break;
		case RD_DOT: acc[S.z] = fma(v[S.x], v[S.y], acc[S.z]); break;
		case RD_ASUM: acc[S.z] += fabs(v[S.x]); break;
		case RD_AMAX: acc[S.z] = fmax(acc[S.z], fabs(v[S.x])); break;
		case RD_MIN: acc[S.z] = fmin(acc[S.z], v[S.x]); break;
		case RD_MAX: acc[S.z] = fmax(acc[S.z], v[S.x]); break;
		case RD_POWSUM: acc[S.z] += pow(v[S.x], sc[S.a]); break;
		}
	}
}

__device__ __forceinline__ double fold_rt(int f, double x, double y) { return f == 0 ? x + y : (f == 1 ? fmax(x, y) : fmin(x, y)); }
__device__ __forceinline__ double ident_rt(int f) {
	return f == 0 ? 0.0 : (f == 1 ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL));
}

template<int UNUSED = 0> // template only so the header may be included by several translation units
__global__ void __launch_bounds__(EW_BLOCK) ew_interp_kernel(const __grid_constant__ interp_args ia) {
	const ew_args & a = ia.a;
	const program & P = ia.p;
	grid_dependency_wait();
	tl_begin(a.tl);
	__shared__ int rfold[MAXR];
	if (threadIdx.x == 0)
		for (int i = 0; i < P.n; ++i)
			if (is_reduction(P.st[i].op))
				rfold[P.st[i].z] = fold_of(P.st[i].op);
	__syncthreads();
	double acc[MAXR];
	for (int r = 0; r < P.nr; ++r)
		acc[r] = ident_rt(rfold[r]);
	double sc[MAXSC];
	for (int k = 0; k < P.ns; ++k)
		sc[k] = a.snum[k] < 0 ? a.s[k] : __dmul_rn(a.s[k], __ddiv_rn(a.sdev[a.snum[k]], a.sdev[a.sden[k]]));
	const bool halted = a.halt && *a.halt;

	const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
	if (a.bn0 > 0 && !halted) { // structured-grid vectors (dofs of an narray mesh, reference examples/poisson/mesh.hh:157-161
		// + index_util.hh:33-71): scalar 8-byte accesses, consecutive threads on consecutive x
		for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
			const long long t = i / a.bn0, i0 = i - t * a.bn0, i2 = t / a.bn1, i1 = t - i2 * a.bn1;
			const long long at = a.boff + i0 + a.bE0 * i1 + a.bE01 * i2;
			double v[MAXV];
			for (int k = 0; k < P.nv; ++k)
				if (P.load_mask & (1u << k))
					v[k] = a.v[k][at];
			interp_exec(P, sc, v, acc);
			for (int k = 0; k < P.nv; ++k)
				if (P.store_mask & (1u << k))
					a.v[k][at] = v[k];
		}
	}
	const long long n2 = (halted || a.bn0 > 0) ? 0 : a.n >> 1;
	for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += stride) {
		double lo[MAXV], hi[MAXV];
		for (int k = 0; k < P.nv; ++k)
			if (P.load_mask & (1u << k)) {
				const double2 t = reinterpret_cast<const double2 *>(a.v[k])[i];
				lo[k] = t.x;
				hi[k] = t.y;
			}
		interp_exec(P, sc, lo, acc);
		interp_exec(P, sc, hi, acc);
		for (int k = 0; k < P.nv; ++k)
			if (P.store_mask & (1u << k))
				reinterpret_cast<double2 *>(a.v[k])[i] = make_double2(lo[k], hi[k]);
	}
	if (!halted && a.bn0 == 0 && (a.n & 1) && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
		const long long i = a.n - 1;
		double t[MAXV];
		for (int k = 0; k < P.nv; ++k)
			if (P.load_mask & (1u << k))
				t[k] = a.v[k][i];
		interp_exec(P, sc, t, acc);
		for (int k = 0; k < P.nv; ++k)
			if (P.store_mask & (1u << k))
				a.v[k][i] = t[k];
	}
	if (P.nr > 0) {
		__shared__ double scratch[32];
		__shared__ bool is_last;
		for (int r = 0; r < P.nr; ++r) {
			const int f = rfold[r];
			double t = f == 0 ? block_fold<0>(acc[r], scratch) : (f == 1 ? block_fold<1>(acc[r], scratch) : block_fold<2>(acc[r], scratch));
			if (threadIdx.x == 0)
				a.partials[r * a.partial_stride + blockIdx.x] = t;
		}
		if (threadIdx.x == 0) {
			__threadfence();
			is_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
		}
		__syncthreads();
		if (is_last) {
			__threadfence();
			for (int r = 0; r < P.nr; ++r) {
				const int f = rfold[r];
				double t = ident_rt(f);
				for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x)
					t = fold_rt(f, t, __ldcg(&a.partials[r * a.partial_stride + b]));
				t = f == 0 ? block_fold<0>(t, scratch) : (f == 1 ? block_fold<1>(t, scratch) : block_fold<2>(t, scratch));
				if (a.xr)
					t = f == 0 ? xrank_allreduce<0>(a.xr, t, a.r[r].token, scratch)
					           : (f == 1 ? xrank_allreduce<1>(a.xr, t, a.r[r].token, scratch)
					                     : xrank_allreduce<2>(a.xr, t, a.r[r].token, scratch));
				if (threadIdx.x == 0)
					publish(a.r[r], t);
			}
			if (threadIdx.x == 0)
				*a.counter = 0u;
		}
	}
	tl_end(a.tl);
}

} // namespace fsb
