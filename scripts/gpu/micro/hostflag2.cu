// The armed-launch hand-over pattern: 740 resident CTAs, CTA 0 watches a word in mapped host memory and republishes it in
// device memory for the others; how long from the host's store until ALL CTAs are through?
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <thread>
__global__ void waiter(const unsigned long long * host_word, unsigned long long * dev_word, unsigned long long want,
                       unsigned long long * out, unsigned * done, int mode) {
	__shared__ unsigned long long seen;
	if (threadIdx.x == 0) {
		const bool leader = blockIdx.x == 0 || mode == 1; // mode 1: every CTA polls the host word itself
		unsigned long long w = 0;
		const long long t0 = clock64();
		for (;;) {
			if (leader) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(host_word) : "memory");
			else asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(dev_word) : "memory");
			if (w == want) break;
			if (clock64() - t0 > 400000000LL) { w = ~0ull; break; }
		}
		if (leader && mode == 0) {
			asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dev_word), "l"(w) : "memory");
			__threadfence();
		}
		seen = w;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(done, 1u) == gridDim.x - 1) {
			*done = 0;
			*reinterpret_cast<volatile unsigned long long *>(out) = seen; // all CTAs are through: tell the host
		}
	}
}
int main() {
	unsigned long long *h, *d, *ho, *dout, *dw;
	unsigned * done;
	cudaHostAlloc(&h, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&d, h, 0);
	cudaHostAlloc(&ho, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&dout, ho, 0);
	cudaMalloc(&dw, 8); cudaMemset(dw, 0, 8); cudaMalloc(&done, 4); cudaMemset(done, 0, 4);
	for (int mode = 0; mode < 2; ++mode)
		for (int grid : {1, 148, 740}) {
			double worst = 0, sum = 0;
			for (unsigned long long it = 1; it <= 8; ++it) {
				const unsigned long long want = it + 100 * grid + 100000 * mode;
				*h = 0; *ho = 0;
				waiter<<<grid, 256>>>(d, dw, want, dout, done, mode);
				std::this_thread::sleep_for(std::chrono::milliseconds(2));
				auto t0 = std::chrono::steady_clock::now();
				__atomic_store_n(h, want, __ATOMIC_RELEASE);
				while (*reinterpret_cast<volatile unsigned long long *>(ho) == 0) {}
				double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
				cudaDeviceSynchronize();
				if (it > 2) { sum += us; if (us > worst) worst = us; }
			}
			printf("%s, %3d CTAs: host store -> all CTAs through -> host sees it: avg %.2f us, worst %.2f us\n",
			       mode == 0 ? "CTA 0 republishes" : "every CTA polls host", grid, sum / 6, worst);
			fflush(stdout);
		}
	return 0;
}
