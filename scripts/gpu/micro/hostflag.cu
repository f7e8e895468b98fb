// How fast does a spinning kernel see a word the host stores into mapped pinned memory?  (flavours of the device-side load)
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <thread>
template<int MODE>
__global__ void waiter(const unsigned long long * host_word, unsigned long long want, unsigned long long * out) {
	unsigned long long w = 0;
	long long n = 0;
	const long long t0 = clock64();
	for (;;) {
		if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(host_word) : "memory");
		if (MODE == 1) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(host_word) : "memory");
		if (MODE == 2) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(host_word) : "memory");
		if (MODE == 3) w = atomicAdd_system(const_cast<unsigned long long *>(host_word), 0ull);
		if (MODE == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(w) : "l"(host_word) : "memory");
		++n;
		if (w == want) break;
		if (clock64() - t0 > 400000000LL) { n = -n; break; } // ~0.2 s: never saw it
	}
	out[0] = static_cast<unsigned long long>(n);
	*reinterpret_cast<volatile unsigned long long *>(out + 1) = want; // tell the host (mapped)
}
template<int MODE>
void run(const char * name) {
	unsigned long long *h, *d, *ho, *dout;
	cudaHostAlloc(&h, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&d, h, 0);
	cudaHostAlloc(&ho, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&dout, ho, 0);
	double worst = 0, sum = 0;
	int missed = 0;
	for (unsigned long long it = 1; it <= 8; ++it) {
		*h = 0; ho[0] = ho[1] = 0;
		waiter<MODE><<<1, 1>>>(d, it, dout);
		std::this_thread::sleep_for(std::chrono::milliseconds(2)); // the kernel is resident and polling by now
		auto t0 = std::chrono::steady_clock::now();
		__atomic_store_n(h, it, __ATOMIC_RELEASE);
		while (*reinterpret_cast<volatile unsigned long long *>(ho + 1) != it) {}
		if (static_cast<long long>(*reinterpret_cast<volatile unsigned long long *>(ho)) < 0) ++missed;
		double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
		cudaDeviceSynchronize();
		if (it > 2) { sum += us; if (us > worst) worst = us; }
	}
	printf("%-28s round trip host store -> kernel sees it -> host sees the answer: avg %.2f us, worst %.2f us, never seen %d of 8\n", name, sum / 6, worst, missed);
	fflush(stdout);
}
int main() {
	run<0>("ld.volatile.global");
	run<1>("ld.relaxed.sys.global");
	run<2>("ld.acquire.sys.global");
	run<3>("atomicAdd_system(p, 0)");
	run<4>("ld.global.cv");
	return 0;
}
