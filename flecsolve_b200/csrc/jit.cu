// Run-time instantiation of program kernels (NVRTC).
//
// The ahead-of-time registry (fuser.cu) holds one compiled kernel per statement group the reference's own
// Krylov loops produce.  Any other group -- integrator stages, multi-component variants, user diagnostics,
// structured-grid layouts -- used to fall back to ew_interp_kernel, which keeps its slots in local memory
// (about half the bandwidth).  Here such a group gets the SAME template as the registered ones,
// ew_program_body<Program, DEV, BOX>, instantiated for its canonical program by NVRTC the first time it is
// seen (tens of milliseconds, once per process and program), loaded through the runtime's library API and
// cached.  The kernel sources are the very files the ahead-of-time build compiles (program.h,
// ew_kernels.cuh), embedded as strings by build.py, so both paths cannot drift apart.
//
// libnvrtc is looked up with dlopen at first use: no link-time dependency, and when it is absent the caller
// simply keeps using the generic kernel.  Compiling needs no GPU (tests/test_jit_compile.py runs here);
// loading and launching do.  Opt-in in this version: FSB_JIT=1 / FSB_OPT_JIT.
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ew_kernels.cuh"
#include "fsb_internal.h"

namespace fsb {

namespace {
#include "_jit_headers.inc" // program_h_src[], ew_kernels_cuh_src[]

struct nvrtc_api {
	void * so = nullptr;
	decltype(&nvrtcCreateProgram) create = nullptr;
	decltype(&nvrtcCompileProgram) compile = nullptr;
	decltype(&nvrtcDestroyProgram) destroy = nullptr;
	decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
	decltype(&nvrtcGetCUBIN) cubin = nullptr;
	decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
	decltype(&nvrtcGetProgramLog) log = nullptr;
	bool ok = false;
};

nvrtc_api & nvrtc() {
	static nvrtc_api api = [] {
		nvrtc_api a;
		for (const char * name : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
			a.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
			if (a.so)
				break;
		}
		if (!a.so)
			return a;
		auto sym = [&](const char * n) { return dlsym(a.so, n); };
		a.create = reinterpret_cast<decltype(a.create)>(sym("nvrtcCreateProgram"));
		a.compile = reinterpret_cast<decltype(a.compile)>(sym("nvrtcCompileProgram"));
		a.destroy = reinterpret_cast<decltype(a.destroy)>(sym("nvrtcDestroyProgram"));
		a.cubin_size = reinterpret_cast<decltype(a.cubin_size)>(sym("nvrtcGetCUBINSize"));
		a.cubin = reinterpret_cast<decltype(a.cubin)>(sym("nvrtcGetCUBIN"));
		a.log_size = reinterpret_cast<decltype(a.log_size)>(sym("nvrtcGetProgramLogSize"));
		a.log = reinterpret_cast<decltype(a.log)>(sym("nvrtcGetProgramLog"));
		a.ok = a.create && a.compile && a.destroy && a.cubin_size && a.cubin && a.log_size && a.log;
		return a;
	}();
	return api;
}

// the translation unit for one canonical program
std::string source_of(const program & p, bool dev, bool box) {
	std::string s = "#include \"ew_kernels.cuh\"\nnamespace fsb {\nstruct jit_prog {\n\tstatic constexpr program value = program{";
	s += std::to_string(p.n) + ", {";
	for (int i = 0; i < p.n; ++i) {
		const stmt & t = p.st[i];
		s += "{" + std::to_string(t.op) + "," + std::to_string(t.z) + "," + std::to_string(t.x) + "," + std::to_string(t.y) + "," +
		     std::to_string(t.a) + "," + std::to_string(t.b) + "}" + (i + 1 < p.n ? ", " : "");
	}
	s += "}, " + std::to_string(p.nv) + ", " + std::to_string(p.ns) + ", " + std::to_string(p.nr) + ", " +
	     std::to_string(p.load_mask) + "u, " + std::to_string(p.store_mask) + "u};\n};\n}\n";
	s += "extern \"C\" __global__ void __launch_bounds__(fsb::EW_BLOCK) fsb_jit_entry(const __grid_constant__ fsb::ew_args a) {\n";
	s += std::string("\tfsb::ew_program_body<fsb::jit_prog, ") + (dev ? "true" : "false") + ", " + (box ? "true" : "false") + ">(a);\n}\n";
	return s;
}

struct jit_entry {
	cudaLibrary_t lib = nullptr;
	cudaKernel_t kernel = nullptr;
	int resident = 0; // CTAs of one wave
};

std::mutex g_mutex;
std::map<std::string, jit_entry> g_cache; // key: program bytes + flags; kernel == nullptr: failed, do not retry

std::string key_of(const program & p, bool dev, bool box) {
	std::string k(reinterpret_cast<const char *>(&p.n), sizeof(int));
	k.append(reinterpret_cast<const char *>(p.st), sizeof(stmt) * p.n);
	k.push_back(dev ? 'd' : 'i');
	k.push_back(box ? 'b' : 'f');
	return k;
}
}

// Compile the kernel of one program to a cubin for sm_100a.  Works without a GPU.
bool jit_compile(const program & p, bool dev, bool box, std::vector<char> & cubin, std::string & log) {
	nvrtc_api & rt = nvrtc();
	if (!rt.ok) {
		log = "libnvrtc not found";
		return false;
	}
	const std::string src = source_of(p, dev, box);
	const char * headers[] = {ew_kernels_cuh_src, program_h_src};
	const char * names[] = {"ew_kernels.cuh", "program.h"};
	nvrtcProgram prog = nullptr;
	if (rt.create(&prog, src.c_str(), "fsb_jit.cu", 2, headers, names) != NVRTC_SUCCESS) {
		log = "nvrtcCreateProgram failed";
		return false;
	}
	const char * opts[] = {"--gpu-architecture=sm_100a", "--std=c++20", "-default-device", "-lineinfo"};
	const nvrtcResult rc = rt.compile(prog, 4, opts);
	size_t n = 0;
	if (rt.log_size(prog, &n) == NVRTC_SUCCESS && n > 1) {
		log.resize(n);
		rt.log(prog, log.data());
	}
	bool ok = rc == NVRTC_SUCCESS;
	if (ok) {
		ok = rt.cubin_size(prog, &n) == NVRTC_SUCCESS && n > 0;
		if (ok) {
			cubin.resize(n);
			ok = rt.cubin(prog, cubin.data()) == NVRTC_SUCCESS;
			if (const char * dump = std::getenv("FSB_JIT_DUMP")) // inspection: cuobjdump -sass / -res-usage <file>
				if (FILE * f = ok ? fopen(dump, "wb") : nullptr) {
					fwrite(cubin.data(), 1, cubin.size(), f);
					fclose(f);
				}
		}
	}
	rt.destroy(&prog);
	return ok;
}

// kernel for this program, compiled and loaded on first use; nullptr when the run-time path is unavailable
// (the caller then uses the generic kernel)
const void * jit_kernel(const program & p, bool dev, bool box, int * resident) {
	std::lock_guard<std::mutex> lock(g_mutex);
	const std::string key = key_of(p, dev, box);
	auto it = g_cache.find(key);
	if (it == g_cache.end()) {
		jit_entry e;
		std::vector<char> cubin;
		std::string log;
		if (jit_compile(p, dev, box, cubin, log)) {
			if (cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
			    cudaLibraryGetKernel(&e.kernel, e.lib, "fsb_jit_entry") == cudaSuccess) {
				int nb = 0;
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, reinterpret_cast<const void *>(e.kernel), EW_BLOCK, 0) !=
				        cudaSuccess ||
				    nb < 1)
					nb = 4;
				e.resident = nb * SM_COUNT;
			}
			else {
				cudaGetLastError();
				e.kernel = nullptr;
			}
		}
		else if (std::getenv("FSB_JIT_DEBUG"))
			fprintf(stderr, "[fsb] jit compile failed:\n%s\n", log.c_str());
		it = g_cache.emplace(key, e).first;
	}
	if (resident)
		*resident = it->second.resident;
	return reinterpret_cast<const void *>(it->second.kernel);
}

void jit_launch(const void * kernel, int resident, const ew_args & a, int want, cudaStream_t s) {
	const int grid = want < resident ? want : resident;
	void * args[] = {const_cast<ew_args *>(&a)};
	FSB_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(EW_BLOCK), args, 0, s));
}

}
