// REFCHECK (test infrastructure): runs the REFERENCE's own serial code -- mat::csr +
// compressed_ops::spmv (flecsolve/matrices/seq.hh), vec::seq_vec + seq_ops (flecsolve/vectors/seq.hh)
// and the Krylov solver templates (flecsolve/solvers/{cg,gmres,bicgstab}.hh) -- compiled from the
// sources where they lie under /root/reference, against stub FleCSI/Boost headers (oracle/refcheck/stubs).
// Only the serial, FleCSI-free paths are executed; this is how the reference's own tests drive the
// solvers (solvers/test/cg.cc:62-91).  Output: one JSON object on stdout; vectors in a raw file.
//
//   refcheck <kind 5|7|27> <nx> <ny> <nz> <solver cg|gmres|bicgstab|spmv> <precond 0|1> <rtol> <maxiter>
//            <use_zero_guess 0|1> <max_krylov_dim> <restart 0|1> <bseed> <xseed> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "flecsolve/matrices/seq.hh"
#include "flecsolve/solvers/bicgstab.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/gmres.hh"
#include "flecsolve/vectors/seq.hh"

using namespace flecsolve;

// synthetic operators of SURVEY.md section 8d (same definition as oracle/oracle.cpp: visit_stencil)
static mat::csr<double> stencil(int kind, long nx, long ny, long nz) {
	const long n = nx * ny * nz;
	const double dv = kind == 27 ? 26.0 : (kind == 7 ? 6.0 : 4.0), ov = -1.0;
	std::vector<std::size_t> rp(1, 0), ci;
	std::vector<double> va;
	for (long g = 0; g < n; ++g) {
		const long i = g % nx, j = (g / nx) % ny, k = g / (nx * ny);
		auto put = [&](long c, double v) {
			ci.push_back(c);
			va.push_back(v);
		};
		if (kind == 27) {
			for (int dk = -1; dk <= 1; ++dk)
				for (int dj = -1; dj <= 1; ++dj)
					for (int di = -1; di <= 1; ++di) {
						const long ii = i + di, jj = j + dj, kk = k + dk;
						if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz)
							continue;
						put(ii + nx * (jj + ny * kk), (di == 0 && dj == 0 && dk == 0) ? dv : ov);
					}
		}
		else {
			if (kind == 7 && k > 0) put(g - nx * ny, ov);
			if (j > 0) put(g - nx, ov);
			if (i > 0) put(g - 1, ov);
			put(g, dv);
			if (i < nx - 1) put(g + 1, ov);
			if (j < ny - 1) put(g + nx, ov);
			if (kind == 7 && k < nz - 1) put(g + nx * ny, ov);
		}
		rp.push_back(ci.size());
	}
	mat::csr<double> A(n, n);
	A.resize(ci.size());
	auto [rowptr, colind, values] = A.rep();
	for (std::size_t r = 0; r < rp.size(); ++r)
		rowptr[r] = rp[r];
	for (std::size_t e = 0; e < ci.size(); ++e) {
		colind[e] = ci[e];
		values[e] = va[e];
	}
	return A;
}

// the reference test fixture's Dinv(): diagonal CSR holding 1/a_ii (util/test/mesh.hh:123-140)
static mat::csr<double> dinv_of(const mat::csr<double> & A) {
	mat::csr<double> out{A.rows(), A.cols()};
	out.resize(A.rows());
	for (std::size_t i = 0; i < A.rows(); i++) {
		for (std::size_t off = A.data.offsets()[i]; off < A.data.offsets()[i + 1]; off++) {
			if (A.data.indices()[off] == i) {
				out.data.values()[i] = 1.0 / A.data.values()[off];
				out.data.indices()[i] = i;
			}
		}
		out.data.offsets()[i + 1] = i + 1;
	}
	return out;
}

struct recorder {
	std::vector<double> hist;
	template<class V>
	bool operator()(const V &, double r) {
		hist.push_back(r);
		return false;
	}
};

int main(int argc, char ** argv) {
	if (argc < 15) {
		std::fprintf(stderr, "usage: see header comment\n");
		return 2;
	}
	const int kind = std::atoi(argv[1]);
	const long nx = std::atol(argv[2]), ny = std::atol(argv[3]), nz = std::atol(argv[4]);
	const std::string solver = argv[5];
	const bool precond = std::atoi(argv[6]) != 0;
	const float rtol = std::strtof(argv[7], nullptr);
	const int maxiter = std::atoi(argv[8]);
	const bool zero_guess = std::atoi(argv[9]) != 0;
	const int kdim = std::atoi(argv[10]);
	const bool restart = std::atoi(argv[11]) != 0;
	const unsigned bseed = std::atoi(argv[12]), xseed = std::atoi(argv[13]);
	const char * outfile = argv[14];

	op::core<mat::csr<double>> A(stencil(kind, nx, ny, nz));
	op::core<mat::csr<double>> Dinv(dinv_of(A));
	const std::size_t n = A.rows();
	vec::seq_vec<double> b{n}, x{n}, t{n};
	// b = A * u with u = set_random(bseed) keeps the system consistent; x0 = set_random(xseed)
	t.set_random(bseed);
	A.apply(t, b);
	x.set_random(xseed);

	solve_info info;
	recorder rec;
	if (solver == "spmv") {
		// y = A x only
		A.apply(x, t);
	}
	else if (solver == "cg") {
		cg::settings st{maxiter, rtol, 0.f, zero_guess};
		if (precond)
			info = cg::solver(st, vec::seq_work<double, cg::nwork>{b})(op::ref(A), op::ref(Dinv), std::ref(rec))(b, x);
		else
			info = cg::solver(st, vec::seq_work<double, cg::nwork>{b})(op::ref(A), op::I, std::ref(rec))(b, x);
	}
	else if (solver == "bicgstab") {
		bicgstab::settings st{{maxiter, rtol, 0.f, zero_guess}};
		if (precond)
			info = bicgstab::solver(st, vec::seq_work<double, bicgstab::nwork>{b})(op::ref(A), op::ref(Dinv),
			                                                                       std::ref(rec))(b, x);
		else
			info = bicgstab::solver(st, vec::seq_work<double, bicgstab::nwork>{b})(op::ref(A), op::I, std::ref(rec))(b, x);
	}
	else if (solver == "gmres") {
		gmres::settings st{{maxiter, rtol, 0.f, zero_guess}, kdim, gmres::precond_side::right, restart};
		// gmres indexes its workspace (work[i], work.data()): an array of serial vectors
		std::array<vec::seq_vec<double>, gmres::nwork> work;
		for (auto & w : work)
			w.data.resize(n);
		if (precond)
			info = gmres::solver(st, std::move(work))(op::ref(A), op::ref(Dinv), std::ref(rec))(b, x);
		else
			info = gmres::solver(st, std::move(work))(op::ref(A), op::I, std::ref(rec))(b, x);
	}
	else {
		std::fprintf(stderr, "unknown solver %s\n", solver.c_str());
		return 2;
	}

	// raw output: b, then x (solution, or y = A x0 for "spmv"), then the residual history
	FILE * f = std::fopen(outfile, "wb");
	if (!f)
		return 3;
	auto dump = [&](const vec::seq_vec<double> & v) {
		for (std::size_t i = 0; i < n; ++i) {
			const double d = v.data[i];
			std::fwrite(&d, sizeof(double), 1, f);
		}
	};
	dump(b);
	dump(solver == "spmv" ? t : x);
	std::fwrite(rec.hist.data(), sizeof(double), rec.hist.size(), f);
	std::fclose(f);
	std::printf("{\"n\": %zu, \"status\": %d, \"iters\": %d, \"restarts\": %d, \"res_norm_initial\": %.9g, "
	            "\"res_norm_final\": %.9g, \"sol_norm_initial\": %.9g, \"sol_norm_final\": %.9g, \"rhs_norm\": %.9g, "
	            "\"history_len\": %zu}\n",
	            n, static_cast<int>(info.status), info.iters, info.restarts, info.res_norm_initial, info.res_norm_final,
	            info.sol_norm_initial, info.sol_norm_final, info.rhs_norm, rec.hist.size());
	return 0;
}
