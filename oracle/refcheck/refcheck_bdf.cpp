// REFCHECK-BDF (test infrastructure): drives the REFERENCE's bdf::integrator
// (flecsolve/time-integrators/bdf.hh, bdf.cc from /root/reference, stub FleCSI/Boost headers) on the
// scalar decay problem x' = lambda x of the reference's own test (time-integrators/test/implicit.cc:
// operator `rate`, exact "solver" rate_solver) with serial vectors, and prints the whole step history.
//
//   refcheck_bdf <method BDF2..BDF6|BE|CN> <rtol> <atol> <initial_dt> <max_dt> <min_dt> <final_time>
//                <lambda> <ic> <n> <pi 0|1> <controller PC.4.7|H211b|PC11|Deadbeat> <predictor leapfrog|ab2>
//                [heat <nx> <ny> <nz> <scale> <solver cg|gmres> <inner_rtol> <inner_maxiter> <kdim> <max_attempts> <out.bin>]
// With the `heat` tail the problem is u_t = F u, F = scale * (7-point stencil, diag 6 / off -1) -- i.e.
// scale = -alpha/h^2 gives the heat equation of examples/heat_equation -- driven exactly like that
// example: operator_adapter<F> + bdf::integrator + the reference's own Krylov solver, u0 = box of `ic`.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>

#include "flecsolve/matrices/seq.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/gmres.hh"
#include "flecsolve/time-integrators/bdf.hh"
#include "flecsolve/time-integrators/operator_adapter.hh"
#include "flecsolve/vectors/seq.hh"

using namespace flecsolve;
using namespace flecsolve::time_integrator;

// The reference's serial Ops policy has no `apply` member (vec::apply needs one; only the FleCSI-backed
// topo_view policy provides it, vectors/operations/topo_view.hh:286-289).  Same one-liner added on top of
// the reference's own seq_ops so that every vector operation the integrator performs is the reference's.
template<class Data>
struct seq_ops_with_apply : vec::seq_ops<Data> {
	template<class F, class... Vecs>
	static constexpr decltype(auto) apply(F && f, Vecs &&... vecs) {
		return std::forward<F>(f)(std::forward<Vecs>(vecs)...);
	}
};
struct host_vec : vec::core<vec::seq_data_vec, seq_ops_with_apply, vec::seq_config<double>> {
	using base = vec::core<vec::seq_data_vec, seq_ops_with_apply, vec::seq_config<double>>;
	host_vec() : base{vec::vec_t<double>{}} {}
	explicit host_vec(std::size_t n) : base{vec::vec_t<double>(n)} {}
};

struct rate_params {
	double lambda;
	double gamma;
};

// F(x) = lambda x wrapped like operator_adapter: apply() is x - gamma F(x)
struct rate : op::base<rate_params> {
	rate(double lambda) : op::base<rate_params>(rate_params{lambda, 1.}) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		apply_rhs(x, y);
		y.axpy(-params.gamma, y, x);
	}
	template<class D, class R>
	void apply_rhs(const D & x, R & y) const {
		y.scale(params.lambda, x);
	}
	template<class V>
	bool is_valid(const V &) {
		return true;
	}
	double get_scaling() const { return params.gamma; }
	void set_scaling(double s) { params.gamma = s; }
	double get_rate() const { return params.lambda; }
};

// (I - gamma lambda) x = b solved exactly
struct rate_solver : op::base<> {
	rate_solver(op::handle<op::core<rate>> h) : F(h) {}
	template<class D, class R>
	solve_info apply(const D & b, R & x) const {
		auto rhs = b.min().get();
		const auto & o = F.get();
		x.set_scalar(rhs / (1. - o.get_rate() * o.get_scaling()));
		solve_info info;
		info.status = solve_info::stop_reason::converged_atol;
		return info;
	}
	op::handle<op::core<rate>> F;
};

// F(u) = A u for a serial CSR matrix; operator_adapter turns it into u - gamma F(u)
struct csr_rhs : op::base<> {
	const op::core<mat::csr<double>> * A;
	explicit csr_rhs(const op::core<mat::csr<double>> * a) : A(a) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		A->mult(x, y);
	}
};

static mat::csr<double> stencil7(long nx, long ny, long nz, double scale) {
	const long n = nx * ny * nz;
	std::vector<std::size_t> rp(1, 0), ci;
	std::vector<double> va;
	for (long g = 0; g < n; ++g) {
		const long i = g % nx, j = (g / nx) % ny, k = g / (nx * ny);
		auto put = [&](long c, double v) {
			ci.push_back(c);
			va.push_back(v * scale);
		};
		if (k > 0) put(g - nx * ny, -1.0);
		if (j > 0) put(g - nx, -1.0);
		if (i > 0) put(g - 1, -1.0);
		put(g, 6.0);
		if (i < nx - 1) put(g + 1, -1.0);
		if (j < ny - 1) put(g + nx, -1.0);
		if (k < nz - 1) put(g + nx * ny, -1.0);
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

static int run_heat(time_integrator::bdf::settings s, long nx, long ny, long nz, double scale, const std::string & solver,
                    float inner_rtol, int inner_maxiter, int kdim, int max_attempts, double ic, const char * outfile) {
	op::core<mat::csr<double>> A(stencil7(nx, ny, nz, scale));
	const std::size_t n = A.rows();
	auto F = op::make_shared<operator_adapter<csr_rhs>>(&A);
	host_vec u{n}, unew{n};
	// initial condition: `ic` inside the middle fifth of the box, 0 elsewhere (3-D analogue of
	// examples/heat_equation/heat.cc:10-27)
	for (long g = 0; g < static_cast<long>(n); ++g) {
		const long i = g % nx, j = (g / nx) % ny, k = g / (nx * ny);
		auto mid = [](long a, long m) { return 5 * a >= 2 * m && 5 * a < 3 * m; };
		u.data[g] = (mid(i, nx) && mid(j, ny) && mid(k, nz)) ? ic : 0.0;
	}
	std::array<host_vec, bdf::work_size> work;
	for (auto & w : work)
		w.data.resize(n);
	int last_iters = 0, total_iters = 0, attempts = 0;
	auto count = [&](const auto &, double) {
		++last_iters;
		return false;
	};
	std::printf("{\"steps\": [");
	auto drive = [&](auto & ti) {
		auto dt = ti.get_current_dt();
		bool first_step = true, comma = false;
		while (ti.get_current_time() < ti.get_final_time() && (max_attempts <= 0 || attempts < max_attempts)) {
			ti.advance(dt, first_step, u, unew);
			const bool good = ti.check_solution();
			std::printf("%s[\"%a\", %d, %d]", comma ? ", " : "", dt, good ? 1 : 0, last_iters);
			comma = true;
			total_iters += last_iters;
			last_iters = 0;
			++attempts;
			if (good) {
				ti.update();
				std::swap(u, unew);
				first_step = false;
			}
			dt = ti.get_next_dt(good);
		}
		std::printf("], \"final_time\": \"%a\", \"nsteps\": %d, \"rejects\": %d, \"value_max\": \"%a\", \"value_l2\": \"%a\", "
		            "\"inner_iterations\": %d, \"n\": %zu}\n",
		            ti.get_current_time(), ti.get_current_step(), ti.num_step_rejects(), u.max().get(), u.l2norm().get(),
		            total_iters, n);
	};
	if (solver == "gmres") {
		gmres::settings st{{inner_maxiter, inner_rtol, 0.f, false}, kdim, gmres::precond_side::right, true};
		std::array<host_vec, gmres::nwork> kwork;
		for (auto & w : kwork)
			w.data.resize(n);
		auto slv = op::make_shared(gmres::solver(st, std::move(kwork))(F, op::I, std::ref(count)));
		bdf::integrator ti(bdf::parameters(s, F, std::move(work), slv));
		drive(ti);
	}
	else {
		cg::settings st{inner_maxiter, inner_rtol, 0.f, false};
		std::array<host_vec, cg::nwork> kwork;
		for (auto & w : kwork)
			w.data.resize(n);
		auto slv = op::make_shared(cg::solver(st, std::move(kwork))(F, op::I, std::ref(count)));
		bdf::integrator ti(bdf::parameters(s, F, std::move(work), slv));
		drive(ti);
	}
	FILE * f = std::fopen(outfile, "wb");
	if (!f)
		return 3;
	for (std::size_t i = 0; i < n; ++i) {
		const double d = u.data[i];
		std::fwrite(&d, sizeof(double), 1, f);
	}
	std::fclose(f);
	return 0;
}

template<class T>
static T parse(const char * s) {
	std::istringstream in(s);
	T v;
	in >> v;
	return v;
}

int main(int argc, char ** argv) {
	if (argc < 14) {
		std::fprintf(stderr, "usage: see header comment\n");
		return 2;
	}
	bdf::settings s{};
	s.initial_time = 0.0;
	s.integrator = parse<bdf::method>(argv[1]);
	s.time_rtol = std::atof(argv[2]);
	s.time_atol = std::atof(argv[3]);
	s.initial_dt = std::atof(argv[4]);
	s.max_dt = std::atof(argv[5]);
	s.min_dt = std::atof(argv[6]);
	s.final_time = std::atof(argv[7]);
	const double lambda = std::atof(argv[8]), ic = std::atof(argv[9]);
	const std::size_t n = std::atol(argv[10]);
	s.use_pi_controller = std::atoi(argv[11]) != 0;
	s.pi_controller_type = parse<bdf::controller>(argv[12]);
	s.predictor = parse<bdf::predictor>(argv[13]);
	s.max_steps = 1000;
	s.timestep_strategy = bdf::strategy::truncation_error;
	s.starting_integrator = bdf::method::cn;
	s.calculate_time_trunc_error = false;
	s.time_trunc_err_norm = vec::norm_type::inf;
	s.use_predictor = true;
	s.use_initial_predictor = true; // option defaults of bdf_parameters.hh:121-146
	s.has_source_term = false;
	s.combine_timestep_estimators = false;
	s.dt_cut_lower_bound = 0.58754407;
	s.dt_growth_upper_bound = 1.702;
	s.number_of_time_intervals = 100;
	s.control_timestep_variation = false;
	s.time_error_scaling = bdf::error_scaling::fixed_resolution;
	s.problem_scales = {1.};
	s.target_relative_change = 0;

	if (argc > 14 && std::string(argv[14]) == "heat") {
		const long nx = std::atol(argv[15]), ny = std::atol(argv[16]), nz = std::atol(argv[17]);
		const double scale = std::atof(argv[18]);
		const std::string solver = argv[19];
		const float inner_rtol = std::strtof(argv[20], nullptr);
		const int inner_maxiter = std::atoi(argv[21]), kdim = std::atoi(argv[22]), max_attempts = std::atoi(argv[23]);
		const char * outfile = argv[24];
		return run_heat(s, nx, ny, nz, scale, solver, inner_rtol, inner_maxiter, kdim, max_attempts, ic, outfile);
	}

	auto F = op::make_shared<rate>(lambda);
	auto solver = op::make_shared<rate_solver>(F);
	host_vec x{n}, xnew{n};
	std::array<host_vec, bdf::work_size> work;
	for (auto & w : work)
		w.data.resize(n);
	bdf::integrator ti(bdf::parameters(s, F, std::move(work), solver));

	x.set_scalar(ic);
	auto dt = ti.get_current_dt();
	bool first_step = true;
	std::printf("{\"steps\": [");
	bool comma = false;
	while (ti.get_current_time() < ti.get_final_time()) {
		ti.advance(dt, first_step, x, xnew);
		const bool good = ti.check_solution();
		std::printf("%s[\"%a\", %d, \"%a\"]", comma ? ", " : "", dt, good ? 1 : 0, xnew.max().get());
		comma = true;
		if (good) {
			ti.update();
			std::swap(x, xnew);
			first_step = false;
		}
		dt = ti.get_next_dt(good);
	}
	const double exact = ic * std::exp(lambda * ti.get_final_time());
	std::printf("], \"final_time\": \"%a\", \"error\": \"%a\", \"nsteps\": %d, \"rejects\": %d, \"value\": \"%a\"}\n",
	            ti.get_current_time(), std::fabs(exact - x.max().get()), ti.get_current_step(), ti.num_step_rejects(),
	            x.max().get());
	return 0;
}
