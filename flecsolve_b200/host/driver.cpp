// libfsb_host.so: C entry points that run the flecsolve-shaped C++ host layer
// (flecsolve_b200/include/flecsolve/...) over device vectors.  This is what a flecsolve
// application does -- build an operator, make vectors and work vectors from field definitions,
// bind a solver, call it -- compiled once so tests and bench.py can drive it through ctypes.
// All arithmetic happens behind include/fsb.h; nothing here touches vector data on the host
// except the explicit host-buffer copies of fsbh_solve.
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <algorithm>
#include <array>
#include <vector>
#include <variant>
#include <sstream>
#include <fstream>

#include "flecsolve/matrices/parcsr.hh"
#include "flecsolve/operators/shell.hh"
#include "flecsolve/solvers/bicgstab.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/cg_device.hh"
#include "flecsolve/solvers/cg_sr.hh"
#include "flecsolve/solvers/factory.hh"
#include "flecsolve/matrices/io/matrix_market.hh"
#include "flecsolve/topo/narray.hh"
#include "flecsolve/solvers/gmres.hh"
#include "flecsolve/solvers/mg/jacobi.hh"
#include "flecsolve/time-integrators/bdf.hh"
#include "flecsolve/time-integrators/operator_adapter.hh"
#include "flecsolve/time-integrators/rk45.hh"
#include "flecsolve/vectors/multi.hh"

using namespace flecsolve;

namespace {

using parcsr = mat::parcsr<double>;
using csr_topo = parcsr::topo_t;
using vec_def = csr_topo::vec_def<csr_topo::cols>;

// field definitions of this "application" (static, like the reference's examples and tests)
const vec_def bd, xd, diag_work_d;
const vec_def b2d, x2d; // second component for the multi-vector driver

thread_local std::string g_error;

template<class F>
int guarded(F && f) noexcept {
	try {
		f();
		return 0;
	}
	catch (const device::error & e) {
		g_error = e.what();
		return e.code;
	}
	catch (const std::exception & e) {
		g_error = e.what();
		return 100;
	}
}

struct session {
	device::context ctx;
	op::core<parcsr> A;
	decltype(vec::make(bd(std::declval<csr_topo::topology &>()))) b, x;
	// preconditioners are built once per session (their fields live on the topology)
	std::unique_ptr<op::core<op::diagonal_inverse<double, std::size_t>>> dinv;

	session(fsb_ctx_t c, fsb_parcsr_t a)
		: ctx(c), A(ctx, a, false), b(vec::make(bd(A.data.topo()))), x(vec::make(xd(A.data.topo()))) {}
};

// per-iteration hook: residual history + CUDA-event marks for the timing window
struct recorder {
	fsb_ctx_t ctx;
	double * history;
	int cap;
	int ev_start, ev_stop; // iteration numbers (1-based count of completed iterations) or -1
	int count = 0;
	std::int64_t launches_start = 0, launches_stop = 0;

	template<class V>
	bool operator()(const V &, double rnorm) {
		if (history && count < cap)
			history[count] = rnorm;
		++count;
		if (count == ev_start) {
			fsb_ctx_event_record(ctx, 0);
			fsb_ctx_get_stat(ctx, FSB_STAT_KERNEL_LAUNCHES, &launches_start);
		}
		if (count == ev_stop) {
			fsb_ctx_event_record(ctx, 1);
			fsb_ctx_get_stat(ctx, FSB_STAT_KERNEL_LAUNCHES, &launches_stop);
		}
		return false;
	}
};

}

extern "C" {
struct fsbh_info {
	int status, iters, restarts;
	float res_norm_initial, res_norm_final, sol_norm_initial, sol_norm_final, rhs_norm;
	int callbacks; // times the diagnostic ran
	int window_launches; // kernels launched between the two event marks
	double solve_ms; // wall clock around the solver call alone, device idle on both sides
	double h2d_ms, d2h_ms; // fsbh_solve with host buffers: upload of b and x0 / download of x (wall clock, synchronous copies)
};

struct fsbh_options {
	int solver; // 0 cg, 1 gmres, 2 bicgstab, 3 fcg, 4 cg with device-resident scalars (cg_device.hh), 5 single-reduction cg (cg_sr.hh)
	int precond; // 0 identity (op::I), 1 diagonal inverse, 2 weighted Jacobi relaxation
	float omega; // precond 2
	int nrelax; // precond 2
	int maxiter;
	float rtol, atol;
	int use_zero_guess;
	int max_krylov_dim, restart, pre_side_left; // gmres
	int ev_start, ev_stop; // record CUDA events 0 / 1 after this many iterations (-1: never)
	int lag; // solver 4: iterations the host runs ahead of the residual norm it inspects
};

// settings of time_integrator::bdf (enum values in the order of bdf_parameters.hh)
struct fsbh_bdf_options {
	int method; // 0 CN, 1 BE, 2..6 BDF2..BDF6
	int starting; // 0 CN, 1 BE
	int predictor; // 0 ab2, 1 leapfrog
	int strategy; // 0 truncation-error, 1 constant, 2 final-constant, 3 limit-relative-change
	int controller; // 0 H211b, 1 PC.4.7, 2 PC11, 3 Deadbeat
	int use_pi_controller;
	int norm; // 0 inf, 1 l2
	int error_scaling; // 0 fixed-resolution, 1 fixed-scaling
	double time_rtol, time_atol, problem_scale;
	double initial_time, final_time, initial_dt, max_dt, min_dt;
	int max_steps;
	int max_attempts; // stop after this many advance() calls (0: until final_time)
};

struct fsbh_rk_options {
	int order; // 23 or 45
	double initial_time, final_time, initial_dt, max_dt, min_dt;
	int max_steps;
	float safety_factor, atol;
	int use_fixed_dt;
};

struct fsbh_bdf_result {
	int attempts, steps, rejects;
	double final_time, value_max, value_l2; // max and l2 norm of the final solution
	int inner_iterations; // total Krylov iterations over all attempts (heat driver)
	double solve_ms; // heat driver: wall clock of the integration alone, device idle on both sides
};

}

namespace {

enum class var { first, second };

// policy applying a parallel CSR matrix to vectors of any variable tag
struct matrix_rhs : op::base<> {
	const op::core<parcsr> * A;
	explicit matrix_rhs(const op::core<parcsr> * a) : A(a) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		A->mult(x, y);
	}
};

void fill(fsbh_info * out, const solve_info & i, int callbacks, std::int64_t window_launches = 0) {
	out->window_launches = static_cast<int>(window_launches);
	out->status = static_cast<int>(i.status);
	out->iters = i.iters;
	out->restarts = i.restarts;
	out->res_norm_initial = i.res_norm_initial;
	out->res_norm_final = i.res_norm_final;
	out->sol_norm_initial = i.sol_norm_initial;
	out->sol_norm_final = i.sol_norm_final;
	out->rhs_norm = i.rhs_norm;
	out->callbacks = callbacks;
	out->solve_ms = 0;
	out->h2d_ms = out->d2h_ms = 0;
}

template<class S, class P>
solve_info run_solver(session & S_, const fsbh_options & o, P precond_handle, recorder & rec) {
	auto & A = S_.A;
	auto & b = S_.b;
	auto & x = S_.x;
	(void)sizeof(S);
	if (o.solver == 0) {
		auto slv = cg::solver(cg::settings{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0},
		                      cg::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
		return slv(b, x);
	}
	if (o.solver == 4) {
		cg_device::settings st{{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0}, o.lag};
		auto slv = cg_device::solver(st, cg_device::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
		return slv(b, x);
	}
	if (o.solver == 5) {
		cg_sr::settings st{{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0}, o.lag};
		auto slv = cg_sr::solver(st, cg_sr::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
		return slv(b, x);
	}
	if (o.solver == 3) {
		auto slv = fcg::solver(fcg::settings{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0},
		                       fcg::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
		return slv(b, x);
	}
	if (o.solver == 2) {
		bicgstab::settings st{{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0}};
		auto slv = bicgstab::solver(st, bicgstab::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
		return slv(b, x);
	}
	gmres::settings st{{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0},
	                   o.max_krylov_dim,
	                   o.pre_side_left ? gmres::precond_side::left : gmres::precond_side::right,
	                   o.restart != 0};
	auto slv = gmres::solver(st, gmres::make_work(x))(op::ref(A), precond_handle, std::ref(rec));
	return slv(b, x);
}

}

extern "C" {

const char * fsbh_last_error(void) { return g_error.c_str(); }

int fsbh_session_create(fsb_ctx_t ctx, fsb_parcsr_t A, void ** out) {
	return guarded([&] { *out = new session(ctx, A); });
}

int fsbh_session_destroy(void * s) {
	return guarded([&] { delete static_cast<session *>(s); });
}

fsb_vec_t fsbh_session_b(void * s) { return static_cast<session *>(s)->b.data.handle(); }
fsb_vec_t fsbh_session_x(void * s) { return static_cast<session *>(s)->x.data.handle(); }

// Solve A x = b.  b_host / x_host may be null: then the session's device vectors (fsbh_session_b/x)
// are used as they are.  Otherwise b is uploaded, x is uploaded (initial guess) and downloaded.
int fsbh_solve(void * sv, const fsbh_options * o, const double * b_host, double * x_host, fsbh_info * info,
               double * history, int history_cap) {
	return guarded([&] {
		session & S = *static_cast<session *>(sv);
		const std::int64_t n = fsb_vec_local_size(S.b.data.handle());
		using clk = std::chrono::steady_clock;
		auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
		const auto t_h2d = clk::now();
		if (b_host)
			device::check(fsb_vec_upload(S.b.data.handle(), b_host, n, 0));
		if (x_host && !o->use_zero_guess)
			device::check(fsb_vec_upload(S.x.data.handle(), x_host, n, 0));
		const double h2d_ms = ms_since(t_h2d);
		const auto t_solve = clk::now();
		recorder rec{S.ctx.handle(), history, history_cap, o->ev_start, o->ev_stop};
		solve_info si;
		// inside the solver call nothing but this library talks to the device: let it launch the statement groups of the
		// Krylov loop ahead of the reductions their coefficients come from (FSB_OPT_SPECULATE; FSB_SPECULATE=0: off)
		struct speculation_scope {
			fsb_ctx_t ctx;
			bool on;
			explicit speculation_scope(fsb_ctx_t c) : ctx(c) {
				// measured (profiles/r2_speculation.txt): -15 % ... -1 % per iteration on one GPU (64^3 ... 256^3), -5 % on two
				// at 1-2 M rows per rank, but +5 % on eight (256^3: 102.5 against 97.3 us) -- on by default on one rank only
				const char * e = std::getenv("FSB_SPECULATE");
				on = e ? std::atoi(e) != 0 : fsb_ctx_nranks(ctx) == 1;
				if (on)
					fsb_ctx_set_option(ctx, FSB_OPT_SPECULATE, 1);
			}
			~speculation_scope() {
				if (on)
					fsb_ctx_set_option(ctx, FSB_OPT_SPECULATE, 0);
			}
		} speculating(S.ctx.handle());
		if (o->precond == 1) {
			if (!S.dinv)
				S.dinv = std::make_unique<op::core<op::diagonal_inverse<double, std::size_t>>>(S.A);
			si = run_solver<int>(S, *o, op::ref(*S.dinv), rec);
		}
		else if (o->precond == 2) {
			auto P = mg::jacobi{{o->omega, static_cast<std::size_t>(o->nrelax)}}(op::ref(S.A));
			// relaxation does not zero its output first (mg/jacobi.hh:58-93); as a preconditioner it
			// is applied to a zeroed vector like the multigrid cycles do
			auto Pz = op::make_shell([&P](const auto & r, auto & z) {
				z.set_scalar(0.0);
				P.apply(r, z);
			});
			si = run_solver<int>(S, *o, op::ref(Pz), rec);
		}
		else {
			si = run_solver<int>(S, *o, op::I, rec);
		}
		S.ctx.sync();
		const double solve_ms = ms_since(t_solve);
		const auto t_d2h = clk::now();
		if (x_host)
			device::check(fsb_vec_download(S.x.data.handle(), x_host, n, 0));
		fill(info, si, rec.count, rec.launches_stop - rec.launches_start);
		info->solve_ms = solve_ms;
		info->h2d_ms = h2d_ms;
		info->d2h_ms = x_host ? ms_since(t_d2h) : 0.0;
	});
}

// y = (I - gamma A) x through time_integrator::operator_adapter (host buffers in/out)
int fsbh_adapter_apply(void * sv, double gamma, const double * x_host, double * y_host) {
	return guarded([&] {
		session & S = *static_cast<session *>(sv);
		const std::int64_t n = fsb_vec_local_size(S.b.data.handle());
		device::check(fsb_vec_upload(S.x.data.handle(), x_host, n, 0));
		op::core<time_integrator::operator_adapter<matrix_rhs>> F(&S.A);
		F.set_scaling(gamma);
		F.apply(S.x, S.b);
		device::check(fsb_vec_download(S.b.data.handle(), y_host, n, 0));
	});
}

// CG on a two-component vec::multi with the block-diagonal operator diag(A, A2): the shape of the
// reference's cgmulti test (solvers/test/cgmulti.cc:28-77) and of examples/equilibrium_diffusion.
// solver: 0 cg, 2 bicgstab.  Host buffers b = [b0 | b1], x = [x0 | x1].
int fsbh_solve_multi2(fsb_ctx_t ctx_h, fsb_parcsr_t A0h, fsb_parcsr_t A1h, const fsbh_options * o, const double * b_host,
                      double * x_host, fsbh_info * info, double * history, int history_cap) {
	return guarded([&] {
		device::context ctx(ctx_h);
		op::core<parcsr> A0(ctx, A0h, false), A1(ctx, A1h, false);
		auto b0 = vec::make(variable<var::first>, bd(A0.data.topo()));
		auto x0 = vec::make(variable<var::first>, xd(A0.data.topo()));
		auto b1 = vec::make(variable<var::second>, b2d(A1.data.topo()));
		auto x1 = vec::make(variable<var::second>, x2d(A1.data.topo()));
		const std::int64_t n0 = fsb_vec_local_size(b0.data.handle()), n1 = fsb_vec_local_size(b1.data.handle());
		device::check(fsb_vec_upload(b0.data.handle(), b_host, n0, 0));
		device::check(fsb_vec_upload(b1.data.handle(), b_host + n0, n1, 0));
		device::check(fsb_vec_upload(x0.data.handle(), x_host, n0, 0));
		device::check(fsb_vec_upload(x1.data.handle(), x_host + n0, n1, 0));
		vec::multi b(b0, b1), x(x0, x1);
		fsb_ctx_sync(ctx_h);
		const auto t_begin = std::chrono::steady_clock::now();

		auto blockdiag = op::make_shell(
			[&](const auto & xin, auto & yout) {
				A0.mult(xin.subset(variable<var::first>), yout.subset(variable<var::first>));
				A1.mult(xin.subset(variable<var::second>), yout.subset(variable<var::second>));
			},
			multivariable<var::first, var::second>, multivariable<var::first, var::second>);

		recorder rec{ctx_h, history, history_cap, -1, -1};
		solve_info si;
		if (o->solver == 2) {
			bicgstab::settings st{{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0}};
			auto slv = bicgstab::solver(st, bicgstab::make_work(x))(
				op::ref(blockdiag),
				op::make_identity(multivariable<var::first, var::second>, multivariable<var::first, var::second>),
				std::ref(rec));
			si = slv(b, x);
		}
		else {
			auto slv = cg::solver(cg::settings{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0}, cg::make_work(x))(
				op::ref(blockdiag),
				op::make_identity(multivariable<var::first, var::second>, multivariable<var::first, var::second>),
				std::ref(rec));
			si = slv(b, x);
		}
		fsb_ctx_sync(ctx_h);
		const double solve_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
		device::check(fsb_vec_download(x0.data.handle(), x_host, n0, 0));
		device::check(fsb_vec_download(x1.data.handle(), x_host + n0, n1, 0));
		fill(info, si, rec.count);
		info->solve_ms = solve_ms;
	});
}

// CG bound to ONE variable of a two-component multi-vector: slv(bm, xm) must pick the component
// through subset() and iterate exactly like the single-vector solve (solvers/test/cgmulti.cc:63-76).
// Both components live on A's topology; b = [b_first | b_second], x likewise; `which` selects.
int fsbh_solve_subset(fsb_ctx_t ctx_h, fsb_parcsr_t Ah, int which, const fsbh_options * o, const double * b_host,
                      double * x_host, fsbh_info * info) {
	return guarded([&] {
		device::context ctx(ctx_h);
		op::core<parcsr> A(ctx, Ah, false);
		auto b0 = vec::make(variable<var::first>, bd(A.data.topo()));
		auto x0 = vec::make(variable<var::first>, xd(A.data.topo()));
		auto b1 = vec::make(variable<var::second>, b2d(A.data.topo()));
		auto x1 = vec::make(variable<var::second>, x2d(A.data.topo()));
		const std::int64_t n = fsb_vec_local_size(b0.data.handle());
		device::check(fsb_vec_upload(b0.data.handle(), b_host, n, 0));
		device::check(fsb_vec_upload(b1.data.handle(), b_host + n, n, 0));
		device::check(fsb_vec_upload(x0.data.handle(), x_host, n, 0));
		device::check(fsb_vec_upload(x1.data.handle(), x_host + n, n, 0));
		auto bm = vec::make(b0, b1);
		auto xm = vec::make(x0, x1);
		const cg::settings st{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0};
		solve_info si;
		auto run = [&](auto tag) {
			auto op1 = op::make_shell([&A](const auto & xin, auto & yout) { A.mult(xin, yout); }, tag, tag);
			auto slv = cg::solver(st, cg::make_work(bm.subset(tag)))(op::ref(op1));
			si = slv(bm, xm);
		};
		if (which == 0)
			run(variable<var::first>);
		else
			run(variable<var::second>);
		device::check(fsb_vec_download(x0.data.handle(), x_host, n, 0));
		device::check(fsb_vec_download(x1.data.handle(), x_host + n, n, 0));
		fill(info, si, 0);
	});
}

// ---- time integration (time-integrators/bdf.hh) ------------------------------------------------
} // extern "C"

namespace {

time_integrator::bdf::settings bdf_settings(const fsbh_bdf_options & o) {
	namespace b = time_integrator::bdf;
	b::settings s{};
	s.initial_time = o.initial_time;
	s.final_time = o.final_time;
	s.max_steps = o.max_steps;
	s.max_dt = o.max_dt;
	s.min_dt = o.min_dt;
	s.initial_dt = o.initial_dt;
	s.integrator = static_cast<b::method>(o.method);
	s.starting_integrator = static_cast<b::method>(o.starting);
	s.predictor = static_cast<b::predictor>(o.predictor);
	s.timestep_strategy = static_cast<b::strategy>(o.strategy);
	s.pi_controller_type = static_cast<b::controller>(o.controller);
	s.use_pi_controller = o.use_pi_controller != 0;
	s.time_trunc_err_norm = o.norm == 0 ? vec::norm_type::inf : vec::norm_type::l2;
	s.time_error_scaling = static_cast<b::error_scaling>(o.error_scaling);
	s.time_rtol = o.time_rtol;
	s.time_atol = o.time_atol;
	s.problem_scales = {o.problem_scale};
	return s;
}

// the reference test's scalar decay operator (time-integrators/test/implicit.cc:15-50): F(x) = lambda x,
// apply() = x - gamma F(x)
struct rate_params {
	double lambda, gamma;
};
struct rate : op::base<rate_params> {
	explicit rate(double lambda) : op::base<rate_params>(rate_params{lambda, 1.}) {}
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
	void set_scaling(double g) { params.gamma = g; }
	double get_rate() const { return params.lambda; }
};
// plain right-hand side F(x) = lambda x for the explicit integrators (test/explicit.cc)
struct decay : op::base<> {
	double lambda;
	explicit decay(double l) : lambda(l) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		y.scale(lambda, x);
	}
};
struct rate_solver : op::base<> {
	explicit rate_solver(op::handle<op::core<rate>> h) : F(h) {}
	template<class D, class R>
	solve_info apply(const D & b, R & x) const {
		const double rhs = b.min().get();
		const auto & o = F.get();
		x.set_scalar(rhs / (1. - o.get_rate() * o.get_scaling()));
		solve_info info;
		info.status = solve_info::stop_reason::converged_atol;
		return info;
	}
	op::handle<op::core<rate>> F;
};

// the protocol every flecsolve application uses to drive an integrator (examples/heat_equation/implicit.cc:38-53)
template<class TI, class Vec, class OnAttempt>
void integrate(TI & ti, Vec & u, Vec & unew, int max_attempts, fsbh_bdf_result & res, OnAttempt && on_attempt) {
	double dt = ti.get_current_dt();
	bool first_step = true;
	res.attempts = 0;
	while (ti.get_current_time() < ti.get_final_time() && (max_attempts <= 0 || res.attempts < max_attempts)) {
		ti.advance(dt, first_step, u, unew);
		const bool good = ti.check_solution();
		on_attempt(res.attempts, dt, good, unew);
		++res.attempts;
		if (good) {
			ti.update();
			std::swap(u, unew);
			first_step = false;
		}
		dt = ti.get_next_dt(good);
	}
	res.steps = ti.get_current_step();
	res.rejects = ti.num_step_rejects();
	res.final_time = ti.get_current_time();
	res.value_max = u.max().get();
	res.value_l2 = u.l2norm().get();
}

}

extern "C" {

// x' = lambda x, x(0) = ic on every entry of a vector living on A's topology, exact inner solver.
// step_*[k] describe attempt k: the dt tried, whether it was accepted, max of the candidate solution.
int fsbh_bdf_rate(fsb_ctx_t ctx_h, fsb_parcsr_t Ah, const fsbh_bdf_options * o, double lambda, double ic,
                  fsbh_bdf_result * res, double * step_dt, int * step_good, double * step_val, int cap) {
	return guarded([&] {
		using namespace time_integrator;
		device::context ctx(ctx_h);
		op::core<parcsr> A(ctx, Ah, false);
		static const vec_def ud, unewd;
		auto u = vec::make(ud(A.data.topo()));
		auto unew = vec::make(unewd(A.data.topo()));
		auto F = op::make_shared<rate>(lambda);
		auto solver = op::make_shared<rate_solver>(F);
		bdf::integrator ti(bdf::parameters(bdf_settings(*o), F, bdf::make_work(u), solver));
		u.set_scalar(ic);
		*res = fsbh_bdf_result{};
		integrate(ti, u, unew, o->max_attempts, *res, [&](int k, double dt, bool good, auto & cand) {
			if (k < cap) {
				step_dt[k] = dt;
				step_good[k] = good ? 1 : 0;
				step_val[k] = cand.max().get();
			}
		});
	});
}

// x' = lambda x with the explicit integrators rk23 / rk45, driven like time-integrators/test/explicit.cc:52-68
int fsbh_rk_rate(fsb_ctx_t ctx_h, fsb_parcsr_t Ah, const fsbh_rk_options * o, double lambda, double ic,
                 fsbh_bdf_result * res, double * step_dt, int * step_good, double * step_val, int cap) {
	return guarded([&] {
		using namespace time_integrator;
		device::context ctx(ctx_h);
		op::core<parcsr> A(ctx, Ah, false);
		static const vec_def ud, unewd;
		auto u = vec::make(ud(A.data.topo()));
		auto unew = vec::make(unewd(A.data.topo()));
		op::core<decay> F(lambda);
		rk45::settings s{};
		s.initial_time = o->initial_time;
		s.final_time = o->final_time;
		s.initial_dt = o->initial_dt;
		s.max_dt = o->max_dt;
		s.min_dt = o->min_dt;
		s.max_steps = o->max_steps;
		s.safety_factor = o->safety_factor;
		s.atol = o->atol;
		s.use_fixed_dt = o->use_fixed_dt != 0;
		*res = fsbh_bdf_result{};
		auto drive = [&](auto & ti) {
			u.set_scalar(ic);
			double dt = ti.get_current_dt();
			while (ti.get_current_time() < ti.get_final_time()) {
				ti.advance(dt, u, unew);
				const bool good = ti.check_solution();
				if (res->attempts < cap) {
					step_dt[res->attempts] = dt;
					step_good[res->attempts] = good ? 1 : 0;
					step_val[res->attempts] = unew.max().get();
				}
				++res->attempts;
				if (good || ti.fixed_dt()) {
					ti.update();
					std::swap(u, unew);
				}
				dt = ti.get_next_dt(good);
			}
			res->steps = ti.get_current_step();
			res->final_time = ti.get_current_time();
			res->value_max = u.max().get();
			res->value_l2 = u.l2norm().get();
		};
		if (o->order == 23) {
			rk23::integrator ti(rk23::parameters(static_cast<const rk23::settings &>(s), op::ref(F), rk23::make_work(u)));
			drive(ti);
		}
		else {
			rk45::integrator ti(rk45::parameters(s, op::ref(F), rk45::make_work(u)));
			drive(ti);
		}
	});
}

// u_t = F u with F = the session's matrix (e.g. alpha/h^2 times the negative 7-point stencil), implicit
// BDF steps with a Krylov inner solve of (I - gamma F) u = rhs (examples/heat_equation: operator_adapter +
// bdf::integrator + Krylov solver).  Host buffers: u0 in, u out.
int fsbh_bdf_heat(void * sv, const fsbh_bdf_options * o, const fsbh_options * so, const double * u0_host, double * u_host,
                  fsbh_bdf_result * res, double * step_dt, int * step_good, int * step_iters, int cap) {
	return guarded([&] {
		using namespace time_integrator;
		session & S = *static_cast<session *>(sv);
		static const vec_def ud, unewd;
		auto u = vec::make(ud(S.A.data.topo()));
		auto unew = vec::make(unewd(S.A.data.topo()));
		const std::int64_t n = fsb_vec_local_size(u.data.handle());
		device::check(fsb_vec_upload(u.data.handle(), u0_host, n, 0));
		fsb_ctx_sync(S.ctx.handle());
		const auto t_begin = std::chrono::steady_clock::now();

		auto F = op::make_shared<operator_adapter<matrix_rhs>>(&S.A);
		int total_iters = 0, last_iters = 0;
		auto count = [&](const auto &, double) {
			++last_iters;
			return false;
		};
		*res = fsbh_bdf_result{};
		auto run = [&](auto slv_op) {
			auto solver = op::make_shared(std::move(slv_op));
			bdf::integrator ti(bdf::parameters(bdf_settings(*o), F, bdf::make_work(u), solver));
			integrate(ti, u, unew, o->max_attempts, *res, [&](int k, double dt, bool good, auto &) {
				if (k < cap) {
					step_dt[k] = dt;
					step_good[k] = good ? 1 : 0;
					step_iters[k] = last_iters;
				}
				total_iters += last_iters;
				last_iters = 0;
			});
		};
		if (so->solver == 1) {
			gmres::settings st{{so->maxiter, so->rtol, so->atol, so->use_zero_guess != 0},
			                   so->max_krylov_dim,
			                   gmres::precond_side::right,
			                   so->restart != 0};
			run(gmres::solver(st, gmres::make_work(u))(F, op::I, std::ref(count)));
		}
		else if (so->solver == 2) {
			bicgstab::settings st{{so->maxiter, so->rtol, so->atol, so->use_zero_guess != 0}};
			run(bicgstab::solver(st, bicgstab::make_work(u))(F, op::I, std::ref(count)));
		}
		else {
			run(cg::solver(cg::settings{so->maxiter, so->rtol, so->atol, so->use_zero_guess != 0}, cg::make_work(u))(
				F, op::I, std::ref(count)));
		}
		res->inner_iterations = total_iters;
		fsb_ctx_sync(S.ctx.handle());
		res->solve_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
		device::check(fsb_vec_download(u.data.handle(), u_host, n, 0));
	});
}

// The vector-operation sequence of the reference's unit test (vectors/test/flecsi_vector.cc:324-392)
// on device vectors of global size n with x = gid, y = 2 gid, z = 3 gid.  err[k] = sum_i |got - expected|
// for each checked operation (order: add, subtract, multiply, add_scalar, divide, scale, reciprocal,
// linear_sum, axpy, axpby, abs), then min, max, dot error, l1 error, l2 error.
int fsbh_vector_selftest(void * sv, double * err16) {
	return guarded([&] {
		session & S = *static_cast<session *>(sv);
		auto & topo = S.A.data.topo();
		static const vec_def xd_, yd_, zd_, tmpd_;
		auto [x, y, z, tmp] = vec::make(topo)(xd_, yd_, zd_, tmpd_);
		const std::int64_t n = fsb_vec_local_size(x.data.handle());
		const std::int64_t g0 = topo.meta().rows.beg;
		std::vector<double> hx(n), hy(n), hz(n), got(n);
		for (std::int64_t i = 0; i < n; ++i) {
			const double gid = static_cast<double>(g0 + i);
			hx[i] = gid;
			hy[i] = 2 * gid;
			hz[i] = 3 * gid;
		}
		device::check(fsb_vec_upload(x.data.handle(), hx.data(), n, 0));
		device::check(fsb_vec_upload(y.data.handle(), hy.data(), n, 0));
		device::check(fsb_vec_upload(z.data.handle(), hz.data(), n, 0));
		int k = 0;
		auto check = [&](auto & v, auto expect) {
			device::check(fsb_vec_download(v.data.handle(), got.data(), n, 0));
			double e = 0;
			for (std::int64_t i = 0; i < n; ++i)
				e += std::abs(expect(static_cast<double>(g0 + i)) - got[i]);
			err16[k++] = e;
		};
		auto X = [](double g) { return g; };
		auto Y = [](double g) { return 2 * g; };
		auto Z = [](double g) { return 3 * g; };
		tmp.add(x, z);
		check(tmp, [&](double g) { return X(g) + Z(g); });
		tmp.subtract(x, z);
		check(tmp, [&](double g) { return X(g) - Z(g); });
		tmp.multiply(x, z);
		check(tmp, [&](double g) { return X(g) * Z(g); });
		x.add_scalar(x, 1);
		check(x, [&](double g) { return X(g) + 1; });
		tmp.divide(y, x);
		check(tmp, [&](double g) { return Y(g) / (X(g) + 1); });
		x.add_scalar(x, -1);
		tmp.scale(2, x);
		check(tmp, [&](double g) { return X(g) * 2; });
		y.add_scalar(y, 1);
		tmp.reciprocal(y);
		check(tmp, [&](double g) { return 1.0 / (Y(g) + 1); });
		y.add_scalar(y, -1);
		tmp.linear_sum(8, y, 9, z);
		check(tmp, [&](double g) { return Y(g) * 8 + Z(g) * 9; });
		tmp.axpy(7, x, y);
		check(tmp, [&](double g) { return X(g) * 7 + Y(g); });
		tmp.copy(y);
		tmp.axpby(4, 11, z);
		check(tmp, [&](double g) { return Z(g) * 4 + Y(g) * 11; });
		tmp.add_scalar(y, -4);
		tmp.abs(tmp);
		check(tmp, [&](double g) { return std::abs(Y(g) - 4); });
		tmp.add_scalar(y, -7);
		err16[k++] = tmp.min().get(); // -7 on the rank owning gid 0
		err16[k++] = z.max().get(); // 3 (n_global - 1)
		const double N = static_cast<double>(x.global_size().get());
		x.set_scalar(1.5);
		y.set_scalar(3.8);
		err16[k++] = std::abs(x.dot(y).get() - N * 1.5 * 3.8);
		x.set_scalar(-3.141719);
		err16[k++] = std::abs(x.l1norm().get() - N * 3.141719);
		err16[k++] = std::abs(x.l2norm().get() - std::sqrt(N * 3.141719 * 3.141719));
	});
}

// ---- on-disk formats and configuration -----------------------------------------

// Matrix Market file -> CSR with the reference reader's rules.  First call with rowptr == null to get
// the sizes (nnz counts mirrored entries of a symmetric file), then with buffers to fill.
int fsbh_mtx_read(const char * fname, std::int64_t * nrows, std::int64_t * ncols, std::int64_t * nnz, int * symmetric,
                  std::int64_t * rowptr, std::int64_t * col, double * val) {
	return guarded([&] {
		using mm = mat::io::matrix_market<double, std::size_t>;
		std::ifstream fh(fname);
		if (!fh)
			throw std::runtime_error(std::string("cannot open ") + fname);
		const auto hdr = mm::read_header(fh);
		const auto csr = mm::read(fh, hdr).tocsr();
		*nrows = static_cast<std::int64_t>(csr.nrows);
		*ncols = static_cast<std::int64_t>(csr.ncols);
		*nnz = static_cast<std::int64_t>(csr.nnz());
		*symmetric = hdr.symmetric ? 1 : 0;
		if (rowptr) {
			std::copy(csr.offsets.begin(), csr.offsets.end(), rowptr);
			std::copy(csr.indices.begin(), csr.indices.end(), col);
			std::copy(csr.values.begin(), csr.values.end(), val);
		}
	});
}

// this rank's equal row block of the file as a device matrix (what mat::parcsr does with a
// matrix_market::definition, matrices/parcsr.hh:101-177)
int fsbh_mtx_create(fsb_ctx_t ctx_h, const char * fname, fsb_parcsr_t * out) {
	return guarded([&] {
		mat::io::matrix_market<double, std::size_t>::definition def(fname);
		auto ci = def.init_for<csr_topo::init>(static_cast<std::size_t>(fsb_ctx_rank(ctx_h)),
		                                       static_cast<std::size_t>(fsb_ctx_nranks(ctx_h)));
		device::check(fsb_parcsr_create(ctx_h, static_cast<std::int64_t>(ci.nrows), ci.row_part.offsets.data(), ci.offsets.data(),
		                                ci.indices.data(), ci.values.data(), out));
	});
}

namespace {
void put(std::ostream & os, const solver_settings & s) {
	os << "\"maxiter\": " << s.maxiter << ", \"rtol\": " << s.rtol << ", \"atol\": " << s.atol
	   << ", \"use_zero_guess\": " << (s.use_zero_guess ? "true" : "false");
}
}

// Parse `fname` with one option set and print the resulting settings as JSON:
//   kind 0: solver_options(prefix)   1: krylov_factory::options(prefix)   2: bdf::options(prefix)
//   kind 3: bdf::options("time-integrator") and krylov_factory::options("linear-solver") in ONE read_config call,
//           as examples/heat_equation/implicit.cc:19-22 does (prefix unused)
int fsbh_config_dump(const char * fname, int kind, const char * prefix, char * out, int cap) {
	return guarded([&] {
		std::ostringstream os;
		os.precision(17);
		os << "{";
		std::optional<krylov_factory::settings> ks;
		std::optional<time_integrator::bdf::settings> bs;
		if (kind == 3) {
			auto [ti, ls] = read_config(fname, time_integrator::bdf::options("time-integrator"),
			                            krylov_factory::options("linear-solver"));
			bs = ti;
			ks = ls;
		}
		if (kind == 0) {
			put(os, read_config(fname, solver_options(prefix)));
		}
		if (kind == 1 || kind == 3) {
			auto s = kind == 1 ? read_config(fname, krylov_factory::options(prefix)) : *ks;
			static const char * names[] = {"cg", "gmres", "bicgstab", "cg-device", "cg-sr"};
			os << "\"type\": \"" << names[static_cast<int>(s.target_id.value())] << "\", ";
			std::visit(
				[&](const auto & t) {
					put(os, t);
					using T = std::decay_t<decltype(t)>;
					if constexpr (std::is_same_v<T, gmres::settings>)
						os << ", \"max_krylov_dim\": " << t.max_krylov_dim << ", \"pre_side\": \"" << t.pre_side
						   << "\", \"restart\": " << (t.restart ? "true" : "false");
					if constexpr (std::is_same_v<T, cg_device::settings> || std::is_same_v<T, cg_sr::settings>)
						os << ", \"lag\": " << t.lag;
				},
				s.target_settings);
		}
		if (kind == 3)
			os << ", ";
		if (kind == 2 || kind == 3) {
			auto s = kind == 2 ? read_config(fname, time_integrator::bdf::options(prefix)) : *bs;
			os << "\"initial_time\": " << s.initial_time << ", \"final_time\": " << s.final_time << ", \"max_steps\": " << s.max_steps
			   << ", \"max_dt\": " << s.max_dt << ", \"min_dt\": " << s.min_dt << ", \"initial_dt\": " << s.initial_dt
			   << ", \"integrator\": " << static_cast<int>(s.integrator) << ", \"starting_integrator\": "
			   << static_cast<int>(s.starting_integrator) << ", \"predictor\": " << static_cast<int>(s.predictor)
			   << ", \"strategy\": " << static_cast<int>(s.timestep_strategy) << ", \"controller\": "
			   << static_cast<int>(s.pi_controller_type) << ", \"norm\": " << static_cast<int>(s.time_trunc_err_norm)
			   << ", \"error_scaling\": " << static_cast<int>(s.time_error_scaling) << ", \"time_rtol\": " << s.time_rtol
			   << ", \"time_atol\": " << s.time_atol << ", \"use_predictor\": " << (s.use_predictor ? "true" : "false")
			   << ", \"use_pi_controller\": " << (s.use_pi_controller ? "true" : "false") << ", \"problem_scales\": [";
			for (std::size_t i = 0; i < s.problem_scales.size(); ++i)
				os << (i ? ", " : "") << s.problem_scales[i];
			os << "]";
		}
		os << "}";
		const std::string text = os.str();
		if (static_cast<int>(text.size()) + 1 > cap)
			throw std::runtime_error("config dump: buffer too small");
		std::memcpy(out, text.c_str(), text.size() + 1);
	});
}

// Solve with the solver a configuration file names: krylov_factory::options(prefix) -> make(settings, x, A, P, diag)
// (examples/heat_equation/implicit.cc:19-28 uses the same two calls).  precond: 0 identity, 1 1/diag.
int fsbh_solve_config(void * sv, const char * fname, const char * prefix, int precond, const double * b_host, double * x_host,
                      fsbh_info * info, double * history, int history_cap) {
	return guarded([&] {
		session & S = *static_cast<session *>(sv);
		const std::int64_t n = fsb_vec_local_size(S.b.data.handle());
		device::check(fsb_vec_upload(S.b.data.handle(), b_host, n, 0));
		device::check(fsb_vec_upload(S.x.data.handle(), x_host, n, 0));
		auto settings = read_config(fname, krylov_factory::options(prefix));
		recorder rec{S.ctx.handle(), history, history_cap, -1, -1};
		solve_info si;
		if (precond == 1) {
			if (!S.dinv)
				S.dinv = std::make_unique<op::core<op::diagonal_inverse<double, std::size_t>>>(S.A);
			auto slv = krylov_factory::make(settings, S.x, op::ref(S.A), op::ref(*S.dinv), std::ref(rec));
			si = slv(S.b, S.x);
		}
		else {
			auto slv = krylov_factory::make(settings, S.x, op::ref(S.A), op::I, std::ref(rec));
			si = slv(S.b, S.x);
		}
		device::check(fsb_vec_download(S.x.data.handle(), x_host, n, 0));
		fill(info, si, rec.count);
	});
}

// ---- structured-grid fields ----------------------------------------------------

// examples/poisson on its own data layout: u and f are fields of a 2-D narray mesh with one boundary layer,
// A is the 5-point operator {4, -1} over the padded arrays (homogeneous Dirichlet data in the layer),
// f = 8 pi^2 sin(2 pi x) sin(2 pi y) h^2, x0 = mt19937(seed) drawn in dof order, unpreconditioned CG
// (poisson.cc:27-84, poisson.cfg).  Returns u (m*m dofs, x fastest) and the usual solve_info.
int fsbh_poisson_narray(fsb_ctx_t ctx_h, int m, const fsbh_options * o, unsigned seed, double * u_host, fsbh_info * info,
                        double * history, int history_cap) {
	return guarded([&] {
		using mesh_t = topo::narray<double, 2>;
		static const mesh_t::vec_def<mesh_t::vertices> ud, fd;
		device::context ctx(ctx_h);
		mesh_t::topology mesh(ctx, {m, m});
		mesh.set_geometry({{{0.0, 1.0}, {0.0, 1.0}}});
		auto u = vec::make(ud(mesh));
		auto f = vec::make(fd(mesh));
		const double h = mesh.delta[0], pi = 3.14159265358979323846;
		std::vector<double> rhs(static_cast<std::size_t>(m) * m);
		for (int j = 0; j < m; ++j)
			for (int i = 0; i < m; ++i)
				rhs[static_cast<std::size_t>(j) * m + i] =
					8 * pi * pi * std::sin(2 * pi * (i + 1) * h) * std::sin(2 * pi * (j + 1) * mesh.delta[1]) * h * h;
		device::check(fsb_vec_upload(f.data.handle(), rhs.data(), static_cast<std::int64_t>(rhs.size()), 0));
		u.set_random(seed);
		op::core<mat::box_stencil<double, 2>> A(mesh, 4.0, std::array<double, 2>{-1.0, -1.0});
		recorder rec{ctx_h, history, history_cap, -1, -1};
		solve_info si;
		if (o->solver == 4) {
			cg_device::settings st{{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0}, o->lag};
			si = cg_device::solver(st, cg_device::make_work(u))(op::ref(A), op::I, std::ref(rec))(f, u);
		}
		else if (o->solver == 2) {
			bicgstab::settings st{{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0}};
			si = bicgstab::solver(st, bicgstab::make_work(u))(op::ref(A), op::I, std::ref(rec))(f, u);
		}
		else {
			si = cg::solver(cg::settings{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0}, cg::make_work(u))(
				op::ref(A), op::I, std::ref(rec))(f, u);
		}
		device::check(fsb_vec_download(u.data.handle(), u_host, static_cast<std::int64_t>(rhs.size()), 0));
		fill(info, si, rec.count);
	});
}

// The closed forms of the reference's multi-vector test (vectors/test/flecsi_multivector.cc:84-241) on a
// four-component vec::multi over the session's topology: component i of x, y, z holds (i+1) gid, (i+2) gid,
// (i+3) gid.  out[0..10]: max |error| of add, subtract, multiply, add_scalar, divide, scale, reciprocal,
// linear_sum, axpy, axpby, abs; out[11]: min check (== -7 expected, returns the value); out[12..16]: 1 when the
// combined max / l1 / l2 / inf norm / dot equals the combination of the component reductions EXACTLY, else 0;
// out[17]: 1 when subset() by variable and by multivariable hands back the named components (by field id).
int fsbh_multivector_selftest(void * sv, double * out18) {
	return guarded([&] {
		session & S = *static_cast<session *>(sv);
		auto & topo = S.A.data.topo();
		static const std::array<vec_def, 4> xs, ys, zs, ts;
		static const vec_def pd, td, dd;
		const std::int64_t n = fsb_vec_local_size(S.b.data.handle());
		const std::int64_t g0 = topo.meta().rows.beg;
		auto family = [&](const std::array<vec_def, 4> & defs, int offset, bool fill) {
			auto mv = vec::multi(vec::make(defs[0](topo)), vec::make(defs[1](topo)), vec::make(defs[2](topo)),
			                     vec::make(defs[3](topo)));
			if (fill) {
				std::vector<double> h(static_cast<std::size_t>(n));
				int index = 0;
				std::apply(
					[&](auto &... comp) {
						(([&] {
							 for (std::int64_t i = 0; i < n; ++i)
								 h[static_cast<std::size_t>(i)] = static_cast<double>((index + offset + 1) * (g0 + i));
							 device::check(fsb_vec_upload(comp.data.handle(), h.data(), n, 0));
							 ++index;
						 }()),
						 ...);
					},
					mv.data.components);
			}
			return mv;
		};
		auto x = family(xs, 0, true), y = family(ys, 1, true), z = family(zs, 2, true), tmp = family(ts, 0, false);

		int k = 0;
		std::vector<double> got(static_cast<std::size_t>(n));
		auto check = [&](auto & mv, auto expect) {
			double worst = 0;
			int index = 0;
			std::apply(
				[&](auto &... comp) {
					(([&] {
						 device::check(fsb_vec_download(comp.data.handle(), got.data(), n, 0));
						 for (std::int64_t i = 0; i < n; ++i)
							 worst = std::max(worst, std::abs(expect(static_cast<double>(g0 + i), static_cast<double>(index)) -
							                                  got[static_cast<std::size_t>(i)]));
						 ++index;
					 }()),
					 ...);
				},
				mv.data.components);
			out18[k++] = worst;
		};
		tmp.add(x, z);
		check(tmp, [](double g, double i) { return (i + 1) * g + (i + 3) * g; });
		tmp.subtract(y, z);
		check(tmp, [](double g, double i) { return (i + 2) * g - (i + 3) * g; });
		tmp.multiply(x, z);
		check(tmp, [](double g, double i) { return (i + 1) * g * (i + 3) * g; });
		x.add_scalar(x, 1);
		check(x, [](double g, double i) { return (i + 1) * g + 1; });
		tmp.divide(y, x);
		check(tmp, [](double g, double i) { return ((i + 2) * g) / ((i + 1) * g + 1); });
		x.add_scalar(x, -1);
		tmp.scale(3.5, x);
		check(tmp, [](double g, double i) { return (i + 1) * g * 3.5; });
		y.add_scalar(y, 2);
		tmp.reciprocal(y);
		check(tmp, [](double g, double i) { return 1.0 / ((i + 2) * g + 2); });
		y.add_scalar(y, -2);
		tmp.linear_sum(8, y, 9, z);
		check(tmp, [](double g, double i) { return ((i + 2) * g) * 8 + ((i + 3) * g) * 9; });
		tmp.axpy(7, x, y);
		check(tmp, [](double g, double i) { return (i + 1) * g * 7 + ((i + 2) * g); });
		tmp.copy(y);
		tmp.axpby(4, 11, z);
		check(tmp, [](double g, double i) { return ((i + 3) * g) * 4 + ((i + 2) * g) * 11; });
		tmp.add_scalar(y, -4.3);
		tmp.abs(tmp);
		check(tmp, [](double g, double i) { return std::abs((i + 2) * g - 4.3); });

		tmp.add_scalar(y, -7);
		out18[k++] = tmp.min().get();
		{
			auto & [t0, t1, t2, t3] = tmp.data.components;
			out18[k++] = tmp.max().get() == std::max({t0.max().get(), t1.max().get(), t2.max().get(), t3.max().get()});
			tmp.add_scalar(z, -43);
			out18[k++] = tmp.l1norm().get() == t0.l1norm().get() + t1.l1norm().get() + t2.l1norm().get() + t3.l1norm().get();
			out18[k++] = tmp.l2norm().get() == std::sqrt(std::pow(t0.l2norm().get(), 2) + std::pow(t1.l2norm().get(), 2) +
			                                             std::pow(t2.l2norm().get(), 2) + std::pow(t3.l2norm().get(), 2));
			out18[k++] = tmp.inf_norm().get() ==
			             std::max({t0.inf_norm().get(), t1.inf_norm().get(), t2.inf_norm().get(), t3.inf_norm().get()});
		}
		{
			auto & [x0, x1, x2, x3] = x.data.components;
			auto & [y0, y1, y2, y3] = y.data.components;
			out18[k++] = x.dot(y).get() == (x0.dot(y0).get() + x1.dot(y1).get() + x2.dot(y2).get() + x3.dot(y3).get());
		}
		// named variables and subsets
		enum class vars { pressure, temperature, density };
		auto pvec = vec::make(variable<vars::pressure>, pd(topo));
		auto tvec = vec::make(variable<vars::temperature>, td(topo));
		auto dvec = vec::make(variable<vars::density>, dd(topo));
		auto mv = vec::multi(pvec, tvec, dvec);
		auto sub = mv.subset(multivariable<vars::density, vars::pressure>);
		auto & [d1, p1] = sub.data.components;
		auto sub1 = mv.subset(multivariable<vars::temperature, vars::density>);
		auto & [t1, d2] = sub1.data.components;
		auto & t2 = mv.subset(variable<vars::temperature>);
		out18[k++] = (p1.data.fid() == pvec.data.fid() && d1.data.fid() == dvec.data.fid() && t1.data.fid() == tvec.data.fid() &&
		              d2.data.fid() == dvec.data.fid() && t2.data.fid() == tvec.data.fid())
		                 ? 1.0
		                 : 0.0;
	});
}

} // extern "C"
