// DROP-IN PROOF (test infrastructure; built in this container only, the .so travels to the GPU box):
// the REFERENCE's own headers, unmodified, where they lie under /root/reference --
//     flecsolve/vectors/core.hh, vectors/multi.hh, operators/{core,handle,shell}.hh,
//     solvers/{cg,gmres,bicgstab,krylov_parameters,solver_settings}.hh,
//     time-integrators/{bdf,operator_adapter}.hh (+ bdf.cc, bdf_parameters.cc, vectors/util.cc)
// -- instantiated on device vectors through the policy classes of include/fsb_flecsolve/b200.hh and linked
// against libfsb.so.  FleCSI / Boost are the ~300 lines of stubs in oracle/refcheck/stubs (the parallel
// task bodies they stand for are never instantiated: all vector work goes through the C ABI).
// Nothing of flecsolve_b200/include/flecsolve/ (this repo's own host layer) is on the include path.
//
// Entry points mirror flecsolve_b200/host/driver.cpp (same option / info structs) so the tests can run one
// problem through either host layer.
#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "fsb_flecsolve/b200.hh" // before the solver headers (see the note there)

#include "flecsolve/operators/shell.hh"
#include "flecsolve/solvers/bicgstab.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/gmres.hh"
#include "flecsolve/time-integrators/bdf.hh"
#include "flecsolve/time-integrators/operator_adapter.hh"
#include "flecsolve/vectors/multi.hh"

using namespace flecsolve;

namespace {

// field definitions of this "application" (static objects, like the reference's examples and tests)
const b200::field_definition bd, xd, dinvd, ud, unewd;
const b200::field_definition b2d, x2d;

thread_local std::string g_error;

template<class F>
int guarded(F && f) noexcept {
	try {
		f();
		return 0;
	}
	catch (const b200::failure & e) {
		g_error = e.what();
		return e.code;
	}
	catch (const std::exception & e) {
		g_error = e.what();
		return 100;
	}
}

// per-iteration diagnostic: residual history + CUDA-event marks for a timing window
struct recorder {
	fsb_ctx_t ctx;
	double * history;
	int cap;
	int ev_start, ev_stop;
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

// layout shared with flecsolve_b200/host/driver.cpp (flecsolve_b200/host.py: Options / Info)
struct fsbd_info {
	int status, iters, restarts;
	float res_norm_initial, res_norm_final, sol_norm_initial, sol_norm_final, rhs_norm;
	int callbacks;
	int window_launches;
	double solve_ms;
	double h2d_ms, d2h_ms;
};
struct fsbd_options {
	int solver; // 0 cg, 1 gmres, 2 bicgstab, 3 fcg
	int precond; // 0 op::I, 1 diagonal inverse
	float omega;
	int nrelax;
	int maxiter;
	float rtol, atol;
	int use_zero_guess;
	int max_krylov_dim, restart, pre_side_left;
	int ev_start, ev_stop;
	int lag;
};
struct fsbd_bdf_options {
	int method, starting, predictor, strategy, controller, use_pi_controller, norm, error_scaling;
	double time_rtol, time_atol, problem_scale;
	double initial_time, final_time, initial_dt, max_dt, min_dt;
	int max_steps;
	int max_attempts;
};
struct fsbd_bdf_result {
	int attempts, steps, rejects;
	double final_time, value_max, value_l2;
	int inner_iterations;
	double solve_ms;
};
}

namespace {

void fill(fsbd_info * out, const solve_info & i, const recorder & rec) {
	out->status = static_cast<int>(i.status);
	out->iters = i.iters;
	out->restarts = i.restarts;
	out->res_norm_initial = i.res_norm_initial;
	out->res_norm_final = i.res_norm_final;
	out->sol_norm_initial = i.sol_norm_initial;
	out->sol_norm_final = i.sol_norm_final;
	out->rhs_norm = i.rhs_norm;
	out->callbacks = rec.count;
	out->window_launches = static_cast<int>(rec.launches_stop - rec.launches_start);
	out->solve_ms = 0;
	out->h2d_ms = out->d2h_ms = 0;
}

// bind the reference's solver templates exactly as solvers/test/cg.cc:62-91 and examples/poisson do
template<class Op, class Vec, class P>
solve_info run(const fsbd_options & o, Op & A, P precond, Vec & b, Vec & x, recorder & rec) {
	const solver_settings base{o.maxiter, o.rtol, o.atol, o.use_zero_guess != 0};
	switch (o.solver) {
	case 0:
		return cg::solver(cg::settings{base}, cg::make_work(x))(op::ref(A), precond, std::ref(rec))(b, x);
	case 3:
		return fcg::solver(fcg::settings{base}, fcg::make_work(x))(op::ref(A), precond, std::ref(rec))(b, x);
	case 2:
		return bicgstab::solver(bicgstab::settings{base}, bicgstab::make_work(x))(op::ref(A), precond, std::ref(rec))(b, x);
	case 1: {
		gmres::settings st{base, o.max_krylov_dim, o.pre_side_left ? gmres::precond_side::left : gmres::precond_side::right,
		                   o.restart != 0};
		return gmres::solver(st, gmres::make_work(x))(op::ref(A), precond, std::ref(rec))(b, x);
	}
	default:
		throw std::runtime_error("dropin: unknown solver");
	}
}

enum class var { first, second };

// F(u) = A u; time_integrator::operator_adapter turns it into u - gamma F(u)
template<class M>
struct matrix_rhs : op::base<> {
	const M * A;
	explicit matrix_rhs(const M * a) : A(a) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		A->mult(x, y);
	}
};

}

extern "C" {

const char * fsbd_last_error(void) { return g_error.c_str(); }

// which reference files this library was compiled from (the tests print it)
const char * fsbd_provenance(void) {
	return "reference headers: flecsolve/vectors/{core,multi}.hh, operators/{core,handle,shell}.hh, "
	       "solvers/{cg,gmres,bicgstab,krylov_parameters,solver_settings}.hh, time-integrators/{bdf,operator_adapter}.hh "
	       "(+ bdf.cc, bdf_parameters.cc, vectors/util.cc), unmodified from " FSBD_REFERENCE_ROOT
	       "; policies: include/fsb_flecsolve/b200.hh";
}

// A x = b with host buffers: b uploaded, x uploaded (initial guess) and downloaded
int fsbd_solve(fsb_ctx_t ctx, fsb_parcsr_t Ah, const fsbd_options * o, const double * b_host, double * x_host,
               fsbd_info * info, double * history, int history_cap) {
	return guarded([&] {
		op::core<mat::b200_parcsr<double>> A(ctx, Ah);
		auto [b, x] = vec::make(A.data.topo())(bd, xd);
		const std::int64_t n = fsb_vec_local_size(b.data.h);
		b200::ok(fsb_vec_upload(b.data.h, b_host, n, 0));
		if (!o->use_zero_guess)
			b200::ok(fsb_vec_upload(x.data.h, x_host, n, 0));
		recorder rec{ctx, history, history_cap, o->ev_start, o->ev_stop};
		solve_info si;
		if (o->precond == 1) {
			op::core<op::b200_dinv<decltype(x)>> Dinv(A, A.vec(dinvd));
			si = run(*o, A, op::ref(Dinv), b, x, rec);
		}
		else
			si = run(*o, A, op::I, b, x, rec);
		b200::ok(fsb_vec_download(x.data.h, x_host, n, 0));
		fill(info, si, rec);
	});
}

// CG (solver 0) or BiCGStab (2) on a two-component vec::multi with the block-diagonal operator diag(A0, A1):
// the shape of solvers/test/cgmulti.cc:28-77 and examples/equilibrium_diffusion.  b = [b0 | b1], x = [x0 | x1].
int fsbd_solve_multi2(fsb_ctx_t ctx, fsb_parcsr_t A0h, fsb_parcsr_t A1h, const fsbd_options * o, const double * b_host,
                      double * x_host, fsbd_info * info, double * history, int history_cap) {
	return guarded([&] {
		mat::b200_parcsr<double> A0(ctx, A0h), A1(ctx, A1h);
		auto b0 = vec::make(variable<var::first>, bd(A0.data.topo()));
		auto x0 = vec::make(variable<var::first>, xd(A0.data.topo()));
		auto b1 = vec::make(variable<var::second>, b2d(A1.data.topo()));
		auto x1 = vec::make(variable<var::second>, x2d(A1.data.topo()));
		const std::int64_t n0 = fsb_vec_local_size(b0.data.h), n1 = fsb_vec_local_size(b1.data.h);
		b200::ok(fsb_vec_upload(b0.data.h, b_host, n0, 0));
		b200::ok(fsb_vec_upload(b1.data.h, b_host + n0, n1, 0));
		b200::ok(fsb_vec_upload(x0.data.h, x_host, n0, 0));
		b200::ok(fsb_vec_upload(x1.data.h, x_host + n0, n1, 0));
		vec::multi b(b0, b1), x(x0, x1);
		auto blockdiag = op::make_shell(
			[&](const auto & xin, auto & yout) {
				A0.mult(xin.subset(variable<var::first>), yout.subset(variable<var::first>));
				A1.mult(xin.subset(variable<var::second>), yout.subset(variable<var::second>));
			},
			multivariable<var::first, var::second>, multivariable<var::first, var::second>);
		auto ident = op::make_identity(multivariable<var::first, var::second>, multivariable<var::first, var::second>);
		recorder rec{ctx, history, history_cap, -1, -1};
		const solver_settings base{o->maxiter, o->rtol, o->atol, o->use_zero_guess != 0};
		solve_info si;
		fsb_ctx_sync(ctx);
		const auto t0 = std::chrono::steady_clock::now();
		if (o->solver == 2)
			si = bicgstab::solver(bicgstab::settings{base}, bicgstab::make_work(x))(op::ref(blockdiag), ident, std::ref(rec))(b, x);
		else
			si = cg::solver(cg::settings{base}, cg::make_work(x))(op::ref(blockdiag), ident, std::ref(rec))(b, x);
		fsb_ctx_sync(ctx);
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		b200::ok(fsb_vec_download(x0.data.h, x_host, n0, 0));
		b200::ok(fsb_vec_download(x1.data.h, x_host + n0, n1, 0));
		fill(info, si, rec);
		info->solve_ms = ms;
	});
}

// u_t = A u by the reference's bdf::integrator with the reference's GMRES (solver 1) or CG (0) on (I - gamma A)
// bound through operator_adapter, driven exactly like examples/heat_equation/implicit.cc:10-56.
int fsbd_bdf_heat(fsb_ctx_t ctx, fsb_parcsr_t Ah, const fsbd_bdf_options * bo, const fsbd_options * so, const double * u0,
                  double * u_out, fsbd_bdf_result * res, double * dts, int * good_flags, int * inner_iters, int cap) {
	using namespace flecsolve::time_integrator;
	return guarded([&] {
		using matrix = mat::b200_parcsr<double>;
		matrix A(ctx, Ah);
		auto [u, unew] = vec::make(A.data.topo())(ud, unewd);
		const std::int64_t n = fsb_vec_local_size(u.data.h);
		b200::ok(fsb_vec_upload(u.data.h, u0, n, 0));

		bdf::settings s{};
		static const bdf::method methods[] = {bdf::method::cn,   bdf::method::be,   bdf::method::bdf2, bdf::method::bdf3,
		                                      bdf::method::bdf4, bdf::method::bdf5, bdf::method::bdf6};
		s.integrator = methods[bo->method];
		s.starting_integrator = methods[bo->starting];
		s.predictor = bo->predictor == 0 ? bdf::predictor::ab2 : bdf::predictor::leapfrog;
		static const bdf::strategy strategies[] = {bdf::strategy::truncation_error, bdf::strategy::constant,
		                                           bdf::strategy::final_constant, bdf::strategy::limit_relative_change};
		s.timestep_strategy = strategies[bo->strategy];
		static const bdf::controller controllers[] = {bdf::controller::H211b, bdf::controller::pc4_7, bdf::controller::pc11,
		                                              bdf::controller::deadbeat};
		s.pi_controller_type = controllers[bo->controller];
		s.use_pi_controller = bo->use_pi_controller != 0;
		s.time_trunc_err_norm = bo->norm == 0 ? vec::norm_type::inf : vec::norm_type::l2;
		s.time_error_scaling = bo->error_scaling == 0 ? bdf::error_scaling::fixed_resolution : bdf::error_scaling::fixed_scaling;
		s.time_rtol = bo->time_rtol;
		s.time_atol = bo->time_atol;
		s.problem_scales = {bo->problem_scale};
		s.initial_time = bo->initial_time;
		s.final_time = bo->final_time;
		s.initial_dt = bo->initial_dt;
		s.max_dt = bo->max_dt;
		s.min_dt = bo->min_dt;
		s.max_steps = bo->max_steps;
		// option defaults of time-integrators/bdf_parameters.hh:121-146
		s.calculate_time_trunc_error = false;
		s.use_predictor = true;
		s.use_initial_predictor = true;
		s.has_source_term = false;
		s.combine_timestep_estimators = false;
		s.dt_cut_lower_bound = 0.58754407;
		s.dt_growth_upper_bound = 1.702;
		s.number_of_time_intervals = 100;
		s.control_timestep_variation = false;
		s.target_relative_change = 0;

		auto F = op::make_shared<operator_adapter<matrix_rhs<matrix>>>(&A);
		int last_iters = 0, total = 0, attempts = 0;
		auto count = [&](const auto &, double) {
			++last_iters;
			return false;
		};
		auto drive = [&](auto & ti) {
			auto dt = ti.get_current_dt();
			bool first_step = true;
			while (ti.get_current_time() < ti.get_final_time() && (bo->max_attempts <= 0 || attempts < bo->max_attempts)) {
				ti.advance(dt, first_step, u, unew);
				const bool good = ti.check_solution();
				if (attempts < cap) {
					dts[attempts] = dt;
					good_flags[attempts] = good ? 1 : 0;
					inner_iters[attempts] = last_iters;
				}
				total += last_iters;
				last_iters = 0;
				++attempts;
				if (good) {
					ti.update();
					std::swap(u, unew);
					first_step = false;
				}
				dt = ti.get_next_dt(good);
			}
			res->attempts = attempts;
			res->steps = ti.get_current_step();
			res->rejects = ti.num_step_rejects();
			res->final_time = ti.get_current_time();
		};
		const solver_settings base{so->maxiter, so->rtol, so->atol, false};
		fsb_ctx_sync(ctx);
		const auto t0 = std::chrono::steady_clock::now();
		if (so->solver == 1) {
			gmres::settings st{base, so->max_krylov_dim, gmres::precond_side::right, so->restart != 0};
			auto slv = op::make_shared(gmres::solver(st, gmres::make_work(u))(F, op::I, std::ref(count)));
			bdf::integrator ti(bdf::parameters(s, F, bdf::make_work(u), slv));
			drive(ti);
		}
		else {
			auto slv = op::make_shared(cg::solver(cg::settings{base}, cg::make_work(u))(F, op::I, std::ref(count)));
			bdf::integrator ti(bdf::parameters(s, F, bdf::make_work(u), slv));
			drive(ti);
		}
		fsb_ctx_sync(ctx);
		res->solve_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		res->inner_iterations = total;
		res->value_max = u.max().get();
		res->value_l2 = u.l2norm().get();
		b200::ok(fsb_vec_download(u.data.h, u_out, n, 0));
	});
}

// closed forms of flecsolve/vectors/test/flecsi_vector.cc:99-307 through the reference's vec::core on device
// fields: out[k] = |device - closed form| for the element-wise checks, then the exact reductions
int fsbd_vector_selftest(fsb_ctx_t ctx, fsb_parcsr_t Ah, double * out, int cap) {
	return guarded([&] {
		static const b200::field_definition xd_, yd_, zd_, tmpd_;
		mat::b200_parcsr<double> A(ctx, Ah);
		auto [x, y, z, tmp] = vec::make(A.data.topo())(xd_, yd_, zd_, tmpd_);
		const std::int64_t n = fsb_vec_local_size(x.data.h);
		std::vector<double> h(n);
		auto put = [&](auto & v, auto f) {
			for (std::int64_t i = 0; i < n; ++i)
				h[i] = f(i);
			b200::ok(fsb_vec_upload(v.data.h, h.data(), n, 0));
		};
		auto maxdiff = [&](auto & v, auto f) {
			b200::ok(fsb_vec_download(v.data.h, h.data(), n, 0));
			double m = 0;
			for (std::int64_t i = 0; i < n; ++i)
				m = std::max(m, std::fabs(h[i] - f(i)));
			return m;
		};
		int k = 0;
		auto emit = [&](double v) {
			if (k < cap)
				out[k] = v;
			++k;
		};
		put(x, [](auto i) { return double(i); });
		put(y, [](auto i) { return 2.0 * i; });
		put(z, [](auto i) { return 3.0 * i; });
		tmp.add(x, y); emit(maxdiff(tmp, [](auto i) { return 3.0 * i; }));
		tmp.subtract(z, x); emit(maxdiff(tmp, [](auto i) { return 2.0 * i; }));
		tmp.multiply(x, y); emit(maxdiff(tmp, [](auto i) { return 2.0 * i * i; }));
		tmp.add_scalar(x, 1.0); tmp.divide(y, tmp); emit(maxdiff(tmp, [](auto i) { return 2.0 * i / (i + 1.0); }));
		tmp.scale(2.5, x); emit(maxdiff(tmp, [](auto i) { return 2.5 * i; }));
		tmp.scale(2.0); emit(maxdiff(tmp, [](auto i) { return 5.0 * i; }));
		tmp.linear_sum(2.0, x, -1.0, y); emit(maxdiff(tmp, [](auto) { return 0.0; }));
		tmp.axpy(3.0, x, z); emit(maxdiff(tmp, [](auto i) { return 6.0 * i; }));
		tmp.copy(x); tmp.axpby(2.0, 3.0, y); emit(maxdiff(tmp, [](auto i) { return 7.0 * i; }));
		tmp.add_scalar(x, 1.0); tmp.reciprocal(tmp); emit(maxdiff(tmp, [](auto i) { return 1.0 / (i + 1.0); }));
		tmp.add_scalar(x, -7.0); emit(tmp.min().get() - (-7.0));
		tmp.abs(tmp); emit(tmp.max().get() - std::max(7.0, double(n - 1) - 7.0));
		// aliased forms (the _self task variants of operations/topo_view.hh:78-210)
		tmp.copy(x); tmp.add(tmp, y); emit(maxdiff(tmp, [](auto i) { return 3.0 * i; }));
		tmp.copy(x); tmp.subtract(y, tmp); emit(maxdiff(tmp, [](auto i) { return 1.0 * i; }));
		tmp.copy(x); tmp.axpy(2.0, tmp, tmp); emit(maxdiff(tmp, [](auto i) { return 3.0 * i; }));
		// reductions
		const double N = double(n);
		emit(x.l1norm().get() - N * (N - 1) / 2);
		emit(x.l2norm().get() - std::sqrt((N - 1) * N * (2 * N - 1) / 6));
		emit(x.inf_norm().get() - (N - 1));
		emit(x.dot(y).get() - 2 * (N - 1) * N * (2 * N - 1) / 6);
		emit(double(x.local_size()) - N);
		emit(double(x.global_size().get()) - N);
		emit((x == x && x != y) ? 0.0 : 1.0);
		if (k < cap)
			out[k] = -1e300; // end marker
	});
}

} // extern "C"
