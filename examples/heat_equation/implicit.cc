// Heat equation u_t = alpha Laplace(u) on (0, 10)^3, implicit BDF steps with a Krylov inner solve -- the
// device counterpart of the reference's examples/heat_equation (heat.cc:10-27, implicit.cc:10-58): the
// integrator and the linear solver are both chosen by implicit.cfg (bdf::options + krylov_factory::options
// read in one read_config call), the operator (I - gamma L) comes from time_integrator::operator_adapter.
//
//   usage: implicit [n = 64] [implicit.cfg]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "flecsolve/matrices/parcsr.hh"
#include "flecsolve/solvers/factory.hh"
#include "flecsolve/time-integrators/bdf.hh"
#include "flecsolve/time-integrators/operator_adapter.hh"
#include "flecsolve/util/config.hh"

using namespace flecsolve;
using namespace flecsolve::time_integrator;
using parcsr = mat::parcsr<double>;
using csr_topo = parcsr::topo_t;

static const csr_topo::vec_def<csr_topo::cols> ud, unewd;

// F(u) = L u with L the scaled 7-point matrix
struct laplacian : op::base<> {
	const op::core<parcsr> * L;
	explicit laplacian(const op::core<parcsr> * l) : L(l) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		L->mult(x, y);
	}
};

int main(int argc, char ** argv) {
	const int n = argc > 1 ? std::atoi(argv[1]) : 64;
	const char * cfg = argc > 2 ? argv[2] : "implicit.cfg";
	try {
		device::context ctx(0);
		const double length = 10.0, h = length / (n + 1), alpha = 1.0;
		// stencil {6, -1} times -alpha / h^2  =  alpha * Laplacian with homogeneous Dirichlet closure
		op::core<parcsr> L(parcsr::stencil(ctx, 7, n, n, n, 0.0, -alpha / (h * h)));
		auto u = vec::make(ud(L.data.topo()));
		auto unew = vec::make(unewd(L.data.topo()));

		std::vector<double> ic(static_cast<std::size_t>(n) * n * n, 0.0);
		auto inside = [&](int i) {
			const double x = (i + 1) * h;
			return x >= 4.0 && x <= 6.0;
		};
		for (int k = 0; k < n; ++k)
			for (int j = 0; j < n; ++j)
				for (int i = 0; i < n; ++i)
					if (inside(i) && inside(j) && inside(k))
						ic[(static_cast<std::size_t>(k) * n + j) * n + i] = 50.0;
		device::check(fsb_vec_upload(u.data.handle(), ic.data(), static_cast<std::int64_t>(ic.size()), 0));
		const double heat0 = u.l1norm().get() * h * h * h;

		auto [ti_settings, slv_settings] =
			read_config(cfg, bdf::options("time-integrator"), krylov_factory::options("linear-solver"));
		auto F = op::make_shared<operator_adapter<laplacian>>(&L);
		bdf::integrator ti(bdf::parameters(ti_settings, F, bdf::make_work(u), krylov_factory::make_shared(slv_settings, u, F)));

		double dt = ti.get_current_dt();
		bool first_step = true;
		int attempts = 0;
		while (ti.get_current_time() < ti.get_final_time()) {
			ti.advance(dt, first_step, u, unew);
			const bool good = ti.check_solution();
			++attempts;
			if (good) {
				std::printf("step %3d advanced %.4e s to time %.5f s\n", ti.get_current_step(), dt, ti.get_current_time());
				ti.update();
				std::swap(u, unew);
				first_step = false;
			}
			dt = ti.get_next_dt(good);
		}
		std::printf("%d^3: %d steps (%d attempts, %d rejected), max u = %.6f, heat %.6f -> %.6f\n", n, ti.get_current_step(),
		            attempts, ti.num_step_rejects(), u.max().get(), heat0, u.l1norm().get() * h * h * h);
		return 0;
	}
	catch (const std::exception & e) {
		std::fprintf(stderr, "implicit: %s\n", e.what());
		return 2;
	}
}
