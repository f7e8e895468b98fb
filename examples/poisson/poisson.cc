// 2-D Poisson on a structured mesh, unpreconditioned CG -- the device counterpart of the reference's
// examples/poisson (poisson.cc:27-84, 150-180): fields live on an narray mesh with one boundary layer,
// the operator is the 5-point stencil over the padded arrays, settings come from poisson.cfg.
//
//   usage: poisson [n = 256] [poisson.cfg]
//
// Everything below is flecsolve's API surface (vec::make, op::core, cg::solver, read_config,
// diagnostics, solve_info); the arithmetic runs on the GPU behind include/fsb.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "flecsolve/solvers/cg.hh"
#include "flecsolve/topo/narray.hh"
#include "flecsolve/util/config.hh"
#include "flecsolve/vectors/topo_view.hh"

using namespace flecsolve;
using mesh_t = topo::narray<double, 2>;

static const mesh_t::vec_def<mesh_t::vertices> ud, fd, exactd;

int main(int argc, char ** argv) {
	const int n = argc > 1 ? std::atoi(argv[1]) : 256;
	const char * cfg = argc > 2 ? argv[2] : "poisson.cfg";
	try {
		device::context ctx(0);
		mesh_t::topology mesh(ctx, {n, n});
		mesh.set_geometry({{{0.0, 1.0}, {0.0, 1.0}}});
		auto u = vec::make(ud(mesh));
		auto f = vec::make(fd(mesh));
		auto exact = vec::make(exactd(mesh));

		const double pi = 3.14159265358979323846, hx = mesh.delta[0], hy = mesh.delta[1];
		std::vector<double> rhs(static_cast<std::size_t>(n) * n), sol(rhs.size());
		for (int j = 0; j < n; ++j)
			for (int i = 0; i < n; ++i) {
				const double s = std::sin(2 * pi * (i + 1) * hx) * std::sin(2 * pi * (j + 1) * hy);
				sol[static_cast<std::size_t>(j) * n + i] = s;
				rhs[static_cast<std::size_t>(j) * n + i] = 8 * pi * pi * s * hx * hy;
			}
		device::check(fsb_vec_upload(f.data.handle(), rhs.data(), static_cast<std::int64_t>(rhs.size()), 0));
		device::check(fsb_vec_upload(exact.data.handle(), sol.data(), static_cast<std::int64_t>(sol.size()), 0));
		u.set_random(7);

		// -(u_xx + u_yy) scaled by hx hy: 2 (hy/hx + hx/hy) on the diagonal, -hy/hx and -hx/hy beside it
		op::core<mat::box_stencil<double, 2>> A(mesh, 2 * (hy / hx + hx / hy), std::array<double, 2>{-hy / hx, -hx / hy});

		auto settings = read_config(cfg, cg::options("solver"));
		int shown = 0;
		auto diagnostic = [&](const auto &, double rnorm) {
			if (shown++ % 100 == 0)
				std::printf("  iteration %4d  |r| = %.6e\n", shown - 1, rnorm);
			return false;
		};
		auto solver = cg::solver(settings, cg::make_work(u))(op::ref(A), op::I, diagnostic);
		const solve_info info = solver(f, u);

		exact.subtract(exact, u);
		std::printf("%dx%d: %s after %d iterations, |r| = %.3e, max error vs sin(2 pi x) sin(2 pi y) = %.3e\n", n, n,
		            info.success() ? "converged" : "not converged", info.iters, static_cast<double>(info.res_norm_final),
		            exact.inf_norm().get());
		return info.success() ? 0 : 1;
	}
	catch (const std::exception & e) {
		std::fprintf(stderr, "poisson: %s\n", e.what());
		return 2;
	}
}
