// Two diffusing fields on one 2-D mesh solved as ONE system on a vec::multi -- the device counterpart of the reference's
// examples/equilibrium_diffusion (include/equilibrium_diffusion.hh:268-361): component v1 sees Dirichlet boundaries,
// component v2 zero-flux (Neumann) boundaries, both blocks are the finite-volume diffusion operator
// -beta div(b grad u) + alpha vol a u of physics/volume_diffusion/diffusion.hh with beta = 1, alpha = 0 and constant face
// coefficients, the right-hand side vanishes and both fields start at 2: v1 drains into its boundary, v2 -- a constant
// is in the null space of its block -- stays where it is.  Settings come from diffusion.cfg, solver: CG.
//
//   usage: diffusion [n = 64] [diffusion.cfg]
//
// The boundary conditions live in the operator's coefficient fields: a Dirichlet side keeps its boundary faces (the
// boundary layer of every Krylov vector is zero = the reference's 1e-9 boundary value), a zero-flux side has b = 0 on its
// boundary faces.  Everything else is flecsolve's API surface; the arithmetic runs on the GPU behind include/fsb.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "flecsolve/operators/shell.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/topo/narray.hh"
#include "flecsolve/util/config.hh"
#include "flecsolve/vectors/multi.hh"
#include "flecsolve/vectors/topo_view.hh"

using namespace flecsolve;
using mesh_t = topo::narray<double, 2>;
enum class diffusion_var { v1, v2 };

static const mesh_t::vec_def<mesh_t::vertices> x1d, x2d, rhs1d, rhs2d;

int main(int argc, char ** argv) {
	const int n = argc > 1 ? std::atoi(argv[1]) : 64;
	const char * cfg = argc > 2 ? argv[2] : "diffusion.cfg";
	try {
		device::context ctx(0);
		mesh_t::topology mesh(ctx, {n, n});
		const double dx = 1.0 / n, dy = 1.0 / n, vol = dx * dy;
		const std::array<double, 2> kface{dy * (1.0 / dx), dx * (1.0 / dy)}; // dA_axis / dx_axis (diffusion.hh:126-132)

		auto v1 = vec::make(variable<diffusion_var::v1>, x1d(mesh));
		auto v2 = vec::make(variable<diffusion_var::v2>, x2d(mesh));
		auto r1 = vec::make(variable<diffusion_var::v1>, rhs1d(mesh));
		auto r2 = vec::make(variable<diffusion_var::v2>, rhs2d(mesh));
		vec::multi X(v1, v2), RHS(r1, r2);
		X.set_scalar(2.0);
		RHS.set_scalar(0.0);

		// coefficient fields over the padded array: a (cells, unused with alpha = 0), b_x / b_y (the face above each cell)
		const std::int64_t E0 = mesh.ext[0], E1 = mesh.ext[1];
		std::vector<double> a(static_cast<std::size_t>(E0 * E1), 1.0), bx(a.size(), 1.0), by(a.size(), 1.0);
		std::vector<double> bx_closed = bx, by_closed = by;
		for (std::int64_t j = 0; j < E1; ++j) { // zero-flux sides of v2: no flux through the faces that touch a boundary layer
			bx_closed[static_cast<std::size_t>(j * E0 + mesh.lo[0] - 1)] = 0.0;
			bx_closed[static_cast<std::size_t>(j * E0 + mesh.hi[0] - 1)] = 0.0;
		}
		for (std::int64_t i = 0; i < E0; ++i) {
			by_closed[static_cast<std::size_t>((mesh.lo[1] - 1) * E0 + i)] = 0.0;
			by_closed[static_cast<std::size_t>((mesh.hi[1] - 1) * E0 + i)] = 0.0;
		}
		const double beta = 1.0, alpha = 0.0;
		// the two diagonal blocks (operator policies; the shell below is the operator the solver sees)
		mat::box_fvm<double, 2> A1(mesh, beta, alpha, vol, kface, a.data(), std::array<const double *, 2>{bx.data(), by.data()});
		mat::box_fvm<double, 2> A2(mesh, beta, alpha, vol, kface, a.data(), std::array<const double *, 2>{bx_closed.data(), by_closed.data()});
		auto A = op::make_shell(
			[&](const auto & xin, auto & yout) {
				A1.apply(xin.subset(variable<diffusion_var::v1>), yout.subset(variable<diffusion_var::v1>));
				A2.apply(xin.subset(variable<diffusion_var::v2>), yout.subset(variable<diffusion_var::v2>));
			},
			multivariable<diffusion_var::v1, diffusion_var::v2>, multivariable<diffusion_var::v1, diffusion_var::v2>);

		auto settings = read_config(cfg, cg::options("solver"));
		auto solver = cg::solver(settings, cg::make_work(RHS))(op::ref(A));
		const solve_info info = solver(RHS, X);

		std::printf("%dx%d, 2 components: %s after %d iterations, |r| = %.3e; v1 in [%.3e, %.3e], v2 in [%.6f, %.6f]\n", n, n,
		            info.success() ? "converged" : "not converged", info.iters, static_cast<double>(info.res_norm_final),
		            v1.min().get(), v1.max().get(), v2.min().get(), v2.max().get());
		return info.success() ? 0 : 1;
	}
	catch (const std::exception & e) {
		std::fprintf(stderr, "diffusion: %s\n", e.what());
		return 2;
	}
}
