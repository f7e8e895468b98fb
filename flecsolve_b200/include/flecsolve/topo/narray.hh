// Structured-grid topology: fields are padded N-d arrays, vectors act on the interior sub-box.
//
// Reference: the example applications specialise flecsi::topo::narray (examples/poisson/mesh.hh:25-215,
// examples/heat_equation/mesh.hh, examples/equilibrium_diffusion mesh): per axis a number of interior
// ("logical") points plus boundary/ghost layers, `dofs<space>()` = the interior sub-box in colexicographic
// order (index_util.hh:33-71), geometry (xdelta, ydelta) from the bounding box.  This header is the
// device counterpart (one padded array per rank; several ranks = slabs along the last axis whose facing pad planes are ghost
// planes the operators refresh, see fsb.h): storage per field = the padded array (fsb_vec_create_box), every vec::core
// operation runs over the dofs only, and operators read the boundary layers (mat::box_stencil below).
#ifndef FLECSOLVE_B200_TOPO_NARRAY_HH
#define FLECSOLVE_B200_TOPO_NARRAY_HH

#include <array>
#include <cstdint>
#include <unordered_map>

#include "flecsolve/device/data.hh"
#include "flecsolve/operators/core.hh"

namespace flecsolve::topo {

template<class Scalar, unsigned Dim>
struct narray {
	static_assert(Dim >= 1 && Dim <= 3, "1-, 2- or 3-dimensional meshes");
	using scalar = Scalar;
	static constexpr unsigned dimension = Dim;
	enum index_space { vertices };
	template<index_space S>
	using vec_def = data::field_definition<Scalar, narray, S>;
	using extent_t = std::array<std::int64_t, Dim>;

	struct topology {
		// interior[a] points per axis, `layers` boundary layers on each side (the reference's meshes: 1)
		topology(device::context & c, extent_t interior, std::int64_t layers = 1) : ctx(c) {
			for (unsigned a = 0; a < Dim; ++a) {
				lo[a] = layers;
				hi[a] = layers + interior[a];
				ext[a] = interior[a] + 2 * layers;
			}
		}
		topology(const topology &) = delete;
		~topology() {
			for (auto & kv : fields)
				fsb_vec_destroy(kv.second);
		}

		std::int64_t dofs() const {
			std::int64_t n = 1;
			for (unsigned a = 0; a < Dim; ++a)
				n *= hi[a] - lo[a];
			return n;
		}
		std::int64_t storage_size() const {
			std::int64_t n = 1;
			for (unsigned a = 0; a < Dim; ++a)
				n *= ext[a];
			return n;
		}
		// grid spacing of a box [lo, hi] per axis: boundary points sit on the box faces
		// (mesh.hh:171-179: delta = |hi - lo| / (extent + 1))
		void set_geometry(const std::array<std::array<double, 2>, Dim> & box) {
			for (unsigned a = 0; a < Dim; ++a)
				delta[a] = (box[a][1] > box[a][0] ? box[a][1] - box[a][0] : box[a][0] - box[a][1]) /
				           static_cast<double>(hi[a] - lo[a] + 1);
		}

		template<index_space S>
		fsb_vec_t storage(data::field_id fid) {
			auto it = fields.find(fid);
			if (it != fields.end())
				return it->second;
			fsb_vec_t v = nullptr;
			device::check(fsb_vec_create_box(ctx.handle(), static_cast<int>(Dim), ext.data(), lo.data(), hi.data(), &v));
			fields.emplace(fid, v);
			return v;
		}

		device::context & ctx;
		extent_t ext{}, lo{}, hi{};
		std::array<double, Dim> delta{};

	private:
		std::unordered_map<data::field_id, fsb_vec_t> fields;
	};
};

}

namespace flecsolve::mat {

// constant-coefficient (2 Dim + 1)-point operator on an narray topology, assembled as CSR over the
// padded array (fsb_parcsr_create_box_stencil): the role of the stencil operators of
// examples/poisson/mesh.hh:92-134 / poisson.cc:44-82
template<class Scalar, unsigned Dim>
struct box_stencil : op::base<> {
	using topo_t = topo::narray<Scalar, Dim>;

	box_stencil(typename topo_t::topology & t, Scalar center, std::array<Scalar, Dim> off) : topo(&t) {
		device::check(fsb_parcsr_create_box_stencil(t.ctx.handle(), static_cast<int>(Dim), t.ext.data(), t.lo.data(), t.hi.data(),
		                                            center, off.data(), &handle_));
	}
	box_stencil(const box_stencil &) = delete;
	box_stencil(box_stencil && o) noexcept : topo(o.topo), handle_(o.handle_) { o.handle_ = nullptr; }
	~box_stencil() {
		if (handle_)
			fsb_parcsr_destroy(handle_);
	}

	template<class D, class R>
	void apply(const D & x, R & y) const {
		device::check(fsb_parcsr_spmv(handle_, x.data.handle(), y.data.handle()));
	}
	fsb_parcsr_t handle() const { return handle_; }

	typename topo_t::topology * topo;

private:
	fsb_parcsr_t handle_ = nullptr;
};

// finite-volume diffusion operator  -beta div(b grad u) + alpha vol a u  of physics/volume_diffusion/diffusion.hh:84-207 on
// an narray topology, assembled from its coefficient fields (fsb_parcsr_create_box_fvm): a, b[axis] are host arrays over
// the padded array (cells; faces between a cell and its upper neighbour), kface[axis] = dA_axis / dx_axis
template<class Scalar, unsigned Dim>
struct box_fvm : op::base<> {
	using topo_t = topo::narray<Scalar, Dim>;

	box_fvm(typename topo_t::topology & t, Scalar beta, Scalar alpha, Scalar vol, std::array<Scalar, Dim> kface, const Scalar * a,
	        std::array<const Scalar *, Dim> b)
		: topo(&t) {
		device::check(fsb_parcsr_create_box_fvm(t.ctx.handle(), static_cast<int>(Dim), t.ext.data(), t.lo.data(), t.hi.data(), beta,
		                                        alpha, vol, kface.data(), a, b.data(), &handle_));
	}
	box_fvm(const box_fvm &) = delete;
	box_fvm(box_fvm && o) noexcept : topo(o.topo), handle_(o.handle_) { o.handle_ = nullptr; }
	~box_fvm() {
		if (handle_)
			fsb_parcsr_destroy(handle_);
	}

	template<class D, class R>
	void apply(const D & x, R & y) const {
		device::check(fsb_parcsr_spmv(handle_, x.data.handle(), y.data.handle()));
	}
	fsb_parcsr_t handle() const { return handle_; }

	typename topo_t::topology * topo;

private:
	fsb_parcsr_t handle_ = nullptr;
};

}
#endif
