// mat::parcsr: the row-block distributed CSR matrix on the device.
//
// Reference: flecsolve/matrices/parcsr.hh:29-177.  The reference multiplies in three task
// launches -- spmv_remote (offd block, needs ghost x => implicit halo exchange), spmv_local
// (diag block), y.add(y, tmp) (:61-91).  Here Ops::spmv is one call of the C ABI
// (fsb_parcsr_spmv, include/fsb.h): the halo exchange of x is started on a communication
// stream, the diag block runs meanwhile, the offd block then accumulates into y; no tmp
// vector and no separate add pass.
//
// Construction:  parcsr(ctx, init&&)              this process' rows in global numbering
//                                                  (the reference's parcsr(s, topo_t::init&&), :121-124)
//                parcsr::stencil(ctx, kind, nx, ny, nz)  synthetic operators generated on the device
// The Matrix-Market constructor (:112-119) is not part of the hot path and is not provided.
#ifndef FLECSOLVE_B200_MATRICES_PARCSR_HH
#define FLECSOLVE_B200_MATRICES_PARCSR_HH

#include <memory>

#include "flecsolve/matrices/sparse.hh"
#include "flecsolve/topo/csr.hh"
#include "flecsolve/vectors/topo_view.hh"

namespace flecsolve::mat {

template<class Config>
struct parcsr_data {
	using config = Config;
	using topo_t = topo::csr<typename config::scalar, typename config::size>;

	typename topo_t::topology & topo() const { return *topo_ptr; }
	fsb_parcsr_t handle() const { return topo_ptr->mat; }
	std::size_t nrows() const { return static_cast<std::size_t>(fsb_parcsr_global_rows(topo_ptr->mat)); }

	void allocate(device::context & ctx, const typename topo_t::init & ci) {
		std::vector<std::int64_t> part = ci.row_part.offsets;
		fsb_parcsr_t h = nullptr;
		device::check(fsb_parcsr_create(ctx.handle(), static_cast<std::int64_t>(ci.nrows), part.data(), ci.offsets.data(),
		                                ci.indices.data(), ci.values.data(), &h));
		topo_ptr = std::make_unique<typename topo_t::topology>(ctx, h, true);
	}
	void adopt(device::context & ctx, fsb_parcsr_t h, bool take_ownership) {
		topo_ptr = std::make_unique<typename topo_t::topology>(ctx, h, take_ownership);
	}

protected:
	typename topo_t::ptr topo_ptr;
};

template<class Data>
struct parcsr_ops {
	using data_t = Data;
	template<class D, class R>
	static void spmv(const D & x, const data_t & data, R & y) {
		device::check(fsb_parcsr_spmv(data.handle(), x.data.handle(), y.data.handle()));
	}
};

template<class scalar_t, class size_t_>
struct parcsr_config {
	using scalar = scalar_t;
	using size = size_t_;
};

template<class scalar, class size = std::size_t>
struct parcsr : sparse<parcsr_data, parcsr_ops, parcsr_config<scalar, size>> {
	using base = sparse<parcsr_data, parcsr_ops, parcsr_config<scalar, size>>;
	using data_t = typename base::data_t;
	using topo_t = typename data_t::topo_t;
	using base::data;
	using scalar_type = scalar;
	using size_type = size;

	parcsr(device::context & ctx, typename topo_t::init && init) { data.allocate(ctx, init); }
	parcsr(device::context & ctx, const typename topo_t::init & init) { data.allocate(ctx, init); }
	// wrap a matrix created through the C ABI directly
	parcsr(device::context & ctx, fsb_parcsr_t h, bool take_ownership = false) { data.adopt(ctx, h, take_ownership); }

	static parcsr stencil(device::context & ctx, int kind, std::int64_t nx, std::int64_t ny, std::int64_t nz,
	                      double diag_shift = 0.0, double scale = 1.0) {
		fsb_parcsr_t h = nullptr;
		device::check(fsb_parcsr_create_stencil(ctx.handle(), kind, nx, ny, nz, diag_shift, scale, &h));
		return parcsr(ctx, h, true);
	}

	template<typename topo_t::index_space S>
	auto vec(const data::field_definition<scalar, topo_t, S> & def) {
		return vec::make(def(data.topo()));
	}
};

}
#endif
