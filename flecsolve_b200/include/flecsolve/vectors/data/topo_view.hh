// Data policy of vec::topo_view: a field reference plus the device handle behind it.
// Reference: flecsolve/vectors/data/topo_view.hh:24-81 (ref(), fid(), topo(), equality by fid).
#ifndef FLECSOLVE_B200_VECTORS_DATA_TOPO_VIEW_HH
#define FLECSOLVE_B200_VECTORS_DATA_TOPO_VIEW_HH

#include "flecsolve/device/data.hh"

namespace flecsolve::vec::data {

template<class Config>
struct topo_view {
	using config = Config;
	using scalar = typename Config::scalar;
	using topo_t = typename Config::topo_t;
	static constexpr typename topo_t::index_space space = Config::space;
	using field_definition = ::flecsolve::data::field_definition<scalar, topo_t, space>;
	using field_reference = ::flecsolve::data::field_reference<scalar, topo_t, space>;

	explicit topo_view(field_reference r) : reference(r), handle_(r.topology().template storage<space>(r.fid())) {}

	field_reference reference;

	auto ref() const { return reference; }
	auto fid() const { return reference.fid(); }
	auto & topo() const { return reference.topology(); }
	fsb_vec_t handle() const { return handle_; }
	fsb_ctx_t ctx() const { return reference.topology().ctx.handle(); }

private:
	fsb_vec_t handle_;
};

template<class C>
bool operator==(const topo_view<C> & a, const topo_view<C> & b) {
	return a.fid() == b.fid();
}
template<class C>
bool operator!=(const topo_view<C> & a, const topo_view<C> & b) {
	return a.fid() != b.fid();
}

}
#endif
