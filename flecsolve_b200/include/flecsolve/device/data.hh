// Mini data model: field definitions, references and device storage keyed by field id.
//
// The reference's vectors are views of FleCSI fields: a static `definition` object names a field,
// `definition(topology)` yields a field_reference {fid(), topology()}, two vectors alias iff their
// fids match, and storage is owned by the runtime, never allocated during a solve
// (vectors/data/topo_view.hh:24-81, solvers/solver_settings.hh:123-149, README "fields cannot be
// constructed dynamically during a solve").  This header reproduces exactly those three notions
// on top of the C ABI so user code keeps its shape:
//
//     static const topo_t::vec_def<topo_t::cols> xd, yd;        // definitions
//     auto [x, y] = vec::make(A.data.topo())(xd, yd);           // references -> vectors
//
// Storage for a field on a topology is created the first time it is referenced and lives as
// long as the topology (the work vectors of a solver are therefore allocated once, at
// make_work time, not inside apply()).
#ifndef FLECSOLVE_B200_DEVICE_DATA_HH
#define FLECSOLVE_B200_DEVICE_DATA_HH

#include <atomic>
#include <cstddef>
#include <unordered_map>

#include "flecsolve/device/runtime.hh"

namespace flecsolve::data {

using field_id = std::size_t;

inline field_id next_field_id() {
	static std::atomic<field_id> counter{1};
	return counter++;
}

template<class T, class Topo, typename Topo::index_space Space>
struct field_reference {
	field_id id;
	typename Topo::topology * slot;

	field_id fid() const { return id; }
	typename Topo::topology & topology() const { return *slot; }
};

template<class T, class Topo, typename Topo::index_space Space>
struct field_definition {
	using value_type = T;
	field_definition() : fid(next_field_id()) {}
	field_definition(const field_definition &) = delete; // a definition *is* its identity

	field_reference<T, Topo, Space> operator()(typename Topo::topology & t) const { return {fid, &t}; }

	const field_id fid;
};

// device vectors of one topology instance, by (field id); all have the shape of the index space
class field_store {
public:
	explicit field_store(fsb_ctx_t ctx) : ctx_(ctx) {}
	field_store(const field_store &) = delete;
	~field_store() {
		for (auto & kv : vecs_)
			fsb_vec_destroy(kv.second);
	}
	fsb_vec_t get(field_id fid, std::int64_t n_owned, std::int64_t n_ghost) {
		auto it = vecs_.find(fid);
		if (it != vecs_.end())
			return it->second;
		fsb_vec_t v = nullptr;
		device::check(fsb_vec_create(ctx_, n_owned, n_ghost, &v));
		vecs_.emplace(fid, v);
		return v;
	}
	std::size_t size() const { return vecs_.size(); }

private:
	fsb_ctx_t ctx_;
	std::unordered_map<field_id, fsb_vec_t> vecs_;
};

}
#endif
