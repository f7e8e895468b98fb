// b200.hh -- the policy classes a flecsolve maintainer adds to bind the B200 back end (libfsb, include/fsb.h).
//
// This header is compiled against flecsolve's OWN tree (-I <flecsolve source root>): it includes the reference's
// flecsolve/vectors/core.hh, flecsolve/matrices/seq.hh (mat::sparse) and flecsolve/operators/core.hh and only
// supplies what those seams leave open:
//
//   vec::data::b200 / vec::ops::b200     Data / Ops policies of vec::core<Data, Ops, Config>
//                                        (flecsolve/vectors/core.hh:26-353; the members listed in SURVEY.md 8b)
//   b200::fields                         stand-in for the FleCSI topology the vector code touches: static field
//                                        definitions, definition(topology) -> reference {fid(), topology()},
//                                        storage owned by the topology and keyed by field id -- exactly what
//                                        cg::make_work(x) needs (flecsolve/solvers/solver_settings.hh:123-149)
//   vec::make(...)                       the factory overload set of flecsolve/vectors/topo_view.hh:49-82
//   mat::b200_parcsr                     mat::sparse<Data, Ops, Config> whose Ops::spmv is fsb_parcsr_spmv
//                                        (replaces mat::parcsr_ops, flecsolve/matrices/parcsr.hh:52-91)
//   op::b200_dinv                        the Dinv() operator of flecsolve/util/test/mesh.hh:123-140 as an
//                                        element-wise multiply
//
// Every solver, integrator, operator handle and diagnostic of flecsolve then compiles and runs unchanged on device
// vectors: tests/dropin/dropin.cpp does that with the unmodified solvers/{cg,gmres,bicgstab}.hh, vectors/multi.hh
// and time-integrators/bdf.hh of the reference.
//
// Include this header BEFORE flecsolve/solvers/*.hh: the work factory calls vec::make(...) by qualified name, which
// binds to the overloads declared at that point.
#ifndef FSB_FLECSOLVE_B200_HH
#define FSB_FLECSOLVE_B200_HH

#include <atomic>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>

#include <fsb.h>

#include "flecsolve/matrices/seq.hh" // mat::sparse
#include "flecsolve/operators/core.hh"
#include "flecsolve/util/future.hh"
#include "flecsolve/vectors/core.hh"
#include "flecsolve/vectors/variable.hh"

namespace flecsolve::b200 {

struct failure : std::runtime_error {
	int code;
	failure(int c, const char * what) : std::runtime_error(what), code(c) {}
};
inline void ok(int rc) {
	if (rc != FSB_OK)
		throw failure(rc, fsb_last_error());
}

// One index space of device fields, all shaped [n_owned | n_ghost]: the role topo::csr's `cols` space plays for
// FleCSI-backed vectors.  A field comes into being the first time a definition is referenced on the topology
// and lives as long as the topology, so a solver's work vectors are allocated by make_work(), never inside apply().
struct fields {
	enum index_space { cols };
	struct topology {
		topology(fsb_ctx_t c, std::int64_t owned, std::int64_t ghost) : ctx(c), n_owned(owned), n_ghost(ghost) {}
		topology(const topology &) = delete;
		~topology() {
			for (auto & f : store)
				fsb_vec_destroy(f.second);
		}
		fsb_vec_t field(std::size_t fid) {
			auto at = store.find(fid);
			if (at == store.end()) {
				fsb_vec_t v = nullptr;
				ok(fsb_vec_create(ctx, n_owned, n_ghost, &v));
				at = store.emplace(fid, v).first;
			}
			return at->second;
		}
		fsb_ctx_t ctx;
		std::int64_t n_owned, n_ghost;
		std::map<std::size_t, fsb_vec_t> store;
	};
};

struct field_reference {
	std::size_t id;
	fields::topology * where;
	std::size_t fid() const { return id; }
	fields::topology & topology() const { return *where; }
};

// a static object of this type IS a field (its address-independent id is the field id)
struct field_definition {
	field_definition() : id(counter()++) {}
	field_definition(const field_definition &) = delete;
	field_reference operator()(fields::topology & t) const { return {id, &t}; }
	const std::size_t id;

private:
	static std::atomic<std::size_t> & counter() {
		static std::atomic<std::size_t> c{1};
		return c;
	}
};

// the value a reduction returns: get() blocks until the all-rank value is on the host (flecsi::future::get)
struct future {
	fsb_ctx_t ctx;
	fsb_token_t token;
	double get() {
		double v;
		ok(fsb_red_get(ctx, token, &v));
		return v;
	}
	void wait() { ok(fsb_red_wait(ctx, token)); }
};
template<class T>
struct ready {
	T value;
	T get() { return value; }
	void wait() {}
};

}

namespace flecsolve::vec {

template<class Scalar, auto V = anon_var::anonymous>
struct b200_config {
	using scalar = Scalar;
	using real = typename num_traits<scalar>::real;
	using len_t = std::size_t;
	static constexpr auto var = variable<V>;
	using var_t = decltype(V);
	static constexpr std::size_t num_components = 1;
};

namespace data {
template<class Config>
struct b200 {
	using config = Config;
	using scalar = typename Config::scalar;
	using topo_t = ::flecsolve::b200::fields;
	using field_definition = ::flecsolve::b200::field_definition;
	using field_reference = ::flecsolve::b200::field_reference;

	explicit b200(field_reference r) : reference(r), h(r.topology().field(r.fid())) {}

	auto ref() const { return reference; }
	auto fid() const { return reference.fid(); }
	auto & topo() const { return reference.topology(); }
	fsb_ctx_t ctx() const { return reference.topology().ctx; }

	field_reference reference;
	fsb_vec_t h; // n_owned + n_ghost doubles on the device
};
template<class C>
bool operator==(const b200<C> & a, const b200<C> & b) {
	return a.fid() == b.fid();
}
template<class C>
bool operator!=(const b200<C> & a, const b200<C> & b) {
	return a.fid() != b.fid();
}
}

namespace ops {
template<class Data>
struct b200 {
	using scalar = typename Data::scalar;
	using fut = ::flecsolve::b200::future;
	static void ok(int rc) { ::flecsolve::b200::ok(rc); }

	template<class Other>
	static void copy(const Other & x, Data & z) { ok(fsb_vec_copy(z.h, x.h)); }
	static void zero(Data & x) { ok(fsb_vec_set(x.h, 0.0)); }
	static void set_to_scalar(scalar a, Data & x) { ok(fsb_vec_set(x.h, a)); }
	static void scale(scalar a, Data & x) { ok(fsb_vec_scale(x.h, a, x.h)); }
	template<class X>
	static void scale(scalar a, const X & x, Data & y) { ok(fsb_vec_scale(y.h, a, x.h)); }
	template<class X, class Y>
	static void add(const X & x, const Y & y, Data & z) { ok(fsb_vec_add(z.h, x.h, y.h)); }
	template<class X, class Y>
	static void subtract(const X & x, const Y & y, Data & z) { ok(fsb_vec_sub(z.h, x.h, y.h)); }
	template<class X, class Y>
	static void multiply(const X & x, const Y & y, Data & z) { ok(fsb_vec_mul(z.h, x.h, y.h)); }
	template<class X, class Y>
	static void divide(const X & x, const Y & y, Data & z) { ok(fsb_vec_div(z.h, x.h, y.h)); }
	template<class X>
	static void reciprocal(const X & x, Data & y) { ok(fsb_vec_recip(y.h, x.h)); }
	template<class X, class Y>
	static void linear_sum(scalar a, const X & x, scalar b, const Y & y, Data & z) {
		ok(fsb_vec_linear_sum(z.h, a, x.h, b, y.h));
	}
	template<class X, class Y>
	static void axpy(scalar a, const X & x, const Y & y, Data & z) { ok(fsb_vec_axpy(z.h, a, x.h, y.h)); }
	template<class X>
	static void axpby(scalar a, scalar b, const X & x, Data & z) { ok(fsb_vec_axpby(z.h, a, b, x.h)); }
	template<class X>
	static void abs(const X & x, Data & y) { ok(fsb_vec_abs(y.h, x.h)); }
	template<class X>
	static void add_scalar(const X & x, scalar a, Data & y) { ok(fsb_vec_add_scalar(y.h, x.h, a)); }
	static void set_random(Data & x, unsigned seed) { ok(fsb_vec_set_random(x.h, seed)); }
	static void dump(std::string_view prefix, const Data & x) { ok(fsb_vec_dump(x.h, std::string(prefix).c_str())); }

	template<class Y>
	static fut dot(const Data & x, const Y & y) {
		fut f{x.ctx(), 0};
		ok(fsb_vec_dot(x.h, y.h, &f.token));
		return f;
	}
	static fut min(const Data & x) { return red(fsb_vec_min, x); }
	static fut max(const Data & x) { return red(fsb_vec_max, x); }
	static fut inf_norm(const Data & x) { return red(fsb_vec_amax, x); }
	template<unsigned short p>
	static fut lp_norm_local(const Data & x) {
		if constexpr (p == 1)
			return red(fsb_vec_asum, x);
		else if constexpr (p == 2)
			return red(fsb_vec_sumsq, x);
		else {
			fut f{x.ctx(), 0};
			ok(fsb_vec_powsum(x.h, p, &f.token));
			return f;
		}
	}
	// the root is taken when the value is read, as vectors/operations/topo_view.hh:243-257 does
	template<unsigned short p>
	static auto lp_norm(const Data & x) {
		auto f = lp_norm_local<p>(x);
		if constexpr (p == 1)
			return f;
		else if constexpr (p == 2)
			return future_transform{std::move(f), [](double v) { return std::sqrt(v); }};
		else
			return future_transform{std::move(f), [](double v) { return std::pow(v, 1. / p); }};
	}
	static std::size_t local_size(const Data & x) { return static_cast<std::size_t>(fsb_vec_local_size(x.h)); }
	static auto global_size(const Data & x) {
		std::int64_t n = 0;
		ok(fsb_vec_global_size(x.h, &n));
		return ::flecsolve::b200::ready<std::size_t>{static_cast<std::size_t>(n)};
	}
	template<class F, class... Vecs>
	static constexpr decltype(auto) apply(F && f, Vecs &&... vecs) {
		return std::forward<F>(f)(std::forward<Vecs>(vecs)...);
	}

private:
	static fut red(int (*call)(fsb_vec_t, fsb_token_t *), const Data & x) {
		fut f{x.ctx(), 0};
		ok(call(x.h, &f.token));
		return f;
	}
};
}

template<class Scalar = double, auto V = anon_var::anonymous>
using b200_vec = core<data::b200, ops::b200, b200_config<Scalar, V>>;

// ---- vec::make, the overload set of flecsolve/vectors/topo_view.hh:49-82 for device fields
template<auto V>
auto make(variable_t<V>, ::flecsolve::b200::field_reference ref) {
	using vec_t = b200_vec<double, V>;
	return vec_t{data::b200<typename vec_t::config>{ref}};
}
inline auto make(::flecsolve::b200::field_reference ref) { return make(variable<anon_var::anonymous>, ref); }
template<auto V>
auto make(variable_t<V> var, ::flecsolve::b200::fields::topology & topo) {
	return [&topo, var](auto &... fd) {
		if constexpr (sizeof...(fd) == 1)
			return (make(var, fd(topo)), ...);
		else
			return std::tuple(make(var, fd(topo))...);
	};
}
inline auto make(::flecsolve::b200::fields::topology & topo) { return make(variable<anon_var::anonymous>, topo); }

}

namespace flecsolve::mat {

template<class Scalar>
struct b200_parcsr_config {
	using scalar = Scalar;
	using size = std::size_t;
};

template<class Config>
struct b200_parcsr_data {
	using config = Config;
	using topo_t = ::flecsolve::b200::fields;
	fsb_parcsr_t handle = nullptr;
	std::shared_ptr<topo_t::topology> fields; // the `cols` space of this matrix: n_local owned + n_ghost ghost entries
	auto & topo() const { return *fields; }
	auto nrows() const { return static_cast<std::size_t>(fsb_parcsr_global_rows(handle)); }
};

template<class Data>
struct b200_parcsr_ops {
	// ghost exchange of x (if stale) overlapped with the diag block, then the offd block accumulates: replaces
	// execute<spmv_remote>, execute<spmv_local>, y.add(y, tmp) of flecsolve/matrices/parcsr.hh:61-68
	template<class D, class R>
	static void spmv(const D & x, const Data & data, R & y) {
		::flecsolve::b200::ok(fsb_parcsr_spmv(data.handle, x.data.h, y.data.h));
	}
};

template<class Scalar = double>
struct b200_parcsr : sparse<b200_parcsr_data, b200_parcsr_ops, b200_parcsr_config<Scalar>> {
	using base = sparse<b200_parcsr_data, b200_parcsr_ops, b200_parcsr_config<Scalar>>;
	using base::data;
	// adopts a matrix made by fsb_parcsr_create / fsb_parcsr_create_stencil (not destroyed here)
	b200_parcsr(fsb_ctx_t ctx, fsb_parcsr_t A) {
		data.handle = A;
		data.fields = std::make_shared<::flecsolve::b200::fields::topology>(ctx, fsb_parcsr_local_rows(A),
		                                                                   fsb_parcsr_num_ghosts(A));
	}
	auto vec(const ::flecsolve::b200::field_definition & def) { return vec::make(def(data.topo())); }
	std::size_t rows() const { return data.nrows(); }
};

}

namespace flecsolve::op {

// D^-1 held as a vector and applied element-wise; util/test/mesh.hh:123-140 stores the same numbers as a diagonal CSR
template<class Vec>
struct b200_dinv : base<> {
	template<class Matrix>
	b200_dinv(const Matrix & A, Vec storage) : d(std::move(storage)) {
		::flecsolve::b200::ok(fsb_parcsr_extract_dinv(A.data.handle, d.data.h));
	}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		y.multiply(d, x);
	}
	Vec d;
};

}
#endif
