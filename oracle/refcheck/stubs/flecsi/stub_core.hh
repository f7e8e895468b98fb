// STUB (test infrastructure): the minimum of FleCSI's surface needed to *parse* the reference's
// headers and to *run* its serial paths (vec::seq_vec, mat::csr, the Krylov solver templates).
// Nothing here implements FleCSI; every parallel construct is declared but never instantiated.
#pragma once
#include <array>
#include <cassert>
#include <cstddef>
#include <iostream>
#include <map>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#define flog(severity) std::cerr
#define flog_assert(cond, msg)                                                       \
	do {                                                                             \
		if (!(cond)) {                                                               \
			std::cerr << "flog_assert failed: " << msg << std::endl;                 \
			std::abort();                                                            \
		}                                                                            \
	} while (0)
#define flog_fatal(msg)                                                              \
	do {                                                                             \
		std::cerr << msg << std::endl;                                               \
		std::abort();                                                                \
	} while (0)
#define flog_error(msg)                                                              \
	do {                                                                             \
		std::cerr << msg << std::endl;                                               \
	} while (0)
#define flog_warn(msg)                                                               \
	do {                                                                             \
		std::cerr << msg << std::endl;                                               \
	} while (0)
#define FLECSI_INLINE_TARGET inline

// forall / reduceall: only ever appear inside uninstantiated task templates
#define forall(it, range) stub_forall(range)->*[=](auto && it)
#define reduceall(it, up, range, fold, type) template stub_reduce<type>(range)->*[=](auto && it, auto && up)

namespace flecsi {

using Color = std::size_t;
using PrivilegeCount = unsigned short;
using Privileges = unsigned;
enum privilege : unsigned { na = 0, ro = 1, wo = 2, rw = 3 };
template<privilege P, PrivilegeCount N>
inline constexpr Privileges privilege_repeat = 0;
template<Privileges... P>
inline constexpr Privileges privilege_cat = 0;
struct mpi_t {};
inline constexpr mpi_t mpi{};

namespace util {
using id = std::size_t;
using gid = std::size_t;

template<class T>
struct span {
	using element_type = T;
	using value_type = std::remove_cv_t<T>;
	constexpr span() : p(nullptr), n(0) {}
	constexpr span(T * ptr, std::size_t count) : p(ptr), n(count) {}
	template<class C, class = decltype(std::declval<C &>().data())>
	constexpr span(C & c) : p(c.data()), n(c.size()) {}
	constexpr T * data() const { return p; }
	constexpr std::size_t size() const { return n; }
	constexpr T & operator[](std::size_t i) const { return p[i]; }
	constexpr T * begin() const { return p; }
	constexpr T * end() const { return p + n; }
	constexpr T & front() const { return p[0]; }
	constexpr T & back() const { return p[n - 1]; }

private:
	T * p;
	std::size_t n;
};

// column-major multi-dimensional view (only rank 2 is used: solvers/gmres.hh)
template<class T, unsigned short D>
struct mdcolex {
	mdcolex(T * ptr, std::array<std::size_t, D> ext) : p(ptr), e(ext) {}
	T & operator()(std::size_t i, std::size_t j) const { return p[i + j * e[0]]; }
	T * p;
	std::array<std::size_t, D> e;
};

// row-major multi-dimensional view (solvers/nka.hh only; never instantiated by refcheck)
template<class T, unsigned short D>
struct mdspan {
	using size_type = std::size_t;
	mdspan(T * ptr, std::array<std::size_t, D> ext) : p(ptr), e(ext) {}
	T & operator()(std::size_t i, std::size_t j) const { return p[i * e[1] + j]; }
	T * p;
	std::array<std::size_t, D> e;
};

template<class I>
struct iota_view {
	iota_view(I b, I e) : b_(b), e_(e) {}
	I b_, e_;
};

template<auto... V>
struct constants {
	static constexpr std::size_t size = sizeof...(V);
};
template<auto V>
struct constant {};

namespace serial {
template<class T, class = void>
struct traits;
template<class P, class... T>
void put(P &, const T &...);
template<class T>
T get(const std::byte *&);
}
}

namespace exec {
struct on_t {};
inline constexpr on_t on{};
namespace fold {
struct sum {};
struct min {};
struct max {};
}
// shapes only: the task bodies that use them are never instantiated by the serial paths
struct stub_loop {
	template<class F>
	void operator->*(F &&) const {}
};
template<class T>
struct stub_reduction {
	template<class F>
	T operator->*(F &&) const {
		return T{};
	}
};
struct stub_executor {
	stub_executor & named(const char *) { return *this; }
	template<class R>
	stub_loop stub_forall(R &&) {
		return {};
	}
	template<class T, class R>
	stub_reduction<T> stub_reduce(R &&) {
		return {};
	}
};
struct launch_info {
	std::size_t index = 0;
};
struct accelerator {
	stub_executor executor() const { return {}; }
	launch_info launch() const { return {}; }
};
using cpu = accelerator;
}

namespace data {
enum layout { dense };
template<class T, layout L, class Topo, typename Topo::index_space S>
struct field_reference;
template<class Topo, Privileges P>
struct topology_accessor;
}

template<class T, data::layout L = data::dense>
struct field {
	template<class Topo, typename Topo::index_space S>
	struct definition;
	template<class Topo, typename Topo::index_space S>
	struct Reference;
	template<Privileges P>
	struct accessor1;
	template<privilege... P>
	struct accessor;
};

template<class T>
struct topology;

struct stub_future {
	double get() { return 0; }
	void wait() {}
};
struct scheduler {
	static inline scheduler * instance = nullptr;
	template<auto & Task, class... Tags, class... Args>
	void execute(Args &&...) {}
	template<auto & Task, class Fold, class... Args>
	stub_future reduce(Args &&...) {
		return {};
	}
};

namespace run {
struct context {
	static context & instance() {
		static context c;
		return c;
	}
	std::size_t process() const { return 0; }
};
}

}
