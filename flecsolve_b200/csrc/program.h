// Statement programs: the unit the deferred-execution queue fuses into one launch.
//
// A program is a short straight-line sequence of element-wise statements and
// reductions over "vector slots".  The same constexpr canonicaliser is used
//   (a) at compile time, to instantiate one kernel per registered program, and
//   (b) at run time, to turn a run of queued statements into a lookup key,
// so a queued group matches a registered kernel iff their canonical forms are
// byte-identical.
#pragma once
// (no standard headers: this file is also compiled by NVRTC at run time, csrc/jit.cu)

namespace fsb {

enum ew_op : int {
	OP_SET = 0, // z = s[a]
	OP_SCALE = 1, // z = x * s[a]                 (copy is scale by 1.0)
	OP_LIN2 = 2, // z = s[a]*x + s[b]*y          (add, subtract, axpy, axpby, linear_sum)
	OP_MUL = 3, // z = x * y
	OP_DIV = 4, // z = x / y
	OP_RECIP = 5, // z = 1 / x
	OP_ABS = 6, // z = |x|
	OP_ADDS = 7, // z = x + s[a]
	RD_FIRST = 16,
	RD_DOT = 16, // r[z] += x*y
	RD_ASUM = 17, // r[z] += |x|
	RD_AMAX = 18, // r[z] = max(r[z], |x|)
	RD_MIN = 19, // r[z] = min(r[z], x)
	RD_MAX = 20, // r[z] = max(r[z], x)
	RD_POWSUM = 21 // r[z] += pow(x, s[a])
};

constexpr int MAXS = 12; // statements per program
constexpr int MAXV = 12; // distinct vectors per program
constexpr int MAXSC = 24; // scalars per program
constexpr int MAXR = 4; // reductions per program

constexpr bool is_reduction(int op) { return op >= RD_FIRST; }
constexpr int scalars_of(int op) {
	return op == OP_LIN2 ? 2 : (op == OP_SET || op == OP_SCALE || op == OP_ADDS || op == RD_POWSUM) ? 1 : 0;
}
constexpr bool reads_x(int op) { return op != OP_SET; }
constexpr bool reads_y(int op) { return op == OP_LIN2 || op == OP_MUL || op == OP_DIV || op == RD_DOT; }
// 0 sum, 1 max, 2 min
constexpr int fold_of(int op) { return (op == RD_AMAX || op == RD_MAX) ? 1 : (op == RD_MIN ? 2 : 0); }

// statement before canonicalisation: operands are arbitrary small non-negative ids
struct raw_stmt {
	int op;
	int z, x, y; // -1 when unused; for reductions z is ignored
};

typedef signed char slot_t; // small indices; -1 = unused
struct stmt {
	slot_t op, z, x, y, a, b;
};

// compile-time integer sequences (std::index_sequence without <utility>)
template<int... I>
struct iseq {};
template<int N, int... I>
struct make_iseq_impl : make_iseq_impl<N - 1, N - 1, I...> {};
template<int... I>
struct make_iseq_impl<0, I...> {
	using type = iseq<I...>;
};
template<int N>
using make_iseq = typename make_iseq_impl<N>::type;

struct program {
	int n = 0;
	stmt st[MAXS] = {};
	int nv = 0, ns = 0, nr = 0;
	unsigned load_mask = 0, store_mask = 0;
};

struct canon_result {
	program p;
	int id_of_slot[MAXV] = {}; // raw vector id bound to each slot
	bool ok = true; // false if limits exceeded
};

// Canonical form: slots are numbered in order of first appearance scanning
// statements in order, operands in (x, y, z) order; scalars and reduction
// outputs are numbered in statement order.
constexpr canon_result canonicalize(const raw_stmt * rs, int n) {
	canon_result c;
	if (n > MAXS) {
		c.ok = false;
		return c;
	}
	int nv = 0;
	bool touched[MAXV] = {};
	auto slot = [&](int id) -> int {
		for (int k = 0; k < nv; ++k)
			if (c.id_of_slot[k] == id)
				return k;
		if (nv == MAXV) {
			c.ok = false;
			return 0;
		}
		c.id_of_slot[nv] = id;
		return nv++;
	};
	auto rd = [&](int id) -> slot_t {
		int k = slot(id);
		if (!touched[k]) {
			c.p.load_mask |= 1u << k; // first touch is a read: value comes from memory
			touched[k] = true;
		}
		return static_cast<slot_t>(k);
	};
	for (int i = 0; i < n; ++i) {
		const raw_stmt & r = rs[i];
		stmt s{};
		s.op = static_cast<slot_t>(r.op);
		s.x = s.y = s.z = s.a = s.b = -1;
		if (reads_x(r.op))
			s.x = rd(r.x);
		if (reads_y(r.op))
			s.y = rd(r.y);
		if (is_reduction(r.op)) {
			if (c.p.nr == MAXR) {
				c.ok = false;
				return c;
			}
			s.z = static_cast<slot_t>(c.p.nr++);
		}
		else {
			int k = slot(r.z);
			touched[k] = true;
			c.p.store_mask |= 1u << k;
			s.z = static_cast<slot_t>(k);
		}
		int k = scalars_of(r.op);
		if (c.p.ns + k > MAXSC) {
			c.ok = false;
			return c;
		}
		if (k >= 1)
			s.a = static_cast<slot_t>(c.p.ns++);
		if (k >= 2)
			s.b = static_cast<slot_t>(c.p.ns++);
		if (!c.ok)
			return c;
		c.p.st[i] = s;
	}
	c.p.n = n;
	c.p.nv = nv;
	return c;
}

constexpr bool same_program(const program & a, const program & b) {
	if (a.n != b.n)
		return false;
	for (int i = 0; i < a.n; ++i) {
		const stmt &s = a.st[i], &t = b.st[i];
		if (s.op != t.op || s.z != t.z || s.x != t.x || s.y != t.y || s.a != t.a || s.b != t.b)
			return false;
	}
	return true;
}

} // namespace fsb
