// TEST INFRASTRUCTURE ONLY -- never built into, linked with or loaded by anything under flecsolve_b200/.
//
// A sequential, eager CPU stand-in for the subset of the C ABI (include/fsb.h) that the C++ host layer
// calls.  Purpose: run the flecsolve-shaped templates of flecsolve_b200/include/flecsolve/** (solvers,
// integrators, factory, multi-vectors, narray) in the CPU test suite, where there is no GPU, and hold them
// against the golden vectors produced by the reference's own code (tests/golden/).  It is linked with a
// second copy of flecsolve_b200/host/driver.cpp into tests/hostcheck/_build/libhostcheck.so by
// tests/hostcheck/build.py and driven by tests/test_hostcheck.py.  It says nothing about the device
// kernels -- those are checked on the GPU against the oracle (tests/*_gpu.py).
//
// Arithmetic follows the reference's serial vector operations (vectors/operations/topo_tasks.hh,
// vectors/seq.hh) and serial SpMV (matrices/seq.hh:178-194): one statement per call, products and sums rounded
// separately (build with -ffp-contract=off), reductions accumulated left to right.  One rank, no deferral:
// a reduction token is an index into a table of finished values.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "fsb.h"

struct fsb_ctx_s {
	std::vector<double> results{0.0}; // token -> value (token 0 unused)
	double scalars[64] = {1.0};
	bool scalar_used[64] = {true};
	bool armed = false, halted = false;
	int64_t stats[8] = {};
};

struct fsb_vec_s {
	fsb_ctx_s * ctx;
	std::vector<double> store;
	std::vector<int64_t> map; // dof -> storage offset; empty: identity over the first n entries
	int64_t n = 0, n_ghost = 0;
	int64_t ext[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, cnt[3] = {0, 0, 0};
	bool box = false;
	double & at(int64_t k) { return store[map.empty() ? k : map[k]]; }
	const double & at(int64_t k) const { return store[map.empty() ? k : map[k]]; }
};

struct fsb_parcsr_s {
	fsb_ctx_s * ctx;
	int64_t n = 0;
	std::vector<int64_t> rowptr, col; // box operators: col = storage offsets
	std::vector<double> val;
	std::vector<int64_t> row_store; // box operators: storage offset of each row
	bool box = false;
};

namespace {
// FSB_STANDIN_TRACE=<file>: append one line per C-ABI call that the real library would queue or that flushes its
// queue, so that tests/hostcheck/groups.py can replay the fuser's grouping on the host (which statement groups
// does a solver produce, and do they have ahead-of-time kernels?)
FILE * trace_file() {
	static FILE * f = [] {
		const char * p = std::getenv("FSB_STANDIN_TRACE");
		return p ? std::fopen(p, "a") : nullptr;
	}();
	return f;
}
std::intptr_t tid(const void * p) { return reinterpret_cast<std::intptr_t>(p); }
extern bool g_quiet;
void trace_ew(int op, const fsb_vec_s * z, const fsb_vec_s * x, const fsb_vec_s * y, bool dev = false);
void trace_red(int op, const fsb_vec_s * x, const fsb_vec_s * y, int store);
void trace_mark(const char * what) {
	if (FILE * f = trace_file()) {
		std::fprintf(f, "%s\n", what);
		std::fflush(f);
	}
}
thread_local std::string g_error;
int fail(int code, const char * msg) {
	g_error = msg;
	return code;
}
bool same_space(const fsb_vec_s * a, const fsb_vec_s * b) {
	return a->ctx == b->ctx && a->n == b->n && a->box == b->box && a->map == b->map;
}
bool skip(fsb_ctx_s * c) { return c->armed && c->halted; }
bool g_quiet = false;
void trace_ew(int op, const fsb_vec_s * z, const fsb_vec_s * x, const fsb_vec_s * y, bool dev) {
	if (g_quiet)
		return;
	if (FILE * f = trace_file())
		std::fprintf(f, "EW %d %ld %ld %ld %ld %d\n", op, (long)tid(z), (long)tid(x), (long)tid(y), (long)z->n,
		             (dev || z->ctx->armed) ? 1 : 0);
}
void trace_red(int op, const fsb_vec_s * x, const fsb_vec_s * y, int store) {
	if (FILE * f = trace_file())
		std::fprintf(f, "RED %d %ld %ld %ld %d %d\n", op, (long)tid(x), (long)tid(y), (long)x->n, store, x->ctx->armed ? 1 : 0);
}

template<class F>
int each(fsb_vec_t z, F && f) {
	if (!z)
		return fail(FSB_ERR_ARG, "null vector");
	if (skip(z->ctx))
		return FSB_OK;
	for (int64_t k = 0; k < z->n; ++k)
		f(k);
	return FSB_OK;
}
int finish(fsb_ctx_s * c, double v, fsb_token_t * tok, const fsb_red_opts * o = nullptr) {
	c->results.push_back(v);
	*tok = static_cast<fsb_token_t>(c->results.size() - 1);
	if (o) {
		if (o->store > 0)
			c->scalars[o->store] = v;
		if (o->halt_mode == FSB_HALT_IF_SQRT_LT && std::sqrt(v) < o->halt_threshold)
			c->halted = true;
		if (o->halt_mode == FSB_HALT_IF_LT && v < o->halt_threshold)
			c->halted = true;
		for (int k = 0; k < o->n_post; ++k) {
			const fsb_scalar_op & s = o->post[k];
			const double a = c->scalars[s.a], b = c->scalars[s.b];
			c->scalars[s.dst] = s.op == FSB_SOP_ADD ? a + b : s.op == FSB_SOP_SUB ? a - b : s.op == FSB_SOP_MUL ? a * b : s.op == FSB_SOP_DIV ? a / b : a;
		}
	}
	return FSB_OK;
}
double coef(const fsb_ctx_s * c, fsb_coef k) {
	if (k.num == 0 && k.den == 0)
		return k.scale;
	return k.scale * (c->scalars[k.num] / c->scalars[k.den]);
}
}

extern "C" {

const char * fsb_last_error(void) { return g_error.c_str(); }
int fsb_version(void) { return 1; }
int fsb_device_count(void) { return 0; }

int fsb_ctx_create(int, int rank, int nranks, const void *, fsb_ctx_t * out) {
	if (rank != 0 || nranks != 1)
		return fail(FSB_ERR_ARG, "stand-in: one rank only");
	*out = new fsb_ctx_s;
	return FSB_OK;
}
int fsb_ctx_destroy(fsb_ctx_t c) {
	delete c;
	return FSB_OK;
}
int fsb_ctx_set_option(fsb_ctx_t, int, int64_t) { return FSB_OK; } // tuning knobs of the device library: nothing to do here
int fsb_ctx_flush(fsb_ctx_t) {
	trace_mark("FLUSH");
	return FSB_OK;
}
int fsb_ctx_sync(fsb_ctx_t) {
	trace_mark("FLUSH");
	return FSB_OK;
}
int fsb_ctx_rank(fsb_ctx_t) { return 0; }
int fsb_ctx_nranks(fsb_ctx_t) { return 1; }
int fsb_ctx_get_stat(fsb_ctx_t c, int s, int64_t * out) {
	*out = c->stats[s & 7];
	return FSB_OK;
}
int fsb_ctx_event_record(fsb_ctx_t, int) {
	trace_mark("FLUSH");
	return FSB_OK;
}
int fsb_ctx_halt_arm(fsb_ctx_t c) {
	trace_mark("FLUSH");
	c->armed = true;
	c->halted = false;
	return FSB_OK;
}
int fsb_ctx_halt_disarm(fsb_ctx_t c, int * was) {
	trace_mark("FLUSH");
	if (was)
		*was = c->halted ? 1 : 0;
	c->armed = c->halted = false;
	return FSB_OK;
}

// ---- vectors
int fsb_vec_create(fsb_ctx_t c, int64_t n, int64_t g, fsb_vec_t * out) {
	auto * v = new fsb_vec_s;
	v->ctx = c;
	v->n = n;
	v->n_ghost = g;
	v->store.assign(static_cast<size_t>(n + g), 0.0);
	*out = v;
	return FSB_OK;
}
int fsb_vec_create_box(fsb_ctx_t c, int dim, const int64_t * ext, const int64_t * lo, const int64_t * hi, fsb_vec_t * out) {
	auto * v = new fsb_vec_s;
	v->ctx = c;
	v->box = true;
	int64_t e[3] = {1, 1, 1}, l[3] = {0, 0, 0}, m[3] = {1, 1, 1};
	for (int a = 0; a < dim; ++a) {
		e[a] = ext[a];
		l[a] = lo[a];
		m[a] = hi[a] - lo[a];
	}
	v->store.assign(static_cast<size_t>(e[0] * e[1] * e[2]), 0.0);
	for (int64_t k = 0; k < m[2]; ++k)
		for (int64_t j = 0; j < m[1]; ++j)
			for (int64_t i = 0; i < m[0]; ++i)
				v->map.push_back((l[0] + i) + e[0] * ((l[1] + j) + e[1] * (l[2] + k)));
	v->n = static_cast<int64_t>(v->map.size());
	v->n_ghost = static_cast<int64_t>(v->store.size()) - v->n;
	std::copy(e, e + 3, v->ext);
	std::copy(l, l + 3, v->lo);
	std::copy(m, m + 3, v->cnt);
	*out = v;
	return FSB_OK;
}
int fsb_vec_destroy(fsb_vec_t v) {
	delete v;
	return FSB_OK;
}
int64_t fsb_vec_local_size(fsb_vec_t v) { return v->n; }
int64_t fsb_vec_ghost_size(fsb_vec_t v) { return v->n_ghost; }
int fsb_vec_upload(fsb_vec_t v, const double * h, int64_t n, int64_t off) {
	trace_mark("FLUSH");
	for (int64_t k = 0; k < n; ++k)
		v->at(off + k) = h[k];
	return FSB_OK;
}
int fsb_vec_download(fsb_vec_t v, double * h, int64_t n, int64_t off) {
	trace_mark("FLUSH");
	for (int64_t k = 0; k < n; ++k)
		h[k] = v->at(off + k);
	return FSB_OK;
}
int fsb_vec_global_size(fsb_vec_t v, int64_t * out) {
	*out = v->n;
	return FSB_OK;
}

#define NEED_SAME(a, b)                                                    \
	if (!(a) || !(b) || !same_space(a, b))                                 \
	return fail(FSB_ERR_ARG, "stand-in: operands live on different index sets")

int fsb_vec_copy(fsb_vec_t z, fsb_vec_t x) {
	NEED_SAME(z, x);
	if (z != x)
		trace_ew(1, z, x, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k); });
}
int fsb_vec_set(fsb_vec_t z, double a) {
	trace_ew(0, z, nullptr, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = a; });
}
int fsb_vec_scale(fsb_vec_t z, double a, fsb_vec_t x) {
	NEED_SAME(z, x);
	trace_ew(1, z, x, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) * a; });
}
int fsb_vec_add(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(2, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) + y->at(k); });
}
int fsb_vec_sub(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(2, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) - y->at(k); });
}
int fsb_vec_mul(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(3, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) * y->at(k); });
}
int fsb_vec_div(fsb_vec_t z, fsb_vec_t x, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(4, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) / y->at(k); });
}
int fsb_vec_recip(fsb_vec_t z, fsb_vec_t x) {
	NEED_SAME(z, x);
	trace_ew(5, z, x, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = 1.0 / x->at(k); });
}
int fsb_vec_linear_sum(fsb_vec_t z, double a, fsb_vec_t x, double b, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(2, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = a * x->at(k) + b * y->at(k); });
}
int fsb_vec_linear_sum_c(fsb_vec_t z, fsb_coef a, fsb_vec_t x, fsb_coef b, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(2, z, x, y, a.num != 0 || a.den != 0 || b.num != 0 || b.den != 0);
	g_quiet = true;
	if (skip(z->ctx)) { // coefficients of a halted solve may be meaningless (0/0): do not even form them
		g_quiet = false;
		return FSB_OK;
	}
	const int rc = fsb_vec_linear_sum(z, coef(z->ctx, a), x, coef(z->ctx, b), y);
	g_quiet = false;
	return rc;
}
// a x + y: the reference's axpy body has no multiplication on y (topo_tasks.hh:174-191); 1.0 * y is exact anyway
int fsb_vec_axpy(fsb_vec_t z, double a, fsb_vec_t x, fsb_vec_t y) {
	NEED_SAME(z, x);
	NEED_SAME(z, y);
	trace_ew(2, z, x, y);
	return each(z, [&](int64_t k) { z->at(k) = a * x->at(k) + y->at(k); });
}
int fsb_vec_axpby(fsb_vec_t z, double a, double b, fsb_vec_t x) {
	NEED_SAME(z, x);
	trace_ew(2, z, x, z);
	return each(z, [&](int64_t k) { z->at(k) = a * x->at(k) + b * z->at(k); });
}
int fsb_vec_abs(fsb_vec_t z, fsb_vec_t x) {
	NEED_SAME(z, x);
	trace_ew(6, z, x, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = std::fabs(x->at(k)); });
}
int fsb_vec_add_scalar(fsb_vec_t z, fsb_vec_t x, double a) {
	NEED_SAME(z, x);
	trace_ew(7, z, x, nullptr);
	return each(z, [&](int64_t k) { z->at(k) = x->at(k) + a; });
}
int fsb_vec_set_random(fsb_vec_t z, unsigned seed) {
	trace_mark("FLUSH");
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> dis(0., 1.);
	for (int64_t k = 0; k < z->n; ++k)
		z->at(k) = dis(gen);
	return FSB_OK;
}
int fsb_vec_dump(fsb_vec_t x, const char * prefix) {
	trace_mark("FLUSH");
	std::ofstream f(std::string(prefix) + "-0");
	for (int64_t k = 0; k < x->n; ++k)
		f << x->at(k) << '\n';
	return FSB_OK;
}

// ---- reductions (a halted context delivers the fold identity, like a skipped pass on the device)
static double dot_of(fsb_vec_t x, fsb_vec_t y) {
	double s = 0.0;
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s += x->at(k) * y->at(k);
	return s;
}
int fsb_vec_dot(fsb_vec_t x, fsb_vec_t y, fsb_token_t * t) {
	NEED_SAME(x, y);
	trace_red(16, x, y, 0);
	return finish(x->ctx, dot_of(x, y), t);
}
int fsb_vec_dot_opts(fsb_vec_t x, fsb_vec_t y, const fsb_red_opts * o, fsb_token_t * t) {
	NEED_SAME(x, y);
	trace_red(16, x, y, o ? o->store : 0);
	if (skip(x->ctx))
		return finish(x->ctx, 0.0, t); // nothing is stored or tested after the halt
	return finish(x->ctx, dot_of(x, y), t, o);
}
int fsb_vec_sumsq(fsb_vec_t x, fsb_token_t * t) {
	trace_red(16, x, x, 0);
	return finish(x->ctx, dot_of(x, x), t);
}
int fsb_vec_asum(fsb_vec_t x, fsb_token_t * t) {
	trace_red(17, x, nullptr, 0);
	double s = 0.0;
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s += std::fabs(x->at(k));
	return finish(x->ctx, s, t);
}
int fsb_vec_powsum(fsb_vec_t x, int p, fsb_token_t * t) {
	trace_red(21, x, nullptr, 0);
	double s = 0.0;
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s += std::pow(x->at(k), p);
	return finish(x->ctx, s, t);
}
int fsb_vec_amax(fsb_vec_t x, fsb_token_t * t) {
	trace_red(18, x, nullptr, 0);
	double s = -std::numeric_limits<double>::infinity();
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s = std::max(s, std::fabs(x->at(k)));
	return finish(x->ctx, s, t);
}
int fsb_vec_min(fsb_vec_t x, fsb_token_t * t) {
	trace_red(19, x, nullptr, 0);
	double s = std::numeric_limits<double>::infinity();
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s = std::min(s, x->at(k));
	return finish(x->ctx, s, t);
}
int fsb_vec_max(fsb_vec_t x, fsb_token_t * t) {
	trace_red(20, x, nullptr, 0);
	double s = -std::numeric_limits<double>::infinity();
	if (!skip(x->ctx))
		for (int64_t k = 0; k < x->n; ++k)
			s = std::max(s, x->at(k));
	return finish(x->ctx, s, t);
}
int fsb_red_get(fsb_ctx_t c, fsb_token_t t, double * out) {
	trace_mark("FLUSH");
	if (t <= 0 || t >= static_cast<fsb_token_t>(c->results.size()))
		return fail(FSB_ERR_ARG, "unknown reduction token");
	*out = c->results[static_cast<size_t>(t)];
	return FSB_OK;
}
int fsb_red_wait(fsb_ctx_t, fsb_token_t) {
	trace_mark("FLUSH");
	return FSB_OK;
}

// ---- device scalars
int fsb_scalar_create(fsb_ctx_t c, fsb_scalar_t * out) {
	for (int k = 1; k < 64; ++k)
		if (!c->scalar_used[k]) {
			c->scalar_used[k] = true;
			*out = k;
			return FSB_OK;
		}
	return fail(FSB_ERR_STATE, "out of scalars");
}
int fsb_scalar_destroy(fsb_ctx_t c, fsb_scalar_t s) {
	c->scalar_used[s] = false;
	return FSB_OK;
}
int fsb_scalar_set(fsb_ctx_t c, fsb_scalar_t s, double v) {
	trace_mark("FLUSH");
	c->scalars[s] = v;
	return FSB_OK;
}
int fsb_scalar_get(fsb_ctx_t c, fsb_scalar_t s, double * out) {
	trace_mark("FLUSH");
	*out = c->scalars[s];
	return FSB_OK;
}

// ---- matrices (one rank: the whole matrix is the diag block, rows kept in the order given)
int fsb_parcsr_create(fsb_ctx_t c, int64_t n, const int64_t * part, const int64_t * rowptr, const int64_t * col,
                      const double * val, fsb_parcsr_t * out) {
	if (part[0] != 0 || part[1] != n)
		return fail(FSB_ERR_ARG, "stand-in: one rank only");
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->n = n;
	A->rowptr.assign(rowptr, rowptr + n + 1);
	A->col.assign(col, col + rowptr[n]);
	A->val.assign(val, val + rowptr[n]);
	*out = A;
	return FSB_OK;
}
// the synthetic operators of SURVEY 8(d): rows g = i + nx (j + ny k), Dirichlet truncation, columns ascending,
// centre (4 | 6 | 26 + diag_shift) * scale, neighbours -scale
int fsb_parcsr_create_stencil(fsb_ctx_t c, int kind, int64_t nx, int64_t ny, int64_t nz, double diag_shift, double scale,
                              fsb_parcsr_t * out) {
	if (kind != 5 && kind != 7 && kind != 27)
		return fail(FSB_ERR_ARG, "stencil kind must be 5, 7 or 27");
	if (kind == 5)
		nz = 1;
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->n = nx * ny * nz;
	const double dv = ((kind == 27 ? 26.0 : kind == 7 ? 6.0 : 4.0) + diag_shift) * scale, ov = -1.0 * scale;
	A->rowptr.push_back(0);
	for (int64_t g = 0; g < A->n; ++g) {
		const int64_t i = g % nx, j = (g / nx) % ny, k = g / (nx * ny);
		for (int dk = -1; dk <= 1; ++dk)
			for (int dj = -1; dj <= 1; ++dj)
				for (int di = -1; di <= 1; ++di) {
					const int64_t ii = i + di, jj = j + dj, kk = k + dk;
					if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz)
						continue;
					const int nonzero = (di != 0) + (dj != 0) + (dk != 0);
					if (kind != 27 && nonzero > 1)
						continue; // 5- and 7-point: axis neighbours only
					A->col.push_back(ii + nx * (jj + ny * kk));
					A->val.push_back(nonzero == 0 ? dv : ov);
				}
		A->rowptr.push_back(static_cast<int64_t>(A->col.size()));
	}
	*out = A;
	return FSB_OK;
}
int fsb_parcsr_create_box_stencil(fsb_ctx_t c, int dim, const int64_t * ext, const int64_t * lo, const int64_t * hi,
                                  double center, const double * off, fsb_parcsr_t * out) {
	fsb_vec_t shape = nullptr;
	fsb_vec_create_box(c, dim, ext, lo, hi, &shape);
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->box = true;
	A->n = shape->n;
	A->row_store = shape->map;
	const int64_t stride[3] = {1, shape->ext[0], shape->ext[0] * shape->ext[1]};
	A->rowptr.push_back(0);
	for (int64_t r = 0; r < A->n; ++r) {
		const int64_t at = shape->map[r];
		for (int a = dim - 1; a >= 0; --a) {
			A->col.push_back(at - stride[a]);
			A->val.push_back(off[a]);
		}
		A->col.push_back(at);
		A->val.push_back(center);
		for (int a = 0; a < dim; ++a) {
			A->col.push_back(at + stride[a]);
			A->val.push_back(off[a]);
		}
		A->rowptr.push_back(static_cast<int64_t>(A->col.size()));
	}
	delete shape;
	*out = A;
	return FSB_OK;
}
// same entries as the device library documents (include/fsb.h): towards c +- e  -(beta (b kface)), centre  beta sum (b kface) +
// (alpha vol) a
int fsb_parcsr_create_box_fvm(fsb_ctx_t c, int dim, const int64_t * ext, const int64_t * lo, const int64_t * hi, double beta,
                              double alpha, double vol, const double * kface, const double * a, const double * const * bface,
                              fsb_parcsr_t * out) {
	fsb_vec_t shape = nullptr;
	fsb_vec_create_box(c, dim, ext, lo, hi, &shape);
	auto * A = new fsb_parcsr_s;
	A->ctx = c;
	A->box = true;
	A->n = shape->n;
	A->row_store = shape->map;
	const int64_t stride[3] = {1, shape->ext[0], shape->ext[0] * shape->ext[1]};
	auto face = [&](int64_t at, int ax) { return bface[ax][at] * kface[ax]; };
	A->rowptr.push_back(0);
	for (int64_t r = 0; r < A->n; ++r) {
		const int64_t at = shape->map[r];
		for (int ax = dim - 1; ax >= 0; --ax) {
			A->col.push_back(at - stride[ax]);
			A->val.push_back(-(beta * face(at - stride[ax], ax)));
		}
		double sum = 0.0;
		for (int ax = 0; ax < dim; ++ax)
			sum += face(at, ax) + face(at - stride[ax], ax);
		A->col.push_back(at);
		A->val.push_back(beta * sum + (alpha * vol) * a[at]);
		for (int ax = 0; ax < dim; ++ax) {
			A->col.push_back(at + stride[ax]);
			A->val.push_back(-(beta * face(at, ax)));
		}
		A->rowptr.push_back(static_cast<int64_t>(A->col.size()));
	}
	delete shape;
	*out = A;
	return FSB_OK;
}
int fsb_parcsr_destroy(fsb_parcsr_t A) {
	delete A;
	return FSB_OK;
}
int64_t fsb_parcsr_local_rows(fsb_parcsr_t A) { return A->n; }
int64_t fsb_parcsr_global_rows(fsb_parcsr_t A) { return A->n; }
int64_t fsb_parcsr_num_ghosts(fsb_parcsr_t) { return 0; }
int64_t fsb_parcsr_row_begin(fsb_parcsr_t) { return 0; }
// matrices/seq.hh:178-194: y[i] = 0; y[i] += val * x[col] in index order.  Not subject to the halt flag
// (the device does not skip SpMV either: it only writes work vectors).
int fsb_parcsr_spmv(fsb_parcsr_t A, fsb_vec_t x, fsb_vec_t y) {
	if (x == y)
		return fail(FSB_ERR_ARG, "spmv: x and y must be different vectors");
	if (x->n != A->n || y->n != A->n || x->box != A->box || y->box != A->box)
		return fail(FSB_ERR_ARG, "spmv: vector does not fit the matrix");
	if (FILE * f = trace_file())
		std::fprintf(f, "SPMV %ld %ld %ld\n", (long)tid(x), (long)tid(y), (long)x->n);
	for (int64_t r = 0; r < A->n; ++r) {
		double s = 0.0;
		for (int64_t p = A->rowptr[r]; p < A->rowptr[r + 1]; ++p)
			s += A->val[p] * (A->box ? x->store[A->col[p]] : x->store[A->col[p]]);
		(A->box ? y->store[A->row_store[r]] : y->store[r]) = s;
	}
	return FSB_OK;
}
int fsb_parcsr_extract_dinv(fsb_parcsr_t A, fsb_vec_t d) {
	trace_mark("FLUSH");
	if (A->box)
		return fail(FSB_ERR_ARG, "extract_dinv: not available for structured-grid operators");
	for (int64_t r = 0; r < A->n; ++r) {
		double a = 0.0;
		for (int64_t p = A->rowptr[r]; p < A->rowptr[r + 1]; ++p)
			if (A->col[p] == r)
				a = A->val[p];
		d->store[r] = 1.0 / a;
	}
	return FSB_OK;
}
int fsb_parcsr_jacobi_relax(fsb_parcsr_t, double, int64_t, fsb_vec_t, fsb_vec_t, fsb_vec_t) {
	return fail(FSB_ERR_STATE, "stand-in: jacobi_relax is a device kernel, checked on the GPU");
}

} // extern "C"
