// Device-resident scalars for solver variants that never read a reduction on the host
// (include/fsb.h, "device scalars").  No counterpart in the reference: its loops read every
// reduction through future.get() (solvers/cg.hh:98-131); SURVEY 8(f) N1 asks for variants that
// keep alpha/beta on the device and look at the residual norm late.
#ifndef FLECSOLVE_B200_DEVICE_SCALAR_HH
#define FLECSOLVE_B200_DEVICE_SCALAR_HH

#include "flecsolve/device/runtime.hh"
#include "flecsolve/util/future.hh"

namespace flecsolve::device {

struct scalar {
	explicit scalar(fsb_ctx_t c) : ctx_(c) { check(fsb_scalar_create(c, &id_)); }
	scalar(const scalar &) = delete;
	scalar & operator=(const scalar &) = delete;
	~scalar() { fsb_scalar_destroy(ctx_, id_); }

	fsb_scalar_t id() const { return id_; }
	void set(double v) { check(fsb_scalar_set(ctx_, id_, v)); }
	double get() const {
		double v = 0;
		check(fsb_scalar_get(ctx_, id_, &v));
		return v;
	}

private:
	fsb_ctx_t ctx_;
	fsb_scalar_t id_ = 0;
};

// coefficient  s * num / den  evaluated on the device when the statement runs
inline fsb_coef ratio(double s, const scalar & num, const scalar & den) { return fsb_coef{s, num.id(), den.id()}; }
inline fsb_coef number(double v) { return fsb_coef{v, 0, 0}; }

// z = a x + b y  on single (topology) vectors
template<class Z, class X, class Y>
void linear_sum(Z & z, fsb_coef a, const X & x, fsb_coef b, const Y & y) {
	check(fsb_vec_linear_sum_c(z.data.handle(), a, x.data.handle(), b, y.data.handle()));
}

// scalar statements a reduction carries: evaluated on the device, in order, once its all-rank value is stored
struct scalar_program {
	fsb_scalar_op ops[FSB_MAX_POST_OPS];
	int n = 0;
	scalar_program & then(int op, const scalar & dst, const scalar & a, const scalar & b) {
		if (n == FSB_MAX_POST_OPS)
			throw error(FSB_ERR_ARG, "too many scalar statements on one reduction");
		ops[n++] = fsb_scalar_op{op, dst.id(), a.id(), b.id()};
		return *this;
	}
	scalar_program & div(const scalar & dst, const scalar & a, const scalar & b) { return then(FSB_SOP_DIV, dst, a, b); }
	scalar_program & mul(const scalar & dst, const scalar & a, const scalar & b) { return then(FSB_SOP_MUL, dst, a, b); }
	scalar_program & sub(const scalar & dst, const scalar & a, const scalar & b) { return then(FSB_SOP_SUB, dst, a, b); }
	scalar_program & copy(const scalar & dst, const scalar & a) { return then(FSB_SOP_COPY, dst, a, a); }
};

// <x, y>, optionally kept in `store`, tested against the halt threshold, and followed by scalar statements
template<class X, class Y>
device_future dot(const X & x, const Y & y, const scalar * store, int halt_mode = FSB_HALT_NEVER, double threshold = 0,
                  const scalar_program * post = nullptr) {
	fsb_red_opts o{};
	o.store = store ? store->id() : 0;
	o.halt_mode = halt_mode;
	o.halt_threshold = threshold;
	if (post) {
		o.n_post = post->n;
		for (int k = 0; k < post->n; ++k)
			o.post[k] = post->ops[k];
	}
	device_future f{x.data.ctx(), 0};
	check(fsb_vec_dot_opts(x.data.handle(), y.data.handle(), &o, &f.token));
	return f;
}

// while alive, element-wise statements become no-ops once a reduction's halt test has passed
struct halt_scope {
	explicit halt_scope(fsb_ctx_t c) : ctx_(c) { check(fsb_ctx_halt_arm(c)); }
	halt_scope(const halt_scope &) = delete;
	halt_scope & operator=(const halt_scope &) = delete;
	bool release() { // returns whether the flag was raised
		int was = 0;
		if (ctx_)
			check(fsb_ctx_halt_disarm(ctx_, &was));
		ctx_ = nullptr;
		return was != 0;
	}
	~halt_scope() {
		if (ctx_)
			fsb_ctx_halt_disarm(ctx_, nullptr);
	}

private:
	fsb_ctx_t ctx_;
};

}
#endif
