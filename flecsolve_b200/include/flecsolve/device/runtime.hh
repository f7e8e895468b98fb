// RAII access to the C ABI (include/fsb.h) for the header layer.
//
// device::context stands where the reference passes `flecsi::scheduler &` and an MPI
// communicator (matrices/parcsr.hh:112-124): it names the GPU this process drives and its
// rank in the row-block partition.  Failures of the C ABI become exceptions here
// (the reference aborts through flog_assert).
#ifndef FLECSOLVE_B200_DEVICE_RUNTIME_HH
#define FLECSOLVE_B200_DEVICE_RUNTIME_HH

#include <stdexcept>
#include <string>

#include "fsb.h"

namespace flecsolve::device {

struct error : std::runtime_error {
	int code;
	error(int c, const char * msg) : std::runtime_error(std::string("fsb: ") + msg), code(c) {}
};

inline void check(int rc) {
	if (rc != FSB_OK)
		throw error(rc, fsb_last_error());
}

struct context {
	context(int device_index = 0, int rank = 0, int nranks = 1, const void * nccl_unique_id = nullptr) : owned_(true) {
		check(fsb_ctx_create(device_index, rank, nranks, nccl_unique_id, &h_));
	}
	// adopt a context created elsewhere (not destroyed here)
	explicit context(fsb_ctx_t h) : h_(h), owned_(false) {}
	context(const context &) = delete;
	context & operator=(const context &) = delete;
	~context() {
		if (owned_ && h_)
			fsb_ctx_destroy(h_);
	}

	fsb_ctx_t handle() const { return h_; }
	int rank() const { return fsb_ctx_rank(h_); }
	int processes() const { return fsb_ctx_nranks(h_); }
	void sync() const { check(fsb_ctx_sync(h_)); }
	void flush() const { check(fsb_ctx_flush(h_)); }

private:
	fsb_ctx_t h_ = nullptr;
	bool owned_;
};

}
#endif
