// topo::csr: the distributed-CSR topology a parallel matrix and its vectors live on.
//
// Reference: flecsolve/topo/csr.hh.  Kept: the index spaces {rows, cols, nnz_diag, nnz_offd,
// rowsp1} (:420-421), `cols` = owned columns followed by ghosts (:429-431, 524-526), vec_def<S>
// (:424-427), the `init` description of this process' rows (:474-480) and the metadata
// accessors (:40-50).  Replaced: everything FleCSI did underneath (colouring, repartitions, copy
// plan) now happens inside fsb_parcsr_create (include/fsb.h), which performs color() and
// init_mats() (:482-618) and builds the halo plan (:116-187) on the device side.
#ifndef FLECSOLVE_B200_TOPO_CSR_HH
#define FLECSOLVE_B200_TOPO_CSR_HH

#include <cstdint>
#include <memory>
#include <vector>

#include "flecsolve/device/data.hh"
#include "flecsolve/device/runtime.hh"

namespace flecsolve::topo {

// contiguous row-block partition: offsets[c] .. offsets[c+1] belongs to colour c
struct partition {
	std::vector<std::int64_t> offsets{0, 0};

	// flecsi::util::equal_map(n, bins): the first n % bins blocks get one extra entry
	void set_block_map(std::size_t n, std::size_t bins) {
		offsets.assign(bins + 1, 0);
		const std::size_t q = n / bins, r = n % bins;
		for (std::size_t b = 0; b < bins; ++b)
			offsets[b + 1] = offsets[b] + static_cast<std::int64_t>(q + (b < r ? 1 : 0));
	}
	void set_offsets(std::vector<std::int64_t> off) { offsets = std::move(off); }
	std::size_t size() const { return offsets.size() - 1; }
	std::size_t bin(std::size_t i) const {
		std::size_t lo = 0, hi = size();
		while (hi - lo > 1) {
			const std::size_t mid = (lo + hi) / 2;
			(static_cast<std::int64_t>(i) < offsets[mid] ? hi : lo) = mid;
		}
		return lo;
	}
};

template<class Scalar, class Size = std::size_t>
struct csr {
	using scalar = Scalar;
	using scalar_type = Scalar;
	using size_type = Size;
	enum index_space { rows, cols, nnz_diag, nnz_offd, rowsp1 };
	static constexpr index_space column_space = cols;
	template<index_space S>
	static constexpr int privilege_count = (S == cols) ? 2 : 1;

	template<index_space S>
	using vec_def = data::field_definition<scalar, csr, S>;

	// this process' rows in global numbering (the reference ships one mat::csr per colour in
	// init::proc_mats; one colour per process is the only configuration it instantiates,
	// matrices/parcsr.hh:115)
	struct init {
		std::size_t nrows = 0, ncols = 0;
		partition row_part, col_part;
		std::vector<std::int64_t> offsets; // local rows + 1
		std::vector<std::int64_t> indices; // global column ids
		std::vector<scalar> values;
	};

	struct metadata {
		std::int64_t nrows, ncols;
		struct rng {
			std::int64_t beg, end; // inclusive, as in the reference
			constexpr std::int64_t size() const { return end - beg + 1; }
		} rows, cols;
	};

	// runtime object: owns the device matrix handle and the fields defined on it
	struct topology {
		topology(device::context & c, fsb_parcsr_t m, bool owns_matrix)
			: ctx(c), mat(m), owns(owns_matrix), store(c.handle()) {}
		topology(const topology &) = delete;
		~topology() {
			if (owns && mat)
				fsb_parcsr_destroy(mat);
		}

		std::int64_t local_rows() const { return fsb_parcsr_local_rows(mat); }
		std::int64_t ghosts() const { return fsb_parcsr_num_ghosts(mat); }
		std::size_t colors() const { return static_cast<std::size_t>(ctx.processes()); }
		metadata meta() const {
			const std::int64_t b = fsb_parcsr_row_begin(mat), n = local_rows(), g = fsb_parcsr_global_rows(mat);
			return {g, g, {b, b + n - 1}, {b, b + n - 1}};
		}
		template<index_space S>
		fsb_vec_t storage(data::field_id fid) {
			static_assert(S == rows || S == cols, "vectors live on the rows or cols index space");
			return store.get(fid, local_rows(), S == cols ? ghosts() : 0);
		}

		device::context & ctx;
		fsb_parcsr_t mat;
		bool owns;
		data::field_store store;
	};
	using ptr = std::unique_ptr<topology>;
};

}
#endif
