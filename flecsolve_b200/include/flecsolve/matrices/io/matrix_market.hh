// Matrix Market coordinate files -> host CSR -> this rank's rows of a parallel CSR matrix.
//
// Behaviour kept from the reference reader (flecsolve/matrices/io/matrix_market.hh:39-181) because
// the SuiteSparse iteration-count goldens of its tests depend on it:
//   * the header is the first line that does not start with '%': three integers separated by
//     single blanks; the file is symmetric iff any line before it contains "symmetric";
//   * an entry is "row col value" (1-based indices) split at single blanks; the value is converted
//     with std::stof, i.e. rounded to float first, then widened to the matrix scalar;
//   * for a symmetric file every off-diagonal entry is followed by its mirror image;
//   * COO -> CSR is a stable sort by row (seq.hh:324-352): inside a row the file order is kept and
//     repeated entries are not summed.
// `definition` is the lazily read matrix from which each colour takes the rows of its range
// (reference :108-158); init_for() packages that as the topo::csr::init a parcsr is built from.
#ifndef FLECSOLVE_B200_MATRICES_IO_MATRIX_MARKET_HH
#define FLECSOLVE_B200_MATRICES_IO_MATRIX_MARKET_HH

#include <algorithm>
#include <array>
#include <cstdlib>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace flecsolve::mat::io {

template<class scalar, class size>
struct host_csr {
	size nrows = 0, ncols = 0;
	std::vector<size> offsets{0};
	std::vector<size> indices;
	std::vector<scalar> values;
	size nnz() const { return static_cast<size>(indices.size()); }
};

template<class scalar, class size>
struct host_coo {
	size nrows = 0, ncols = 0;
	std::vector<size> I, J;
	std::vector<scalar> V;

	host_csr<scalar, size> tocsr() const {
		std::vector<std::size_t> order(I.size());
		std::iota(order.begin(), order.end(), std::size_t{0});
		std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return I[a] < I[b]; });
		host_csr<scalar, size> out;
		out.nrows = nrows;
		out.ncols = ncols;
		out.offsets.assign(static_cast<std::size_t>(nrows) + 1, 0);
		out.indices.resize(I.size());
		out.values.resize(I.size());
		for (std::size_t k = 0; k < order.size(); ++k) {
			++out.offsets[static_cast<std::size_t>(I[order[k]]) + 1];
			out.indices[k] = J[order[k]];
			out.values[k] = V[order[k]];
		}
		std::partial_sum(out.offsets.begin(), out.offsets.end(), out.offsets.begin());
		return out;
	}
};

template<class scalar = double, class size = std::size_t>
struct matrix_market {
	struct header {
		size nrows = 0, ncols = 0, nnz = 0;
		bool symmetric = false;
	};

	static header read_header(std::ifstream & fh) {
		header hdr;
		std::string line;
		bool found = false;
		while (std::getline(fh, line)) {
			if (line.find("symmetric") != std::string::npos)
				hdr.symmetric = true;
			if (line.empty() || line[0] != '%') {
				std::istringstream in(line);
				std::array<size, 3> dims{};
				for (auto & d : dims) {
					std::string tok;
					std::getline(in, tok, ' ');
					d = static_cast<size>(std::atoi(tok.c_str()));
				}
				hdr.nrows = dims[0];
				hdr.ncols = dims[1];
				hdr.nnz = dims[2];
				found = true;
				break;
			}
		}
		if (!found)
			throw std::runtime_error("matrix market: no size line");
		return hdr;
	}

	static host_coo<scalar, size> read(const char * fname) {
		std::ifstream fh(fname);
		if (!fh)
			throw std::runtime_error(std::string("matrix market: cannot open ") + fname);
		const header hdr = read_header(fh);
		return read(fh, hdr);
	}

	static host_coo<scalar, size> read(std::ifstream & fh, const header & hdr) {
		host_coo<scalar, size> out;
		out.nrows = hdr.nrows;
		out.ncols = hdr.ncols;
		const std::size_t expect = hdr.symmetric ? 2 * static_cast<std::size_t>(hdr.nnz) : static_cast<std::size_t>(hdr.nnz);
		out.I.reserve(expect);
		out.J.reserve(expect);
		out.V.reserve(expect);
		std::string line;
		for (size k = 0; k < hdr.nnz; ++k) {
			if (!std::getline(fh, line))
				throw std::runtime_error("matrix market: fewer entries than the size line promises");
			std::istringstream in(line);
			std::string tok;
			std::getline(in, tok, ' ');
			const size row = static_cast<size>(std::stoi(tok) - 1);
			std::getline(in, tok, ' ');
			const size col = static_cast<size>(std::stoi(tok) - 1);
			std::getline(in, tok, ' ');
			const scalar val = static_cast<scalar>(std::stof(tok)); // float rounding, as the reference
			out.I.push_back(row);
			out.J.push_back(col);
			out.V.push_back(val);
			if (hdr.symmetric && row != col) {
				out.I.push_back(col);
				out.J.push_back(row);
				out.V.push_back(val);
			}
		}
		return out;
	}

	struct definition {
		explicit definition(const char * fname) : fh(fname) {
			if (!fh)
				throw std::runtime_error(std::string("matrix market: cannot open ") + fname);
			hdr = read_header(fh);
		}

		size num_rows() const { return hdr.nrows; }
		size num_cols() const { return hdr.ncols; }

		// rows [first, last] (inclusive) with global column ids
		host_csr<scalar, size> matrix(size first, size last) {
			load();
			host_csr<scalar, size> out;
			out.nrows = last + 1 - first;
			out.ncols = mat.ncols;
			const size z0 = mat.offsets[first], z1 = mat.offsets[last + 1];
			out.offsets.resize(static_cast<std::size_t>(out.nrows) + 1);
			for (size r = first; r <= last + 1; ++r)
				out.offsets[r - first] = mat.offsets[r] - z0;
			out.indices.assign(mat.indices.begin() + z0, mat.indices.begin() + z1);
			out.values.assign(mat.values.begin() + z0, mat.values.begin() + z1);
			return out;
		}
		const host_csr<scalar, size> & whole() {
			load();
			return mat;
		}

		// what topo::csr::init carries for colour `color` of `colors` (equal row blocks, parcsr.hh:170-172)
		template<class Init>
		Init init_for(std::size_t color, std::size_t colors) {
			load();
			Init ci;
			ci.nrows = static_cast<std::size_t>(hdr.nrows);
			ci.ncols = static_cast<std::size_t>(hdr.ncols);
			ci.row_part.set_block_map(ci.nrows, colors);
			ci.col_part.set_block_map(ci.ncols, colors);
			const auto lo = static_cast<size>(ci.row_part.offsets[color]), hi = static_cast<size>(ci.row_part.offsets[color + 1]);
			const size z0 = mat.offsets[lo], z1 = mat.offsets[hi];
			ci.offsets.resize(static_cast<std::size_t>(hi - lo) + 1);
			for (size r = lo; r <= hi; ++r)
				ci.offsets[r - lo] = static_cast<std::int64_t>(mat.offsets[r] - z0);
			ci.indices.assign(mat.indices.begin() + z0, mat.indices.begin() + z1);
			ci.values.assign(mat.values.begin() + z0, mat.values.begin() + z1);
			return ci;
		}

	protected:
		void load() {
			if (!loaded) {
				mat = read(fh, hdr).tocsr();
				loaded = true;
			}
		}
		host_csr<scalar, size> mat;
		header hdr;
		std::ifstream fh;
		bool loaded = false;
	};
};

}
#endif
