// Runs the REFERENCE'S Matrix Market reader (flecsolve/matrices/io/matrix_market.hh, compiled where
// it lies under /root/reference) on a file and prints the CSR it produces as JSON: the golden the
// repo's own reader (flecsolve_b200/include/flecsolve/matrices/io/matrix_market.hh) is pinned against.
// Test infrastructure only.
#include <cstdio>
#include <string>

#include "flecsolve/matrices/io/matrix_market.hh"

int main(int argc, char ** argv) {
	if (argc < 2) {
		std::fprintf(stderr, "usage: refcheck_mtx file.mtx\n");
		return 2;
	}
	using mm = flecsolve::mat::io::matrix_market<double, std::size_t>;
	std::ifstream fh(argv[1]);
	auto hdr = mm::read_header(fh);
	auto csr = mm::read(fh, hdr).tocsr();
	auto [rowptr, colind, values] = csr.rep();
	std::printf("{\"nrows\": %zu, \"ncols\": %zu, \"nnz\": %zu, \"symmetric\": %s, \"rowptr\": [", csr.rows(), csr.cols(),
	            static_cast<std::size_t>(csr.nnz()), hdr.symmetric ? "true" : "false");
	for (std::size_t i = 0; i <= csr.rows(); ++i)
		std::printf("%s%zu", i ? ", " : "", static_cast<std::size_t>(rowptr[i]));
	std::printf("], \"col\": [");
	for (std::size_t i = 0; i < csr.nnz(); ++i)
		std::printf("%s%zu", i ? ", " : "", static_cast<std::size_t>(colind[i]));
	std::printf("], \"val\": [");
	for (std::size_t i = 0; i < csr.nnz(); ++i)
		std::printf("%s\"%a\"", i ? ", " : "", static_cast<double>(values[i]));
	std::printf("]}\n");
	return 0;
}
