// CPU-only harness for ginkgo_b200/host/gko_b200_io.hpp (no CUDA, no library to link):
//   io_check read   <in>  <f64|f32> <i32|i64>          print "rows cols nnz" + "r c value" lines
//   io_check write  <in>  <out> <coordinate|array|binary>   read (generic) and write back
//   io_check csr    <in>  <out>                          through CSR arrays, written as binary
#include <fstream>
#include <iomanip>
#include <iostream>

#include "../../ginkgo_b200/host/gko_b200_io.hpp"

using namespace gko_b200;

template <typename V, typename I>
int do_read(const char* path)
{
    std::ifstream is(path, std::ios::binary);
    auto d = read_generic_raw<V, I>(is);
    std::cout << d.size.rows << ' ' << d.size.cols << ' ' << d.nonzeros.size() << '\n';
    std::cout << std::setprecision(17);
    for (const auto& e : d.nonzeros)
        std::cout << e.row << ' ' << e.column << ' ' << static_cast<double>(e.value) << '\n';
    return 0;
}

int main(int argc, char** argv)
{
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "read" && argc == 5) {
            const std::string vt = argv[3], it = argv[4];
            if (vt == "f64" && it == "i32") return do_read<double, int32>(argv[2]);
            if (vt == "f32" && it == "i32") return do_read<float, int32>(argv[2]);
            if (vt == "f64" && it == "i64") return do_read<double, int64>(argv[2]);
            if (vt == "f32" && it == "i64") return do_read<float, int64>(argv[2]);
        } else if (mode == "write" && argc == 5) {
            std::ifstream is(argv[2], std::ios::binary);
            auto d = read_generic_raw<double, int32>(is);
            std::ofstream os(argv[3], std::ios::binary);
            os << std::setprecision(17);
            const std::string layout = argv[4];
            if (layout == "binary")
                write_binary_raw(os, d);
            else
                write_raw(os, d, layout == "array" ? layout_type::array : layout_type::coordinate);
            return 0;
        }
        if (mode == "csr" && argc == 4) {
            // file -> matrix_data -> CSR arrays -> matrix_data -> binary file (the host half of
            // Csr::read / Csr::write)
            std::ifstream is(argv[2], std::ios::binary);
            auto d = read_generic_raw<double, int32>(is);
            std::vector<double> va;
            std::vector<int32> ci, rp;
            csr_arrays_from_matrix_data(d, rp, ci, va);
            auto back = matrix_data_from_csr_arrays(d.size, rp, ci, va);
            if (!(back.nonzeros == d.nonzeros)) return 4;
            std::ofstream os(argv[3], std::ios::binary);
            write_binary_raw(os, back);
            return 0;
        }
        std::cerr << "usage: io_check read <in> <f64|f32> <i32|i64> | write <in> <out> <layout>\n";
        return 2;
    } catch (const NotSupported& e) {
        std::cerr << "NotSupported: " << e.what() << '\n';
        return 3;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << '\n';
        return 1;
    }
}
