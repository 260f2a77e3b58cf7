// gko_b200_io.hpp -- the on-disk formats on the input side of the path:
// gko::matrix_data and the MatrixMarket / Ginkgo-binary readers and writers
// (include/ginkgo/core/base/matrix_data.hpp:455-510, include/ginkgo/core/base/mtx_io.hpp,
// core/base/mtx_io.cpp).  Pure host code, no CUDA: a matrix is parsed into matrix_data, and
// matrix::Csr::read / write (gko_b200.hpp) move it to and from the device.
//
// Behaviour mirrored from the reference:
//  * header "%%MatrixMarket matrix <coordinate|array> <real|integer|pattern> <general|
//    symmetric|skew-symmetric>", case-insensitive, leading empty lines and comment lines
//    skipped (mtx_io.cpp:676-731); complex / hermitian files are rejected (this path is real);
//  * coordinate entries are 1-based, the rest of each line is ignored; symmetric storage adds
//    the transposed entry, skew-symmetric the negated one, never for the diagonal
//    (:272-400); array layout is column-major with the row start of the modifier (:563-600);
//  * values are parsed as double and converted (:129-134); pattern entries are 1;
//  * the result is sorted row-major (:80);
//  * write_raw emits "general" files, coordinate in storage order or array column-major with
//    explicit zeros (:537-640, :980-989);
//  * binary format: 32-byte header {magic "GINKGO" + value char + index char, rows, cols,
//    entries} followed by (row, col, value) records; any stored value / index type is
//    converted on read (:762-978); read_generic_raw picks by the first character (:930-939).
#pragma once

#include <algorithm>
#include <array>
#include <cctype>
#include <cstring>
#include <istream>
#include <limits>
#include <ostream>
#include <sstream>
#include <tuple>
#include <vector>

#include "gko_b200_types.hpp"

namespace gko_b200 {

template <typename V, typename I>
struct matrix_data_entry {
    I row;
    I column;
    V value;
    bool operator==(const matrix_data_entry& o) const
    {
        return row == o.row && column == o.column && value == o.value;
    }
};

template <typename V = double, typename I = int32>
struct matrix_data {
    using value_type = V;
    using index_type = I;
    using nonzero_type = matrix_data_entry<V, I>;
    dim2 size;
    std::vector<nonzero_type> nonzeros;

    matrix_data() = default;
    explicit matrix_data(dim2 s) : size(s) {}
    matrix_data(dim2 s, std::vector<nonzero_type> nz) : size(s), nonzeros(std::move(nz)) {}

    void sort_row_major()
    {
        std::sort(nonzeros.begin(), nonzeros.end(), [](const nonzero_type& x, const nonzero_type& y) {
            return std::tie(x.row, x.column) < std::tie(y.row, y.column);
        });
    }
    void remove_zeros()
    {
        nonzeros.erase(std::remove_if(nonzeros.begin(), nonzeros.end(),
                                      [](const nonzero_type& nz) { return nz.value == V(0); }),
                       nonzeros.end());
    }
    void sum_duplicates()
    {
        sort_row_major();
        std::vector<nonzero_type> out;
        if (!nonzeros.empty()) {
            out.push_back({nonzeros.front().row, nonzeros.front().column, V(0)});
            for (const auto& e : nonzeros) {
                if (e.row != out.back().row || e.column != out.back().column)
                    out.push_back({e.row, e.column, V(0)});
                out.back().value += e.value;
            }
            nonzeros = std::move(out);
        }
    }
};

enum class layout_type { array, coordinate };

namespace detail {

inline void stream_check(bool ok, const std::string& msg)
{
    if (!ok) throw StreamError(msg);
}

struct mtx_header {
    bool coordinate = true;
    int entry = 0;     // 0 real / integer, 1 pattern
    int modifier = 0;  // 0 general, 1 symmetric, 2 skew-symmetric
    std::string dimensions_line;
};

inline mtx_header read_mtx_header(std::istream& is)
{
    mtx_header h;
    std::string line;
    do {
        stream_check((bool)std::getline(is, line), "error when reading the header line");
    } while (line.empty());
    std::transform(line.begin(), line.end(), line.begin(),
                   [](unsigned char c) { return (char)std::tolower(c); });
    std::istringstream ls(line);
    std::string banner, object, layout, entry, modifier;
    ls >> banner >> object >> layout >> entry >> modifier;
    const bool ok = banner == "%%matrixmarket" && object == "matrix" &&
                    (layout == "coordinate" || layout == "array") &&
                    (entry == "real" || entry == "integer" || entry == "pattern" ||
                     entry == "complex") &&
                    (modifier == "general" || modifier == "symmetric" ||
                     modifier == "skew-symmetric" || modifier == "hermitian");
    stream_check(ok,
                 "error parsing the header line, expected %%MatrixMarket matrix "
                 "<coordinate|array> <real|integer|complex|pattern> "
                 "<general|symmetric|skew-symmetric|hermitian>, found: " + line);
    if (entry == "complex" || modifier == "hermitian")
        throw NotSupported("complex / hermitian MatrixMarket files are not on this path");
    h.coordinate = layout == "coordinate";
    h.entry = entry == "pattern" ? 1 : 0;
    h.modifier = modifier == "general" ? 0 : modifier == "symmetric" ? 1 : 2;
    do {
        stream_check((bool)std::getline(is, h.dimensions_line),
                     "error when reading the dimensions line");
    } while (!h.dimensions_line.empty() && h.dimensions_line[0] == '%');
    return h;
}

template <typename V>
inline V read_entry(std::istream& is, int entry_kind)
{
    if (entry_kind == 1) return V(1);
    double v{};
    stream_check((bool)(is >> v), "error while reading matrix entry");
    return static_cast<V>(v);
}

template <typename V, typename I>
inline void insert_entry(matrix_data<V, I>& data, int modifier, I row, I col, V value)
{
    data.nonzeros.push_back({row, col, value});
    if (modifier != 0 && row != col)
        data.nonzeros.push_back({col, row, modifier == 1 ? value : static_cast<V>(-value)});
}

template <typename V, typename I>
constexpr std::uint64_t binary_magic()
{
    const char vc = std::is_same<V, double>::value ? 'D' : 'S';
    const char ic = sizeof(I) == 4 ? 'I' : 'L';
    const char m[8] = {'G', 'I', 'N', 'K', 'G', 'O', vc, ic};
    std::uint64_t r = 0;
    for (int i = 7; i >= 0; --i) r = r * 256 + (unsigned char)m[i];
    return r;
}

template <typename FV, typename FI, typename V, typename I>
matrix_data<V, I> read_binary_convert(std::istream& is, std::uint64_t rows, std::uint64_t cols,
                                      std::uint64_t entries)
{
    stream_check(rows <= (std::uint64_t)std::numeric_limits<I>::max() &&
                     cols <= (std::uint64_t)std::numeric_limits<I>::max(),
                 "cannot read into this format, its dimensions would overflow");
    matrix_data<V, I> result(dim2{(size_type)rows, (size_type)cols});
    result.nonzeros.resize(entries);
    constexpr size_type block = sizeof(FV) + 2 * sizeof(FI);
    for (std::uint64_t i = 0; i < entries; ++i) {
        std::array<char, block> buf;
        stream_check((bool)is.read(buf.data(), block), "failed reading entry " + std::to_string(i));
        FI r{}, c{};
        FV v{};
        std::memcpy(&r, &buf[0], sizeof(FI));
        std::memcpy(&c, &buf[sizeof(FI)], sizeof(FI));
        std::memcpy(&v, &buf[2 * sizeof(FI)], sizeof(FV));
        result.nonzeros[i] = {static_cast<I>(r), static_cast<I>(c), static_cast<V>(v)};
    }
    result.sort_row_major();
    return result;
}

}  // namespace detail

// gko::read_raw (MatrixMarket text)
template <typename V = double, typename I = int32>
matrix_data<V, I> read_raw(std::istream& is)
{
    const auto h = detail::read_mtx_header(is);
    std::istringstream dims(h.dimensions_line);
    size_type rows{}, cols{};
    matrix_data<V, I> data;
    if (h.coordinate) {
        size_type nnz{};
        detail::stream_check((bool)(dims >> rows >> cols >> nnz),
                             "error when determining matrix size, expected: rows cols nnz");
        data = matrix_data<V, I>(dim2{rows, cols});
        data.nonzeros.reserve(h.modifier == 0 ? nnz : 2 * nnz);
        for (size_type i = 0; i < nnz; ++i) {
            I r{}, c{};
            detail::stream_check((bool)(is >> r >> c),
                                 "error when reading coordinates of matrix entry " + std::to_string(i));
            const V v = detail::read_entry<V>(is, h.entry);
            detail::insert_entry(data, h.modifier, static_cast<I>(r - 1), static_cast<I>(c - 1), v);
            is.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
        }
    } else {
        detail::stream_check((bool)(dims >> rows >> cols),
                             "error when determining matrix size, expected: rows cols");
        data = matrix_data<V, I>(dim2{rows, cols});
        for (size_type c = 0; c < cols; ++c) {
            const size_type start = h.modifier == 0 ? 0 : h.modifier == 1 ? c : c + 1;
            for (size_type r = start; r < rows; ++r) {
                const V v = detail::read_entry<V>(is, h.entry);
                detail::insert_entry(data, h.modifier, static_cast<I>(r), static_cast<I>(c), v);
                is.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
            }
        }
    }
    data.sort_row_major();
    return data;
}

// gko::read_binary_raw
template <typename V = double, typename I = int32>
matrix_data<V, I> read_binary_raw(std::istream& is)
{
    std::array<char, 32> header{};
    detail::stream_check((bool)is.read(header.data(), 32), "failed reading header");
    std::uint64_t magic{}, rows{}, cols{}, entries{};
    std::memcpy(&magic, &header[0], 8);
    std::memcpy(&rows, &header[8], 8);
    std::memcpy(&cols, &header[16], 8);
    std::memcpy(&entries, &header[24], 8);
    if (magic == detail::binary_magic<double, int32>())
        return detail::read_binary_convert<double, int32, V, I>(is, rows, cols, entries);
    if (magic == detail::binary_magic<float, int32>())
        return detail::read_binary_convert<float, int32, V, I>(is, rows, cols, entries);
    if (magic == detail::binary_magic<double, int64>())
        return detail::read_binary_convert<double, int64, V, I>(is, rows, cols, entries);
    if (magic == detail::binary_magic<float, int64>())
        return detail::read_binary_convert<float, int64, V, I>(is, rows, cols, entries);
    throw StreamError("binary header has an unknown or unsupported (complex / half) magic number");
}

// gko::read_generic_raw: MatrixMarket if the stream starts with '%', binary otherwise
template <typename V = double, typename I = int32>
matrix_data<V, I> read_generic_raw(std::istream& is)
{
    const auto first = is.peek();
    detail::stream_check((bool)is, "failed reading from stream");
    return first == '%' ? read_raw<V, I>(is) : read_binary_raw<V, I>(is);
}

// gko::write_raw: "general" MatrixMarket file in the given layout
template <typename V, typename I>
void write_raw(std::ostream& os, const matrix_data<V, I>& data,
               layout_type layout = layout_type::coordinate)
{
    os << "%%MatrixMarket matrix " << (layout == layout_type::array ? "array" : "coordinate")
       << " real general\n";
    if (layout == layout_type::coordinate) {
        os << data.size.rows << ' ' << data.size.cols << ' ' << data.nonzeros.size() << '\n';
        for (const auto& nz : data.nonzeros)
            os << nz.row + 1 << ' ' << nz.column + 1 << ' ' << static_cast<double>(nz.value) << '\n';
    } else {
        auto nonzeros = data.nonzeros;
        using nt = typename matrix_data<V, I>::nonzero_type;
        std::sort(nonzeros.begin(), nonzeros.end(), [](const nt& x, const nt& y) {
            return std::tie(x.column, x.row) < std::tie(y.column, y.row);
        });
        size_type pos = 0;
        os << data.size.rows << ' ' << data.size.cols << '\n';
        for (size_type j = 0; j < data.size.cols; ++j)
            for (size_type i = 0; i < data.size.rows; ++i) {
                if (pos >= nonzeros.size() || (size_type)nonzeros[pos].row != i ||
                    (size_type)nonzeros[pos].column != j) {
                    os << static_cast<double>(V(0)) << '\n';
                } else {
                    os << static_cast<double>(nonzeros[pos].value) << '\n';
                    ++pos;
                }
            }
    }
    detail::stream_check((bool)os, "error when writing matrix data");
}

// gko::write_binary_raw
template <typename V, typename I>
void write_binary_raw(std::ostream& os, const matrix_data<V, I>& data)
{
    const std::uint64_t magic = detail::binary_magic<V, I>(), rows = data.size.rows,
                        cols = data.size.cols, entries = data.nonzeros.size();
    std::array<char, 32> header{};
    std::memcpy(&header[0], &magic, 8);
    std::memcpy(&header[8], &rows, 8);
    std::memcpy(&header[16], &cols, 8);
    std::memcpy(&header[24], &entries, 8);
    detail::stream_check((bool)os.write(header.data(), 32), "failed writing header");
    constexpr size_type block = sizeof(V) + 2 * sizeof(I);
    for (const auto& nz : data.nonzeros) {
        std::array<char, block> buf;
        std::memcpy(&buf[0], &nz.row, sizeof(I));
        std::memcpy(&buf[sizeof(I)], &nz.column, sizeof(I));
        std::memcpy(&buf[2 * sizeof(I)], &nz.value, sizeof(V));
        detail::stream_check((bool)os.write(buf.data(), block), "failed writing entry");
    }
    os.flush();
}

// matrix_data (row-major sorted) <-> CSR host arrays: the host half of Csr::read / Csr::write
// (core/matrix/csr.cpp:558-582 aos_to_soa + convert_idxs_to_ptrs, :633-650)
template <typename V, typename I>
void csr_arrays_from_matrix_data(const matrix_data<V, I>& data, std::vector<I>& row_ptrs,
                                 std::vector<I>& col_idxs, std::vector<V>& values)
{
    const size_type nnz = data.nonzeros.size();
    values.assign(nnz, V(0));
    col_idxs.assign(nnz, I(0));
    row_ptrs.assign(data.size.rows + 1, I(0));
    for (size_type k = 0; k < nnz; ++k) {
        const auto& e = data.nonzeros[k];
        if (e.row < 0 || (size_type)e.row >= data.size.rows || e.column < 0 ||
            (size_type)e.column >= data.size.cols)
            throw BadDimension("Csr::read: entry outside the matrix");
        if (k > 0 && e.row < data.nonzeros[k - 1].row)
            throw BadDimension("Csr::read: matrix_data is not in row-major order");
        values[k] = e.value;
        col_idxs[k] = e.column;
        ++row_ptrs[e.row + 1];
    }
    for (size_type r = 0; r < data.size.rows; ++r) row_ptrs[r + 1] += row_ptrs[r];
}

template <typename V, typename I>
matrix_data<V, I> matrix_data_from_csr_arrays(dim2 size, const std::vector<I>& row_ptrs,
                                              const std::vector<I>& col_idxs,
                                              const std::vector<V>& values)
{
    matrix_data<V, I> data(size);
    data.nonzeros.reserve(values.size());
    for (size_type row = 0; row < size.rows; ++row)
        for (auto k = row_ptrs[row]; k < row_ptrs[row + 1]; ++k)
            data.nonzeros.push_back({(I)row, col_idxs[k], values[k]});
    return data;
}

}  // namespace gko_b200
