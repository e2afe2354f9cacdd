// wg_grid.h — grid containers and UNF binary I/O of the host side.
//
// Mirrors the interface of the reference's Grid<T> / VariableChannelGrid<C,T> (grid.h:58-474)
// and the on-disk contract of GridUNFAdapter (grid_io_adapters.h:56-127): big-endian raw
// arrays, element type given by the file suffix (UNF0 f32, UNF1 i8, UNF2 i16, UNF4 i32),
// multi-channel grids stored cell-major [cell][channel], and Grid<double> stored as float32.
// The cell count is a run-time value here (the reference hard-codes ng = 67420, def.h:10).
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace wg {

template <class T> inline T byteswap(T v) {
    unsigned char b[sizeof(T)];
    std::memcpy(b, &v, sizeof(T));
    for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
    std::memcpy(&v, b, sizeof(T));
    return v;
}

template <class T> void read_unf_raw(const std::string &file, T *dst, size_t n) {
    std::ifstream f(file, std::ios::binary);
    if (!f) throw std::runtime_error(file + " not found.");  // grid_io_adapters.h:94-98 exits here
    f.read(reinterpret_cast<char *>(dst), (std::streamsize)(n * sizeof(T)));
    if ((size_t)f.gcount() != n * sizeof(T)) throw std::runtime_error(file + ": short read");
    if (sizeof(T) > 1)
        for (size_t i = 0; i < n; i++) dst[i] = byteswap(dst[i]);
}

template <class T> void write_unf_raw(const std::string &file, const T *src, size_t n) {
    std::ofstream f(file, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + file);
    std::vector<T> tmp(src, src + n);
    if (sizeof(T) > 1)
        for (auto &v : tmp) v = byteswap(v);
    f.write(reinterpret_cast<const char *>(tmp.data()), (std::streamsize)(n * sizeof(T)));
}

// Channels == 1: plain per-cell grid.  operator()(cell, channel) as grid.h:471-474.
template <class T = double, int Channels = 1> class Grid {
  public:
    Grid() = default;
    explicit Grid(size_t ncell) { initialize(ncell); }
    void initialize(size_t ncell) { ncell_ = ncell; data_.assign(ncell * Channels, T()); }
    size_t cells() const { return ncell_; }
    size_t size() const { return data_.size(); }
    bool initialized() const { return !data_.empty(); }
    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }
    T &operator[](size_t i) { return data_[i]; }
    const T &operator[](size_t i) const { return data_[i]; }
    T &operator()(size_t cell, size_t channel) { return data_[cell * Channels + channel]; }
    const T &operator()(size_t cell, size_t channel) const { return data_[cell * Channels + channel]; }
    void fill(T v) { std::fill(data_.begin(), data_.end(), v); }
    // UNF: doubles live on disk as float32 (grid_io_adapters.h:64-72, 100-115)
    void read(const std::string &file) {
        if (std::is_same<T, double>::value) {
            std::vector<float> tmp(data_.size());
            read_unf_raw(file, tmp.data(), tmp.size());
            for (size_t i = 0; i < tmp.size(); i++) data_[i] = static_cast<T>(tmp[i]);
        } else {
            read_unf_raw(file, data_.data(), data_.size());
        }
    }
    void write(const std::string &file) const {
        if (std::is_same<T, double>::value) {
            std::vector<float> tmp(data_.size());
            for (size_t i = 0; i < tmp.size(); i++) tmp[i] = static_cast<float>(data_[i]);
            write_unf_raw(file, tmp.data(), tmp.size());
        } else {
            write_unf_raw(file, data_.data(), data_.size());
        }
    }

  private:
    size_t ncell_ = 0;
    std::vector<T> data_;
};

}  // namespace wg
