// arrayfile.hpp -- tiny named-array container (".mocflat") shared by the host
// plugin, the oracle tools and the Python side (mocc_b200/flatfile.py).
//
// Layout (little endian):
//   char[8]  "MOCFLAT1"
//   u32      number of arrays
//   per array:  u32 name length, name bytes, u32 dtype, u32 ndim, u64 dims[ndim],
//               zero padding to an 8-byte boundary, raw row-major data,
//               zero padding to an 8-byte boundary
// dtype: 0 = f64, 1 = i32, 2 = i64, 3 = u32
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mocc_b200 {

enum class DType : uint32_t { F64 = 0, I32 = 1, I64 = 2, U32 = 3 };

inline size_t dtype_size(DType t)
{
    return (t == DType::F64 || t == DType::I64) ? 8 : 4;
}

template <class T> struct dtype_of;
template <> struct dtype_of<double> {
    static const DType value = DType::F64;
};
template <> struct dtype_of<int32_t> {
    static const DType value = DType::I32;
};
template <> struct dtype_of<int64_t> {
    static const DType value = DType::I64;
};
template <> struct dtype_of<uint32_t> {
    static const DType value = DType::U32;
};

struct NamedArray {
    DType dtype;
    std::vector<uint64_t> dims;
    std::vector<unsigned char> bytes;

    size_t count() const
    {
        size_t n = 1;
        for (auto d : dims)
            n *= d;
        return n;
    }
    template <class T> const T *as() const
    {
        if (dtype_of<T>::value != dtype)
            throw std::runtime_error("arrayfile: dtype mismatch");
        return reinterpret_cast<const T *>(bytes.data());
    }
};

class ArrayFile {
public:
    template <class T>
    void put(const std::string &name, const T *data, std::vector<uint64_t> dims)
    {
        NamedArray a;
        a.dtype = dtype_of<T>::value;
        a.dims  = dims;
        size_t n = a.count();
        a.bytes.resize(n * sizeof(T));
        if (n)
            std::memcpy(a.bytes.data(), data, n * sizeof(T));
        if (!arrays_.count(name))
            order_.push_back(name);
        arrays_[name] = std::move(a);
    }
    template <class T> void put(const std::string &name, const std::vector<T> &v)
    {
        put(name, v.data(), {(uint64_t)v.size()});
    }
    template <class T> void put_scalar(const std::string &name, T v)
    {
        put(name, &v, {1});
    }
    bool has(const std::string &name) const
    {
        return arrays_.count(name) != 0;
    }
    const NamedArray &get(const std::string &name) const
    {
        auto it = arrays_.find(name);
        if (it == arrays_.end())
            throw std::runtime_error("arrayfile: no array named " + name);
        return it->second;
    }
    template <class T> T scalar(const std::string &name) const
    {
        return get(name).as<T>()[0];
    }
    const std::vector<std::string> &names() const
    {
        return order_;
    }

    void save(const std::string &path) const
    {
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f)
            throw std::runtime_error("arrayfile: cannot open for writing: " + path);
        const char magic[8] = {'M', 'O', 'C', 'F', 'L', 'A', 'T', '1'};
        std::fwrite(magic, 1, 8, f);
        uint32_t n = (uint32_t)order_.size();
        std::fwrite(&n, 4, 1, f);
        size_t pos = 12;
        auto pad   = [&]() {
            static const char zeros[8] = {0};
            size_t r = (8 - pos % 8) % 8;
            std::fwrite(zeros, 1, r, f);
            pos += r;
        };
        for (const auto &name : order_) {
            const NamedArray &a = arrays_.at(name);
            uint32_t len = (uint32_t)name.size();
            std::fwrite(&len, 4, 1, f);
            std::fwrite(name.data(), 1, len, f);
            uint32_t dt = (uint32_t)a.dtype, nd = (uint32_t)a.dims.size();
            std::fwrite(&dt, 4, 1, f);
            std::fwrite(&nd, 4, 1, f);
            std::fwrite(a.dims.data(), 8, nd, f);
            pos += 4 + len + 8 + 8 * (size_t)nd;
            pad();
            std::fwrite(a.bytes.data(), 1, a.bytes.size(), f);
            pos += a.bytes.size();
            pad();
        }
        std::fclose(f);
    }

    static ArrayFile load(const std::string &path)
    {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f)
            throw std::runtime_error("arrayfile: cannot open: " + path);
        ArrayFile af;
        char magic[8];
        size_t pos = 0;
        auto rd    = [&](void *p, size_t n) {
            if (std::fread(p, 1, n, f) != n) {
                std::fclose(f);
                throw std::runtime_error("arrayfile: truncated file: " + path);
            }
            pos += n;
        };
        auto skip_pad = [&]() {
            char tmp[8];
            size_t r = (8 - pos % 8) % 8;
            if (r)
                rd(tmp, r);
        };
        rd(magic, 8);
        if (std::memcmp(magic, "MOCFLAT1", 8) != 0) {
            std::fclose(f);
            throw std::runtime_error("arrayfile: bad magic: " + path);
        }
        uint32_t n;
        rd(&n, 4);
        for (uint32_t i = 0; i < n; i++) {
            uint32_t len, dt, nd;
            rd(&len, 4);
            std::string name(len, '\0');
            rd(&name[0], len);
            rd(&dt, 4);
            rd(&nd, 4);
            NamedArray a;
            a.dtype = (DType)dt;
            a.dims.resize(nd);
            rd(a.dims.data(), 8 * (size_t)nd);
            skip_pad();
            a.bytes.resize(a.count() * dtype_size(a.dtype));
            if (!a.bytes.empty())
                rd(a.bytes.data(), a.bytes.size());
            skip_pad();
            af.order_.push_back(name);
            af.arrays_[name] = std::move(a);
        }
        std::fclose(f);
        return af;
    }

private:
    std::map<std::string, NamedArray> arrays_;
    std::vector<std::string> order_;
};
}
