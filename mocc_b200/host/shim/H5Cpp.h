// No-op stand-in for the HDF5 C++ API surface that MOCC's H5Node wrapper
// (src/util/h5file.hpp:101-409, h5file.cpp:24-268) touches.
//
// BUILD SHIM: HDF5 is an external dependency of the reference
// (CMakeLists.txt:103) that is not installed in this image. Writes are
// discarded, reads throw (so any code path that needs real HDF5 input fails
// loudly instead of silently producing garbage). Results are compared
// in-process (oracle/ref_tool.cpp), never through .h5 files.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>

typedef unsigned long long hsize_t;
typedef std::string H5std_string;

#define H5F_ACC_RDONLY 0x0000u
#define H5F_ACC_RDWR 0x0001u
#define H5F_ACC_TRUNC 0x0002u
#define H5T_VARIABLE ((size_t)(-1))

enum H5G_link_t { H5G_LINK_HARD = 0, H5G_LINK_SOFT = 1 };
enum H5S_class_t { H5S_SCALAR = 0, H5S_SIMPLE = 1 };

namespace H5 {

class Exception : public std::runtime_error {
public:
    Exception(const std::string &what) : std::runtime_error(what)
    {
    }
    static void dontPrint()
    {
    }
};

class DataType {
public:
    virtual ~DataType()
    {
    }
};

class PredType : public DataType {
public:
    static const PredType NATIVE_DOUBLE;
    static const PredType NATIVE_INT;
    static const PredType NATIVE_ULONG;
};

class StrType : public DataType {
public:
    StrType()
    {
    }
    StrType(int, size_t)
    {
    }
};

class DataSpace {
public:
    DataSpace()
    {
    }
    DataSpace(H5S_class_t)
    {
    }
    DataSpace(int, const hsize_t *)
    {
    }
    int getSimpleExtentNdims() const
    {
        throw Exception("HDF5 stub: no readable datasets");
    }
    long long getSimpleExtentNpoints() const
    {
        throw Exception("HDF5 stub: no readable datasets");
    }
    int getSimpleExtentDims(hsize_t *) const
    {
        throw Exception("HDF5 stub: no readable datasets");
    }
};

class DataSet {
public:
    void write(const void *, const DataType &) const
    {
    }
    void write(const H5std_string &, const DataType &) const
    {
    }
    void read(void *, const DataType &) const
    {
        throw Exception("HDF5 stub: no readable datasets");
    }
    DataSpace getSpace() const
    {
        return DataSpace();
    }
};

class Group;

class CommonFG {
public:
    virtual ~CommonFG()
    {
    }
    inline Group createGroup(const std::string &) const;
    inline Group createGroup(const char *) const;
    inline Group openGroup(const std::string &) const;
    DataSet createDataSet(const std::string &, const DataType &,
                          const DataSpace &) const
    {
        return DataSet();
    }
    DataSet openDataSet(const std::string &path) const
    {
        throw Exception("HDF5 stub: cannot open dataset " + path);
    }
    void link(H5G_link_t, const char *, const char *) const
    {
    }
};

class Group : public CommonFG {
};

class H5File : public CommonFG {
public:
    H5File(const std::string &, unsigned int)
    {
    }
};

inline Group CommonFG::createGroup(const std::string &) const
{
    return Group();
}
inline Group CommonFG::createGroup(const char *) const
{
    return Group();
}
inline Group CommonFG::openGroup(const std::string &path) const
{
    throw Exception("HDF5 stub: cannot open group " + path);
}

} // namespace H5
