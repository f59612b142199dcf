// Definitions for the static members declared in the HDF5 stub header.
// TEST INFRASTRUCTURE ONLY (see H5Cpp.h in this directory).
#include "H5Cpp.h"
namespace H5 {
const PredType PredType::NATIVE_DOUBLE;
const PredType PredType::NATIVE_INT;
const PredType PredType::NATIVE_ULONG;
}
