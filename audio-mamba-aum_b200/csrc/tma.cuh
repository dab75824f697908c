// Host-side helper for cuTensorMapEncodeTiled (fetched through the runtime so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace aum {

// 2-D row-major tensor [rows, cols] with row pitch ld_elems; box = [box_rows x box_cols]; swizzle128: 128-byte swizzle
// (inner box must then be 128 bytes).  dt: aum_dtype.  Returns 0 on success.
int tma_encode_2d(CUtensorMap* tm, const void* base, int dt, int64_t rows, int64_t cols, int64_t ld_elems,
                  int box_rows, int box_cols, bool swizzle128, const char* what);
bool tma_available();

}  // namespace aum
