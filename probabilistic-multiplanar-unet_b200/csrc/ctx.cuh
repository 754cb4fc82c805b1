// pmu_ctx — per-device launch context of the C-ABI (include/pmu_b200.h: pmu_ctx_create / destroy / bind / stats).
//
// What a launch of the tcgen05 kernels needs besides its arguments is host-side state that does not change from call to
// call: the device's SM count and compute capability, the driver's cuTensorMapEncodeTiled entry point, the dynamic
// shared-memory attribute of each kernel, and the TMA descriptors — a pure function of (pointer, extents, strides, box).
// Without a context every convolution launch asks the runtime for the device, re-sets the kernel attribute and encodes up
// to seven tensor maps (2 130 launches per 256^3 volume).  A context bound to the calling thread caches all of it:
// descriptors are looked up by their defining tuple (torch's caching allocator hands the same addresses back every
// step, so in steady state every lookup hits); the cache is bounded and simply cleared when full.
#pragma once

#include <cudaTypedefs.h>

#include "pmu_common.cuh"

namespace pmu {

struct TensorMapSpec {
  const void* ptr;
  uint64_t dims[4];
  uint64_t strides[3];      // bytes, dims 1..rank-1
  uint32_t box[4];
  int rank;
  int dtype;                // CUtensorMapDataType
  int swizzle;              // CUtensorMapSwizzle
  int l2promo;              // CUtensorMapL2promotion
  int oob;                  // CUtensorMapFloatOOBfill
};

// Encode (or fetch from the bound context's cache) the tiled tensor map `spec` describes; element strides 1, no interleave.
// `what` names the tensor in the error message.
int tensor_map(CUtensorMap* out, const TensorMapSpec& spec, const char* what);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), once per kernel and context
int set_max_dyn_smem(const void* kernel, int bytes);
// compute capability major of the current device (cached in the bound context)
int device_cc_major(int* cc_major);

}  // namespace pmu
