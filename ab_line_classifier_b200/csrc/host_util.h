// Host-side helpers shared by the launchers: error codes, TMA descriptor encoding through the
// driver entry point (no link-time dependency on libcuda, so the library dlopen()s on a CPU box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abl_b200.h"

namespace abl {


int sm_count();
int check_launch();  // returns ABL_OK or ABL_ERR_CUDA (and keeps the message for abl_last_error)
void set_error(const char* msg);

// rank<=5 bf16/fp32 tiled tensor map; dims/strides innermost first, strides in BYTES for dims 1..rank-1.
int encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

}  // namespace abl
