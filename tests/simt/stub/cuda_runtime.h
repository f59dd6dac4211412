// TEST INFRASTRUCTURE ONLY -- stands in for <cuda_runtime.h> when kernel SOURCE is compiled for the host
// by tests/simt (see simt.h).  Only what the kernel translation units mention outside BF_SIMT_HOST guards.
#pragma once
#include "../simt.h"
